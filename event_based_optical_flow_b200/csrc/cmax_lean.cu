// Strip kernels: the event kernels of the CM iteration for batches the plan could cut into STRIPS (cmax_events.cu,
// pack_strips_kernel): events in source-pixel order, every pixel's run padded to whole strips of 8 events, so a strip
// never mixes two source pixels.  One thread walks one strip; a warp-tile of 32 strips (1152 bytes: 32 headers + 256
// times) arrives by one TMA bulk copy, double buffered per warp.
//
// Why: the run kernels of cmax_fused.cu are ISSUE bound (profiles/README.md: ~67 (K1) / ~87 (K3) SASS instructions per
// event, 66 % issue-active, DRAM at 30 %), and over a third of those instructions serve something that happens once in
// ~50 events per lane -- noticing that the source pixel changed and re-deriving its float coordinates, flat index and
// flow vector.  ptxas if-converts that block, so it is issued for every event.  With strips the source pixel is a
// per-strip constant: its set-up runs once per 8 events, unconditionally, with both flow loads in flight together; an
// event is a 4-byte time (4.5 bytes per event with the header instead of 8), fetched four at a time with LDS.128.  On
// top of that the warp is 2 scalar multiplies + 1 packed subtract (the multiplies must stay scalar, see mul2_rounded),
// the four bilinear weights are accumulated with two FFMA2, and K3 reads the gradient quad of an out-of-image or padding
// event from an all-zero extra cell, which removes every validity select from its accumulation.
// Results equal the run kernels' up to fp32 summation order; warped coordinates stay bit-exact.
#include <stddef.h>

#include "cmax_runs.cuh"

namespace cmax {

// (a.lo * b, a.hi * b) with each product rounded on its own.  The products stay SCALAR on purpose: ptxas contracts a packed
// multiply (mul.rn.f32x2, and even fma.rn.f32x2 with a -0 addend) with the packed add / sub that consumes it into one
// FFMA2 (observed, CUDA 12.9, despite -fmad=false), which would skip the rounding of dt * f that the reference performs;
// __fmul_rn is never contracted.
__device__ __forceinline__ f32x2 mul2_rounded(f32x2 a, float b) {
  float lo, hi;
  upk2(a, lo, hi);
  return pk2(__fmul_rn(lo, b), __fmul_rn(hi, b));
}
// v = pred ? 0 : v, in place
__device__ __forceinline__ void clear_if(bool pred, f32x2& v) {
  asm("{\n .reg .pred p;\n setp.ne.u32 p, %1, 0;\n @p mov.b64 %0, 0;\n}" : "+l"(v) : "r"((unsigned)pred));
}

__device__ __forceinline__ float small_int_to_float(unsigned v) { return __fsub_rn(__uint_as_float(0x4B000000u | v), 8388608.0f); }

// One warp-tile of strips as it lands in shared memory.
struct StripTile {
  uint32_t head[32];   // row << 17 | col << 4 | count (0..8 events of this strip are real)
  float4 t[2][32];     // t[h][lane] = times 4h .. 4h+3 of lane's strip (normalised dt for a single reference time)
  uint2 bins[CMAX_MAX_REFS][32];  // time-aware plans only, first n_ref entries: byte k of bins[r][lane] = time bin of event k (255: none)
};
static_assert(offsetof(StripTile, bins) == kStripTileBytes && sizeof(uint2) * 32 == kStripBinBytes, "strip tile layout");
// bytes of one warp-tile as the plan packed it for this kernel instantiation
template <int MODEL, int NREF>
struct StripTileBytes {
  static constexpr uint32_t value = kStripTileBytes + (MODEL == CMAX_MOTION_VOXEL ? NREF * kStripBinBytes : 0);
};

// Per-strip constants.
struct StripHead {
  int count, src, row, col;
  f32x2 xy;
};
__device__ __forceinline__ StripHead strip_head(uint32_t h, int W) {
  StripHead s;
  const unsigned row = h >> 17, col = (h >> 4) & 0x1FFFu;
  s.row = (int)row;
  s.col = (int)col;
  s.count = (int)(h & 0xFu);
  s.src = (int)(row * (unsigned)W + col);
  s.xy = pk2(small_int_to_float(row), small_int_to_float(col));
  return s;
}

// ((row + 1, col + 1) of the floor pixel in the padded image, fractions) of a warped coordinate pair
struct LeanGeom {
  int ri, ci;
  f32x2 fr;
};
__device__ __forceinline__ LeanGeom lean_geometry(f32x2 w, int off_r, int off_c) {
  const f32x2 bias = pk2(kFloorBias, kFloorBias);
  const f32x2 t = add2_rd(add2(w, pk2(1e-6f, 1e-6f)), bias);  // floor(x' + 1e-6) in the mantissa   event_image_converter.py:340-345
  LeanGeom g;
  g.fr = sub2(w, sub2(t, bias));
  float tx, ty;
  upk2(t, tx, ty);
  g.ri = __float_as_int(tx) + off_r;  // off = pad + 1 - 0x4B400000
  g.ci = __float_as_int(ty) + off_c;
  return g;
}
// accumulator cell of the vote, or `outside` when the event touches no pixel / is strip padding (k >= count).
// Three chained predicate compares and ONE select (ptxas turns the C expression into three selects).
__device__ __forceinline__ int lean_cell(const LeanGeom& g, int k, int count, int Hp, int Wp, int outside) {
  int c;
  const int idx = g.ri * (Wp + 1) + g.ci;
  asm("{\n .reg .pred p;\n setp.le.u32 p, %1, %2;\n setp.le.and.u32 p, %3, %4, p;\n setp.lt.and.s32 p, %5, %6, p;\n selp.s32 %0, %7, %8, p;\n}"
      : "=r"(c)
      : "r"(g.ri), "r"(Hp), "r"(g.ci), "r"(Wp), "r"(k), "r"(count), "r"(idx), "r"(outside));
  return c;
}
// warped coordinate pair of one event for reference time r (dense / 2-dof: the flow vector f is a per-strip constant)
template <int MODEL, int NREF, bool PRE_DT>
__device__ __forceinline__ f32x2 lean_warp(f32x2 xy, float tz, f32x2 f, const RefRegs<NREF>& rr, int r, float& dt) {
  dt = PRE_DT ? tz : __fdiv_rn(__fsub_rn(tz, rr.ref[r]), rr.period[r]);
  if (MODEL == CMAX_MOTION_2DOF) return add2(xy, mul2_rounded(f, dt));  // src/warp.py:507-514
  return sub2(xy, mul2_rounded(f, dt));                                   // src/warp.py:306-307 (dense flow, or the tile flow at this pixel)
}

// Tile flow (CMAX_MOTION_TILE): the dense flow is never materialised -- every strip evaluates the patch grid at ITS source
// pixel with the up-sampling kernel's own expression (tile_value, bit-identical), times t_scale (the reference multiplies the
// up-sampled flow by it, src/solver/patch_contrast_pyramid.py:452-453).  The <= 2 KB grid stays in L1.
constexpr int kTileMaxNodes = 1024;  // patch grids up to 32 x 32
__device__ __forceinline__ f32x2 tile_flow_at(const FusedArgs& a, const StripHead& h, TileTaps& t) {
  t = tile_taps(a.tile, h.row, h.col);
  const int np = a.tile.hp * a.tile.wp;
  return pk2(__fmul_rn(tile_value(a.motion, a.tile, t), a.t_scale), __fmul_rn(tile_value(a.motion + np, a.tile, t), a.t_scale));
}

// Time-aware (voxel) model: the flow vector depends on the event's time bin (src/warp.py:346-357).  The bin of every event
// for every reference time is iteration-invariant, so the plan packed it next to the times (one byte each); the events
// of a strip are in time order, so the bin changes at most a few times along a strip and the flow vector is re-fetched
// only then.
struct VoxelWalk {
  int bin;    // -1 = none yet / in no bin
  f32x2 f;
};
template <int NREF, bool PRE_DT>
__device__ __forceinline__ f32x2 voxel_warp(f32x2 xy, float tz, int k, bool real, uint2 bins, int src, int HW, const float* __restrict__ motion,
                                            const RefRegs<NREF>& rr, int r, VoxelWalk& vw, float& dt) {
  dt = PRE_DT ? tz : __fdiv_rn(__fsub_rn(tz, rr.ref[r]), rr.period[r]);
  const unsigned word = k < 4 ? bins.x : bins.y;
  int b = (int)((word >> (8 * (k & 3))) & 0xFFu);
  b = (b == 255) ? -1 : b;
  if (real && b != vw.bin) {  // (strip padding is masked by the caller: leave the walk alone)
    vw.bin = b;
    if (b >= 0) {
      const float* fb = motion + (int64_t)b * 2 * HW;
      vw.f = pk2(__ldg(fb + src), __ldg(fb + HW + src));
    }
  }
  return vw.bin >= 0 ? sub2(xy, mul2_rounded(vw.f, dt)) : xy;
}

// ------------------------------------------------------------------------------------------------ K1
template <int NREF>
struct StripVote {
  int cell[NREF];
  f32x2 w01[NREF], w23[NREF];  // (w00, w10), (w01, w11) of the current accumulator cell
  VoxelWalk vw[NREF];          // voxel model only
};

template <int MODEL, int NREF, bool PRE_DT>
__device__ __forceinline__ void strip_vote_step(float tz, int k, const StripHead& h, f32x2 f, const uint2 (&bins)[NREF], StripVote<NREF>& st,
                                                const FusedArgs& a, int HW, int off_r, int off_c, const RefRegs<NREF>& rr,
                                                float4* __restrict__ acc) {
#pragma unroll
  for (int r = 0; r < NREF; ++r) {
    float dt;
    f32x2 w;
    if constexpr (MODEL == CMAX_MOTION_VOXEL) w = voxel_warp<NREF, PRE_DT>(h.xy, tz, k, k < h.count, bins[r], h.src, HW, a.motion, rr, r, st.vw[r], dt);
    else w = lean_warp<MODEL, NREF, PRE_DT>(h.xy, tz, f, rr, r, dt);
    const LeanGeom g = lean_geometry(w, off_r, off_c);
    const int c = lean_cell(g, k, h.count, a.Hp, a.Wp, -1);
    const bool change = c != st.cell[r];
    {  // a new accumulator cell: flush the old one (predicated)
      float w0, w1, w2, w3;
      upk2(st.w01[r], w0, w1);
      upk2(st.w23[r], w2, w3);
      red_add_v4_if(change && st.cell[r] >= 0, acc + r * a.cells + st.cell[r], w0, w1, w2, w3);
      st.cell[r] = c;
      clear_if(change, st.w01[r]);  // (ptxas has no predicated packed arithmetic: a predicated FMUL2 / FFMA2 pair becomes
      clear_if(change, st.w23[r]);  //  both operations plus four selects, measured in SASS -- clearing is cheaper)
    }
    // (w00, w10) += (1-fx, fx) * (1-fy);  (w01, w11) += (1-fx, fx) * fy        event_image_converter.py:365-369
    // (what an out-of-image event adds here is dropped by the next cell change: a cell of -1 is never flushed)
    float fx, fy;
    upk2(g.fr, fx, fy);
    const f32x2 p = pk2(__fsub_rn(1.0f, fx), fx);
    const float ay = __fsub_rn(1.0f, fy);
    st.w01[r] = fma2(p, pk2(ay, ay), st.w01[r]);
    st.w23[r] = fma2(p, pk2(fy, fy), st.w23[r]);
  }
}

template <int MODEL, int NREF, bool PRE_DT>
__global__ void __launch_bounds__(kRunThreads) vote_strips_kernel(FusedArgs a, float4* __restrict__ acc) {
  constexpr uint32_t kTile = StripTileBytes<MODEL, NREF>::value;
  __shared__ TilePipe<kTile> pipes[kRunWarps];
  if (a.zero256 != nullptr && blockIdx.x == 0 && threadIdx.x < 64) a.zero256[threadIdx.x] = 0u;  // StatAcc block + CTA counter
  const RefRegs<NREF> rr = load_refs<NREF>(a.tp);
  const int HW = a.H * a.W;
  const int lane = threadIdx.x & 31;
  const int off_r = a.pad_h + 1 - 0x4B400000, off_c = a.pad_w + 1 - 0x4B400000;
  TilePipe<kTile>& pipe = pipes[threadIdx.x >> 5];
  pipe_init(pipe, lane);
  const int64_t n_tiles = (a.n_strips + 31) / 32;
  const int64_t warp0 = (int64_t)blockIdx.x * kRunWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kRunWarps;
  if (warp0 < n_tiles) pipe_issue(pipe, 0, a.strips, warp0, lane);
  pdl_trigger();  // the fold may be scheduled as soon as this grid drains
  f32x2 f = 0ull;
  if (MODEL == CMAX_MOTION_2DOF) f = pk2(__ldg(a.motion), __ldg(a.motion + 1));
  int it = 0;
  for (int64_t tile = warp0; tile < n_tiles; tile += n_warps, ++it) {
    const int stage = it & 1;
    __syncwarp();  // every lane is done with the other buffer (walked in the previous iteration)
    if (tile + n_warps < n_tiles) pipe_issue(pipe, stage ^ 1, a.strips, tile + n_warps, lane);
    mbar_wait(&pipe.bar[stage], (it >> 1) & 1);
    const StripTile* T = reinterpret_cast<const StripTile*>(pipe.buf[stage]);
    const StripHead h = strip_head(T->head[lane], a.W);
    if (MODEL == CMAX_MOTION_DENSE) f = pk2(__ldg(a.motion + h.src), __ldg(a.motion + HW + h.src));
    if (MODEL == CMAX_MOTION_TILE) {
      TileTaps taps;
      f = tile_flow_at(a, h, taps);
    }
    const float4 ta = T->t[0][lane], tb = T->t[1][lane];
    const float tz[kRunE] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
    uint2 bins[NREF];
#pragma unroll
    for (int r = 0; r < NREF; ++r) bins[r] = (MODEL == CMAX_MOTION_VOXEL) ? T->bins[r][lane] : make_uint2(0u, 0u);
    StripVote<NREF> st;
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      st.cell[r] = -1;
      st.w01[r] = st.w23[r] = 0ull;
      st.vw[r].bin = -1;
      st.vw[r].f = 0ull;
    }
#pragma unroll
    for (int k = 0; k < kRunE; ++k) strip_vote_step<MODEL, NREF, PRE_DT>(tz[k], k, h, f, bins, st, a, HW, off_r, off_c, rr, acc);
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      float w0, w1, w2, w3;
      upk2(st.w01[r], w0, w1);
      upk2(st.w23[r], w2, w3);
      red_add_v4_if(st.cell[r] >= 0, acc + r * a.cells + st.cell[r], w0, w1, w2, w3);
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3
template <int MODEL, int NREF>
struct StripGrad {
  static constexpr int NACC = (MODEL == CMAX_MOTION_VOXEL) ? NREF : 1;
  int cell[NREF];
  float d_x0[NREF], d_c0[NREF], d_r[NREF];  // corner differences of the current cell's gradient quad
  int slot[NACC];                           // voxel model: flat index of the (bin, pixel) slot being accumulated, -1 = none
  f32x2 g[NACC];                            // sum of -dt * (dL/dx', dL/dy')
  VoxelWalk vw[NREF];                       // voxel model only
};

template <int MODEL, int NREF, bool PRE_DT>
__device__ __forceinline__ void strip_grad_step(float tz, int k, const StripHead& h, f32x2 f, const uint2 (&bins)[NREF],
                                                StripGrad<MODEL, NREF>& st, const FusedArgs& a, int HW, int off_r, int off_c, int outside,
                                                const RefRegs<NREF>& rr, const float4* __restrict__ gq, float* __restrict__ gmotion) {
#pragma unroll
  for (int r = 0; r < NREF; ++r) {
    float dt;
    f32x2 w;
    if constexpr (MODEL == CMAX_MOTION_VOXEL) w = voxel_warp<NREF, PRE_DT>(h.xy, tz, k, k < h.count, bins[r], h.src, HW, a.motion, rr, r, st.vw[r], dt);
    else w = lean_warp<MODEL, NREF, PRE_DT>(h.xy, tz, f, rr, r, dt);
    const int bin = (MODEL == CMAX_MOTION_VOXEL) ? st.vw[r].bin : 0;
    const LeanGeom g = lean_geometry(w, off_r, off_c);
    const int c = lean_cell(g, k, h.count, a.Hp, a.Wp, outside);  // `outside` = the extra all-zero cell
    if (c != st.cell[r]) {
      st.cell[r] = c;
      const float4 q = __ldg(gq + r * a.cells + c);
      // dL/dx' = (1-fy)(g10-g00) + fy(g11-g01) = d_x0 + fy*d_r ;  dL/dy' = (1-fx)(g01-g00) + fx(g11-g10) = d_c0 + fx*d_r
      st.d_x0[r] = q.y - q.x;
      st.d_c0[r] = q.z - q.x;
      st.d_r[r] = (q.w - q.z) - st.d_x0[r];
    }
    float fx, fy;
    upk2(g.fr, fx, fy);
    const f32x2 d = pk2(fmaf(fy, st.d_r[r], st.d_x0[r]), fmaf(fx, st.d_r[r], st.d_c0[r]));
    const float ndt = -dt;
    if (MODEL == CMAX_MOTION_VOXEL) {
      const int slot = (bin >= 0 && k < h.count) ? bin * 2 * HW + h.src : -1;
      if (slot != st.slot[r]) {
        if (st.slot[r] >= 0) {
          float g0, g1;
          upk2(st.g[r], g0, g1);
          atomicAdd(gmotion + st.slot[r], g0);
          atomicAdd(gmotion + st.slot[r] + HW, g1);
        }
        st.slot[r] = slot;
        st.g[r] = 0ull;
      }
      st.g[r] = fma2(pk2(ndt, ndt), d, st.g[r]);
    } else {
      st.g[0] = fma2(pk2(ndt, ndt), d, st.g[0]);
    }
  }
}

template <int MODEL, int NREF, bool PRE_DT>
__global__ void __launch_bounds__(kRunThreads) grad_strips_kernel(FusedArgs a, const float4* __restrict__ gq, float* __restrict__ gmotion) {
  constexpr uint32_t kTile = StripTileBytes<MODEL, NREF>::value;
  __shared__ TilePipe<kTile> pipes[kRunWarps];
  __shared__ double red2[2][kRunWarps];
  // tile flow: dL/d(patch grid) of this CTA (the adjoint of the up-sampling: every source pixel hands its flow gradient to
  // its <= 4 grid nodes).  A CTA of this instantiation walks a CONTIGUOUS range of warp-tiles, i.e. a few neighbouring pixels'
  // worth of strips, so it touches a handful of nodes: they are summed here and only the non-zero ones go to global memory.
  // (The same reductions issued per pixel, or per CTA over strided tiles, serialise in the L2 at ~100 ns per same-address
  // operation: 757 k of them onto 512 addresses made a 200 us kernel -- measured.)
  __shared__ float sgrad[MODEL == CMAX_MOTION_TILE ? 2 * kTileMaxNodes : 1];
  const RefRegs<NREF> rr = load_refs<NREF>(a.tp);
  const int HW = a.H * a.W;
  const int lane = threadIdx.x & 31;
  const int off_r = a.pad_h + 1 - 0x4B400000, off_c = a.pad_w + 1 - 0x4B400000;
  const int outside = (int)a.cells - 1;
  const int np = a.tile.hp * a.tile.wp;
  if (MODEL == CMAX_MOTION_TILE) {
    for (int k = threadIdx.x; k < 2 * np; k += kRunThreads) sgrad[k] = 0.f;
    __syncthreads();
  }
  TilePipe<kTile>& pipe = pipes[threadIdx.x >> 5];
  pipe_init(pipe, lane);
  // warps stride over all tiles (dense / voxel / 2-dof), or over this CTA's contiguous range of tiles (tile flow)
  int64_t n_tiles = (a.n_strips + 31) / 32;
  int64_t warp0 = (int64_t)blockIdx.x * kRunWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kRunWarps;
  if (MODEL == CMAX_MOTION_TILE) {
    const int64_t per_cta = (n_tiles + gridDim.x - 1) / gridDim.x;
    const int64_t begin = (int64_t)blockIdx.x * per_cta;
    n_tiles = min(n_tiles, begin + per_cta);
    warp0 = begin + (threadIdx.x >> 5);
    n_warps = kRunWarps;
  }
  if (warp0 < n_tiles) pipe_issue(pipe, 0, a.strips, warp0, lane);  // the strips do not depend on the predecessor
  f32x2 f = 0ull;
  if (MODEL == CMAX_MOTION_2DOF) f = pk2(__ldg(a.motion), __ldg(a.motion + 1));
  double t0 = 0.0, t1 = 0.0;  // 2-dof: per-thread fp64 sums over the strips
  pdl_wait();  // gradient quads (and the zeroed gradient buffer) of the predecessor kernel are complete
  int it = 0;
  for (int64_t tile = warp0; tile < n_tiles; tile += n_warps, ++it) {
    const int stage = it & 1;
    __syncwarp();
    if (tile + n_warps < n_tiles) pipe_issue(pipe, stage ^ 1, a.strips, tile + n_warps, lane);
    mbar_wait(&pipe.bar[stage], (it >> 1) & 1);
    const StripTile* T = reinterpret_cast<const StripTile*>(pipe.buf[stage]);
    const StripHead h = strip_head(T->head[lane], a.W);
    if (MODEL == CMAX_MOTION_DENSE) f = pk2(__ldg(a.motion + h.src), __ldg(a.motion + HW + h.src));
    TileTaps taps;
    if (MODEL == CMAX_MOTION_TILE) f = tile_flow_at(a, h, taps);
    const float4 ta = T->t[0][lane], tb = T->t[1][lane];
    const float tz[kRunE] = {ta.x, ta.y, ta.z, ta.w, tb.x, tb.y, tb.z, tb.w};
    uint2 bins[NREF];
#pragma unroll
    for (int r = 0; r < NREF; ++r) bins[r] = (MODEL == CMAX_MOTION_VOXEL) ? T->bins[r][lane] : make_uint2(0u, 0u);
    StripGrad<MODEL, NREF> st;
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      st.cell[r] = -2;
      st.d_x0[r] = st.d_c0[r] = st.d_r[r] = 0.f;
      st.vw[r].bin = -1;
      st.vw[r].f = 0ull;
    }
#pragma unroll
    for (int q = 0; q < StripGrad<MODEL, NREF>::NACC; ++q) {
      st.slot[q] = -1;
      st.g[q] = 0ull;
    }
#pragma unroll
    for (int k = 0; k < kRunE; ++k)
      strip_grad_step<MODEL, NREF, PRE_DT>(tz[k], k, h, f, bins, st, a, HW, off_r, off_c, outside, rr, gq, gmotion);
    if (MODEL == CMAX_MOTION_DENSE) {  // one flush per strip: the strip IS one source pixel
      float g0, g1;
      upk2(st.g[0], g0, g1);
      if (a.seg_reduce) {
        // Dense batches (tens of strips per pixel, e.g. a spatially compact shard of a multi-GPU run): the lanes of a warp
        // hold consecutive strips of the SAME source pixel, and 32 reductions onto one address serialise in the L2 (measured:
        // K3 11 -> 19 us at 440 events per pixel).  Segmented suffix sum over the runs of equal pixel inside the warp, one
        // pair of reductions per run.  (warp-uniform branch; sparse batches skip the ~30 shuffle instructions)
        const int key = h.count > 0 ? h.src : -1 - lane;  // padding strips never merge
        const int prev = __shfl_up_sync(0xffffffffu, key, 1);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float u0 = __shfl_down_sync(0xffffffffu, g0, o), u1 = __shfl_down_sync(0xffffffffu, g1, o);
          const int uk = __shfl_down_sync(0xffffffffu, key, o);
          if (lane + o < 32 && uk == key) {
            g0 += u0;
            g1 += u1;
          }
        }
        if (h.count > 0 && (lane == 0 || prev != key)) {
          atomicAdd(gmotion + h.src, g0);
          atomicAdd(gmotion + HW + h.src, g1);
        }
      } else if (h.count > 0) {
        atomicAdd(gmotion + h.src, g0);
        atomicAdd(gmotion + HW + h.src, g1);
      }
    } else if (MODEL == CMAX_MOTION_TILE) {
      // dense[c] = -t_scale * sum_ab Wr[a] Wc[b] m[c,a,b]  =>  dL/dm[c,a,b] += -t_scale * Wr[a] Wc[b] * dL/ddense[c].
      // (1) the strips of one source pixel are consecutive lanes: segmented suffix sum of their flow gradients;
      float g0, g1;
      upk2(st.g[0], g0, g1);
      const int key = h.count > 0 ? h.src : -1 - lane;  // padding strips never merge
      const int prev = __shfl_up_sync(0xffffffffu, key, 1);
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float u0 = __shfl_down_sync(0xffffffffu, g0, o), u1 = __shfl_down_sync(0xffffffffu, g1, o);
        const int uk = __shfl_down_sync(0xffffffffu, key, o);
        if (lane + o < 32 && uk == key) {
          g0 += u0;
          g1 += u1;
        }
      }
      const bool head = h.count > 0 && (lane == 0 || prev != key);
      // (2) the head of every pixel run turns the pixel's gradient into its 8 node contributions; the ~5 pixels of a warp-tile
      // nearly always sit in ONE patch cell (same 4 nodes): then the 8 values are summed over the warp and one lane adds them
      // to the CTA's accumulators (a float atomicAdd in shared memory is a compare-and-swap loop: keep it uncontended)
      g0 *= -a.t_scale;
      g1 *= -a.t_scale;
      const float wr[2] = {1.0f - taps.lr, taps.lr}, wc[2] = {1.0f - taps.lc, taps.lc};
      const int an[2] = {taps.a0, taps.a1}, bn[2] = {taps.b0, taps.b1};
      float c[8];
      int node[4];
#pragma unroll
      for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const float w = head ? wr[u] * wc[v] : 0.f;
          node[2 * u + v] = an[u] * a.tile.wp + bn[v];
          c[2 * u + v] = w * g0;
          c[4 + 2 * u + v] = w * g1;
        }
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      if (heads != 0u) {
        const int first = __ffs(heads) - 1;
        const int cell_key = (taps.a0 << 20) | (taps.a1 << 10) | 0;  // rows; columns compared separately (grids up to 1024 nodes)
        const int col_key = (taps.b0 << 10) | taps.b1;
        // (both shuffles unconditionally: a short-circuited && would leave lanes out of the second full-mask shuffle)
        const int first_cell = __shfl_sync(0xffffffffu, cell_key, first), first_col = __shfl_sync(0xffffffffu, col_key, first);
        const bool same = (cell_key == first_cell) & (col_key == first_col);
        if (__all_sync(0xffffffffu, !head || same)) {
#pragma unroll
          for (int k = 0; k < 8; ++k) c[k] = warp_sum(c[k]);
          if (lane == first) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              atomicAdd(&sgrad[node[k]], c[k]);
              atomicAdd(&sgrad[np + node[k]], c[4 + k]);
            }
          }
        } else if (head) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            atomicAdd(&sgrad[node[k]], c[k]);
            atomicAdd(&sgrad[np + node[k]], c[4 + k]);
          }
        }
      }
    } else if (MODEL == CMAX_MOTION_VOXEL) {
#pragma unroll
      for (int q = 0; q < NREF; ++q) {
        if (st.slot[q] >= 0) {
          float g0, g1;
          upk2(st.g[q], g0, g1);
          atomicAdd(gmotion + st.slot[q], g0);
          atomicAdd(gmotion + st.slot[q] + HW, g1);
        }
      }
    } else {  // d/d theta of x + dt * theta: the sign of the 2-dof warp is opposite to the flow's
      float g0, g1;
      upk2(st.g[0], g0, g1);
      t0 -= (double)g0;
      t1 -= (double)g1;
    }
  }
  if (MODEL == CMAX_MOTION_TILE) {
    __syncthreads();
    for (int k = threadIdx.x; k < 2 * np; k += kRunThreads) {
      const float v = sgrad[k];
      if (v != 0.f) atomicAdd(gmotion + k, v);
    }
  }
  if (MODEL == CMAX_MOTION_2DOF) {
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    const int wid = threadIdx.x >> 5;
    if (lane == 0) {
      red2[0][wid] = t0;
      red2[1][wid] = t1;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      double tot = 0.0;
      for (int w = 0; w < kRunWarps; ++w) tot += red2[threadIdx.x][w];
      atomicAdd(reinterpret_cast<double*>(gmotion) + threadIdx.x, tot);  // fp64 staging, narrowed by finish_2dof_kernel
    }
  }
}

// ------------------------------------------------------------------------------------------------ dispatch
static int strips_grid_size(int per_sm, int64_t n_strips) {
  // persistent: exactly the CTAs that are resident at once (2 x that / one tile per warp measured no better: profiles/README.md)
  const int64_t tiles = (n_strips + 31) / 32, ctas = (tiles + kRunWarps - 1) / kRunWarps;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)num_sms() * per_sm));
}
template <typename K>
static int strips_grid(K kernel, int64_t n_strips) {
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kRunThreads, 0) != cudaSuccess || v < 1) v = 4;
    per_sm = v;
  }
  return strips_grid_size(per_sm, n_strips);
}

template <int MODEL, int NREF>
static void vote_strips_mn(cudaStream_t s, const FusedArgs& a, float4* acc) {
  auto k = vote_strips_kernel<MODEL, NREF, NREF == 1>;
  k<<<strips_grid(k, a.n_strips), kRunThreads, 0, s>>>(a, acc);
}
template <int MODEL, int NREF>
static void grad_strips_mn(bool pdl, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gm) {
  auto k = grad_strips_kernel<MODEL, NREF, NREF == 1>;
  launch_k(pdl, k, dim3(strips_grid(k, a.n_strips)), dim3(kRunThreads), s, a, gq, gm);
}
template <int MODEL>
static void vote_strips_m(int n_ref, cudaStream_t s, const FusedArgs& a, float4* acc) {
  switch (n_ref) {
    case 1: vote_strips_mn<MODEL, 1>(s, a, acc); break;
#ifndef CMAX_LEAN_DEV
    case 2: vote_strips_mn<MODEL, 2>(s, a, acc); break;
    case 3: vote_strips_mn<MODEL, 3>(s, a, acc); break;
    default: vote_strips_mn<MODEL, 4>(s, a, acc); break;
#endif
  }
}
template <int MODEL>
static void grad_strips_m(int n_ref, bool pdl, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gm) {
  switch (n_ref) {
    case 1: grad_strips_mn<MODEL, 1>(pdl, s, a, gq, gm); break;
#ifndef CMAX_LEAN_DEV
    case 2: grad_strips_mn<MODEL, 2>(pdl, s, a, gq, gm); break;
    case 3: grad_strips_mn<MODEL, 3>(pdl, s, a, gq, gm); break;
    default: grad_strips_mn<MODEL, 4>(pdl, s, a, gq, gm); break;
#endif
  }
}

// tile size the strip kernels of (motion_model, n_ref) expect; the caller checks it against what the plan packed
int strips_tile_bytes_for(int motion_model, int n_ref) {
  return kStripTileBytes + (motion_model == CMAX_MOTION_VOXEL ? n_ref * kStripBinBytes : 0);
}

void launch_vote_strips(int motion_model, int n_ref, cudaStream_t s, const FusedArgs& a, float4* acc) {
  if (motion_model == CMAX_MOTION_DENSE) vote_strips_m<CMAX_MOTION_DENSE>(n_ref, s, a, acc);
#ifndef CMAX_LEAN_DEV
  else if (motion_model == CMAX_MOTION_TILE) vote_strips_m<CMAX_MOTION_TILE>(n_ref, s, a, acc);
  else if (motion_model == CMAX_MOTION_VOXEL) vote_strips_m<CMAX_MOTION_VOXEL>(n_ref, s, a, acc);
  else vote_strips_m<CMAX_MOTION_2DOF>(n_ref, s, a, acc);
#endif
}
void launch_grad_strips(int motion_model, int n_ref, bool pdl, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gmotion) {
  if (motion_model == CMAX_MOTION_DENSE) grad_strips_m<CMAX_MOTION_DENSE>(n_ref, pdl, s, a, gq, gmotion);
#ifndef CMAX_LEAN_DEV
  else if (motion_model == CMAX_MOTION_TILE) grad_strips_m<CMAX_MOTION_TILE>(n_ref, pdl, s, a, gq, gmotion);
  else if (motion_model == CMAX_MOTION_VOXEL) grad_strips_m<CMAX_MOTION_VOXEL>(n_ref, pdl, s, a, gq, gmotion);
  else grad_strips_m<CMAX_MOTION_2DOF>(n_ref, pdl, s, a, gq, gmotion);
#endif
}

}  // namespace cmax
