// The event kernels that read the plan's PACKED copy: the per-event baselines (one thread per event) and the run kernels
// (each thread walks 8 consecutive events), for K1 (warp + bilinear vote of every event, all reference times in one pass) and
// K3 (re-warp, gather dL/dIWE, chain to dL/dmotion), plus the dispatch over motion model / reference times / variant.
// The strip kernels (the default for dense batches) live in cmax_lean.cu, everything image-sized in cmax_mid.cu.
//
// Data layout in HBM / L2 (everything image-sized is L2 resident; only the event stream comes from HBM):
//   events   float4[n]                       one 16-byte streamed load per event per pass
//   acc      float4[n_ref][(Hp+1)*(Wp+1)]    per-corner accumulators: cell (i,j), i in [-1,Hp-1], j in [-1,Wp-1] holds
//                                            the four bilinear weights of all events whose floor pixel is (i,j), so an
//                                            event issues ONE 16-byte vector reduction (red.global.add.v4.f32)
//                                            instead of four scalar ones;  IWE[r,c] = acc[r,c].x + acc[r-1,c].y +
//                                            acc[r,c-1].z + acc[r-1,c-1].w  applies the reference's per-corner masks
//                                            by construction (src/event_image_converter.py:355-372)
//   gq       float4[n_ref][(Hp+1)*(Wp+1)]    the adjoint of that fold: cell (i,j) = dL/dIWE at the four (masked)
//                                            corners, so K3 gathers ONE float4 per event per reference time
#include <stdlib.h>

#include "cmax_objective.cuh"

namespace cmax {

// One event, one reference time: (x', y', dt, bin).        src/warp.py:254-258, 306-307, 346-357, 507-514
template <int MODEL>
__device__ __forceinline__ void warp_ref(const float4 e, int src, int HW, const float* __restrict__ motion, const TimeSmem& s,
                                         int r, float f0, float f1, float& xw, float& yw, float& dt, int& bin) {
  dt = normalised_dt(e.z, s.ref[r], s.period[r], s.normalize_t);
  bin = 0;
  if (MODEL == CMAX_MOTION_2DOF) {
    xw = warp_plus(e.x, dt, f0);
    yw = warp_plus(e.y, dt, f1);
  } else if (MODEL == CMAX_MOTION_DENSE) {
    xw = warp_minus(e.x, dt, f0);
    yw = warp_minus(e.y, dt, f1);
  } else {
    bin = time_bin(dt, s.edges[r], s.n_bins, s.dt_min[r], s.inv_width[r]);
    xw = e.x;
    yw = e.y;
    if (bin >= 0) {
      const float* f = motion + (int64_t)bin * 2 * HW;
      xw = warp_minus(e.x, dt, __ldg(f + src));
      yw = warp_minus(e.y, dt, __ldg(f + HW + src));
    }
  }
}

// ------------------------------------------------------------------------------------------------ K1
// VARIANT 0: one red.v4 per event per reference time into the per-corner accumulators.
// VARIANT 1: four masked scalar red.f32 straight into the IWE (the textbook scatter; kept as the measured baseline).
template <int MODEL, int NREF, int VARIANT>
__global__ void __launch_bounds__(256) vote_fused_kernel(FusedArgs a, float4* __restrict__ acc, float* __restrict__ iwe) {
  __shared__ TimeSmem s;
  stage_time<NREF, MODEL == CMAX_MOTION_VOXEL>(a.tp, s);
  const int HW = a.H * a.W;
  const int64_t HWp = (int64_t)a.Hp * a.Wp;
  float th0 = 0.f, th1 = 0.f;
  if (MODEL == CMAX_MOTION_2DOF) {
    th0 = __ldg(a.motion);
    th1 = __ldg(a.motion + 1);
  }
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += step) {
    const float4 e = __ldcs(a.ev + i);
    int src = 0;
    float f0 = th0, f1 = th1;
    if (MODEL != CMAX_MOTION_2DOF) src = __float2int_rz(e.x) * a.W + __float2int_rz(e.y);  // validated by the plan
    if (MODEL == CMAX_MOTION_DENSE) {
      f0 = __ldg(a.motion + src);
      f1 = __ldg(a.motion + HW + src);
    }
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      float xw, yw, dt;
      int bin;
      warp_ref<MODEL>(e, src, HW, a.motion, s, r, f0, f1, xw, yw, dt, bin);
      const Vote v = vote_geometry(xw, yw, a.pad_h, a.pad_w);
      float w[4];
      vote_weights(v, w);
      if (VARIANT == 0) {
        if (v.row >= -1 && v.row < a.Hp && v.col >= -1 && v.col < a.Wp)
          red_add_v4(acc + r * a.cells + (int64_t)(v.row + 1) * (a.Wp + 1) + (v.col + 1), w[0], w[1], w[2], w[3]);
      } else {
        const bool r0 = v.row >= 0 && v.row < a.Hp, r1 = v.row >= -1 && v.row + 1 < a.Hp;
        const bool c0 = v.col >= 0 && v.col < a.Wp, c1 = v.col >= -1 && v.col + 1 < a.Wp;
        float* p = iwe + r * HWp + (int64_t)v.row * a.Wp + v.col;
        if (r0 && c0) atomicAdd(p, w[0]);
        if (r1 && c0) atomicAdd(p + a.Wp, w[1]);
        if (r0 && c1) atomicAdd(p + 1, w[2]);
        if (r1 && c1) atomicAdd(p + a.Wp + 1, w[3]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ K3
// GVAR 0: scalar red per event.  GVAR 1: events are ordered by source pixel, so equal source pixels are consecutive
// lanes: segmented warp reduction, one red per run.
template <int MODEL, int NREF, int GVAR>
__global__ void __launch_bounds__(256) grad_fused_kernel(FusedArgs a, const float4* __restrict__ gq, float* __restrict__ gmotion) {
  __shared__ TimeSmem s;
  __shared__ double red2[2][8];
  stage_time<NREF, MODEL == CMAX_MOTION_VOXEL>(a.tp, s);
  const int HW = a.H * a.W;
  float th0 = 0.f, th1 = 0.f;
  if (MODEL == CMAX_MOTION_2DOF) {
    th0 = __ldg(a.motion);
    th1 = __ldg(a.motion + 1);
  }
  double t0 = 0.0, t1 = 0.0;  // 2-dof: per-thread fp64 partial sums
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int64_t n_round = (a.n + 31) / 32 * 32;  // whole warps stay in the loop for the shuffles
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
    const bool live = i < a.n;
    float4 e = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) e = __ldcs(a.ev + i);
    int src = 0;
    float f0 = th0, f1 = th1;
    if (MODEL != CMAX_MOTION_2DOF) src = __float2int_rz(e.x) * a.W + __float2int_rz(e.y);
    if (MODEL == CMAX_MOTION_DENSE && live) {
      f0 = __ldg(a.motion + src);
      f1 = __ldg(a.motion + HW + src);
    }
    float g0 = 0.f, g1 = 0.f;  // dense: sum over reference times of -dt * dL/dx'
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      float xw, yw, dt;
      int bin;
      warp_ref<MODEL>(e, src, HW, a.motion, s, r, f0, f1, xw, yw, dt, bin);
      const Vote v = vote_geometry(xw, yw, a.pad_h, a.pad_w);
      float dx = 0.f, dy = 0.f;
      if (live && v.row >= -1 && v.row < a.Hp && v.col >= -1 && v.col < a.Wp) {
        const float4 g = __ldg(gq + r * a.cells + (int64_t)(v.row + 1) * (a.Wp + 1) + (v.col + 1));
        // d w / d x' = (-(1-fy), (1-fy), -fy, fy),  d w / d y' = (-(1-fx), -fx, (1-fx), fx)
        dx = (1.0f - v.fy) * (g.y - g.x) + v.fy * (g.w - g.z);
        dy = (1.0f - v.fx) * (g.z - g.x) + v.fx * (g.w - g.y);
      }
      if (MODEL == CMAX_MOTION_2DOF) {
        t0 += (double)(dt * dx);
        t1 += (double)(dt * dy);
      } else if (MODEL == CMAX_MOTION_DENSE) {
        g0 -= dt * dx;
        g1 -= dt * dy;
      } else if (live && bin >= 0) {
        float* g = gmotion + (int64_t)bin * 2 * HW;
        atomicAdd(g + src, -(dt * dx));
        atomicAdd(g + HW + src, -(dt * dy));
      }
    }
    if (MODEL == CMAX_MOTION_DENSE) {
      if (GVAR == 0) {
        if (live) {
          atomicAdd(gmotion + src, g0);
          atomicAdd(gmotion + HW + src, g1);
        }
      } else {
        // segmented suffix sum over runs of equal `src` (dead lanes carry src = -1 and zeros)
        const int key = live ? src : -1;
        const int lane = threadIdx.x & 31;
        const int prev = __shfl_up_sync(0xffffffffu, key, 1);
        const bool head = (lane == 0) || (prev != key);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float u0 = __shfl_down_sync(0xffffffffu, g0, o);
          const float u1 = __shfl_down_sync(0xffffffffu, g1, o);
          const int uk = __shfl_down_sync(0xffffffffu, key, o);
          if (lane + o < 32 && uk == key) {
            g0 += u0;
            g1 += u1;
          }
        }
        if (head && live) {
          atomicAdd(gmotion + src, g0);
          atomicAdd(gmotion + HW + src, g1);
        }
      }
    }
  }
  if (MODEL == CMAX_MOTION_2DOF) {
    t0 = warp_sum(t0);
    t1 = warp_sum(t1);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
      red2[0][wid] = t0;
      red2[1][wid] = t1;
    }
    __syncthreads();
    if (threadIdx.x < 2) {
      double tot = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red2[threadIdx.x][w];
      atomicAdd(reinterpret_cast<double*>(gmotion) + threadIdx.x, tot);  // fp64 staging, narrowed by finish_2dof_kernel
    }
  }
}

// ------------------------------------------------------------------------------------------------ run kernels
// Events of a plan ordered by source pixel arrive as runs: ~N/HW consecutive events share the source pixel (hence
// the flow vector) and, being time ordered inside the run, walk monotonically along one line of the image.  Each
// THREAD therefore takes kRunE consecutive events and keeps the current accumulator cell in registers:
//   K1: weights of consecutive events that fall into the same cell are summed in registers, one red.v4 per cell change
//   K3: the per-corner gradient quad is re-gathered (one 16-byte load) only on a cell change, and the flow gradient of
//       a source pixel is summed in registers, one pair of reds per source-pixel change
// which divides the number of atomics and gathers by the run length without any shuffles.  The events come from the
// plan's packed copy (x, y, dt|t, src): everything that is constant over the CM iterations -- the source pixel index,
// and for a single reference time the normalised dt including its IEEE division -- is precomputed once per plan, and
// the copy is stored pre-transposed in warp-tile order (cmax_plan.cuh): slot k*32 + lane of a 4 KB tile is lane's k-th
// consecutive event.  Every warp streams its tiles with TMA bulk copies (cp.async.bulk + mbarrier, double buffered in
// shared memory): the copy of tile i+1 is in flight while tile i is walked, the walk reads its events with
// conflict-free LDS.128, and no thread ever issues a global load for an event.
// Correct for ANY event order -- an unordered stream just degenerates to one flush per event.
// When every event has integer pixel coordinates (what a sensor delivers; checked by the plan) the packed copy uses
// 8 bytes per event -- (dt|t, row<<16|col) -- halving the DRAM stream and the shared-memory landing zone, which doubles
// the number of resident warps.
// One packed event, one reference time -> ((x', y') packed, dt, bin).  PRE_DT: tz already is the normalised dt of reference 0.
template <int MODEL, int NREF, bool PRE_DT>
__device__ __forceinline__ f32x2 warp_packed(float x, float y, float tz, int src, int HW, const float* __restrict__ motion,
                                             const RefRegs<NREF>& rr, const TimeSmem& s, int r, float f0, float f1, float& dt, int& bin) {
  dt = PRE_DT ? tz : __fdiv_rn(__fsub_rn(tz, rr.ref[r]), rr.period[r]);
  bin = 0;
  if (MODEL == CMAX_MOTION_2DOF) return warp_plus2(x, y, dt, f0, f1);
  if (MODEL == CMAX_MOTION_DENSE) return warp_minus2(x, y, dt, f0, f1);
  bin = time_bin(dt, s.edges[r], s.n_bins, s.dt_min[r], s.inv_width[r]);
  if (bin < 0) return pk2(x, y);
  const float* f = motion + (int64_t)bin * 2 * HW;
  return warp_minus2(x, y, dt, __ldg(f + src), __ldg(f + HW + src));
}

// ---- K1 walk
template <int NREF>
struct VoteState {
  int cell[NREF];
  float w0[NREF], w1[NREF], w2[NREF], w3[NREF];
  int key, src;  // source pixel of the current run: packed key and flat index
  float f0, f1;
};

template <int MODEL, int NREF, bool PRE_DT, bool COMPACT>
__device__ __forceinline__ void vote_step(float x, float y, float tz, int key, VoteState<NREF>& st, const FusedArgs& a, int HW,
                                          const RefRegs<NREF>& rr, const TimeSmem& s, float4* __restrict__ acc) {
  // a source pixel changes in ~1 of 50 events per lane: skip the whole block unless some lane of the warp needs it
  if (MODEL != CMAX_MOTION_2DOF && __any_sync(0xffffffffu, key != st.key)) {
    if (key != st.key) {
      st.key = key;
      st.src = PackedEv<COMPACT>::src(key, a.W);
      if (MODEL == CMAX_MOTION_DENSE) {
        st.f0 = __ldg(a.motion + st.src);
        st.f1 = __ldg(a.motion + HW + st.src);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NREF; ++r) {
    float dt;
    int bin;
    const f32x2 xyw = warp_packed<MODEL, NREF, PRE_DT>(x, y, tz, st.src, HW, a.motion, rr, s, r, st.f0, st.f1, dt, bin);
    const Vote v = vote_geometry2(xyw, a.pad_h, a.pad_w);
    float w[4];
    vote_weights(v, w);
    const int c = vote_cell(v, a.Hp, a.Wp);
    const bool same = c == st.cell[r];
    red_add_v4_if(!same && st.cell[r] >= 0, acc + r * a.cells + max(st.cell[r], 0), st.w0[r], st.w1[r], st.w2[r], st.w3[r]);
    st.cell[r] = c;
    const float keep = same ? 1.0f : 0.0f;  // acc * 1 + w and acc * 0 + w are exact: one select instead of four
    const f32x2 keep2 = pk2(keep, keep);
    upk2(fma2(pk2(st.w0[r], st.w1[r]), keep2, pk2(w[0], w[1])), st.w0[r], st.w1[r]);
    upk2(fma2(pk2(st.w2[r], st.w3[r]), keep2, pk2(w[2], w[3])), st.w2[r], st.w3[r]);
  }
}

// FB: batch the flow loads of a full tile (dense model) so that they share one L2 round trip -- costs 16 registers.
template <int MODEL, int NREF, bool PRE_DT, bool COMPACT, bool FB>
__global__ void __launch_bounds__(kRunThreads) vote_runs_kernel(FusedArgs a, float4* __restrict__ acc) {
  using PE = PackedEv<COMPACT>;
  __shared__ TimeSmem s;
  __shared__ TilePipe<PE::kTileBytes> pipes[kRunWarps];
  if (MODEL == CMAX_MOTION_VOXEL) stage_time<NREF, true>(a.tp, s);
  if (a.zero256 != nullptr && blockIdx.x == 0 && threadIdx.x < 64) a.zero256[threadIdx.x] = 0u;  // StatAcc block + CTA counter
  const RefRegs<NREF> rr = load_refs<NREF>(a.tp);
  const int HW = a.H * a.W;
  const int lane = threadIdx.x & 31;
  TilePipe<PE::kTileBytes>& pipe = pipes[threadIdx.x >> 5];
  pipe_init(pipe, lane);
  float th0 = 0.f, th1 = 0.f;
  if (MODEL == CMAX_MOTION_2DOF) {
    th0 = __ldg(a.motion);
    th1 = __ldg(a.motion + 1);
  }
  const int64_t n_tiles = (a.n + kWarpTile - 1) / kWarpTile;
  const int64_t warp0 = (int64_t)blockIdx.x * kRunWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kRunWarps;
  const int64_t t_end = n_tiles;  // warps stride over all tiles (a contiguous range per CTA measured no better: profiles/)
  if (warp0 < t_end) pipe_issue(pipe, 0, a.packed, warp0, lane);
  pdl_trigger();  // the fold may be scheduled as soon as this grid drains
  int it = 0;
  for (int64_t tile = warp0; tile < t_end; tile += n_warps, ++it) {
    const int stage = it & 1;
    __syncwarp();  // every lane is done with the other buffer (walked in the previous iteration)
    if (tile + n_warps < t_end) pipe_issue(pipe, stage ^ 1, a.packed, tile + n_warps, lane);
    mbar_wait(&pipe.bar[stage], (it >> 1) & 1);
    const void* buf = pipe.buf[stage];
    VoteState<NREF> st;
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      st.cell[r] = -1;
      st.w0[r] = st.w1[r] = st.w2[r] = st.w3[r] = 0.f;
    }
    st.key = -1;
    st.src = 0;
    st.f0 = th0;
    st.f1 = th1;
    float x, y, tz;
    int key;
    if ((tile + 1) * kWarpTile <= a.n) {  // full tile (warp-uniform)
      if constexpr (MODEL == CMAX_MOTION_DENSE && FB) {
        float f0[kRunE], f1[kRunE];
        int prev = -1;
#pragma unroll
        for (int k = 0; k < kRunE; ++k) {
          const int kk = PE::key_of(buf, k, lane);
          if (kk != prev) {
            const int src = PE::src(kk, a.W);
            f0[k] = __ldg(a.motion + src);
            f1[k] = __ldg(a.motion + HW + src);
          } else {
            f0[k] = f0[k > 0 ? k - 1 : 0];
            f1[k] = f1[k > 0 ? k - 1 : 0];
          }
          prev = kk;
        }
#pragma unroll
        for (int k = 0; k < kRunE; ++k) {
          PE::get(buf, k, lane, x, y, tz, key);
          st.key = key;  // flow already fetched
          st.f0 = f0[k];
          st.f1 = f1[k];
          vote_step<MODEL, NREF, PRE_DT, COMPACT>(x, y, tz, key, st, a, HW, rr, s, acc);
        }
      } else {
#pragma unroll
        for (int k = 0; k < kRunE; ++k) {
          PE::get(buf, k, lane, x, y, tz, key);
          vote_step<MODEL, NREF, PRE_DT, COMPACT>(x, y, tz, key, st, a, HW, rr, s, acc);
        }
      }
    } else {
      const int64_t left = a.n - (tile * kWarpTile + (int64_t)lane * kRunE);
      const int count = left >= kRunE ? kRunE : (left > 0 ? (int)left : 0);
      for (int k = 0; k < count; ++k) {
        PE::get(buf, k, lane, x, y, tz, key);
        vote_step<MODEL, NREF, PRE_DT, COMPACT>(x, y, tz, key, st, a, HW, rr, s, acc);
      }
    }
#pragma unroll
    for (int r = 0; r < NREF; ++r)
      if (st.cell[r] >= 0) red_add_v4(acc + r * a.cells + st.cell[r], st.w0[r], st.w1[r], st.w2[r], st.w3[r]);
  }
}

// K1 with the flow fetched one tile AHEAD (dense model): a 3-deep TMA pipeline means tile i+1 has already landed when
// tile i starts, so its source pixels are known and its flow vectors are requested before tile i is walked -- a whole
// tile walk (thousands of cycles) hides the L2 round trip that the plain walk exposes right after every tile arrival.
template <int NREF, bool PRE_DT, bool COMPACT>
__global__ void __launch_bounds__(kRunThreads) vote_runs_ahead_kernel(FusedArgs a, float4* __restrict__ acc) {
  using PE = PackedEv<COMPACT>;
  constexpr int MODEL = CMAX_MOTION_DENSE;
  constexpr int NS = 3;
  __shared__ TilePipe<PE::kTileBytes, NS> pipes[kRunWarps];
  // the dense model never reads the voxel time table: hand the shared step function a reference it will not touch
  const TimeSmem& s = *reinterpret_cast<const TimeSmem*>(pipes);
  if (a.zero256 != nullptr && blockIdx.x == 0 && threadIdx.x < 64) a.zero256[threadIdx.x] = 0u;
  const RefRegs<NREF> rr = load_refs<NREF>(a.tp);
  const int HW = a.H * a.W;
  const int lane = threadIdx.x & 31;
  TilePipe<PE::kTileBytes, NS>& pipe = pipes[threadIdx.x >> 5];
  pipe_init(pipe, lane);
  const int64_t n_tiles = (a.n + kWarpTile - 1) / kWarpTile;
  const int64_t warp0 = (int64_t)blockIdx.x * kRunWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kRunWarps;
  if (warp0 < n_tiles) pipe_issue(pipe, 0, a.packed, warp0, lane);
  if (warp0 + n_warps < n_tiles) pipe_issue(pipe, 1, a.packed, warp0 + n_warps, lane);
  pdl_trigger();
  float fn0[kRunE], fn1[kRunE];  // flow vectors of the NEXT tile's events (in flight while the current tile is walked)
  auto fetch_flows = [&](const void* buf) {
    int prev = -1;
#pragma unroll
    for (int k = 0; k < kRunE; ++k) {
      const int kk = PE::key_of(buf, k, lane);
      if (kk != prev && kk != -1) {
        const int src = PE::src(kk, a.W);
        fn0[k] = __ldg(a.motion + src);
        fn1[k] = __ldg(a.motion + HW + src);
      } else {
        fn0[k] = fn0[k > 0 ? k - 1 : 0];
        fn1[k] = fn1[k > 0 ? k - 1 : 0];
      }
      prev = kk;
    }
  };
#pragma unroll
  for (int k = 0; k < kRunE; ++k) fn0[k] = fn1[k] = 0.f;
  if (warp0 < n_tiles) {
    mbar_wait(&pipe.bar[0], 0);
    fetch_flows(pipe.buf[0]);
  }
  int it = 0;
  for (int64_t tile = warp0; tile < n_tiles; tile += n_warps, ++it) {
    const int stage = it % NS;
    float f0[kRunE], f1[kRunE];
#pragma unroll
    for (int k = 0; k < kRunE; ++k) {
      f0[k] = fn0[k];
      f1[k] = fn1[k];
    }
    __syncwarp();  // every lane is done with the buffer of the previous iteration: refill it with tile it+2
    if (tile + 2 * n_warps < n_tiles) pipe_issue(pipe, (it + 2) % NS, a.packed, tile + 2 * n_warps, lane);
    if (tile + n_warps < n_tiles) {  // tile it+1 landed during the previous walk: request its flow vectors now
      mbar_wait(&pipe.bar[(it + 1) % NS], ((it + 1) / NS) & 1);
      fetch_flows(pipe.buf[(it + 1) % NS]);
    }
    const void* buf = pipe.buf[stage];
    VoteState<NREF> st;
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      st.cell[r] = -1;
      st.w0[r] = st.w1[r] = st.w2[r] = st.w3[r] = 0.f;
    }
    st.src = 0;
    float x, y, tz;
    int key;
    const int64_t left = a.n - (tile * kWarpTile + (int64_t)lane * kRunE);
    const int count = left >= kRunE ? kRunE : (left > 0 ? (int)left : 0);
#pragma unroll
    for (int k = 0; k < kRunE; ++k) {
      if (k < count) {  // (padding slots of the last tile carry key -1 and are skipped)
        PE::get(buf, k, lane, x, y, tz, key);
        st.key = key;  // flow already fetched
        st.f0 = f0[k];
        st.f1 = f1[k];
        vote_step<MODEL, NREF, PRE_DT, COMPACT>(x, y, tz, key, st, a, HW, rr, s, acc);
      }
    }
#pragma unroll
    for (int r = 0; r < NREF; ++r)
      if (st.cell[r] >= 0) red_add_v4(acc + r * a.cells + st.cell[r], st.w0[r], st.w1[r], st.w2[r], st.w3[r]);
  }
}

// ---- K3 walk
template <int MODEL, int NREF>
struct GradState {
  static constexpr int NACC = (MODEL == CMAX_MOTION_VOXEL) ? NREF : 1;
  int cell[NREF];
  float d_x0[NREF], d_c0[NREF], d_r[NREF];  // corner differences of the current cell's gradient quad
  int slot[NACC];                           // flat index into gmotion of the row-component slot being accumulated (-1: none)
  float g0[NACC], g1[NACC];
  int key, src;  // source pixel of the current run
  float f0, f1;
  double t0, t1;  // 2-dof
};

template <int MODEL, int NREF>
__device__ __forceinline__ void grad_flush(GradState<MODEL, NREF>& st, int q, int HW, float* __restrict__ gmotion) {
  if (st.slot[q] >= 0) {
    atomicAdd(gmotion + st.slot[q], st.g0[q]);
    atomicAdd(gmotion + st.slot[q] + HW, st.g1[q]);
  }
}

template <int MODEL, int NREF, bool PRE_DT, bool COMPACT>
__device__ __forceinline__ void grad_step(float x, float y, float tz, int key, GradState<MODEL, NREF>& st, const FusedArgs& a, int HW,
                                          const RefRegs<NREF>& rr, const TimeSmem& s, const float4* __restrict__ gq,
                                          float* __restrict__ gmotion) {
  if (MODEL != CMAX_MOTION_2DOF && __any_sync(0xffffffffu, key != st.key)) {  // some lane starts a new source pixel
    if (key != st.key) {
      st.key = key;
      st.src = PackedEv<COMPACT>::src(key, a.W);
      if (MODEL == CMAX_MOTION_DENSE) {  // flush the predecessor's gradient, fetch the new flow vector
        grad_flush<MODEL, NREF>(st, 0, HW, gmotion);
        st.slot[0] = st.src;
        st.g0[0] = st.g1[0] = 0.f;
        st.f0 = __ldg(a.motion + st.src);
        st.f1 = __ldg(a.motion + HW + st.src);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < NREF; ++r) {
    float dt;
    int bin;
    const f32x2 xyw = warp_packed<MODEL, NREF, PRE_DT>(x, y, tz, st.src, HW, a.motion, rr, s, r, st.f0, st.f1, dt, bin);
    const Vote v = vote_geometry2(xyw, a.pad_h, a.pad_w);
    const int c = vote_cell(v, a.Hp, a.Wp);
    if (c != st.cell[r]) {
      st.cell[r] = c;
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c >= 0) g = __ldg(gq + r * a.cells + c);
      // dL/dx' = (1-fy)(g10-g00) + fy(g11-g01) = d_x0 + fy*d_r ;  dL/dy' = (1-fx)(g01-g00) + fx(g11-g10) = d_c0 + fx*d_r
      st.d_x0[r] = g.y - g.x;
      st.d_c0[r] = g.z - g.x;
      st.d_r[r] = (g.w - g.z) - st.d_x0[r];
    }
    // (the fractions of an event outside the image are meaningless, see vote_geometry: its quad is zero, keep it zero)
    const float dx = c >= 0 ? fmaf(v.fy, st.d_r[r], st.d_x0[r]) : 0.f;
    const float dy = c >= 0 ? fmaf(v.fx, st.d_r[r], st.d_c0[r]) : 0.f;
    if (MODEL == CMAX_MOTION_2DOF) {
      st.t0 += (double)(dt * dx);
      st.t1 += (double)(dt * dy);
    } else if (MODEL == CMAX_MOTION_DENSE) {
      st.g0[0] = fmaf(-dt, dx, st.g0[0]);
      st.g1[0] = fmaf(-dt, dy, st.g1[0]);
    } else {
      const int kk = bin >= 0 ? bin * 2 * HW + st.src : -1;
      if (kk != st.slot[r]) {
        grad_flush<MODEL, NREF>(st, r, HW, gmotion);
        st.slot[r] = kk;
        st.g0[r] = st.g1[r] = 0.f;
      }
      st.g0[r] = fmaf(-dt, dx, st.g0[r]);
      st.g1[r] = fmaf(-dt, dy, st.g1[r]);
    }
  }
}

// Full tile of the dense-flow model, batched so that the dependent L2 round trips of the walk overlap.  Per batch of
// KB consecutive events: (A) all flow loads, (B) per reference time all warps / cells, then all gradient-quad gathers,
// (C) the accumulation with one flush per source-pixel change.  Two exposed L2 latencies per batch instead of up to
// 2*KB; KB trades registers (occupancy) against memory-level parallelism.
template <int NREF, bool PRE_DT, bool COMPACT, int KB>
__device__ __forceinline__ void grad_tile_dense(const void* __restrict__ buf, int lane, GradState<CMAX_MOTION_DENSE, NREF>& st,
                                                const FusedArgs& a, int HW, const RefRegs<NREF>& rr, const float4* __restrict__ gq,
                                                float* __restrict__ gmotion) {
  using PE = PackedEv<COMPACT>;
#pragma unroll
  for (int b = 0; b < kRunE; b += KB) {
    float f0[KB], f1[KB];
    int keys[KB];
    int prev = -1;
#pragma unroll
    for (int k = 0; k < KB; ++k) {  // (A)
      keys[k] = PE::key_of(buf, b + k, lane);
      if (keys[k] != prev) {
        const int src = PE::src(keys[k], a.W);
        f0[k] = __ldg(a.motion + src);
        f1[k] = __ldg(a.motion + HW + src);
      } else {
        f0[k] = f0[k > 0 ? k - 1 : 0];
        f1[k] = f1[k > 0 ? k - 1 : 0];
      }
      prev = keys[k];
    }
    float gx[KB], gy[KB];  // sum over reference times of -dt * dL/dx', -dt * dL/dy' per event
#pragma unroll
    for (int k = 0; k < KB; ++k) gx[k] = gy[k] = 0.f;
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      float fx[KB], fy[KB], dts[KB];
      int cs[KB];
      float4 g[KB];
      int cprev = -2;
#pragma unroll
      for (int k = 0; k < KB; ++k) {  // (B)
        float x, y, tz;
        int key;
        PE::get(buf, b + k, lane, x, y, tz, key);
        const float dt = PRE_DT ? tz : __fdiv_rn(__fsub_rn(tz, rr.ref[r]), rr.period[r]);
        const Vote v = vote_geometry2(warp_minus2(x, y, dt, f0[k], f1[k]), a.pad_h, a.pad_w);
        fx[k] = v.fx;
        fy[k] = v.fy;
        dts[k] = dt;
        cs[k] = vote_cell(v, a.Hp, a.Wp);
        if (cs[k] != cprev && cs[k] >= 0) g[k] = __ldg(gq + r * a.cells + cs[k]);
        cprev = cs[k];
      }
#pragma unroll
      for (int k = 0; k < KB; ++k) {  // (C)
        if (cs[k] < 0) g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        else if (k > 0 && cs[k] == cs[k - 1]) g[k] = g[k - 1];
        const float d_x0 = g[k].y - g[k].x, d_c0 = g[k].z - g[k].x;
        const float d_r = (g[k].w - g[k].z) - d_x0;
        if (cs[k] >= 0) {  // fractions of an out-of-image event are meaningless (vote_geometry)
          gx[k] = fmaf(-dts[k], fmaf(fy[k], d_r, d_x0), gx[k]);
          gy[k] = fmaf(-dts[k], fmaf(fx[k], d_r, d_c0), gy[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < KB; ++k) {
      if (keys[k] != st.key) {
        grad_flush<CMAX_MOTION_DENSE, NREF>(st, 0, HW, gmotion);
        st.key = keys[k];
        st.slot[0] = PE::src(keys[k], a.W);
        st.g0[0] = st.g1[0] = 0.f;
      }
      st.g0[0] += gx[k];
      st.g1[0] += gy[k];
    }
  }
}

template <int MODEL, int NREF, bool PRE_DT, bool COMPACT, int KB>
__global__ void __launch_bounds__(kRunThreads) grad_runs_kernel(FusedArgs a, const float4* __restrict__ gq, float* __restrict__ gmotion) {
  using PE = PackedEv<COMPACT>;
  __shared__ TimeSmem s;
  __shared__ TilePipe<PE::kTileBytes> pipes[kRunWarps];
  __shared__ double red2[2][kRunWarps];
  if (MODEL == CMAX_MOTION_VOXEL) stage_time<NREF, true>(a.tp, s);
  const RefRegs<NREF> rr = load_refs<NREF>(a.tp);
  const int HW = a.H * a.W;
  const int lane = threadIdx.x & 31;
  TilePipe<PE::kTileBytes>& pipe = pipes[threadIdx.x >> 5];
  pipe_init(pipe, lane);
  GradState<MODEL, NREF> st;
  st.f0 = st.f1 = 0.f;
  st.t0 = st.t1 = 0.0;
  if (MODEL == CMAX_MOTION_2DOF) {
    st.f0 = __ldg(a.motion);
    st.f1 = __ldg(a.motion + 1);
  }
  const int64_t n_tiles = (a.n + kWarpTile - 1) / kWarpTile;
  const int64_t warp0 = (int64_t)blockIdx.x * kRunWarps + (threadIdx.x >> 5), n_warps = (int64_t)gridDim.x * kRunWarps;
  const int64_t t_end = n_tiles;
  if (warp0 < t_end) pipe_issue(pipe, 0, a.packed, warp0, lane);  // the packed events do not depend on the predecessor
  pdl_wait();  // gradient quads (and the zeroed gradient buffer) of the predecessor kernel are complete
  int it = 0;
  for (int64_t tile = warp0; tile < t_end; tile += n_warps, ++it) {
    const int stage = it & 1;
    __syncwarp();
    if (tile + n_warps < t_end) pipe_issue(pipe, stage ^ 1, a.packed, tile + n_warps, lane);
    mbar_wait(&pipe.bar[stage], (it >> 1) & 1);
    const void* buf = pipe.buf[stage];
#pragma unroll
    for (int r = 0; r < NREF; ++r) {
      st.cell[r] = -2;
      st.d_x0[r] = st.d_c0[r] = st.d_r[r] = 0.f;
    }
#pragma unroll
    for (int q = 0; q < GradState<MODEL, NREF>::NACC; ++q) {
      st.slot[q] = -1;
      st.g0[q] = st.g1[q] = 0.f;
    }
    st.key = -1;
    st.src = 0;
    float x, y, tz;
    int key;
    if ((tile + 1) * kWarpTile <= a.n) {
      if constexpr (MODEL == CMAX_MOTION_DENSE && KB > 0) {
        grad_tile_dense<NREF, PRE_DT, COMPACT, KB>(buf, lane, st, a, HW, rr, gq, gmotion);
      } else {
#pragma unroll
        for (int k = 0; k < kRunE; ++k) {
          PE::get(buf, k, lane, x, y, tz, key);
          grad_step<MODEL, NREF, PRE_DT, COMPACT>(x, y, tz, key, st, a, HW, rr, s, gq, gmotion);
        }
      }
    } else {
      const int64_t left = a.n - (tile * kWarpTile + (int64_t)lane * kRunE);
      const int count = left >= kRunE ? kRunE : (left > 0 ? (int)left : 0);
      for (int k = 0; k < count; ++k) {
        PE::get(buf, k, lane, x, y, tz, key);
        grad_step<MODEL, NREF, PRE_DT, COMPACT>(x, y, tz, key, st, a, HW, rr, s, gq, gmotion);
      }
    }
    if (MODEL != CMAX_MOTION_2DOF) {
#pragma unroll
      for (int q = 0; q < GradState<MODEL, NREF>::NACC; ++q) grad_flush<MODEL, NREF>(st, q, HW, gmotion);
    }
  }
  if (MODEL == CMAX_MOTION_2DOF) {
    const double t0 = warp_sum(st.t0), t1 = warp_sum(st.t1);
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
template <int MODEL, int NREF, bool COMPACT>
static void launch_vote_runs(int variant, cudaStream_t s, const FusedArgs& a, float4* acc) {
  if constexpr (MODEL == CMAX_MOTION_DENSE && COMPACT) {  // (3 stages of 16-byte tiles would not fit 48 KB of static shared memory)
    if (variant == 4) {  // flow fetched one tile ahead
      auto k = vote_runs_ahead_kernel<NREF, NREF == 1, COMPACT>;
      k<<<run_grid(k, a.n), kRunThreads, 0, s>>>(a, acc);
      return;
    }
  }
  if (variant == 3) {  // batched flow loads
    auto k = vote_runs_kernel<MODEL, NREF, NREF == 1, COMPACT, true>;
    k<<<run_grid(k, a.n), kRunThreads, 0, s>>>(a, acc);
  } else {
    auto k = vote_runs_kernel<MODEL, NREF, NREF == 1, COMPACT, false>;
    k<<<run_grid(k, a.n), kRunThreads, 0, s>>>(a, acc);
  }
}

template <int MODEL, int NREF>
static void launch_vote(int variant, int grid, cudaStream_t s, const FusedArgs& a, float4* acc, float* iwe) {
  if (variant == 5 && a.strips != nullptr && a.strip_tile_bytes == strips_tile_bytes_for(MODEL, NREF)) {
    launch_vote_strips(MODEL, NREF, s, a, acc);
  } else if (variant >= 2) {
    if (a.compact) launch_vote_runs<MODEL, NREF, true>(variant, s, a, acc);
    else launch_vote_runs<MODEL, NREF, false>(variant, s, a, acc);
  } else if (variant == 1) vote_fused_kernel<MODEL, NREF, 1><<<grid, 256, 0, s>>>(a, acc, iwe);
  else vote_fused_kernel<MODEL, NREF, 0><<<grid, 256, 0, s>>>(a, acc, iwe);
}
template <int MODEL>
static void launch_vote_m(int n_ref, int variant, int grid, cudaStream_t s, const FusedArgs& a, float4* acc, float* iwe) {
  switch (n_ref) {
    case 1: launch_vote<MODEL, 1>(variant, grid, s, a, acc, iwe); break;
    case 2: launch_vote<MODEL, 2>(variant, grid, s, a, acc, iwe); break;
    case 3: launch_vote<MODEL, 3>(variant, grid, s, a, acc, iwe); break;
    default: launch_vote<MODEL, 4>(variant, grid, s, a, acc, iwe); break;
  }
}
template <int MODEL, int NREF, bool COMPACT>
static void launch_grad_runs(int gvar, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gm) {
  if (gvar == 2 || gvar == 5) {  // batches of 4 (also what a plan without strips runs instead of the strip kernel)
    auto k = grad_runs_kernel<MODEL, NREF, NREF == 1, COMPACT, 4>;
    launch_k(pdl_enabled(), k, dim3(run_grid(k, a.n)), dim3(kRunThreads), s, a, gq, gm);
  } else if (gvar == 3) {  // batches of 8
    auto k = grad_runs_kernel<MODEL, NREF, NREF == 1, COMPACT, 8>;
    launch_k(pdl_enabled(), k, dim3(run_grid(k, a.n)), dim3(kRunThreads), s, a, gq, gm);
  } else {  // sequential walk
    auto k = grad_runs_kernel<MODEL, NREF, NREF == 1, COMPACT, 0>;
    launch_k(pdl_enabled(), k, dim3(run_grid(k, a.n)), dim3(kRunThreads), s, a, gq, gm);
  }
}

template <int MODEL, int NREF>
static void launch_grad(int gvar, int grid, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gm) {
  if (gvar == 5 && a.strips != nullptr && a.strip_tile_bytes == strips_tile_bytes_for(MODEL, NREF)) {
    launch_grad_strips(MODEL, NREF, pdl_enabled(), s, a, gq, gm);
  } else if (gvar >= 2) {
    if (a.compact) launch_grad_runs<MODEL, NREF, true>(gvar, s, a, gq, gm);
    else launch_grad_runs<MODEL, NREF, false>(gvar, s, a, gq, gm);
  } else if (gvar == 1 && MODEL == CMAX_MOTION_DENSE) grad_fused_kernel<MODEL, NREF, 1><<<grid, 256, 0, s>>>(a, gq, gm);
  else grad_fused_kernel<MODEL, NREF, 0><<<grid, 256, 0, s>>>(a, gq, gm);
}
template <int MODEL>
static void launch_grad_m(int n_ref, int gvar, int grid, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gm) {
  switch (n_ref) {
    case 1: launch_grad<MODEL, 1>(gvar, grid, s, a, gq, gm); break;
    case 2: launch_grad<MODEL, 2>(gvar, grid, s, a, gq, gm); break;
    case 3: launch_grad<MODEL, 3>(gvar, grid, s, a, gq, gm); break;
    default: launch_grad<MODEL, 4>(gvar, grid, s, a, gq, gm); break;
  }
}

static inline int event_grid(int64_t n, int per_sm) {
  const int64_t want = (n + 255) / 256;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms() * per_sm));
}

FusedArgs fused_args(const cmax_plan* p, const float* motion) {
  FusedArgs a;
  a.ev = reinterpret_cast<const float4*>(p->events);
  a.packed = p->packed;
  a.compact = p->compact;
  a.n = p->n;
  a.H = p->H; a.W = p->W; a.Hp = p->Hp; a.Wp = p->Wp; a.pad_h = p->pad_h; a.pad_w = p->pad_w;
  a.motion = motion;
  a.tp = p->d_params;
  a.cells = (int64_t)(p->Hp + 1) * (p->Wp + 1) + 1;
  a.strips = p->strips;
  a.n_strips = p->n_strips;
  a.strip_tile_bytes = p->strip_tile_bytes;
  a.zero256 = nullptr;
  a.tile = p->tile;
  a.t_scale = p->t_scale;
  // strips per pixel of the rows this batch covers: beyond ~16 the same-address reductions of K3 start to serialise
  const int64_t rows = std::max(1, p->src_row_hi - p->src_row_lo + 1);
  a.seg_reduce = (p->strips != nullptr && p->n_strips >= 16 * rows * (int64_t)p->W) ? 1 : 0;
  return a;
}

void launch_vote_any(const cmax_plan* p, int motion_model, cudaStream_t s, const FusedArgs& a, float4* acc, float* iwe) {
  const int grid = event_grid(p->n, 8);
  if (motion_model == CMAX_MOTION_DENSE) launch_vote_m<CMAX_MOTION_DENSE>(p->n_ref, p->vote_variant, grid, s, a, acc, iwe);
  else if (motion_model == CMAX_MOTION_VOXEL) launch_vote_m<CMAX_MOTION_VOXEL>(p->n_ref, p->vote_variant, grid, s, a, acc, iwe);
  else launch_vote_m<CMAX_MOTION_2DOF>(p->n_ref, p->vote_variant, grid, s, a, acc, iwe);
}

void launch_grad_any(const cmax_plan* p, int motion_model, cudaStream_t s, const FusedArgs& a, const float4* gq, float* target) {
  const int grid = event_grid(p->n, 8);
  int gvar = p->grad_variant;
  if (gvar == 1 && p->order != CMAX_ORDER_PIXEL) gvar = 0;  // the segmented reduction needs source-pixel order
  if (motion_model == CMAX_MOTION_DENSE) launch_grad_m<CMAX_MOTION_DENSE>(p->n_ref, gvar, grid, s, a, gq, target);
  else if (motion_model == CMAX_MOTION_VOXEL) launch_grad_m<CMAX_MOTION_VOXEL>(p->n_ref, gvar, grid, s, a, gq, target);
  else launch_grad_m<CMAX_MOTION_2DOF>(p->n_ref, gvar, grid, s, a, gq, target);
}

}  // namespace cmax
