// The per-iteration hot path: K1 (warp + bilinear vote of every event, all reference times in one pass),
// the fold (per-corner accumulators -> IWE, with the variance sums fused in), K2 glue (blur / statistics / scalar
// cost / per-corner gradient pictures) and K3 (re-warp, gather dL/dIWE, chain to dL/dmotion).
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

#include "cmax_runs.cuh"
#include "cmax_stats.cuh"

namespace cmax {

// ------------------------------------------------------------------------------------------------ workspace
struct ObjLayout {
  size_t off_acc, off_iwe, off_iwe_full, off_blur, off_statacc, off_stats, off_affine, off_misc, off_gxy, off_g, off_g2, off_gq, total;
  int64_t cells, HW;
};

static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static ObjLayout obj_layout(int Hp, int Wp) {
  ObjLayout L;
  const int R = CMAX_MAX_REFS;
  L.cells = (int64_t)(Hp + 1) * (Wp + 1) + 1;  // + one cell no pixel reads: its gradient quad is all zero (the strip K3's 'outside' cell)
  L.HW = (int64_t)Hp * Wp;
  size_t off = 0;
  L.off_acc = off;     off = align256(off + (size_t)R * L.cells * sizeof(float4));
  L.off_iwe = off;     off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_iwe_full = off; off = align256(off + (size_t)R * L.HW * sizeof(float));  // peer exchange: the summed IWE (off_iwe stays the partial peers read)
  L.off_blur = off;    off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_stats = off;   off = align256(off + (size_t)R * 4 * sizeof(double));
  L.off_affine = off;  off = align256(off + (size_t)R * 2 * sizeof(float));
  L.off_misc = off;    off = align256(off + 2 * sizeof(double));  // 2-dof fp64 staging
  // [StatAcc block][Sobel pair] is exactly the workspace layout cmax_image_stats expects (cmax_cost.cu)
  L.off_statacc = off; off = align256(off + (size_t)R * sizeof(StatAcc));
  L.off_gxy = off;     off = align256(off + (size_t)R * 2 * L.HW * sizeof(float));
  L.off_g = off;       off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_g2 = off;      off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_gq = off;      off = align256(off + (size_t)R * L.cells * sizeof(float4));
  L.total = off;
  return L;
}

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

// ------------------------------------------------------------------------------------------------ fold
// acc -> IWE.  Every scalar component of every accumulator cell has exactly ONE reader (pixel (r,c) reads .x of cell
// (r,c), .y of (r-1,c), .z of (r,c-1), .w of (r-1,c-1)), so the reader also zeroes it: the accumulators are clean again
// for the next CM iteration and no memset is ever enqueued (components no pixel reads only ever collect votes of
// out-of-image corners and are never looked at).  Optionally the variance sums of the crop in the same pass (fp64
// accumulators), and optionally (single-GPU fused path) the last CTA to finish also evaluates the scalar cost, so that
// fold + statistics + combine are ONE launch.
__global__ void __launch_bounds__(kStatBlock) fold_kernel(float4* __restrict__ acc, float* __restrict__ iwe, int Hp, int Wp,
                                                          int64_t cells, int want_var, int omit, StatAcc* __restrict__ sacc,
                                                          double* __restrict__ stats, int want_combine, CombineDev cd,
                                                          unsigned int* __restrict__ ctas_done) {
  __shared__ double red[kStatBlock / 32];
  __shared__ bool all_done;
  pdl_trigger();
  pdl_wait();  // K1's reductions are complete and visible
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)Hp * Wp;
  float* A = reinterpret_cast<float*>(acc + img * cells);
  const int Wc = Wp + 1;
  double s = 0.0, q = 0.0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(p / Wp), c = (int)(p % Wp);
    const int64_t k = (int64_t)(r + 1) * Wc + (c + 1);
    float* a00 = A + 4 * k;                  // .x of cell (r, c)
    float* a10 = A + 4 * (k - Wc) + 1;       // .y of cell (r-1, c)
    float* a01 = A + 4 * (k - 1) + 2;        // .z of cell (r, c-1)
    float* a11 = A + 4 * (k - Wc - 1) + 3;   // .w of cell (r-1, c-1)
    const float v = ((*a00 + *a10) + *a01) + *a11;
    *a00 = 0.f;
    *a10 = 0.f;
    *a01 = 0.f;
    *a11 = 0.f;
    iwe[img * HW + p] = v;
    if (want_var && (!omit || (r >= 1 && r <= Hp - 2 && c >= 1 && c <= Wp - 2))) {
      s += (double)v;
      q += (double)v * (double)v;
    }
  }
  if (want_var) {
    const int64_t M = omit ? (int64_t)(Hp - 2) * (Wp - 2) : HW;
    variance_commit(s, q, M, gridDim.x, &sacc[img], stats + 4 * img, red);
    if (want_combine) {
      if (threadIdx.x == 0) {
        __threadfence();
        all_done = (atomicAdd(ctas_done, 1u) == gridDim.x * gridDim.y - 1);
      }
      __syncthreads();
      if (all_done && threadIdx.x == 0) {
        __threadfence();
        combine_eval(stats, cd);
      }
    }
  } else if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < 64) {
    // statistics not fused here: leave the 256-byte statistics block clean for whoever accumulates next
    // (the peer IWE reduction of the multi-GPU path), instead of a memset node
    reinterpret_cast<unsigned int*>(sacc)[threadIdx.x] = 0u;
  }
}

// ------------------------------------------------------------------------------------------------ gq
// Per-corner gradient pictures.  G[p] = a * (src[p] - m) (inside the crop when `crop`, else everywhere), gathered at
// the four corners of every accumulator cell with the per-corner in-bounds masks.  Also zeroes `zero` (the motion
// gradient K3 is about to accumulate into) so that no memset is enqueued for it.
__global__ void __launch_bounds__(256) gq_build_kernel(const float* __restrict__ src, const float* __restrict__ affine, int Hp, int Wp,
                                                       int64_t cells, int crop, float4* __restrict__ gq, float* __restrict__ zero,
                                                       int64_t n_zero) {
  pdl_trigger();  // K3 may start prefetching its event tiles now
  pdl_wait();     // the IWE / statistics / affine pair of the predecessor are complete
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)Hp * Wp;
  const float* I = src + img * HW;
  const float a = affine[2 * img], m = affine[2 * img + 1];
  const int Wc = Wp + 1;
  const int lo = crop ? 1 : 0, hi_r = crop ? Hp - 2 : Hp - 1, hi_c = crop ? Wp - 2 : Wp - 1;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < cells; k += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(k / Wc) - 1, c = (int)(k % Wc) - 1;
    auto g = [&](int rr, int cc) -> float {
      return (rr >= lo && rr <= hi_r && cc >= lo && cc <= hi_c) ? a * (__ldg(I + (int64_t)rr * Wp + cc) - m) : 0.f;
    };
    gq[img * cells + k] = make_float4(g(r, c), g(r + 1, c), g(r, c + 1), g(r + 1, c + 1));
  }
  if (zero != nullptr && img == 0)
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n_zero; k += (int64_t)gridDim.x * blockDim.x) zero[k] = 0.f;
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
        if (a.dbg & 2) {
          st.f0 = 3.0f + (float)(st.src & 7);
          st.f1 = -2.0f;
        } else {
          st.f0 = __ldg(a.motion + st.src);
          st.f1 = __ldg(a.motion + HW + st.src);
        }
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
    red_add_v4_if(!same && st.cell[r] >= 0 && !(a.dbg & 1), acc + r * a.cells + max(st.cell[r], 0), st.w0[r], st.w1[r], st.w2[r], st.w3[r]);
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

// 2-dof gradient: the CTAs accumulate in two doubles (off_misc), narrowed here.
__global__ void finish_2dof_kernel(const double* __restrict__ acc2, float* __restrict__ out) {
  if (threadIdx.x < 2) out[threadIdx.x] = (float)acc2[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------ peer reductions
// Multi-GPU exchange over NVLink peer memory (the workspaces live in symmetric memory, one process per GPU): instead
// of an NCCL all-reduce followed by the statistics kernels, ONE kernel on every rank reads all ranks' partial IWEs
// (P2P loads, summed in rank order so every rank gets the bit-identical image), writes the full IWE and -- exactly like
// the single-GPU fold -- accumulates the variance sums and lets its last CTA evaluate the scalar cost.
struct PeerPtrs {
  const float* p[CMAX_MAX_PEERS];
  int n;
};

// ---- push exchange: flags instead of barrier kernels.  Every rank owns a MAILBOX in symmetric memory: one slot per
// source rank and one 32-bit flag per source rank.  A producer (push_kernel) stores its partial result into its slot of
// EVERY rank's mailbox (posted NVLink stores, no round trip), fences at system scope, and its last CTA then stores the new
// epoch into its flag on every rank.  A consumer kernel starts by waiting until all flags of its own mailbox carry the
// current epoch and then reads only LOCAL memory.  Two mailboxes alternate (IWE, gradient), which is what makes reuse
// safe without a barrier: a rank can only start overwriting its IWE slots for evaluation e+1 after it has seen every
// peer's gradient flag of evaluation e, and a peer raises that flag (stream order) after its IWE reduction of e.
struct PushPtrs {
  float* slot[CMAX_MAX_PEERS];
  uint32_t* flag[CMAX_MAX_PEERS];
  int n;
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Block until flags[0..n) have all reached *epoch (called by every CTA of a consumer kernel).  A peer that never
// arrives (crashed rank) would hang the GPU: after ~4 s the kernel traps instead, which surfaces as a CUDA error.
__device__ __forceinline__ void wait_flags(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ epoch, int n) {
  if (flags == nullptr) return;
  if ((int)threadIdx.x < n) {
    const uint32_t e = *reinterpret_cast<const volatile uint32_t*>(epoch);
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(flags + threadIdx.x) - e) < 0) {
      if (clock64() - t0 > 8000000000ll) __trap();
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256) push_kernel(const float* __restrict__ src, int64_t n, PushPtrs dst, uint32_t* __restrict__ epoch,
                                                   uint32_t* __restrict__ counter) {
  __shared__ bool last;
  if ((n & 3) == 0) {  // (slots and sources are 256-byte aligned)
    const int64_t n4 = n >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      const float4 v = reinterpret_cast<const float4*>(src)[i];
      for (int q = 0; q < dst.n; ++q) reinterpret_cast<float4*>(dst.slot[q])[i] = v;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      const float v = src[i];
      for (int q = 0; q < dst.n; ++q) dst.slot[q][i] = v;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // one system-scope fence per CTA (fences are cumulative: the stores of the whole CTA, ordered before this thread by
    // the barrier, are visible system-wide before the CTA is counted); a fence per thread costs ~20 us per launch
    __threadfence_system();
    last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence_system();
    const uint32_t e = *epoch + 1u;
    *epoch = e;
    *counter = 0u;
    for (int q = 0; q < dst.n; ++q) st_release_sys(dst.flag[q], e);
  }
}

__global__ void wait_kernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ epoch, int n) { wait_flags(flags, epoch, n); }

__global__ void __launch_bounds__(kStatBlock) peer_iwe_kernel(PeerPtrs peers, float* __restrict__ iwe, int Hp, int Wp, int want_var, int omit,
                                                              StatAcc* __restrict__ sacc, double* __restrict__ stats, int want_combine,
                                                              CombineDev cd, unsigned int* __restrict__ ctas_done,
                                                              const uint32_t* __restrict__ flags, const uint32_t* __restrict__ epoch) {
  __shared__ double red[kStatBlock / 32];
  __shared__ bool all_done;
  wait_flags(flags, epoch, peers.n);  // push exchange: every rank's partial image has landed in this rank's mailbox
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)Hp * Wp;
  double s = 0.0, q = 0.0;
  auto account = [&](int64_t p, float v) {
    const int rr = (int)(p / Wp), c = (int)(p % Wp);
    if (want_var && (!omit || (rr >= 1 && rr <= Hp - 2 && c >= 1 && c <= Wp - 2))) {
      s += (double)v;
      q += (double)v * (double)v;
    }
  };
  if ((HW & 3) == 0) {
    // 16-byte loads, all ranks' loads of a thread in flight together: one NVLink round trip per thread instead of 4 n
    for (int64_t p4 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p4 < (HW >> 2); p4 += (int64_t)gridDim.x * blockDim.x) {
      float4 part[CMAX_MAX_PEERS];
#pragma unroll
      for (int r = 0; r < CMAX_MAX_PEERS; ++r)
        if (r < peers.n) part[r] = __ldcg(reinterpret_cast<const float4*>(peers.p[r] + img * HW) + p4);  // L2-coherent: written by another GPU
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < CMAX_MAX_PEERS; ++r)
        if (r < peers.n) {  // rank order: every rank computes the bit-identical sum
          v.x += part[r].x; v.y += part[r].y; v.z += part[r].z; v.w += part[r].w;
        }
      reinterpret_cast<float4*>(iwe + img * HW)[p4] = v;
      account(4 * p4, v.x); account(4 * p4 + 1, v.y); account(4 * p4 + 2, v.z); account(4 * p4 + 3, v.w);
    }
  } else {
    for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
      float v = 0.f;
      for (int r = 0; r < peers.n; ++r) v += __ldcg(peers.p[r] + img * HW + p);
      iwe[img * HW + p] = v;
      account(p, v);
    }
  }
  if (want_var) {
    const int64_t M = omit ? (int64_t)(Hp - 2) * (Wp - 2) : HW;
    variance_commit(s, q, M, gridDim.x, &sacc[img], stats + 4 * img, red);
    if (want_combine) {
      if (threadIdx.x == 0) {
        __threadfence();
        all_done = (atomicAdd(ctas_done, 1u) == gridDim.x * gridDim.y - 1);
      }
      __syncthreads();
      if (all_done && threadIdx.x == 0) {
        __threadfence();
        combine_eval(stats, cd);
      }
    }
  }
}

// out[i] = sum over ranks of peers[r][i], rank order (the motion-gradient exchange)
__global__ void __launch_bounds__(256) peer_sum_kernel(PeerPtrs peers, int64_t n, float* __restrict__ out, const uint32_t* __restrict__ flags,
                                                       const uint32_t* __restrict__ epoch) {
  wait_flags(flags, epoch, peers.n);
  if ((n & 3) == 0) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < (n >> 2); i += (int64_t)gridDim.x * blockDim.x) {
      float4 part[CMAX_MAX_PEERS];
#pragma unroll
      for (int r = 0; r < CMAX_MAX_PEERS; ++r)
        if (r < peers.n) part[r] = __ldcg(reinterpret_cast<const float4*>(peers.p[r]) + i);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < CMAX_MAX_PEERS; ++r)
        if (r < peers.n) {
          v.x += part[r].x; v.y += part[r].y; v.z += part[r].z; v.w += part[r].w;
        }
      reinterpret_cast<float4*>(out)[i] = v;
    }
  } else {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      float v = 0.f;
      for (int r = 0; r < peers.n; ++r) v += __ldcg(peers.p[r] + i);
      out[i] = v;
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
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)kNumSMs * per_sm));
}
static inline int image_grid(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + 255) / 256, (int64_t)kNumSMs * 4));
}

static FusedArgs fused_args(const cmax_plan* p, const float* motion) {
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
  {
    static int dbg = -1;
    if (dbg < 0) {
      const char* e = getenv("CMAX_DEBUG");
      dbg = e ? atoi(e) : 0;
    }
    a.dbg = dbg;
  }
  return a;
}

static int check_model(const char* fn, const cmax_plan* p, int model) {
  CMAX_REQUIRE(p != nullptr, "%s: plan is NULL", fn);
  CMAX_REQUIRE(model == CMAX_MOTION_DENSE || model == CMAX_MOTION_VOXEL || model == CMAX_MOTION_2DOF,
               "%s: motion model %d not supported", fn, model);
  CMAX_REQUIRE(model != CMAX_MOTION_VOXEL || p->n_bins >= 1, "%s: dense-flow-voxel needs cmax_plan_set_refs(..., n_bins >= 1)", fn);
  return CMAX_OK;
}

static bool can_fuse_stats(const cmax_cost_spec* spec) {
  return spec != nullptr && spec->stat == CMAX_STAT_VARIANCE && !(spec->sigma > 0.f);
}

static int check_spec(const char* fn, const cmax_cost_spec* spec, int n_ref) {
  CMAX_REQUIRE(spec != nullptr, "%s: spec is NULL", fn);
  CMAX_REQUIRE(spec->stat == CMAX_STAT_VARIANCE || spec->stat == CMAX_STAT_GRADMAG, "%s: unknown statistic %d", fn, spec->stat);
  CMAX_REQUIRE(spec->form >= CMAX_COST_PLAIN && spec->form <= CMAX_COST_MULTIFOCAL, "%s: unknown cost form %d", fn, spec->form);
  CMAX_REQUIRE(spec->direction_sign == 1 || spec->direction_sign == -1, "%s: direction_sign must be +1 or -1", fn);
  CMAX_REQUIRE(spec->form != CMAX_COST_PLAIN || n_ref == 1, "%s: a plain cost takes exactly one reference time (plan has %d)", fn, n_ref);
  CMAX_REQUIRE(spec->form != CMAX_COST_NORMALIZED || n_ref == 1, "%s: a normalised cost takes exactly one reference time (plan has %d)", fn, n_ref);
  CMAX_REQUIRE(!(spec->sigma < 0.f), "%s: sigma must be >= 0", fn);
  return CMAX_OK;
}

// ------------------------------------------------------------------------------------------------ stages
struct Ws {
  float4* acc; float* iwe; float* iwe_full; float* blur; StatAcc* sacc; double* stats; float* affine; unsigned int* ctas_done;
  float* G; float* G2; float4* gq; char* stats_ws; double* acc2;
};

static Ws carve(void* workspace, const ObjLayout& L) {
  char* ws = static_cast<char*>(workspace);
  Ws w;
  w.acc = reinterpret_cast<float4*>(ws + L.off_acc);
  w.iwe = reinterpret_cast<float*>(ws + L.off_iwe);
  w.iwe_full = reinterpret_cast<float*>(ws + L.off_iwe_full);
  w.blur = reinterpret_cast<float*>(ws + L.off_blur);
  w.sacc = reinterpret_cast<StatAcc*>(ws + L.off_statacc);
  w.stats = reinterpret_cast<double*>(ws + L.off_stats);
  w.affine = reinterpret_cast<float*>(ws + L.off_affine);
  w.ctas_done = reinterpret_cast<unsigned int*>(ws + L.off_statacc + 192);  // inside the 256-byte StatAcc block
  w.G = reinterpret_cast<float*>(ws + L.off_g);
  w.G2 = reinterpret_cast<float*>(ws + L.off_g2);
  w.gq = reinterpret_cast<float4*>(ws + L.off_gq);
  w.stats_ws = ws + L.off_statacc;  // [StatAcc block][Sobel pair], the layout cmax_image_stats expects
  w.acc2 = reinterpret_cast<double*>(ws + L.off_misc);  // 2-dof fp64 staging
  return w;
}

static inline size_t motion_floats(const cmax_plan* p, int motion_model) {
  const size_t HW = (size_t)p->H * p->W;
  if (motion_model == CMAX_MOTION_DENSE) return 2 * HW;
  if (motion_model == CMAX_MOTION_VOXEL) return 2 * (size_t)p->n_bins * HW;
  return 2;
}

// Stage 1.  `fused_combine` (may be NULL): when the statistics can be fused into the fold, also evaluate the scalar
// cost in the fold's last CTA (single-GPU convenience path).
static int vote_stage(const cmax_plan* p, int motion_model, const float* motion, void* workspace, const cmax_cost_spec* fuse_spec,
                      const CombineDev* fused_combine, int32_t* stats_fused, cudaStream_t s) {
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  const int n_ref = p->n_ref;
  FusedArgs a = fused_args(p, motion);
  const int variant = p->vote_variant;
  const bool fuse = can_fuse_stats(fuse_spec) && variant != 1;
  const int mask = p->stage_mask;
  static_assert(sizeof(StatAcc) * CMAX_MAX_REFS <= 192, "StatAcc block and the CTA counter share 256 bytes");
  // the 256-byte statistics block is cleared by K1's first CTA when there is one (variant 2), else by a memset
  const bool k1_clears = fuse && variant >= 2 && p->n > 0 && (mask & 2);
  if (k1_clears) a.zero256 = reinterpret_cast<unsigned int*>(w.sacc);
  if (mask & 1) {
    // the per-corner accumulators are left clean by the fold (see fold_kernel) and by cmax_objective_workspace_init
    if (variant == 1) CMAX_CUDA_CHECK(cudaMemsetAsync(w.iwe, 0, (size_t)n_ref * L.HW * sizeof(float), s));
    if (fuse && !k1_clears) CMAX_CUDA_CHECK(cudaMemsetAsync(w.sacc, 0, 256, s));
  }
  if (p->n > 0 && (mask & 2)) {
    const int grid = event_grid(p->n, 8);
    if (motion_model == CMAX_MOTION_DENSE) launch_vote_m<CMAX_MOTION_DENSE>(n_ref, variant, grid, s, a, w.acc, w.iwe);
    else if (motion_model == CMAX_MOTION_VOXEL) launch_vote_m<CMAX_MOTION_VOXEL>(n_ref, variant, grid, s, a, w.acc, w.iwe);
    else launch_vote_m<CMAX_MOTION_2DOF>(n_ref, variant, grid, s, a, w.acc, w.iwe);
  }
  if (variant != 1 && (mask & 4)) {
    dim3 grid((unsigned)std::min<int64_t>((L.HW + kStatBlock - 1) / kStatBlock, kNumSMs * 4), n_ref);
    CombineDev cd;
    memset(&cd, 0, sizeof(cd));
    if (fuse && fused_combine) cd = *fused_combine;
    launch_k(pdl_enabled(), fold_kernel, grid, dim3(kStatBlock), s, w.acc, w.iwe, p->Hp, p->Wp, L.cells, fuse ? 1 : 0,
             fuse ? fuse_spec->omit_boundary : 0, w.sacc, w.stats, (fuse && fused_combine) ? 1 : 0, cd, w.ctas_done);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  if (stats_fused) *stats_fused = fuse ? 1 : 0;
  return CMAX_OK;
}

static CombineDev combine_for(const cmax_plan* p, const cmax_cost_spec* spec, const double* d_orig_stat, double* d_cost, const Ws& w) {
  const bool explicit_grad = spec->sigma > 0.f || spec->stat == CMAX_STAT_GRADMAG;
  return make_combine(p->n_ref, spec->stat, spec->form, spec->direction_sign, explicit_grad ? 1 : 0, spec->weights, d_orig_stat, d_cost,
                      w.affine);
}

// Stage 2.  combined != 0: the scalar combination already ran inside the fold (cmax_objective's fast path).
// zero_grad (may be NULL): motion-gradient buffer to clear inside the gradient-picture kernel.
static int cost_stage(const cmax_plan* p, const cmax_cost_spec* spec, const double* d_orig_stat, void* workspace, int stats_fused,
                      int combined, int want_grad, double* d_cost, float* zero_grad, size_t n_zero, cmax_stream_t stream,
                      bool use_full = false) {
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  cudaStream_t s = as_stream(stream);
  const int n_ref = p->n_ref;
  const bool blurred = spec->sigma > 0.f;
  const float* img = use_full ? w.iwe_full : w.iwe;
  int rc;
  if (!(p->stage_mask & 4)) return CMAX_OK;
  if (blurred) {
    rc = cmax_blur3(img, w.blur, n_ref, p->Hp, p->Wp, spec->sigma, 0, stream);
    if (rc) return rc;
    img = w.blur;
  }
  // variance without blur needs no explicit gradient image: dL/dIWE is affine in the IWE
  const bool explicit_grad = blurred || spec->stat == CMAX_STAT_GRADMAG;
  if (!stats_fused) {
    static_assert(sizeof(StatAcc) * CMAX_MAX_REFS <= 256, "StatAcc block must fit the 256 bytes before the Sobel pair");
    rc = cmax_image_stats(img, n_ref, p->Hp, p->Wp, spec->stat, spec->omit_boundary, w.stats, (want_grad && explicit_grad) ? w.G : nullptr,
                          w.stats_ws, stream);
    if (rc) return rc;
  }
  if (!combined) launch_combine(w.stats, combine_for(p, spec, d_orig_stat, d_cost, w), s);
  if (want_grad) {
    const float* gsrc = img;
    int crop = spec->omit_boundary ? 1 : 0;
    if (explicit_grad) {
      gsrc = w.G;
      crop = 0;
      if (blurred) {
        rc = cmax_blur3(w.G, w.G2, n_ref, p->Hp, p->Wp, spec->sigma, 1, stream);
        if (rc) return rc;
        gsrc = w.G2;
      }
    }
    dim3 grid((unsigned)image_grid(L.cells), n_ref);
    launch_k(pdl_enabled(), gq_build_kernel, grid, dim3(256), s, gsrc, (const float*)w.affine, p->Hp, p->Wp, L.cells, crop, w.gq, zero_grad,
             (int64_t)n_zero);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

// Stage 3.  pre_zeroed: grad_motion was cleared by stage 2.
static int grad_stage(const cmax_plan* p, int motion_model, const float* motion, void* workspace, float* grad_motion, int pre_zeroed,
                      cudaStream_t s) {
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  const FusedArgs a = fused_args(p, motion);
  if (p->stage_mask & 1) {
    if (!pre_zeroed) CMAX_CUDA_CHECK(cudaMemsetAsync(grad_motion, 0, motion_floats(p, motion_model) * sizeof(float), s));
    if (motion_model == CMAX_MOTION_2DOF) CMAX_CUDA_CHECK(cudaMemsetAsync(w.acc2, 0, 2 * sizeof(double), s));
  }
  if (p->n > 0 && (p->stage_mask & 2)) {
    const int grid = event_grid(p->n, 8);
    int gvar = p->grad_variant;
    if (gvar == 1 && p->order != CMAX_ORDER_PIXEL) gvar = 0;  // the segmented reduction needs source-pixel order
    float* target = (motion_model == CMAX_MOTION_2DOF) ? reinterpret_cast<float*>(w.acc2) : grad_motion;
    if (motion_model == CMAX_MOTION_DENSE) launch_grad_m<CMAX_MOTION_DENSE>(p->n_ref, gvar, grid, s, a, w.gq, target);
    else if (motion_model == CMAX_MOTION_VOXEL) launch_grad_m<CMAX_MOTION_VOXEL>(p->n_ref, gvar, grid, s, a, w.gq, target);
    else launch_grad_m<CMAX_MOTION_2DOF>(p->n_ref, gvar, grid, s, a, w.gq, target);
    if (motion_model == CMAX_MOTION_2DOF) finish_2dof_kernel<<<1, 32, 0, s>>>(w.acc2, grad_motion);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // namespace cmax

using namespace cmax;

extern "C" {

size_t cmax_objective_workspace_bytes(const cmax_plan_t* plan, const cmax_cost_spec* spec) {
  (void)spec;
  if (plan == nullptr) {
    set_error("cmax_objective_workspace_bytes: plan is NULL");
    return 0;
  }
  return obj_layout(plan->Hp, plan->Wp).total;
}

int cmax_objective_workspace_init(const cmax_plan_t* plan, void* workspace, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr && workspace != nullptr, "cmax_objective_workspace_init: NULL argument");
  CMAX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cmax_objective_workspace_init: workspace must be 256-byte aligned");
  CMAX_CUDA_CHECK(cudaMemsetAsync(workspace, 0, obj_layout(plan->Hp, plan->Wp).total, as_stream(stream)));
  return CMAX_OK;
}

int cmax_objective_vote(const cmax_plan_t* plan, int motion_model, const float* motion, void* workspace, float** iwe_out,
                        const cmax_cost_spec* fuse_spec, int32_t* stats_fused, cmax_stream_t stream) {
  int rc = check_model("cmax_objective_vote", plan, motion_model);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && workspace != nullptr, "cmax_objective_vote: NULL motion/workspace");
  CMAX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cmax_objective_vote: workspace must be 256-byte aligned");
  rc = vote_stage(plan, motion_model, motion, workspace, fuse_spec, nullptr, stats_fused, as_stream(stream));
  if (rc) return rc;
  if (iwe_out) *iwe_out = carve(workspace, obj_layout(plan->Hp, plan->Wp)).iwe;
  return CMAX_OK;
}

int cmax_objective_cost(const cmax_plan_t* plan, const cmax_cost_spec* spec, const double* d_orig_stat, void* workspace,
                        int stats_fused, int want_grad, double* d_cost, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr && workspace != nullptr && d_cost != nullptr, "cmax_objective_cost: NULL argument");
  int rc = check_spec("cmax_objective_cost", spec, plan->n_ref);
  if (rc) return rc;
  CMAX_REQUIRE(spec->form == CMAX_COST_PLAIN || d_orig_stat != nullptr, "cmax_objective_cost: normalised costs need d_orig_stat");
  CMAX_REQUIRE(plan->Hp >= 3 && plan->Wp >= 3, "cmax_objective_cost: images must be at least 3x3");
  CMAX_REQUIRE(!stats_fused || can_fuse_stats(spec), "cmax_objective_cost: stats_fused set for a spec that cannot fuse");
  return cost_stage(plan, spec, d_orig_stat, workspace, stats_fused, 0, want_grad, d_cost, nullptr, 0, stream);
}

int cmax_objective_grad(const cmax_plan_t* plan, int motion_model, const float* motion, void* workspace, float* grad_motion,
                        cmax_stream_t stream) {
  int rc = check_model("cmax_objective_grad", plan, motion_model);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && workspace != nullptr && grad_motion != nullptr, "cmax_objective_grad: NULL argument");
  return grad_stage(plan, motion_model, motion, workspace, grad_motion, 0, as_stream(stream));
}

int cmax_objective(const cmax_plan_t* plan, int motion_model, const float* motion, const cmax_cost_spec* spec,
                   const double* d_orig_stat, void* workspace, double* d_cost, float* grad_motion, cmax_stream_t stream) {
  int rc = check_model("cmax_objective", plan, motion_model);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && workspace != nullptr && d_cost != nullptr, "cmax_objective: NULL argument");
  CMAX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cmax_objective: workspace must be 256-byte aligned");
  rc = check_spec("cmax_objective", spec, plan->n_ref);
  if (rc) return rc;
  CMAX_REQUIRE(spec->form == CMAX_COST_PLAIN || d_orig_stat != nullptr, "cmax_objective: normalised costs need d_orig_stat");
  CMAX_REQUIRE(plan->Hp >= 3 && plan->Wp >= 3, "cmax_objective: images must be at least 3x3");
  const Ws w = carve(workspace, obj_layout(plan->Hp, plan->Wp));
  const CombineDev cd = combine_for(plan, spec, d_orig_stat, d_cost, w);
  int32_t fused = 0;
  rc = vote_stage(plan, motion_model, motion, workspace, spec, &cd, &fused, as_stream(stream));
  if (rc) return rc;
  // with everything enabled the gradient buffer is cleared inside the gradient-picture kernel instead of by a memset
  const bool zero_in_gq = grad_motion != nullptr && plan->stage_mask == 7;
  rc = cost_stage(plan, spec, d_orig_stat, workspace, fused, fused, grad_motion != nullptr, d_cost, zero_in_gq ? grad_motion : nullptr,
                  zero_in_gq ? motion_floats(plan, motion_model) : 0, stream);
  if (rc) return rc;
  if (grad_motion != nullptr) rc = grad_stage(plan, motion_model, motion, workspace, grad_motion, zero_in_gq ? 1 : 0, as_stream(stream));
  return rc;
}

int cmax_objective_reduce_iwe(const cmax_plan_t* plan, const cmax_cost_spec* spec, const float* const* h_peer_iwe, int n_peers,
                              const double* d_orig_stat, void* workspace, double* d_cost, int32_t* combined, const uint32_t* d_flags,
                              const uint32_t* d_epoch, cmax_stream_t stream) {
  CMAX_REQUIRE((d_flags == nullptr) == (d_epoch == nullptr), "cmax_objective_reduce_iwe: d_flags and d_epoch go together");
  CMAX_REQUIRE(plan != nullptr && workspace != nullptr && h_peer_iwe != nullptr && d_cost != nullptr, "cmax_objective_reduce_iwe: NULL argument");
  CMAX_REQUIRE(n_peers >= 1 && n_peers <= CMAX_MAX_PEERS, "cmax_objective_reduce_iwe: n_peers must be in [1,%d], got %d", CMAX_MAX_PEERS, n_peers);
  int rc = check_spec("cmax_objective_reduce_iwe", spec, plan->n_ref);
  if (rc) return rc;
  CMAX_REQUIRE(spec->form == CMAX_COST_PLAIN || d_orig_stat != nullptr, "cmax_objective_reduce_iwe: normalised costs need d_orig_stat");
  const cmax_plan* p = plan;
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  cudaStream_t s = as_stream(stream);
  PeerPtrs peers;
  peers.n = n_peers;
  for (int r = 0; r < CMAX_MAX_PEERS; ++r) peers.p[r] = r < n_peers ? h_peer_iwe[r] : nullptr;
  for (int r = 0; r < n_peers; ++r) CMAX_REQUIRE(peers.p[r] != nullptr, "cmax_objective_reduce_iwe: peer %d IWE pointer is NULL", r);
  const bool fuse = can_fuse_stats(spec);
  // the statistics block was left clean by this rank's fold (fold_kernel); variant 1 has no fold
  if (fuse && p->vote_variant == 1) CMAX_CUDA_CHECK(cudaMemsetAsync(w.sacc, 0, 256, s));
  CombineDev cd;
  memset(&cd, 0, sizeof(cd));
  if (fuse) cd = combine_for(p, spec, d_orig_stat, d_cost, w);
  const int64_t work = (L.HW & 3) == 0 ? L.HW / 4 : L.HW;  // threads with work (16-byte loads when the image allows)
  dim3 grid((unsigned)std::min<int64_t>((work + kStatBlock - 1) / kStatBlock, kNumSMs * 4), p->n_ref);
  peer_iwe_kernel<<<grid, kStatBlock, 0, s>>>(peers, w.iwe_full, p->Hp, p->Wp, fuse ? 1 : 0, fuse ? spec->omit_boundary : 0, w.sacc, w.stats,
                                              fuse ? 1 : 0, cd, w.ctas_done, d_flags, d_epoch);
  CMAX_CUDA_CHECK(cudaGetLastError());
  if (combined) *combined = fuse ? 1 : 0;
  return CMAX_OK;
}

int cmax_objective_cost_after_reduce(const cmax_plan_t* plan, const cmax_cost_spec* spec, const double* d_orig_stat, void* workspace,
                                     int combined, int want_grad, double* d_cost, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr && workspace != nullptr && d_cost != nullptr, "cmax_objective_cost_after_reduce: NULL argument");
  int rc = check_spec("cmax_objective_cost_after_reduce", spec, plan->n_ref);
  if (rc) return rc;
  CMAX_REQUIRE(!combined || can_fuse_stats(spec), "cmax_objective_cost_after_reduce: combined set for a spec that cannot fuse");
  return cost_stage(plan, spec, d_orig_stat, workspace, combined, combined, want_grad, d_cost, nullptr, 0, stream, true);
}

int cmax_push(const float* src, int64_t n, float* const* h_peer_slots, uint32_t* const* h_peer_flags, int n_peers, uint32_t* d_epoch,
              uint32_t* d_counter, cmax_stream_t stream) {
  CMAX_REQUIRE(h_peer_slots != nullptr && h_peer_flags != nullptr && d_epoch != nullptr && d_counter != nullptr, "cmax_push: NULL argument");
  CMAX_REQUIRE(n_peers >= 1 && n_peers <= CMAX_MAX_PEERS, "cmax_push: n_peers must be in [1,%d], got %d", CMAX_MAX_PEERS, n_peers);
  CMAX_REQUIRE(n >= 0 && (n == 0 || src != nullptr), "cmax_push: bad source");
  PushPtrs dst;
  dst.n = n_peers;
  for (int r = 0; r < CMAX_MAX_PEERS; ++r) {
    dst.slot[r] = r < n_peers ? h_peer_slots[r] : nullptr;
    dst.flag[r] = r < n_peers ? h_peer_flags[r] : nullptr;
  }
  for (int r = 0; r < n_peers; ++r) CMAX_REQUIRE(dst.flag[r] != nullptr && (n == 0 || dst.slot[r] != nullptr), "cmax_push: peer %d pointer is NULL", r);
  const int grid = n == 0 ? 1 : (int)std::max<int64_t>(1, std::min<int64_t>(((n + 3) / 4 + 255) / 256, (int64_t)kNumSMs * 2));
  push_kernel<<<grid, 256, 0, as_stream(stream)>>>(src, n, dst, d_epoch, d_counter);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

int cmax_reduce_peers(const float* const* h_peer_bufs, int n_peers, int64_t n, float* out, const uint32_t* d_flags, const uint32_t* d_epoch,
                      cmax_stream_t stream) {
  CMAX_REQUIRE((d_flags == nullptr) == (d_epoch == nullptr), "cmax_reduce_peers: d_flags and d_epoch go together");
  if (n == 0 && d_flags != nullptr) {  // nothing to sum: just consume the flags (closes a value-only evaluation)
    CMAX_REQUIRE(n_peers >= 1 && n_peers <= CMAX_MAX_PEERS, "cmax_reduce_peers: n_peers must be in [1,%d], got %d", CMAX_MAX_PEERS, n_peers);
    wait_kernel<<<1, 32, 0, as_stream(stream)>>>(d_flags, d_epoch, n_peers);
    CMAX_CUDA_CHECK(cudaGetLastError());
    return CMAX_OK;
  }
  CMAX_REQUIRE(h_peer_bufs != nullptr && out != nullptr, "cmax_reduce_peers: NULL argument");
  CMAX_REQUIRE(n_peers >= 1 && n_peers <= CMAX_MAX_PEERS, "cmax_reduce_peers: n_peers must be in [1,%d], got %d", CMAX_MAX_PEERS, n_peers);
  CMAX_REQUIRE(n >= 0, "cmax_reduce_peers: n must be >= 0");
  PeerPtrs peers;
  peers.n = n_peers;
  for (int r = 0; r < CMAX_MAX_PEERS; ++r) peers.p[r] = r < n_peers ? h_peer_bufs[r] : nullptr;
  for (int r = 0; r < n_peers; ++r) CMAX_REQUIRE(peers.p[r] != nullptr, "cmax_reduce_peers: peer %d pointer is NULL", r);
  if (n > 0) {
    peer_sum_kernel<<<image_grid((n & 3) == 0 ? n / 4 : n), 256, 0, as_stream(stream)>>>(peers, n, out, d_flags, d_epoch);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

size_t cmax_objective_full_iwe_offset(const cmax_plan_t* plan) {
  if (plan == nullptr) {
    set_error("cmax_objective_full_iwe_offset: plan is NULL");
    return 0;
  }
  return obj_layout(plan->Hp, plan->Wp).off_iwe_full;
}

size_t cmax_objective_iwe_offset(const cmax_plan_t* plan) {
  if (plan == nullptr) {
    set_error("cmax_objective_iwe_offset: plan is NULL");
    return 0;
  }
  return obj_layout(plan->Hp, plan->Wp).off_iwe;
}

}  // extern "C"
