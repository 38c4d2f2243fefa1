// Everything image-sized between the two event passes of a CM iteration, in ONE launch:
//   fold      per-corner accumulators -> IWE (and the accumulators are left clean for the next iteration)
//   exchange  (sharded batches) raise this rank's flag on every peer, wait for theirs, sum all ranks' partial IWEs
//   cost      variance sums (fp64, deterministic order), scalar cost, the affine pair of dL/dIWE
//   gq        per-corner gradient quads K3 gathers from; the motion-gradient buffer is cleared on the way
// `image_kernel` is a persistent grid of at most one CTA per SM with grid-wide barriers between the phases, launched with
// programmatic stream serialisation behind K1 and in front of K3: a CM iteration is K1 -> image_kernel -> K3.
// Costs the fold cannot absorb (gradient magnitude, blurred IWEs) run the operator kernels of cmax_cost.cu / cmax_ops.cu
// between a fold-only and a gq-only launch of the same kernel.  Also here: the stages and the C ABI of the fused path.
#include <stdlib.h>

#include "cmax_objective.cuh"

namespace cmax {

// ------------------------------------------------------------------------------------------------ cross-GPU signalling
// One process per GPU; workspaces, partial gradients and flags live in symmetric (peer-mapped) memory.  A flag is a
// monotonically increasing evaluation counter: rank r stores epoch e into ITS slot of every rank's flag array once its
// partial result of evaluation e is complete and visible (one gpu-scope fence + one posted store per peer, issued by the last
// CTA to finish -- no barrier kernel, no round trip), and a consumer spins on its OWN (local) flag array until every
// source rank has reached e.  Two flag arrays alternate per evaluation (IWE, gradient), which is what makes buffer reuse
// safe without further synchronisation: a rank overwrites its partial IWE for evaluation e+1 only after it has seen
// every peer's gradient flag of e, and a peer raises that flag (stream order) after it has finished reading the IWEs of e.
// Flags are written and polled with RELAXED system-scope accesses and there is NO system-scope fence anywhere: measured on
// 2 x B200, a fence.sys costs 2-3 us where it stands (st.release.sys = one per peer; one per consumer CTA after the poll
// turned a 58 us step into 71 us).  What makes the protocol correct without them: every buffer a peer reads lives in the
// PRODUCER's memory, whose L2 is the point of coherence for its own SMs and for NVLink peer reads alike -- a gpu-scope
// fence (all prior writes performed at that L2) before the posted flag store is enough for a reader that first sees the
// flag and then issues its loads; the consumers' loads are .cg (never served from an L1) and control-dependent on the poll.
__device__ __forceinline__ uint4 ld_relaxed_sys_v4(const uint4* p) {
  uint4 v;
  asm volatile("ld.relaxed.sys.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_sys_v4(uint4* p, uint4 v) {
  asm volatile("st.relaxed.sys.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ unsigned int ld_relaxed_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// A flag is one 16-byte RECORD {epoch, row_lo, row_hi, -}, written with one 16-byte store and polled with 16-byte loads: next
// to "my partial result of evaluation `epoch` is complete" it tells the reader which image rows of that partial result can be
// non-zero at all.  A reader skips the peers whose range misses its rows (exact: what it skips are zeros), so when the
// shards are spatially compact (distributed.reshard_events_by_pixel) the exchange moves the overlaps only instead of N whole
// images per rank; for time-sliced shards the ranges are the whole image and nothing changes.
struct PeerEx {
  int n, rank;                             // n == 0: not sharded
  const float* part[CMAX_MAX_PEERS];       // every rank's partial buffer (rank order)
  uint4* rec_at[CMAX_MAX_PEERS];           // this rank's record in rank q's block
  const uint4* recs;                       // this rank's own block, one record per source rank
  uint32_t* epoch;                         // this rank's evaluation counter (advanced by image_kernel)
};

// Block until every source rank's record has reached `epoch` (every CTA of a consumer calls this); the row ranges land in
// shared memory.  A peer that never arrives (crashed rank) would hang the GPU: after ~4 s the kernel traps instead, which
// surfaces as a CUDA error.
__device__ __forceinline__ void wait_flags(const uint4* __restrict__ recs, uint32_t epoch, int n, int* sh_lo, int* sh_hi) {
  if ((int)threadIdx.x < n) {
    const long long t0 = clock64();
    uint4 r = ld_relaxed_sys_v4(recs + threadIdx.x);
    while ((int32_t)(r.x - epoch) < 0) {
      if (clock64() - t0 > 8000000000ll) __trap();
      r = ld_relaxed_sys_v4(recs + threadIdx.x);
    }
    sh_lo[threadIdx.x] = (int)r.y;
    sh_hi[threadIdx.x] = (int)r.z;
  }
  __syncthreads();
}

// Publish: the calling WARP raises this rank's record on every peer.  Everything this rank wrote before (ordered before the
// caller by barriers / the arrival counter) is performed at this GPU's L2 after the fence; the record stores themselves are
// posted, one lane per peer, in parallel.
__device__ __forceinline__ void raise_flags(const PeerEx& px, uint32_t epoch, int row_lo, int row_hi) {
  __threadfence();
  __syncwarp();
  const int lane = threadIdx.x & 31;
  if (lane < px.n) st_relaxed_sys_v4(px.rec_at[lane], make_uint4(epoch, (uint32_t)row_lo, (uint32_t)row_hi, 0u));
}

// Grid-wide barrier of a co-resident grid: `bar` counts arrivals (cleared before the launch), every CTA arrives once.
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int n_ctas, bool wait) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    if (wait) {
      while (ld_relaxed_gpu(bar) < n_ctas) {  // (an acquire load per poll would invalidate the L1 every time)
      }
      __threadfence();
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------ the image kernel
struct ImageArgs {
  float4* acc;          // per-corner accumulators [n_ref][cells]
  float* iwe;           // [n_ref][Hp*Wp]: written by the fold (this rank's partial when sharded), else the input
  float* iwe_full;      // sharded: the sum over ranks
  const float* gsrc;    // gq_mode 2: the explicit dL/dIWE images
  const float* affine;  // gq_mode 2: (a, m) per image from the combine kernel
  float4* gq;
  float* zero;          // cleared when the gradient quads are built (the motion gradient K3 accumulates into), or NULL
  int64_t n_zero;
  double* zero2;        // the 2-dof fp64 staging pair, cleared likewise
  int Hp, Wp, n_ref;
  int64_t cells;
  int fold;             // acc -> iwe (re-zeroing acc); 0: iwe is already there
  int stats;            // variance sums + scalar cost inside this launch
  int gq_mode;          // 0 none, 1 from the in-kernel affine pair (variance: a * (I - m) inside the crop), 2 from gsrc
  int omit;
  double* slots;        // [gridDim.x][n_ref][2] partial sums
  unsigned int* bar;    // [0] arrivals after the fold (sharded), [1] arrivals after the statistics
  double* stats_out;    // [n_ref][4]
  double inv_M, inv_Mm1;  // 1 / M and 1 / (M - 1), M = pixels the statistic runs over
  CombineDev cd;
  PeerEx px;
};

#ifdef CMAX_MEASURE  // measurement builds: per-CTA globaltimer stamps of the phases (scripts/image_probe.py)
#define STAMP(i) do { if (threadIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); stamps[i] = t_; } } while (0)
#else
#define STAMP(i)
#endif

// IWE[r,c] = acc[r,c].x + acc[r-1,c].y + acc[r,c-1].z + acc[r-1,c-1].w.  Every scalar component of every accumulator
// cell has exactly ONE reader, which also zeroes it: the accumulators are clean again for the next CM iteration and no
// memset is ever enqueued (components no pixel reads only collect votes of out-of-image corners and are never looked at).
struct FoldAddr {
  float *a00, *a10, *a01, *a11;
};
__device__ __forceinline__ FoldAddr fold_addr(float* __restrict__ A, unsigned r, unsigned c, unsigned Wc) {
  const unsigned k = (r + 1u) * Wc + (c + 1u);
  FoldAddr f;
  f.a00 = A + 4u * k;                  // .x of cell (r, c)
  f.a10 = A + 4u * (k - Wc) + 1u;      // .y of cell (r-1, c)
  f.a01 = A + 4u * (k - 1u) + 2u;      // .z of cell (r, c-1)
  f.a11 = A + 4u * (k - Wc - 1u) + 3u;  // .w of cell (r-1, c-1)
  return f;
}

// The kernel is LATENCY bound (a 90 k-pixel image over 148 SMs is one or two pixels per thread), so what counts is the
// length of the dependent chain per phase: 32-bit index arithmetic only (a 64-bit division is ~150 dependent instructions),
// all loads of a thread issued before the first use, statistics reduced by one warp, one grid-wide barrier.
template <bool SHARDED>
__global__ void __launch_bounds__(kMidThreads, 1) image_kernel(ImageArgs a) {
  constexpr int kWarps = kMidThreads / 32;
  __shared__ double red[2][kWarps];
  __shared__ double sh_stats[4 * CMAX_MAX_REFS];
  __shared__ double sh_cost;
  __shared__ float sh_aff[2 * CMAX_MAX_REFS];
  __shared__ uint32_t sh_epoch;
  __shared__ int sh_lo[CMAX_MAX_PEERS], sh_hi[CMAX_MAX_PEERS];  // sharded: rows of every rank's partial IWE that can be non-zero
#ifdef CMAX_MEASURE
  unsigned long long* stamps = reinterpret_cast<unsigned long long*>(a.slots) + 2048 + blockIdx.x * 8;
#endif
  STAMP(0);
  pdl_trigger();  // K3 may be scheduled (it prefetches its first event tile, then waits for this grid to complete)
  pdl_wait();     // K1's reductions are complete and visible
  STAMP(1);
  const unsigned Hp = (unsigned)a.Hp, Wp = (unsigned)a.Wp, Wc = Wp + 1u;
  const unsigned HW = Hp * Wp, cells = (unsigned)a.cells;
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const double M = a.omit ? (double)(Hp - 2u) * (double)(Wp - 2u) : (double)HW;
  uint32_t epoch = 0;
  if (SHARDED) {
    if (threadIdx.x == 0) sh_epoch = *reinterpret_cast<volatile uint32_t*>(a.px.epoch) + 1u;
    __syncthreads();
    epoch = sh_epoch;
  }
  const unsigned lo = a.omit ? 1u : 0u;
  auto in_crop = [&](unsigned r, unsigned c) { return r >= lo && r + lo < Hp && c >= lo && c + lo < Wp; };
  // this CTA's partial sums of image `img` -> its slot (all threads call; one __syncthreads)
  auto commit = [&](int img, double s, double q) {
    s = warp_sum(s);
    q = warp_sum(q);
    if (lane == 0) {
      red[0][wid] = s;
      red[1][wid] = q;
    }
    __syncthreads();
    if (wid == 0) {
      double ts = lane < kWarps ? red[0][lane] : 0.0, tq = lane < kWarps ? red[1][lane] : 0.0;
      ts = warp_sum(ts);
      tq = warp_sum(tq);
      if (lane == 0) {
        double* slot = a.slots + ((size_t)img * gridDim.x + blockIdx.x) * 2;
        slot[0] = ts;
        slot[1] = tq;
      }
    }
  };

  // ---- phase 1: fold (and, on one GPU, the variance sums of the folded image).  Two pixels per round, loads first.
  int nz_hi1 = 0, nz_ilo = 0;  // sharded: max(row + 1) / max(Hp - row) over the non-zero pixels of this rank's partial IWEs
  if (a.fold || (a.stats && !SHARDED)) {
    for (int img = 0; img < a.n_ref; ++img) {
      float* A = reinterpret_cast<float*>(a.acc + (size_t)img * cells);
      float* I = a.iwe + (size_t)img * HW;
      double s = 0.0, q = 0.0;
      for (unsigned p0 = tid; p0 < HW; p0 += 2u * nthr) {
        const unsigned p1 = p0 + nthr;
        const bool two = p1 < HW;
        const unsigned r0 = p0 / Wp, c0 = p0 - r0 * Wp;
        const unsigned r1 = two ? p1 / Wp : 0u, c1 = two ? p1 - r1 * Wp : 0u;
        float v0, v1 = 0.f;
        if (a.fold) {
          const FoldAddr f0 = fold_addr(A, r0, c0, Wc), f1 = fold_addr(A, r1, c1, Wc);
          const float x0 = *f0.a00, y0 = *f0.a10, z0 = *f0.a01, w0 = *f0.a11;
          float x1 = 0.f, y1 = 0.f, z1 = 0.f, w1 = 0.f;
          if (two) {
            x1 = *f1.a00; y1 = *f1.a10; z1 = *f1.a01; w1 = *f1.a11;
          }
          *f0.a00 = 0.f; *f0.a10 = 0.f; *f0.a01 = 0.f; *f0.a11 = 0.f;
          v0 = ((x0 + y0) + z0) + w0;
          I[p0] = v0;
          if (two) {
            *f1.a00 = 0.f; *f1.a10 = 0.f; *f1.a01 = 0.f; *f1.a11 = 0.f;
            v1 = ((x1 + y1) + z1) + w1;
            I[p1] = v1;
          }
          if (SHARDED) {
            if (v0 != 0.f) {
              nz_hi1 = max(nz_hi1, (int)r0 + 1);
              nz_ilo = max(nz_ilo, (int)(Hp - r0));
            }
            if (two && v1 != 0.f) {
              nz_hi1 = max(nz_hi1, (int)r1 + 1);
              nz_ilo = max(nz_ilo, (int)(Hp - r1));
            }
          }
        } else {
          v0 = I[p0];
          if (two) v1 = I[p1];
        }
        if (a.stats && !SHARDED) {
          if (in_crop(r0, c0)) {
            s += (double)v0;
            q += (double)v0 * (double)v0;
          }
          if (two && in_crop(r1, c1)) {
            s += (double)v1;
            q += (double)v1 * (double)v1;
          }
        }
      }
      STAMP(2);
      if (a.stats && !SHARDED) commit(img, s, q);
    }
  }
  STAMP(3);

  // ---- phase 2 (sharded): signal, wait, sum the partial images of all ranks in rank order (bit-identical on every rank)
  if (SHARDED) {
    nz_hi1 = __reduce_max_sync(0xffffffffu, nz_hi1);
    nz_ilo = __reduce_max_sync(0xffffffffu, nz_ilo);
    if (lane == 0 && nz_hi1 > 0) {
      atomicMax(&a.bar[2], (unsigned)nz_hi1);
      atomicMax(&a.bar[3], (unsigned)nz_ilo);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      // The shortest chain from "the slowest CTA has folded" to "the records are on their way": every CTA's writes are
      // performed at this GPU's L2 (fence) before it counts itself in, so when the count completes nothing is left to wait
      // for -- the thread that completes it reads the row range and posts the records itself, no further fence or barrier.
      __threadfence();
      if (atomicAdd(&a.bar[0], 1u) == gridDim.x - 1) {
        const int hi1 = (int)ld_relaxed_gpu(&a.bar[2]), ilo = (int)ld_relaxed_gpu(&a.bar[3]);
        *a.px.epoch = epoch;  // (every CTA has read the old value by now)
        const uint4 rec = make_uint4(epoch, (uint32_t)(hi1 > 0 ? (int)Hp - ilo : (int)Hp), (uint32_t)(hi1 - 1), 0u);  // empty: lo > hi
        for (int q = 0; q < a.px.n; ++q) st_relaxed_sys_v4(a.px.rec_at[q], rec);
      }
    }
    wait_flags(a.px.recs, epoch, a.px.n, sh_lo, sh_hi);
    STAMP(7);
    for (int img = 0; img < a.n_ref; ++img) {
      double s = 0.0, q = 0.0;
      auto account = [&](unsigned r, unsigned c, float v) {
        if (a.stats && in_crop(r, c)) {
          s += (double)v;
          q += (double)v * (double)v;
        }
      };
      if ((HW & 3u) == 0) {
        // 16-byte peer loads, all ranks' loads of a thread in flight together: one NVLink round trip per thread
        for (unsigned p4 = tid; p4 < (HW >> 2); p4 += nthr) {
          const unsigned r0 = (4u * p4) / Wp, c0 = 4u * p4 - r0 * Wp;  // (one division per group of 4 pixels)
          const int row_a = (int)r0, row_b = (int)(c0 + 3u >= Wp ? r0 + 1u : r0);  // rows this group touches
          float4 part[CMAX_MAX_PEERS];
#pragma unroll
          for (int r = 0; r < CMAX_MAX_PEERS; ++r) {
            part[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < a.px.n && row_b >= sh_lo[r] && row_a <= sh_hi[r])  // (a peer whose rows miss these pixels holds zeros there)
              part[r] = __ldcg(reinterpret_cast<const float4*>(a.px.part[r] + (size_t)img * HW) + p4);  // L2-coherent: another GPU wrote it
          }
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < CMAX_MAX_PEERS; ++r)
            if (r < a.px.n) {
              v.x += part[r].x; v.y += part[r].y; v.z += part[r].z; v.w += part[r].w;
            }
          reinterpret_cast<float4*>(a.iwe_full + (size_t)img * HW)[p4] = v;
          const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (unsigned k = 0; k < 4u; ++k) {
            const bool wrap = c0 + k >= Wp;
            account(wrap ? r0 + 1u : r0, wrap ? c0 + k - Wp : c0 + k, vv[k]);
          }
        }
      } else {
        for (unsigned p = tid; p < HW; p += nthr) {
          const unsigned r0 = p / Wp, c0 = p - r0 * Wp;
          float v = 0.f;
          for (int r = 0; r < a.px.n; ++r)
            if ((int)r0 >= sh_lo[r] && (int)r0 <= sh_hi[r]) v += __ldcg(a.px.part[r] + (size_t)img * HW + p);
          a.iwe_full[(size_t)img * HW + p] = v;
          account(r0, c0, v);
        }
      }
      if (a.stats) commit(img, s, q);
    }
  }

  // ---- phase 3: statistics -> scalar cost -> affine pair.  Warp 0 of every CTA sums the per-CTA slots in the same fixed
  // order, so the fp64 totals (hence cost and gradient) are bit-identical in every CTA and on every rank.
  const float* I = SHARDED ? a.iwe_full : a.iwe;
  if (a.stats) {
    const bool need_all = a.gq_mode != 0;  // value only: CTA 0 alone finishes the cost
    grid_barrier(&a.bar[1], gridDim.x, need_all || blockIdx.x == 0);
    STAMP(4);
    if (!need_all && blockIdx.x != 0) return;
    if (wid == 0) {
      for (int img = 0; img < a.n_ref; ++img) {
        double s = 0.0, q = 0.0;
        for (unsigned c = lane; c < gridDim.x; c += 32u) {
          const double2 sq = __ldcg(reinterpret_cast<const double2*>(a.slots + ((size_t)img * gridDim.x + c) * 2));
          s += sq.x;
          q += sq.y;
        }
        s = warp_sum(s);
        q = warp_sum(q);
        if (lane == 0) {
          // (reciprocals from the host: an fp64 division is a ~100-instruction dependent chain on this latency-bound path)
          const double mean = s * a.inv_M;
          sh_stats[4 * img + 0] = (q - s * mean) * a.inv_Mm1;  // unbiased (torch.var default)   src/costs/image_variance.py:47-58
          sh_stats[4 * img + 1] = mean;
          sh_stats[4 * img + 2] = M;
          sh_stats[4 * img + 3] = 0.0;
        }
      }
      if (lane == 0) {
        CombineDev local = a.cd;
        local.cost = &sh_cost;
        local.affine = sh_aff;
        combine_eval(sh_stats, local);
        if (blockIdx.x == 0) {
          for (int k = 0; k < 4 * a.n_ref; ++k) a.stats_out[k] = sh_stats[k];
          for (int k = 0; k < 2 * a.n_ref; ++k) a.cd.affine[k] = sh_aff[k];
          a.cd.cost[0] = sh_cost;
        }
      }
    }
    __syncthreads();
  } else if (a.gq_mode == 2) {
    if (threadIdx.x < 2 * a.n_ref) sh_aff[threadIdx.x] = a.affine[threadIdx.x];
    __syncthreads();
    I = a.gsrc;
  }
  STAMP(5);
  if (a.gq_mode == 0) return;

  // ---- phase 4: per-corner gradient quads.  G[p] = a * (I[p] - m) inside the crop (mode 1) / everywhere (mode 2), gathered at
  // the four corners of every accumulator cell with the per-corner in-bounds masks.
  const unsigned glo = (a.gq_mode == 1) ? lo : 0u;
  for (int img = 0; img < a.n_ref; ++img) {
    const float* Ii = I + (size_t)img * HW;
    const float ga = sh_aff[2 * img], gm = sh_aff[2 * img + 1];
    // two cells per round, all eight loads issued before the first use
    auto corners = [&](unsigned k, float (&v)[4]) {
      const unsigned kr = k / Wc, kc = k - kr * Wc;
      const int r0 = (int)kr - 1, c0 = (int)kc - 1;  // cell (r0, c0): corner pixels rows r0, r0+1 and columns c0, c0+1
      const int gl = (int)glo, rhi = (int)Hp - 1 - gl, chi = (int)Wp - 1 - gl;  // a corner counts iff gl <= row <= rhi and gl <= col <= chi
      const bool rin0 = r0 >= gl && r0 <= rhi, rin1 = r0 + 1 >= gl && r0 + 1 <= rhi;
      const bool cin0 = c0 >= gl && c0 <= chi, cin1 = c0 + 1 >= gl && c0 + 1 <= chi;
      const float* base = Ii + ((int64_t)r0 * (int64_t)Wp + c0);  // (only dereferenced where the masks allow)
      v[0] = (rin0 && cin0) ? __ldcg(base) : gm;
      v[1] = (rin1 && cin0) ? __ldcg(base + Wp) : gm;
      v[2] = (rin0 && cin1) ? __ldcg(base + 1) : gm;
      v[3] = (rin1 && cin1) ? __ldcg(base + Wp + 1) : gm;
    };
    for (unsigned k0 = tid; k0 < cells; k0 += 2u * nthr) {
      const unsigned k1 = k0 + nthr;
      float u[4], v[4];
      corners(k0, u);
      if (k1 < cells) corners(k1, v);
      a.gq[(size_t)img * cells + k0] = make_float4(ga * (u[0] - gm), ga * (u[1] - gm), ga * (u[2] - gm), ga * (u[3] - gm));
      if (k1 < cells) a.gq[(size_t)img * cells + k1] = make_float4(ga * (v[0] - gm), ga * (v[1] - gm), ga * (v[2] - gm), ga * (v[3] - gm));
    }
  }
  if (a.zero != nullptr) {
    if ((a.n_zero & 3) == 0 && (reinterpret_cast<uintptr_t>(a.zero) & 15) == 0) {
      for (int64_t k = tid; k < (a.n_zero >> 2); k += nthr) reinterpret_cast<float4*>(a.zero)[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
      for (int64_t k = tid; k < a.n_zero; k += nthr) a.zero[k] = 0.f;
    }
  }
  if (a.zero2 != nullptr && tid < 2) a.zero2[tid] = 0.0;
  STAMP(6);
}

// ------------------------------------------------------------------------------------------------ gradient exchange
// out[i] = sum over ranks of part[r][i], rank order.  CTA 0 raises this rank's gradient record on every peer first (K3 is
// complete: stream order), then every CTA waits for all ranks' records and pulls.  n == 0 closes a value-only evaluation.
// The motion gradient is `n / plane` planes of [H, W] (plane == 0: no image structure, e.g. the 2-dof model): a rank's
// partial gradient is non-zero only in the rows that hold source pixels of ITS events (row_lo .. row_hi, fixed per plan).
__global__ void __launch_bounds__(256) grad_exchange_kernel(PeerEx px, int64_t n, int plane, int W, int row_lo, int row_hi, float* __restrict__ out,
                                                            unsigned long long* probe) {
  __shared__ int sh_lo[CMAX_MAX_PEERS], sh_hi[CMAX_MAX_PEERS];
#ifdef CMAX_MEASURE
  unsigned long long* stamps = probe + (blockIdx.x & 63) * 8;
#endif
  STAMP(0);
  const uint32_t epoch = *reinterpret_cast<volatile uint32_t*>(px.epoch);  // already advanced by this evaluation's image_kernel
  if (blockIdx.x == 0 && threadIdx.x < 32) raise_flags(px, epoch, row_lo, row_hi);
  wait_flags(px.recs, epoch, px.n, sh_lo, sh_hi);
  STAMP(1);
  // 32-bit index arithmetic (a 64-bit division is a ~150-instruction dependent chain, and every thread has one group to do)
  const unsigned tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
  const unsigned un = (unsigned)n, uplane = (unsigned)plane, uW = (unsigned)W;
  if ((un & 3u) == 0) {
    for (unsigned i = tid; i < (un >> 2); i += nthr) {
      int row_a = 0, row_b = 0;
      if (plane > 0) {
        const unsigned e0 = (4u * i) % uplane;
        row_a = (int)(e0 / uW);
        row_b = (int)((e0 + 3u) / uW);
        if (e0 + 3u >= uplane) {  // the group straddles two planes: take every row
          row_a = 0;
          row_b = 0x7fffffff;
        }
      }
      float4 part[CMAX_MAX_PEERS];
#pragma unroll
      for (int r = 0; r < CMAX_MAX_PEERS; ++r) {
        part[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < px.n && (plane == 0 || (row_b >= sh_lo[r] && row_a <= sh_hi[r]))) part[r] = __ldcg(reinterpret_cast<const float4*>(px.part[r]) + i);
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int r = 0; r < CMAX_MAX_PEERS; ++r)
        if (r < px.n) {
          v.x += part[r].x; v.y += part[r].y; v.z += part[r].z; v.w += part[r].w;
        }
      reinterpret_cast<float4*>(out)[i] = v;
    }
  } else {
    for (unsigned i = tid; i < un; i += nthr) {
      const int row = plane > 0 ? (int)((i % uplane) / uW) : 0;
      float v = 0.f;
      for (int r = 0; r < px.n; ++r)
        if (plane == 0 || (row >= sh_lo[r] && row <= sh_hi[r])) v += __ldcg(px.part[r] + i);
      out[i] = v;
    }
  }
  STAMP(2);
}

// 2-dof gradient: the CTAs accumulate in two doubles (off_misc), narrowed here.
__global__ void finish_2dof_kernel(const double* __restrict__ acc2, float* __restrict__ out) {
  if (threadIdx.x < 2) out[threadIdx.x] = (float)acc2[threadIdx.x];
}

// ------------------------------------------------------------------------------------------------ host side
static int check_model(const char* fn, const cmax_plan* p, int model) {
  CMAX_REQUIRE(p != nullptr, "%s: plan is NULL", fn);
  CMAX_REQUIRE(model == CMAX_MOTION_DENSE || model == CMAX_MOTION_VOXEL || model == CMAX_MOTION_2DOF || model == CMAX_MOTION_TILE,
               "%s: motion model %d not supported", fn, model);
  CMAX_REQUIRE(model != CMAX_MOTION_VOXEL || p->n_bins >= 1, "%s: dense-flow-voxel needs cmax_plan_set_refs(..., n_bins >= 1)", fn);
  if (model == CMAX_MOTION_TILE) {
    CMAX_REQUIRE(p->tile.hp > 0, "%s: the tile-flow model needs cmax_plan_set_tile_flow first", fn);
    CMAX_REQUIRE(p->n == 0 || (p->strips != nullptr && p->vote_variant == 5 && p->grad_variant == 5 &&
                               p->strip_tile_bytes == strips_tile_bytes_for(CMAX_MOTION_DENSE, p->n_ref)),
                 "%s: the tile-flow model runs on the strip kernels only (this plan has no strips, or another variant is selected)", fn);
  }
  return CMAX_OK;
}

// variance of an un-blurred IWE: the statistics ride on the fold, and dL/dIWE is affine in the IWE (no explicit image)
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

static inline size_t motion_floats(const cmax_plan* p, int motion_model) {
  const size_t HW = (size_t)p->H * p->W;
  if (motion_model == CMAX_MOTION_DENSE) return 2 * HW;
  if (motion_model == CMAX_MOTION_VOXEL) return 2 * (size_t)p->n_bins * HW;
  if (motion_model == CMAX_MOTION_TILE) return 2 * (size_t)p->tile.hp * p->tile.wp;
  return 2;
}

static CombineDev combine_for(const cmax_plan* p, const cmax_cost_spec* spec, const double* d_orig_stat, double* d_cost, const Ws& w) {
  const bool explicit_grad = spec->sigma > 0.f || spec->stat == CMAX_STAT_GRADMAG;
  return make_combine(p->n_ref, spec->stat, spec->form, spec->direction_sign, explicit_grad ? 1 : 0, spec->weights, d_orig_stat, d_cost,
                      w.affine);
}

static int image_grid(const ObjLayout& L, int n_ref) {
  // co-resident by construction: at most one CTA per SM (the kernel holds grid-wide barriers)
  const int64_t want = (std::max<int64_t>(L.cells, L.HW) + kMidThreads - 1) / kMidThreads;
  (void)n_ref;
  return (int)std::max<int64_t>(1, std::min<int64_t>({want, (int64_t)num_sms(), (int64_t)kMidMaxCtas}));
}

static ImageArgs image_args(const cmax_plan* p, const ObjLayout& L, const Ws& w) {
  ImageArgs a;
  memset(&a, 0, sizeof(a));
  a.acc = w.acc; a.iwe = w.iwe; a.iwe_full = w.iwe_full; a.gq = w.gq;
  a.Hp = p->Hp; a.Wp = p->Wp; a.n_ref = p->n_ref; a.cells = L.cells;
  a.slots = w.slots; a.bar = w.bar; a.stats_out = w.stats;
  return a;
}

static void set_stats(ImageArgs& a, const cmax_plan* p, const cmax_cost_spec* spec) {
  a.stats = 1;
  a.omit = spec->omit_boundary ? 1 : 0;
  const double M = a.omit ? (double)(p->Hp - 2) * (double)(p->Wp - 2) : (double)p->Hp * (double)p->Wp;
  a.inv_M = 1.0 / M;
  a.inv_Mm1 = 1.0 / (M - 1.0);
}

static PeerEx peer_ex(const cmax_peers* peers, const float* const* part, int flag_block, const Ws& w) {
  PeerEx px;
  memset(&px, 0, sizeof(px));
  if (peers == nullptr) return px;
  px.n = peers->n_peers;
  px.rank = peers->rank;
  for (int r = 0; r < peers->n_peers; ++r) {
    px.part[r] = part[r];
    px.rec_at[r] = reinterpret_cast<uint4*>(peers->flags[r]) + flag_block * CMAX_MAX_PEERS + peers->rank;
  }
  px.recs = reinterpret_cast<const uint4*>(peers->flags[peers->rank]) + flag_block * CMAX_MAX_PEERS;
  px.epoch = w.epoch;
  return px;
}

static int check_peers(const char* fn, const cmax_peers* peers) {
  CMAX_REQUIRE(peers != nullptr, "%s: peers is NULL", fn);
  CMAX_REQUIRE(peers->n_peers >= 1 && peers->n_peers <= CMAX_MAX_PEERS, "%s: n_peers must be in [1,%d], got %d", fn, CMAX_MAX_PEERS, peers->n_peers);
  CMAX_REQUIRE(peers->rank >= 0 && peers->rank < peers->n_peers, "%s: rank %d out of range", fn, peers->rank);
  for (int r = 0; r < peers->n_peers; ++r)
    CMAX_REQUIRE(peers->iwe[r] != nullptr && peers->grad[r] != nullptr && peers->flags[r] != nullptr, "%s: a pointer of peer %d is NULL", fn, r);
  return CMAX_OK;
}

// the 256-byte statistics / barrier block must be clean when image_kernel starts: K1's first CTA clears it when there is
// one (run and strip kernels), else a memset does
static int launch_k1(const cmax_plan* p, int motion_model, const float* motion, const Ws& w, bool clear_block, cudaStream_t s) {
  FusedArgs a = fused_args(p, motion);
  const bool k1_clears = clear_block && p->vote_variant >= 2 && p->n > 0 && (p->stage_mask & 2);
  if (k1_clears) a.zero256 = reinterpret_cast<unsigned int*>(w.sacc);
  if (clear_block && !k1_clears) CMAX_CUDA_CHECK(cudaMemsetAsync(w.sacc, 0, 256, s));
  if (p->vote_variant == 1 && (p->stage_mask & 1))
    CMAX_CUDA_CHECK(cudaMemsetAsync(w.iwe, 0, (size_t)p->n_ref * p->Hp * p->Wp * sizeof(float), s));
  if (p->n > 0 && (p->stage_mask & 2)) {
    const bool strips = p->vote_variant == 5 && p->strips != nullptr && p->strip_tile_bytes == strips_tile_bytes_for(motion_model, p->n_ref);
    if (!strips && p->vote_variant >= 2) {
      const int rc = ensure_packed(p, s);
      if (rc) return rc;
    }
    if (motion_model == CMAX_MOTION_TILE) launch_vote_strips(motion_model, p->n_ref, s, a, w.acc);
    else launch_vote_any(p, motion_model, s, a, w.acc, w.iwe);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

static int launch_image(const ImageArgs& a, const ObjLayout& L, cudaStream_t s) {
  if (a.px.n > 0) launch_k(pdl_enabled(), image_kernel<true>, dim3(image_grid(L, a.n_ref)), dim3(kMidThreads), s, a);
  else launch_k(pdl_enabled(), image_kernel<false>, dim3(image_grid(L, a.n_ref)), dim3(kMidThreads), s, a);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

// Cost of IWEs that are already folded (w.iwe, or w.iwe_full when `use_full`), and -- when want_grad -- the gradient quads.
// zero_grad (may be NULL): motion-gradient buffer to clear on the way.  `bar_clean`: the statistics block is known to be zero.
static int cost_stage(const cmax_plan* p, const cmax_cost_spec* spec, const double* d_orig_stat, void* workspace, int want_grad,
                      double* d_cost, float* zero_grad, size_t n_zero, bool bar_clean, bool use_full, cmax_stream_t stream) {
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  cudaStream_t s = as_stream(stream);
  const int n_ref = p->n_ref;
  if (!(p->stage_mask & 4)) return CMAX_OK;
  ImageArgs a = image_args(p, L, w);
  if (use_full) a.iwe = w.iwe_full;
  a.zero = zero_grad;
  a.n_zero = (int64_t)n_zero;
  a.zero2 = want_grad ? w.acc2 : nullptr;
  if (can_fuse_stats(spec)) {
    if (!bar_clean) CMAX_CUDA_CHECK(cudaMemsetAsync(w.sacc, 0, 256, s));
    set_stats(a, p, spec);
    a.gq_mode = want_grad ? 1 : 0;
    a.cd = combine_for(p, spec, d_orig_stat, d_cost, w);
    a.cd.a.k2 = 2.0 * a.inv_Mm1;
    return launch_image(a, L, s);
  }
  const bool blurred = spec->sigma > 0.f;
  const float* img = a.iwe;
  int rc;
  if (blurred) {
    rc = cmax_blur3(img, w.blur, n_ref, p->Hp, p->Wp, spec->sigma, 0, stream);
    if (rc) return rc;
    img = w.blur;
  }
  rc = cmax_image_stats(img, n_ref, p->Hp, p->Wp, spec->stat, spec->omit_boundary, w.stats, want_grad ? w.G : nullptr, w.stats_ws, stream);
  if (rc) return rc;
  launch_combine(w.stats, combine_for(p, spec, d_orig_stat, d_cost, w), s);
  if (want_grad) {
    const float* gsrc = w.G;
    if (blurred) {
      rc = cmax_blur3(w.G, w.G2, n_ref, p->Hp, p->Wp, spec->sigma, 1, stream);
      if (rc) return rc;
      gsrc = w.G2;
    }
    a.gq_mode = 2;
    a.gsrc = gsrc;
    a.affine = w.affine;
    return launch_image(a, L, s);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

// K3.  pre_zeroed: grad_motion (and the 2-dof staging pair) were cleared by the image kernel.
static int grad_stage(const cmax_plan* p, int motion_model, const float* motion, void* workspace, float* grad_motion, int pre_zeroed,
                      cudaStream_t s) {
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  const FusedArgs a = fused_args(p, motion);
  if ((p->stage_mask & 1) && !pre_zeroed) {
    CMAX_CUDA_CHECK(cudaMemsetAsync(grad_motion, 0, motion_floats(p, motion_model) * sizeof(float), s));
    if (motion_model == CMAX_MOTION_2DOF) CMAX_CUDA_CHECK(cudaMemsetAsync(w.acc2, 0, 2 * sizeof(double), s));
  }
  if (p->n > 0 && (p->stage_mask & 2)) {
    const bool strips = p->grad_variant == 5 && p->strips != nullptr && p->strip_tile_bytes == strips_tile_bytes_for(motion_model, p->n_ref);
    if (!strips && p->grad_variant >= 2) {
      const int rc = ensure_packed(p, s);
      if (rc) return rc;
    }
    float* target = (motion_model == CMAX_MOTION_2DOF) ? reinterpret_cast<float*>(w.acc2) : grad_motion;
    if (motion_model == CMAX_MOTION_TILE) launch_grad_strips(motion_model, p->n_ref, pdl_enabled(), s, a, w.gq, target);
    else launch_grad_any(p, motion_model, s, a, w.gq, target);
  }
  if (motion_model == CMAX_MOTION_2DOF && (p->stage_mask & 2)) finish_2dof_kernel<<<1, 32, 0, s>>>(w.acc2, grad_motion);
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

int cmax_objective_vote(const cmax_plan_t* plan, int motion_model, const float* motion, void* workspace, cmax_stream_t stream) {
  int rc = check_model("cmax_objective_vote", plan, motion_model);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && workspace != nullptr, "cmax_objective_vote: NULL motion/workspace");
  CMAX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cmax_objective_vote: workspace must be 256-byte aligned");
  const Ws w = carve(workspace, obj_layout(plan->Hp, plan->Wp));
  return launch_k1(plan, motion_model, motion, w, true, as_stream(stream));
}

int cmax_objective_fold(const cmax_plan_t* plan, void* workspace, float** iwe_out, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr && workspace != nullptr, "cmax_objective_fold: NULL argument");
  const ObjLayout L = obj_layout(plan->Hp, plan->Wp);
  const Ws w = carve(workspace, L);
  if (plan->vote_variant != 1 && (plan->stage_mask & 4)) {  // (variant 1 votes straight into the image)
    ImageArgs a = image_args(plan, L, w);
    a.fold = 1;
    const int rc = launch_image(a, L, as_stream(stream));
    if (rc) return rc;
  }
  if (iwe_out) *iwe_out = w.iwe;
  return CMAX_OK;
}

int cmax_objective_cost(const cmax_plan_t* plan, const cmax_cost_spec* spec, const double* d_orig_stat, void* workspace, int want_grad,
                        double* d_cost, float* zero_grad, int64_t n_zero, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr && workspace != nullptr && d_cost != nullptr, "cmax_objective_cost: NULL argument");
  int rc = check_spec("cmax_objective_cost", spec, plan->n_ref);
  if (rc) return rc;
  CMAX_REQUIRE(spec->form == CMAX_COST_PLAIN || d_orig_stat != nullptr, "cmax_objective_cost: normalised costs need d_orig_stat");
  CMAX_REQUIRE(plan->Hp >= 3 && plan->Wp >= 3, "cmax_objective_cost: images must be at least 3x3");
  CMAX_REQUIRE(n_zero >= 0 && (n_zero == 0 || zero_grad != nullptr), "cmax_objective_cost: bad zero_grad / n_zero");
  return cost_stage(plan, spec, d_orig_stat, workspace, want_grad, d_cost, want_grad ? zero_grad : nullptr, want_grad ? (size_t)n_zero : 0,
                    false, false, stream);
}

int cmax_objective_grad(const cmax_plan_t* plan, int motion_model, const float* motion, void* workspace, float* grad_motion, int pre_zeroed,
                        cmax_stream_t stream) {
  int rc = check_model("cmax_objective_grad", plan, motion_model);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && workspace != nullptr && grad_motion != nullptr, "cmax_objective_grad: NULL argument");
  return grad_stage(plan, motion_model, motion, workspace, grad_motion, pre_zeroed, as_stream(stream));
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
  const cmax_plan* p = plan;
  cudaStream_t s = as_stream(stream);
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  const bool want_grad = grad_motion != nullptr;
  const size_t n_motion = motion_floats(p, motion_model);
  const bool full = p->stage_mask == 7;
  rc = launch_k1(p, motion_model, motion, w, true, s);
  if (rc) return rc;
  if (can_fuse_stats(spec) && p->vote_variant != 1 && full) {
    // the metric path: K1 -> image_kernel (fold + variance + cost + gradient quads, gradient buffer cleared) -> K3
    ImageArgs a = image_args(p, L, w);
    a.fold = 1;
    set_stats(a, p, spec);
    a.gq_mode = want_grad ? 1 : 0;
    a.cd = combine_for(p, spec, d_orig_stat, d_cost, w);
    a.cd.a.k2 = 2.0 * a.inv_Mm1;
    a.zero = grad_motion;
    a.n_zero = want_grad ? (int64_t)n_motion : 0;
    a.zero2 = want_grad ? w.acc2 : nullptr;
    rc = launch_image(a, L, s);
    if (rc) return rc;
  } else {
    rc = cmax_objective_fold(p, workspace, nullptr, stream);
    if (rc) return rc;
    rc = cost_stage(p, spec, d_orig_stat, workspace, want_grad, d_cost, (want_grad && full) ? grad_motion : nullptr, full ? n_motion : 0, true,
                    false, stream);
    if (rc) return rc;
  }
  if (want_grad) rc = grad_stage(p, motion_model, motion, workspace, grad_motion, full ? 1 : 0, s);
  return rc;
}

int cmax_objective_sharded(const cmax_plan_t* plan, int motion_model, const float* motion, const cmax_cost_spec* spec,
                           const double* d_orig_stat, void* workspace, const cmax_peers* peers, double* d_cost, float* grad_motion,
                           cmax_stream_t stream) {
  int rc = check_model("cmax_objective_sharded", plan, motion_model);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && workspace != nullptr && d_cost != nullptr, "cmax_objective_sharded: NULL argument");
  CMAX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cmax_objective_sharded: workspace must be 256-byte aligned");
  rc = check_spec("cmax_objective_sharded", spec, plan->n_ref);
  if (rc) return rc;
  rc = check_peers("cmax_objective_sharded", peers);
  if (rc) return rc;
  CMAX_REQUIRE(motion_floats(plan, motion_model) < ((size_t)1 << 31), "cmax_objective_sharded: the motion has too many elements for the 32-bit exchange index");
  CMAX_REQUIRE(spec->form == CMAX_COST_PLAIN || d_orig_stat != nullptr, "cmax_objective_sharded: normalised costs need d_orig_stat");
  CMAX_REQUIRE(plan->Hp >= 3 && plan->Wp >= 3, "cmax_objective_sharded: images must be at least 3x3");
  CMAX_REQUIRE(plan->vote_variant != 1, "cmax_objective_sharded: vote variant 1 has no per-corner accumulators to fold");
  CMAX_REQUIRE(plan->stage_mask == 7, "cmax_objective_sharded: partial stage masks are for single-GPU measurements");
  const cmax_plan* p = plan;
  cudaStream_t s = as_stream(stream);
  const ObjLayout L = obj_layout(p->Hp, p->Wp);
  const Ws w = carve(workspace, L);
  CMAX_REQUIRE(peers->iwe[peers->rank] == w.iwe, "cmax_objective_sharded: peers->iwe[rank] must be this workspace's partial IWE "
                                                   "(workspace + cmax_objective_iwe_offset)");
  const bool want_grad = grad_motion != nullptr;
  const size_t n_motion = motion_floats(p, motion_model);
  float* grad_part = const_cast<float*>(peers->grad[peers->rank]);
  rc = launch_k1(p, motion_model, motion, w, true, s);
  if (rc) return rc;
  const bool fuse = can_fuse_stats(spec);
  ImageArgs a = image_args(p, L, w);
  a.fold = 1;
  a.px = peer_ex(peers, peers->iwe, 0, w);
  if (fuse) {
    set_stats(a, p, spec);
    a.gq_mode = want_grad ? 1 : 0;
    a.cd = combine_for(p, spec, d_orig_stat, d_cost, w);
    a.cd.a.k2 = 2.0 * a.inv_Mm1;
    a.zero = want_grad ? grad_part : nullptr;
    a.n_zero = want_grad ? (int64_t)n_motion : 0;
    a.zero2 = want_grad ? w.acc2 : nullptr;
  }
  rc = launch_image(a, L, s);
  if (rc) return rc;
  if (!fuse) {
    rc = cost_stage(p, spec, d_orig_stat, workspace, want_grad, d_cost, want_grad ? grad_part : nullptr, n_motion, false, true, stream);
    if (rc) return rc;
  }
  if (want_grad) {
    rc = grad_stage(p, motion_model, motion, workspace, grad_part, 1, s);
    if (rc) return rc;
  }
  // gradient exchange (value only: flags only, so that no rank overwrites its partial IWE while a slower peer still reads it)
  const PeerEx gx = peer_ex(peers, peers->grad, 1, w);
  const int64_t n = want_grad ? (int64_t)n_motion : 0;
  const int64_t work = (n & 3) == 0 ? n / 4 : n;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((work + 255) / 256, (int64_t)num_sms() * 2));
  const int plane = (motion_model == CMAX_MOTION_2DOF || motion_model == CMAX_MOTION_TILE) ? 0 : p->H * p->W;
  grad_exchange_kernel<<<grid, 256, 0, s>>>(gx, n, plane, p->W, p->src_row_lo, p->src_row_hi, grad_motion,
                                            reinterpret_cast<unsigned long long*>(w.slots) + 2048 + 148 * 8);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

size_t cmax_objective_probe_offset(const cmax_plan_t* plan) {
  if (plan == nullptr) {
    set_error("cmax_objective_probe_offset: plan is NULL");
    return 0;
  }
  return obj_layout(plan->Hp, plan->Wp).off_slots + 2048 * sizeof(double);
}

size_t cmax_objective_iwe_offset(const cmax_plan_t* plan) {
  if (plan == nullptr) {
    set_error("cmax_objective_iwe_offset: plan is NULL");
    return 0;
  }
  return obj_layout(plan->Hp, plan->Wp).off_iwe;
}

size_t cmax_objective_full_iwe_offset(const cmax_plan_t* plan) {
  if (plan == nullptr) {
    set_error("cmax_objective_full_iwe_offset: plan is NULL");
    return 0;
  }
  return obj_layout(plan->Hp, plan->Wp).off_iwe_full;
}

}  // extern "C"
