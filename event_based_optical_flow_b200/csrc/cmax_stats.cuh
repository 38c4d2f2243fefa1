// Accumulators shared by the standalone statistics kernels (cmax_cost.cu) and the fused fold (cmax_fused.cu).
#pragma once
#include "cmax_common.cuh"

namespace cmax {

struct StatAcc {  // per image, in the caller's workspace; zeroed before use
  double sum, sumsq;
  unsigned int done;
  unsigned int pad_;
};

constexpr int kStatBlock = 256;

// Block-wide sum of a double; result valid in warp 0.  `red` holds one double per warp.
__device__ __forceinline__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x < 32) {
    t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.0;
    t = warp_sum(t);
  }
  return t;
}

constexpr int kStatMaxCtas = 512;  // per-CTA slots of the statistics kernels (grids are capped at this)

// Deterministic grid-wide sums: every CTA writes its partial (s, q) to ITS slot of the image and counts itself in; the last
// CTA to arrive sums the slots in a fixed order (thread-strided, then the fixed shuffle tree).  No floating-point atomics:
// the totals -- hence cost and gradient -- are bit-identical from run to run and, for sharded batches, on every rank, which is
// what keeps SPMD optimisers in lock-step.  Returns true in the threads of the CTA that holds the totals (valid in thread 0).
__device__ __forceinline__ bool slots_commit(double& s, double& q, unsigned int n_ctas, unsigned int cta, StatAcc* acc, double* slots,
                                             double* red) {
  s = block_sum(s, red);
  q = block_sum(q, red);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    slots[2 * cta] = s;
    slots[2 * cta + 1] = q;
    __threadfence();
    last = (atomicAdd(&acc->done, 1u) == n_ctas - 1);
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  double ts = 0.0, tq = 0.0;
  for (unsigned int c = threadIdx.x; c < n_ctas; c += blockDim.x) {
    ts += __ldcg(slots + 2 * c);
    tq += __ldcg(slots + 2 * c + 1);
  }
  s = block_sum(ts, red);
  q = block_sum(tq, red);
  return true;
}

// ---- scalar cost combination, shared by combine_cost_kernel (cmax_cost.cu) and the fold kernel's last CTA
struct CombineArgs {
  int n_ref, stat, form, sign, explicit_grad, has_orig;
  float w[CMAX_MAX_REFS];
  double k2;  // 2 / (M - 1) when the caller knows it on the host (0: divide on the device)
};

struct CombineDev {
  CombineArgs a;
  const double* orig;
  double* cost;
  float* affine;    // [2*n_ref]
};

static inline CombineDev make_combine(int n_ref, int stat, int form, int sign, int explicit_grad, const float* h_weights,
                                      const double* orig, double* cost, float* affine) {
  CombineDev cd;
  cd.a.n_ref = n_ref; cd.a.stat = stat; cd.a.form = form; cd.a.sign = sign; cd.a.explicit_grad = explicit_grad ? 1 : 0;
  cd.a.has_orig = orig != nullptr;
  cd.a.k2 = 0.0;
  for (int r = 0; r < CMAX_MAX_REFS; ++r) cd.a.w[r] = (h_weights && r < n_ref) ? h_weights[r] : 1.0f;
  cd.orig = orig; cd.cost = cost; cd.affine = affine;
  return cd;
}

void launch_combine(const double* stats, const CombineDev& cd, cudaStream_t s);

// One thread.  d cost / d stat_r = alpha_r; VARIANCE without an explicit gradient image folds the statistic's own
// image derivative in: dL/dIWE = alpha * 2/(M-1) * (I - mean) =: a * (I - m).
__device__ __forceinline__ void combine_eval(const double* __restrict__ stats, const CombineDev& cd) {
  const CombineArgs& a = cd.a;
  double total = 0.0;
  for (int r = 0; r < a.n_ref; ++r) {
    const double c = stats[4 * r + 0], mean = stats[4 * r + 1], M = stats[4 * r + 2];
    double alpha;
    if (a.form == CMAX_COST_PLAIN) {
      total += -(double)a.sign * c;
      alpha = -(double)a.sign;
    } else {
      const double co = cd.orig[0];
      const double w = (a.form == CMAX_COST_MULTIFOCAL) ? (double)a.w[r] : 1.0;
      if (a.sign > 0) {  // minimize: orig / warped
        total += w * co / c;
        alpha = -w * co / (c * c);
      } else if (a.form == CMAX_COST_NORMALIZED) {  // maximize: warped / orig
        total += c / co;
        alpha = 1.0 / co;
      } else {  // multi-focal "maximize" negates the sum of (warped / orig)
        total += -w * c / co;
        alpha = -w / co;
      }
    }
    if (a.stat == CMAX_STAT_VARIANCE && !a.explicit_grad) {
      cd.affine[2 * r + 0] = (float)(alpha * (a.k2 != 0.0 ? a.k2 : 2.0 / (M - 1.0)));
      cd.affine[2 * r + 1] = (float)mean;
    } else {
      cd.affine[2 * r + 0] = (float)alpha;
      cd.affine[2 * r + 1] = 0.f;
    }
  }
  cd.cost[0] = total;
}

}  // namespace cmax
