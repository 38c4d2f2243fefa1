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

// Adds this CTA's partial (sum, sum of squares) to the image's accumulator; the last CTA of `n_ctas` turns the
// totals into {unbiased variance, mean, M, 0}.  Call from all threads of the CTA.    src/costs/image_variance.py:47-58
__device__ __forceinline__ void variance_commit(double s, double q, int64_t M, unsigned int n_ctas, StatAcc* acc,
                                                double* stats4, double* red) {
  s = block_sum(s, red);
  q = block_sum(q, red);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    atomicAdd(&acc->sum, s);
    atomicAdd(&acc->sumsq, q);
    __threadfence();
    last = (atomicAdd(&acc->done, 1u) == n_ctas - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    const double S = *(volatile double*)&acc->sum, Q = *(volatile double*)&acc->sumsq;
    const double mean = S / (double)M;
    stats4[0] = (Q - S * mean) / (double)(M - 1);  // unbiased (torch.var default)
    stats4[1] = mean;
    stats4[2] = (double)M;
    stats4[3] = 0.0;
  }
}

}  // namespace cmax
