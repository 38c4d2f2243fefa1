// Shared device/host helpers for the cmax_b200 library (sm_100a only).
//
// Arithmetic contract (SURVEY.md section 7 "hard parts" 1, Appendix A): the reference computes in fp32 with every
// operation rounded separately (torch CPU elementwise ops), so the event math below uses explicit round-to-nearest
// intrinsics (__fmul_rn / __fsub_rn / __fadd_rn / __fdiv_rn) that nvcc never contracts into FMA.  This is what makes
// the warped coordinates and floor indices bit-exact with the reference; the library is additionally built with
// -fmad=false so that no other expression can be contracted by accident.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <algorithm>
#include <stdint.h>
#include <stdio.h>

#include "../../include/cmax_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "cmax_b200 is written for sm_100a (B200) only"
#endif

namespace cmax {

void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define CMAX_CUDA_CHECK(call)                                   \
  do {                                                          \
    cudaError_t e__ = (call);                                   \
    if (e__ != cudaSuccess) return ::cmax::cuda_fail(e__, #call); \
  } while (0)

#define CMAX_REQUIRE(cond, ...)        \
  do {                                 \
    if (!(cond)) {                     \
      ::cmax::set_error(__VA_ARGS__);  \
      return CMAX_ERR_ARG;             \
    }                                  \
  } while (0)

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

static inline cudaStream_t as_stream(cmax_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------ device math
struct Vote {
  int row, col;    // floor indices (already offset by the padding)
  float fx, fy;    // fractions measured from the floor: fx along rows, fy along columns; may be in [-1e-6, 1)
};

// Exact floor of v as float AND int without the quarter-rate conversion pipe (FRND / F2I): adding 1.5*2^23 rounds v to
// the nearest integer in the mantissa; one compare-and-decrement turns round-to-nearest into floor.  Exact for
// |v| < 2^22.  Beyond that (or for NaN / inf) the integer is garbage, but provably far outside any image: the biased
// float is then not within [1.5*2^23 - 2^22, 1.5*2^23 + 2^22), so `i` has magnitude >= 2^22 (NaN/inf give ~8.8e8) and
// every bounds check against an image of fewer than 2^22 rows / columns rejects it -- no explicit range guard needed.
__device__ __forceinline__ void floor_exact(float v, float& fl, int& i) {
  const float C = 12582912.0f;  // 1.5 * 2^23
  const float t = __fadd_rn(v, C);
  i = __float_as_int(t) - 0x4B400000;
  fl = __fsub_rn(t, C);
  if (fl > v) {
    fl = __fsub_rn(fl, 1.0f);
    i -= 1;
  }
}

// i = floor(x' + 1e-6), f = x' - i          src/event_image_converter.py:340-345
// For coordinates beyond +-2^22 row / col are garbage-but-out-of-range and fx / fy meaningless: callers must not use the
// fractions of an event that fails its bounds checks.
__device__ __forceinline__ Vote vote_geometry(float xw, float yw, int pad_h, int pad_w) {
  float flx, fly;
  int ix, iy;
  floor_exact(__fadd_rn(xw, 1e-6f), flx, ix);
  floor_exact(__fadd_rn(yw, 1e-6f), fly, iy);
  Vote v;
  v.fx = __fsub_rn(xw, flx);
  v.fy = __fsub_rn(yw, fly);
  v.row = ix + pad_h;
  v.col = iy + pad_w;
  return v;
}

// The 4 bilinear weights in the reference's corner order (r,c), (r+1,c), (r,c+1), (r+1,c+1)
//                                            src/event_image_converter.py:365-369
__device__ __forceinline__ void vote_weights(const Vote& v, float w[4]) {
  const float ax = __fsub_rn(1.0f, v.fx), ay = __fsub_rn(1.0f, v.fy);
  w[0] = __fmul_rn(ax, ay);
  w[1] = __fmul_rn(v.fx, ay);
  w[2] = __fmul_rn(ax, v.fy);
  w[3] = __fmul_rn(v.fx, v.fy);
}

// dt = (t - ref) / period                     src/warp.py:254-258
__device__ __forceinline__ float normalised_dt(float t, float ref, float period, int normalize_t) {
  const float d = __fsub_rn(t, ref);
  return normalize_t ? __fdiv_rn(d, period) : d;
}

// x' = x - dt * f  (product rounded first)   src/warp.py:306-307
__device__ __forceinline__ float warp_minus(float x, float dt, float f) { return __fsub_rn(x, __fmul_rn(dt, f)); }
// x' = x + dt * theta                        src/warp.py:507-514
__device__ __forceinline__ float warp_plus(float x, float dt, float th) { return __fadd_rn(x, __fmul_rn(dt, th)); }

// Time bin of a normalised dt: edges[b] <= dt < edges[b+1], compared in fp32   src/warp.py:346-352.
// Returns -1 if in no bin (NaN).  A closed-form guess is corrected against the exact edges.
__device__ __forceinline__ int time_bin(float dt, const float* __restrict__ edges, int n_bins, float dt_min, float inv_width) {
  int b = __float2int_rd((dt - dt_min) * inv_width);
  b = max(0, min(n_bins - 1, b));
  while (b > 0 && dt < edges[b]) --b;
  while (b < n_bins - 1 && dt >= edges[b + 1]) ++b;
  return (edges[b] <= dt && dt < edges[b + 1]) ? b : -1;
}

__device__ __forceinline__ float4 ld_event(const float* __restrict__ ev, int64_t i) {
  // float4 events: one 16-byte read-only load per event, streamed (they are touched once per pass)
  return __ldcs(reinterpret_cast<const float4*>(ev) + i);
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace cmax
