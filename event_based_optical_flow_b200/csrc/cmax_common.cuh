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

// SM count of the current device (cudaDevAttrMultiProcessorCount, cached per device; 148 on a B200: 2 dies x 74 SMs).
// Grids are sized in multiples of it.
int num_sms();

static inline cudaStream_t as_stream(cmax_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ------------------------------------------------------------------------------------------------ device math
struct Vote {
  int row, col;    // floor indices (already offset by the padding)
  float fx, fy;    // fractions measured from the floor: fx along rows, fy along columns; may be in [-1e-6, 1)
};

// ---- Blackwell packed fp32x2 arithmetic (PTX add/sub/mul/fma .f32x2, SASS FADD2 / FMUL2 / FFMA2): one instruction does the
// row AND the column component of an event, each component IEEE-rounded exactly like its scalar counterpart (the
// library is built with -fmad=false, so ptxas never contracts a mul2 + add2 pair either).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2_rd(f32x2 a, f32x2 b) {  // round toward -inf
  f32x2 r;
  asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}

// Exact floor without the quarter-rate conversion pipe (FRND / F2I) and without a compare: v + 1.5*2^23 ROUNDED DOWN has
// floor(v) in its mantissa (ulp = 1 there), so the biased float minus the bias is floor(v) as a float and its bit
// pattern minus 0x4B400000 is floor(v) as an int.  Exact for |v| < 2^22.  Beyond that (or for NaN / inf) the integer is
// garbage, but provably far outside any image: the biased float is then outside [2^23, 2^24), so `i` has magnitude
// >= 2^22 (NaN/inf give ~8.8e8) and every bounds check against an image of fewer than 2^22 rows / columns rejects it.
constexpr float kFloorBias = 12582912.0f;  // 1.5 * 2^23
__device__ __forceinline__ void floor_exact(float v, float& fl, int& i) {
  const float t = __fadd_rd(v, kFloorBias);
  i = __float_as_int(t) - 0x4B400000;
  fl = __fsub_rn(t, kFloorBias);
}

// i = floor(x' + 1e-6), f = x' - i          src/event_image_converter.py:340-345
// Both components in packed arithmetic.  For coordinates beyond +-2^22 row / col are garbage-but-out-of-range and
// fx / fy meaningless: callers must not use the fractions of an event that fails its bounds checks.
__device__ __forceinline__ Vote vote_geometry2(f32x2 w, int pad_h, int pad_w) {
  const f32x2 bias = pk2(kFloorBias, kFloorBias);
  const f32x2 t = add2_rd(add2(w, pk2(1e-6f, 1e-6f)), bias);
  const f32x2 f = sub2(w, sub2(t, bias));
  float tx, ty;
  upk2(t, tx, ty);
  Vote v;
  upk2(f, v.fx, v.fy);
  v.row = __float_as_int(tx) - 0x4B400000 + pad_h;
  v.col = __float_as_int(ty) - 0x4B400000 + pad_w;
  return v;
}
__device__ __forceinline__ Vote vote_geometry(float xw, float yw, int pad_h, int pad_w) { return vote_geometry2(pk2(xw, yw), pad_h, pad_w); }

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
// Both components.  The products stay SCALAR on purpose: ptxas contracts mul.rn.f32x2 + sub.rn.f32x2 into one FFMA2
// (observed, CUDA 12.9), which would skip the rounding of dt*f that the reference performs; __fmul_rn / __fsub_rn are
// never contracted.
__device__ __forceinline__ f32x2 warp_minus2(float x, float y, float dt, float f0, float f1) {
  return pk2(warp_minus(x, dt, f0), warp_minus(y, dt, f1));
}
__device__ __forceinline__ f32x2 warp_plus2(float x, float y, float dt, float f0, float f1) {
  return pk2(warp_plus(x, dt, f0), warp_plus(y, dt, f1));
}

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
