// Batched 2-dof candidate costs of the per-patch initialiser.
//
// At every pyramid level above the coarsest the reference asks Optuna's TPE sampler for `n_iter / level` translation
// candidates PER PATCH and evaluates each one with a chain of numpy / scipy / OpenCV calls on the events cropped to the patch
// (src/solver/patch_contrast_pyramid.py:320-415): 2-dof warp to the middle of the patch's time span (src/warp.py:483-522),
// bilinear vote into a patch-sized image (src/event_image_converter.py:257-312), scipy.ndimage.gaussian_filter ('reflect'
// border), cv2.Sobel / 8 (REFLECT_101 border), mean squared gradient magnitude (src/costs/gradient_magnitude.py:78-95) and
// the ratio to the un-warped image's (src/costs/normalized_gradient_magnitude.py:81-94).  256 patches x 13 trials at the
// finest level = 3328 such chains per frame, each on ~100 events.
//
// Here ONE call evaluates K candidates for each of P patches; `evaluation` e = p * K + k owns one patch-sized image.
//   patch images that fit twice in shared memory (<= 25 600 pixels: every level of a 260x346 sensor) -- ONE launch, one CTA per
//   evaluation: the patch's events (grouped by patch on the host side, 12 B each: local x, local y, normalised dt) are warped
//   with candidate k and voted into shared memory, blurred there (separable Gaussian, axis 0 then axis 1 like scipy, scipy's
//   taps and 'reflect' rule), Sobel-filtered there (OpenCV's border rule) and reduced in fp64 in a fixed order;
//   larger patches -- the same stages as four kernels over images in global memory (L2 resident), votes by `red.global.add.f32`.
// The candidate arrives as the sampler's doubles; the patch's time span multiplies it on the device; with the un-warped image's
// energy given (computed once per frame and level by the same entry point with a zero candidate) the call returns the
// reference's loss itself, NaN -> 0 included: one H2D of P*K*2 doubles, one launch, one D2H of P*K doubles per TPE trial.
#include "cmax_common.cuh"

namespace cmax {

constexpr int kBlurMaxRadius = 16;

struct BlurTaps {
  int radius;
  float w[2 * kBlurMaxRadius + 1];
};

struct PatchGeom {
  int Hp, Wp, pad_h, pad_w, n_cand;
  int HW;
};

struct PatchJob {
  const float* ev;           // [m,3] grouped by patch
  const int64_t* offsets;    // [P+1]
  const double* cand;        // [P,K,2]
  const double* scale;       // [P] or null
  const double* orig;        // [P] or null: out = orig / energy (NaN -> 0) instead of the energy
  double* out;               // [P,K]
};

// candidate e = (p, k) in pixels per normalised patch time: the double product rounded once, like the host-side float64
// multiplication of the reference followed by the cast the fp32 kernels need (pyramid.py:366-371)
__device__ __forceinline__ void candidate_theta(const PatchJob& j, int e, int p, float& th0, float& th1) {
  const double sc = j.scale ? j.scale[p] : 1.0;
  th0 = (float)(j.cand[2 * e] * sc);
  th1 = (float)(j.cand[2 * e + 1] * sc);
}

// one event's four votes into `out` (global or shared)
__device__ __forceinline__ void vote_event(const float* __restrict__ ev, int64_t i, float th0, float th1, const PatchGeom& g, float* out) {
  const float x = ev[3 * i], y = ev[3 * i + 1], dt = ev[3 * i + 2];
  const float xw = __fadd_rn(x, __fmul_rn(dt, th0));  // src/warp.py:507-514
  const float yw = __fadd_rn(y, __fmul_rn(dt, th1));
  float flx, fly;
  int row, col;
  floor_exact(__fadd_rn(xw, 1e-8f), flx, row);  // floor(x' + 1e-8), src/event_image_converter.py:279
  floor_exact(__fadd_rn(yw, 1e-8f), fly, col);
  const float fx = __fsub_rn(xw, flx), fy = __fsub_rn(yw, fly);
  row += g.pad_h;
  col += g.pad_w;
  const bool r0 = row >= 0 && row < g.Hp, r1 = row + 1 >= 0 && row + 1 < g.Hp;
  const bool c0 = col >= 0 && col < g.Wp, c1 = col + 1 >= 0 && col + 1 < g.Wp;
  const float gx = __fsub_rn(1.0f, fx), gy = __fsub_rn(1.0f, fy);
  if (r0 && c0) atomicAdd(out + row * g.Wp + col, __fmul_rn(gx, gy));
  if (r1 && c0) atomicAdd(out + (row + 1) * g.Wp + col, __fmul_rn(fx, gy));
  if (r0 && c1) atomicAdd(out + row * g.Wp + col + 1, __fmul_rn(gx, fy));
  if (r1 && c1) atomicAdd(out + (row + 1) * g.Wp + col + 1, __fmul_rn(fx, fy));
}

// scipy 'reflect' (d c b a | a b c d | d c b a): the mirror sits on the pixel edge; folded repeatedly for n < radius
__device__ __forceinline__ int reflect_edge(int i, int n) {
  const int period = 2 * n;
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - 1 - i : i;
}
// OpenCV BORDER_REFLECT_101 (c b | a b c | b a): the mirror sits on the border pixel
__device__ __forceinline__ int reflect_101(int i, int n) {
  if (n == 1) return 0;
  const int period = 2 * n - 2;
  i %= period;
  if (i < 0) i += period;
  return i >= n ? period - i : i;
}

template <int AXIS>
__device__ __forceinline__ float blur_at(const float* src, const BlurTaps& taps, const PatchGeom& g, int r, int c) {
  float acc = 0.f;
  for (int k = -taps.radius; k <= taps.radius; ++k) {
    const int q = AXIS == 0 ? reflect_edge(r + k, g.Hp) * g.Wp + c : r * g.Wp + reflect_edge(c + k, g.Wp);
    acc = __fadd_rn(acc, __fmul_rn(taps.w[k + taps.radius], src[q]));
  }
  return acc;
}

// (gx^2 + gy^2) of the Sobel / 8 gradients at pixel (r, c), in fp64
__device__ __forceinline__ double sobel_sq(const float* img, const PatchGeom& g, int r, int c) {
  const int ra = reflect_101(r - 1, g.Hp) * g.Wp, rb = r * g.Wp, rc = reflect_101(r + 1, g.Hp) * g.Wp;
  const int ca = reflect_101(c - 1, g.Wp), cc = reflect_101(c + 1, g.Wp);
  const float aa = img[ra + ca], ab = img[ra + c], ac = img[ra + cc];
  const float ba = img[rb + ca], bc = img[rb + cc];
  const float da = img[rc + ca], db = img[rc + c], dc = img[rc + cc];
  const float d_col = ((ac + 2.f * bc + dc) - (aa + 2.f * ba + da)) * 0.125f;  // cv2.Sobel(dx = 1) / 8
  const float d_row = ((da + 2.f * db + dc) - (aa + 2.f * ab + ac)) * 0.125f;  // cv2.Sobel(dy = 1) / 8
  return (double)d_col * (double)d_col + (double)d_row * (double)d_row;
}

// block sum in a fixed order (deterministic), then the result of evaluation e: np.mean over the whole image
// (omit_boundary = False, pyramid.py:390) or the reference's loss orig / warped with NaN -> 0 (pyramid.py:374-375)
__device__ __forceinline__ void finish_evaluation(double acc, const PatchJob& j, const PatchGeom& g, int e, int p, double* red) {
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    const double energy = s / (double)g.HW;
    if (j.orig) {
      const double loss = j.orig[p] / energy;  // (-orig) / (-warped), normalized_gradient_magnitude.py:90-93, 'minimize'
      j.out[e] = loss != loss ? 0.0 : loss;
    } else {
      j.out[e] = energy;
    }
  }
}

// ---- small patches (two images fit in shared memory: every level of a 260x346 sensor): ONE kernel, one CTA per evaluation
__global__ void __launch_bounds__(512) patch_fused_kernel(PatchJob j, BlurTaps taps, PatchGeom g, float* __restrict__ images) {
  extern __shared__ float sm[];
  __shared__ double red[16];
  float* a = sm;
  float* b = sm + g.HW;
  const int e = blockIdx.x, p = e / g.n_cand;
  for (int q = threadIdx.x; q < g.HW; q += blockDim.x) a[q] = 0.f;
  float th0, th1;
  candidate_theta(j, e, p, th0, th1);
  __syncthreads();
  const int64_t lo = j.offsets[p], hi = j.offsets[p + 1];
  for (int64_t i = lo + threadIdx.x; i < hi; i += blockDim.x) vote_event(j.ev, i, th0, th1, g, a);
  __syncthreads();
  if (taps.radius >= 0) {
    for (int q = threadIdx.x; q < g.HW; q += blockDim.x) b[q] = blur_at<0>(a, taps, g, q / g.Wp, q % g.Wp);
    __syncthreads();
    for (int q = threadIdx.x; q < g.HW; q += blockDim.x) a[q] = blur_at<1>(b, taps, g, q / g.Wp, q % g.Wp);
    __syncthreads();
  }
  double acc = 0.0;
  for (int q = threadIdx.x; q < g.HW; q += blockDim.x) {
    acc += sobel_sq(a, g, q / g.Wp, q % g.Wp);
    if (images) images[(size_t)e * g.HW + q] = a[q];
  }
  finish_evaluation(acc, j, g, e, p, red);
}

// ---- large patches: the same stages as separate kernels over images in global memory (L2 resident)
__global__ void __launch_bounds__(256) patch_vote_kernel(PatchJob j, PatchGeom g, float* __restrict__ img) {
  const int e = blockIdx.y, p = e / g.n_cand;
  float th0, th1;
  candidate_theta(j, e, p, th0, th1);
  const int64_t lo = j.offsets[p], hi = j.offsets[p + 1];
  float* out = img + (size_t)e * g.HW;
  for (int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (int64_t)gridDim.x * blockDim.x)
    vote_event(j.ev, i, th0, th1, g, out);
}

template <int AXIS>
__global__ void __launch_bounds__(256) patch_blur_kernel(const float* __restrict__ in, BlurTaps taps, PatchGeom g, float* __restrict__ out) {
  const float* src = in + (size_t)blockIdx.y * g.HW;
  float* dst = out + (size_t)blockIdx.y * g.HW;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < g.HW; q += gridDim.x * blockDim.x) dst[q] = blur_at<AXIS>(src, taps, g, q / g.Wp, q % g.Wp);
}

__global__ void __launch_bounds__(512) patch_energy_kernel(const float* __restrict__ in, PatchJob j, PatchGeom g) {
  __shared__ double red[16];
  const int e = blockIdx.x;
  const float* img = in + (size_t)e * g.HW;
  double acc = 0.0;
  for (int q = threadIdx.x; q < g.HW; q += blockDim.x) acc += sobel_sq(img, g, q / g.Wp, q % g.Wp);
  finish_evaluation(acc, j, g, e, e / g.n_cand, red);
}

static int blur_radius(float sigma) { return (int)(4.0 * (double)sigma + 0.5); }  // scipy.ndimage.gaussian_filter1d, truncate = 4

constexpr size_t kFusedSmemLimit = 200 * 1024;

}  // namespace cmax

using namespace cmax;

extern "C" {

size_t cmax_patch_candidates_workspace_bytes(int n_patches, int n_candidates, int h, int w, int pad_h, int pad_w) {
  if (n_patches <= 0 || n_candidates <= 0 || h <= 0 || w <= 0 || pad_h < 0 || pad_w < 0) return 0;
  return (size_t)2 * n_patches * n_candidates * (size_t)(h + 2 * pad_h) * (size_t)(w + 2 * pad_w) * sizeof(float);
}

int cmax_patch_candidates(const float* patch_events, const int64_t* patch_offsets, int64_t max_patch_events, int n_patches,
                          const double* candidates, const double* theta_scale, int n_candidates, int h, int w, int pad_h, int pad_w,
                          float sigma, const double* orig_energy, int flags, void* workspace, size_t workspace_bytes, double* out,
                          cmax_stream_t stream) {
  CMAX_REQUIRE(patch_events != nullptr && patch_offsets != nullptr && candidates != nullptr && out != nullptr,
               "cmax_patch_candidates: NULL pointer");
  CMAX_REQUIRE(n_patches > 0 && n_candidates > 0 && h > 0 && w > 0 && pad_h >= 0 && pad_w >= 0,
               "cmax_patch_candidates: need n_patches, n_candidates, h, w > 0 and paddings >= 0");
  CMAX_REQUIRE(sigma >= 0.f && blur_radius(sigma) <= kBlurMaxRadius, "cmax_patch_candidates: sigma %g outside [0, %g]", (double)sigma,
               (kBlurMaxRadius + 0.49) / 4.0);
  CMAX_REQUIRE((flags & ~(CMAX_PATCH_GLOBAL_IMAGES | CMAX_PATCH_KEEP_IMAGES)) == 0, "cmax_patch_candidates: unknown flags %d", flags);
  PatchGeom g;
  g.Hp = h + 2 * pad_h;
  g.Wp = w + 2 * pad_w;
  g.pad_h = pad_h;
  g.pad_w = pad_w;
  g.n_cand = n_candidates;
  const int64_t HW = (int64_t)g.Hp * g.Wp, n_eval = (int64_t)n_patches * n_candidates;
  CMAX_REQUIRE(HW < (1 << 22) && n_eval <= 65535, "cmax_patch_candidates: patch image of %lld pixels / %lld evaluations per call is too large",
               (long long)HW, (long long)n_eval);
  g.HW = (int)HW;
  PatchJob j{patch_events, patch_offsets, candidates, theta_scale, orig_energy, out};
  BlurTaps taps;
  taps.radius = -1;  // sigma == 0: no blur
  if (sigma > 0.f) {
    taps.radius = blur_radius(sigma);
    double wd[2 * kBlurMaxRadius + 1], sum = 0.0;
    for (int k = -taps.radius; k <= taps.radius; ++k) sum += (wd[k + taps.radius] = exp(-0.5 / ((double)sigma * (double)sigma) * (double)k * (double)k));
    for (int k = 0; k <= 2 * taps.radius; ++k) taps.w[k] = (float)(wd[k] / sum);
  }
  cudaStream_t s = as_stream(stream);
  const size_t smem = (size_t)2 * HW * sizeof(float);
  const bool fused = !(flags & CMAX_PATCH_GLOBAL_IMAGES) && smem <= kFusedSmemLimit;
  const bool need_ws = !fused || (flags & CMAX_PATCH_KEEP_IMAGES);
  if (need_ws) {
    CMAX_REQUIRE(workspace != nullptr && workspace_bytes >= cmax_patch_candidates_workspace_bytes(n_patches, n_candidates, h, w, pad_h, pad_w),
                 "cmax_patch_candidates: workspace of %zu bytes is too small", workspace_bytes);
  }
  float* a = static_cast<float*>(workspace);
  if (fused) {
    if (smem > 48 * 1024)  // (per function and device; setting it again costs a host-side table look-up)
      CMAX_CUDA_CHECK(cudaFuncSetAttribute(patch_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemLimit));
    patch_fused_kernel<<<(unsigned)n_eval, 512, smem, s>>>(j, taps, g, (flags & CMAX_PATCH_KEEP_IMAGES) ? a : nullptr);
    CMAX_CUDA_CHECK(cudaGetLastError());
    return CMAX_OK;
  }
  float* b = a + n_eval * HW;
  CMAX_CUDA_CHECK(cudaMemsetAsync(a, 0, (size_t)n_eval * HW * sizeof(float), s));
  const int chunks = (int)std::max<int64_t>(1, std::min<int64_t>((std::max<int64_t>(max_patch_events, 1) + 1023) / 1024, 64));
  patch_vote_kernel<<<dim3(chunks, (unsigned)n_eval), 256, 0, s>>>(j, g, a);
  CMAX_CUDA_CHECK(cudaGetLastError());
  if (taps.radius >= 0) {
    const dim3 grid((unsigned)std::max<int64_t>(1, std::min<int64_t>((HW + 255) / 256, 32)), (unsigned)n_eval);
    patch_blur_kernel<0><<<grid, 256, 0, s>>>(a, taps, g, b);
    patch_blur_kernel<1><<<grid, 256, 0, s>>>(b, taps, g, a);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  patch_energy_kernel<<<(unsigned)n_eval, 512, 0, s>>>(a, j, g);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // extern "C"
