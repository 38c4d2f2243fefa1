// K2: contrast statistics of the IWE(s), their image-space derivatives, and the scalar cost combination.
// Images are tiny next to the event stream (0.36 - 3.7 MB, L2 resident), so these kernels are launch/latency bound;
// accumulation is in float64 so that a single pass over the image meets the 1e-5 cost tolerance
// (SURVEY.md section 7 "hard part" 4).
#include "cmax_stats.cuh"

namespace cmax {

// ---- variance: sum and sum of squares over the crop; last CTA finalises.   src/costs/image_variance.py:37-58
__global__ void __launch_bounds__(kStatBlock) variance_stats_kernel(const float* __restrict__ images, int Hp, int Wp, int omit,
                                                                    StatAcc* __restrict__ acc, double* __restrict__ slots,
                                                                    double* __restrict__ stats) {
  __shared__ double red[kStatBlock / 32];
  const int img = blockIdx.y;
  const float* I = images + (int64_t)img * Hp * Wp;
  const int r0 = omit ? 1 : 0, c0 = omit ? 1 : 0;
  const int hh = Hp - 2 * r0, ww = Wp - 2 * c0;
  const int64_t M = (int64_t)hh * ww;
  double s = 0.0, q = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < M; k += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(k / ww) + r0, c = (int)(k % ww) + c0;
    const double v = (double)__ldg(I + (int64_t)r * Wp + c);
    s += v;
    q += v * v;
  }
  if (slots_commit(s, q, gridDim.x, blockIdx.x, &acc[img], slots + (size_t)img * 2 * kStatMaxCtas, red) && threadIdx.x == 0) {
    const double mean = s / (double)M;
    stats[4 * img + 0] = (q - s * mean) / (double)(M - 1);  // unbiased (torch.var default)
    stats[4 * img + 1] = mean;
    stats[4 * img + 2] = (double)M;
    stats[4 * img + 3] = 0.0;
  }
}

// d var / d I = 2/(M-1) (I - mean) inside the crop, 0 on the border
__global__ void __launch_bounds__(256) variance_grad_kernel(const float* __restrict__ images, int Hp, int Wp, int omit,
                                                            const double* __restrict__ stats, float* __restrict__ grad) {
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)Hp * Wp;
  const double mean = stats[4 * img + 1], M = stats[4 * img + 2];
  const double k2 = 2.0 / (M - 1.0);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < HW; k += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(k / Wp), c = (int)(k % Wp);
    const bool in = !omit || (r >= 1 && r <= Hp - 2 && c >= 1 && c <= Wp - 2);
    grad[img * HW + k] = in ? (float)(k2 * ((double)__ldg(images + img * HW + k) - mean)) : 0.f;
  }
}

// ---- gradient magnitude: Sobel pair / 8 with zero padding, mean of squares over the crop.
//                                       src/utils/stat_utils.py:51-83, src/costs/gradient_magnitude.py:60-76
__device__ __forceinline__ float px(const float* __restrict__ I, int Hp, int Wp, int r, int c) {
  return (r >= 0 && r < Hp && c >= 0 && c < Wp) ? __ldg(I + (int64_t)r * Wp + c) : 0.f;
}

__global__ void __launch_bounds__(kStatBlock) gradmag_stats_kernel(const float* __restrict__ images, int Hp, int Wp, int omit,
                                                                   StatAcc* __restrict__ acc, double* __restrict__ slots,
                                                                   double* __restrict__ stats, float* __restrict__ gxy /* [n_img,2,Hp,Wp] or NULL */) {
  __shared__ double red[kStatBlock / 32];
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)Hp * Wp;
  const float* I = images + img * HW;
  const int64_t M = omit ? (int64_t)(Hp - 2) * (Wp - 2) : HW;
  double s = 0.0;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < HW; k += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(k / Wp), c = (int)(k % Wp);
    const bool in = !omit || (r >= 1 && r <= Hp - 2 && c >= 1 && c <= Wp - 2);
    float gx = 0.f, gy = 0.f;
    if (in) {
      const float a = px(I, Hp, Wp, r - 1, c - 1), b = px(I, Hp, Wp, r - 1, c), d = px(I, Hp, Wp, r - 1, c + 1);
      const float e = px(I, Hp, Wp, r, c - 1), f = px(I, Hp, Wp, r, c + 1);
      const float g = px(I, Hp, Wp, r + 1, c - 1), h = px(I, Hp, Wp, r + 1, c), i = px(I, Hp, Wp, r + 1, c + 1);
      gx = ((g + 2.f * h + i) - (a + 2.f * b + d)) * 0.125f;  // derivative along rows
      gy = ((d + 2.f * f + i) - (a + 2.f * e + g)) * 0.125f;  // derivative along columns
      s += (double)gx * gx + (double)gy * gy;
    }
    if (gxy != nullptr) {
      gxy[(img * 2 + 0) * HW + k] = gx;
      gxy[(img * 2 + 1) * HW + k] = gy;
    }
  }
  double unused = 0.0;
  if (slots_commit(s, unused, gridDim.x, blockIdx.x, &acc[img], slots + (size_t)img * 2 * kStatMaxCtas, red) && threadIdx.x == 0) {
    stats[4 * img + 0] = s / (double)M;
    stats[4 * img + 1] = 0.0;
    stats[4 * img + 2] = (double)M;
    stats[4 * img + 3] = 0.0;
  }
}

// d value / d I = (2/M)/8 * [ Kx^T gx + Ky^T gy ]  (gx, gy already masked to the crop)
__global__ void __launch_bounds__(256) gradmag_grad_kernel(const float* __restrict__ gxy, int Hp, int Wp,
                                                           const double* __restrict__ stats, float* __restrict__ grad) {
  const int img = blockIdx.y;
  const int64_t HW = (int64_t)Hp * Wp;
  const float* GX = gxy + (img * 2 + 0) * HW;
  const float* GY = gxy + (img * 2 + 1) * HW;
  const float k = (float)(2.0 / stats[4 * img + 2] * 0.125);
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(p / Wp), c = (int)(p % Wp);
    // gx[q] = sum_d Kx[d] I[q+d]  =>  dI[p] += Kx[p-q] gx[q];  with q = p - d
    const float x_up = px(GX, Hp, Wp, r - 1, c - 1) + 2.f * px(GX, Hp, Wp, r - 1, c) + px(GX, Hp, Wp, r - 1, c + 1);
    const float x_dn = px(GX, Hp, Wp, r + 1, c - 1) + 2.f * px(GX, Hp, Wp, r + 1, c) + px(GX, Hp, Wp, r + 1, c + 1);
    const float y_lf = px(GY, Hp, Wp, r - 1, c - 1) + 2.f * px(GY, Hp, Wp, r, c - 1) + px(GY, Hp, Wp, r + 1, c - 1);
    const float y_rt = px(GY, Hp, Wp, r - 1, c + 1) + 2.f * px(GY, Hp, Wp, r, c + 1) + px(GY, Hp, Wp, r + 1, c + 1);
    // Kx has +1 rows at d=+1: pixel p is the "+row" neighbour of q = p - (1,*) i.e. the row above p
    grad[img * HW + p] = k * ((x_up - x_dn) + (y_lf - y_rt));
  }
}

// ---- scalar combination                              src/costs/*.py (see cmax_b200.h cmax_cost_form)
__global__ void combine_cost_kernel(const double* __restrict__ stats, CombineDev cd) {
  if (threadIdx.x != 0) return;
  combine_eval(stats, cd);
}

void launch_combine(const double* stats, const CombineDev& cd, cudaStream_t s) { combine_cost_kernel<<<1, 32, 0, s>>>(stats, cd); }

static inline int stat_grid(int64_t n) {
  return (int)std::max<int64_t>(1, std::min<int64_t>({(n + kStatBlock - 1) / kStatBlock, (int64_t)num_sms() * 2, (int64_t)kStatMaxCtas}));
}

}  // namespace cmax

using namespace cmax;

extern "C" {

size_t cmax_stats_workspace_bytes(int n_img, int Hp, int Wp) {
  if (n_img < 1 || Hp < 1 || Wp < 1) return 0;
  const size_t acc = ((size_t)n_img * sizeof(StatAcc) + 255) / 256 * 256;
  const size_t slots = (size_t)n_img * 2 * kStatMaxCtas * sizeof(double);  // per-CTA partial sums (deterministic reduction)
  return acc + slots + (size_t)n_img * 2 * Hp * Wp * sizeof(float);
}

int cmax_image_stats(const float* images, int n_img, int Hp, int Wp, int stat, int omit_boundary, double* d_stats, float* grad,
                     void* workspace, cmax_stream_t stream) {
  CMAX_REQUIRE(images != nullptr && d_stats != nullptr && workspace != nullptr, "cmax_image_stats: NULL pointer");
  CMAX_REQUIRE(n_img >= 1 && n_img <= 64, "cmax_image_stats: n_img must be in [1,64], got %d", n_img);
  CMAX_REQUIRE(Hp >= 3 && Wp >= 3, "cmax_image_stats: images must be at least 3x3 (got %dx%d)", Hp, Wp);
  CMAX_REQUIRE(stat == CMAX_STAT_VARIANCE || stat == CMAX_STAT_GRADMAG, "cmax_image_stats: unknown statistic %d", stat);
  cudaStream_t s = as_stream(stream);
  const size_t acc_bytes = ((size_t)n_img * sizeof(StatAcc) + 255) / 256 * 256;
  StatAcc* acc = static_cast<StatAcc*>(workspace);
  double* slots = reinterpret_cast<double*>(static_cast<char*>(workspace) + acc_bytes);
  float* gxy = reinterpret_cast<float*>(static_cast<char*>(workspace) + acc_bytes + (size_t)n_img * 2 * kStatMaxCtas * sizeof(double));
  CMAX_CUDA_CHECK(cudaMemsetAsync(acc, 0, (size_t)n_img * sizeof(StatAcc), s));
  const int64_t HW = (int64_t)Hp * Wp;
  dim3 grid(stat_grid(HW), n_img);
  if (stat == CMAX_STAT_VARIANCE) {
    variance_stats_kernel<<<grid, kStatBlock, 0, s>>>(images, Hp, Wp, omit_boundary ? 1 : 0, acc, slots, d_stats);
    if (grad) variance_grad_kernel<<<grid, 256, 0, s>>>(images, Hp, Wp, omit_boundary ? 1 : 0, d_stats, grad);
  } else {
    gradmag_stats_kernel<<<grid, kStatBlock, 0, s>>>(images, Hp, Wp, omit_boundary ? 1 : 0, acc, slots, d_stats, grad ? gxy : nullptr);
    if (grad) gradmag_grad_kernel<<<grid, 256, 0, s>>>(gxy, Hp, Wp, d_stats, grad);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

int cmax_combine_cost(const double* d_stats, int n_ref, int stat, int cost_form, const double* d_orig_stat, const float* h_weights,
                      int direction_sign, int explicit_grad, double* d_cost, float* d_affine, cmax_stream_t stream) {
  CMAX_REQUIRE(d_stats && d_cost && d_affine, "cmax_combine_cost: NULL pointer");
  CMAX_REQUIRE(n_ref >= 1 && n_ref <= CMAX_MAX_REFS, "cmax_combine_cost: n_ref must be in [1,%d]", CMAX_MAX_REFS);
  CMAX_REQUIRE(cost_form >= CMAX_COST_PLAIN && cost_form <= CMAX_COST_MULTIFOCAL, "cmax_combine_cost: unknown cost form %d", cost_form);
  CMAX_REQUIRE(cost_form == CMAX_COST_PLAIN || d_orig_stat != nullptr, "cmax_combine_cost: normalised costs need the un-warped statistic");
  CMAX_REQUIRE(cost_form != CMAX_COST_PLAIN || n_ref == 1, "cmax_combine_cost: a plain cost takes exactly one image");
  CMAX_REQUIRE(direction_sign == 1 || direction_sign == -1, "cmax_combine_cost: direction_sign must be +1 (minimize) or -1 (maximize)");
  CombineDev cd = make_combine(n_ref, stat, cost_form, direction_sign, explicit_grad, h_weights, d_orig_stat, d_cost, d_affine);
  launch_combine(d_stats, cd, as_stream(stream));
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // extern "C"
