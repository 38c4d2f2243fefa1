// Tile (patch-grid) flow -> dense flow and its adjoint: the per-iteration map the reference applies in front of the
// warp (src/solver/patch_contrast_base.py:462-506): negate, replicate-pad the grid, bilinear resize by the integer
// sliding window (align_corners = false), central crop to the image.  SURVEY.md section 8(f) row 1.
//
// The map is separable: dense[c,i,j] = - sum_a sum_b Wr[i,a] Wc[j,b] m[c,a,b] with at most two non-zeros per row of Wr /
// Wc (after the replicate padding is folded onto the border nodes).  Forward = one thread per pixel (4 taps per
// channel, the grid is <= a few hundred floats and lives in L1); adjoint = a gather per grid node over its support --
// no atomics, so the gradient of the 512-parameter motion is deterministic.
#include "cmax_common.cuh"

namespace cmax {

struct TileGeom {
  int hp, wp;        // patch grid
  int pad_h, pad_w;  // replicate padding of the grid
  int sh, sw;        // integer up-sampling factors (the sliding window)
  int H, W;          // image
  int h1, w1;        // crop offsets inside the up-sampled padded grid
};

// Source taps of output index `full` (coordinates of the up-sampled padded grid) along one axis, PyTorch bilinear
// align_corners=false semantics: src = max(0, (full + 0.5) / s - 0.5); taps i0 = floor(src), i1 = min(i0 + 1, n_pad - 1)
// with weights (1 - l, l); padded index p -> grid node clamp(p - pad, 0, n - 1) (replicate padding).
__device__ __forceinline__ void axis_taps(int full, int s, int n, int pad, int* a0, int* a1, float* l1) {
  const int n_pad = n + 2 * pad;
  float src = ((float)full + 0.5f) * (1.0f / (float)s) - 0.5f;
  src = fmaxf(src, 0.0f);
  const int i0 = (int)src;
  const int i1 = min(i0 + 1, n_pad - 1);
  *l1 = src - (float)i0;
  *a0 = min(max(i0 - pad, 0), n - 1);
  *a1 = min(max(i1 - pad, 0), n - 1);
}

__global__ void __launch_bounds__(256) tile_flow_upsample_kernel(const float* __restrict__ motion, TileGeom g, float* __restrict__ dense) {
  const int64_t HW = (int64_t)g.H * g.W;
  const int np = g.hp * g.wp;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(p / g.W), j = (int)(p % g.W);
    int a0, a1, b0, b1;
    float lr, lc;
    axis_taps(i + g.h1, g.sh, g.hp, g.pad_h, &a0, &a1, &lr);
    axis_taps(j + g.w1, g.sw, g.wp, g.pad_w, &b0, &b1, &lc);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const float* m = motion + c * np;
      const float v00 = __ldg(m + a0 * g.wp + b0), v01 = __ldg(m + a0 * g.wp + b1);
      const float v10 = __ldg(m + a1 * g.wp + b0), v11 = __ldg(m + a1 * g.wp + b1);
      const float top = (1.0f - lc) * v00 + lc * v01, bot = (1.0f - lc) * v10 + lc * v11;
      dense[c * HW + p] = -((1.0f - lr) * top + lr * bot);
    }
  }
}

// Adjoint: one CTA per (channel, grid row a).  Phase 1: T[j] = sum_i Wr[i,a] G[c,i,j] over the image rows in the support
// of node row a (threads over j, coalesced reads of G).  Phase 2: gm[c,a,b] = - sum_j Wc[j,b] T[j] (one warp per b).
__global__ void __launch_bounds__(256) tile_flow_upsample_backward_kernel(const float* __restrict__ gdense, TileGeom g,
                                                                          float* __restrict__ gmotion) {
  extern __shared__ float T[];  // [W]
  const int c = blockIdx.y, a = blockIdx.x;
  const int64_t HW = (int64_t)g.H * g.W;
  const float* G = gdense + c * HW;
  // rows of the image whose taps can touch node row a: padded rows [a+pad-1, a+pad+1] scaled by sh, widened to the
  // whole padding band for the border nodes (replicate padding folds it onto them)
  int lo_full = (a == 0) ? 0 : (a + g.pad_h - 1) * g.sh;
  int hi_full = (a == g.hp - 1) ? (g.hp + 2 * g.pad_h) * g.sh - 1 : (a + g.pad_h + 2) * g.sh - 1;
  const int i_lo = max(lo_full - g.h1, 0), i_hi = min(hi_full - g.h1, g.H - 1);
  for (int j = threadIdx.x; j < g.W; j += blockDim.x) {
    float acc = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      int a0, a1;
      float l;
      axis_taps(i + g.h1, g.sh, g.hp, g.pad_h, &a0, &a1, &l);
      const float w = (a0 == a ? 1.0f - l : 0.f) + (a1 == a ? l : 0.f);
      if (w != 0.f) acc += w * __ldg(G + (int64_t)i * g.W + j);
    }
    T[j] = acc;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
  for (int b = wid; b < g.wp; b += n_warps) {
    float acc = 0.f;
    for (int j = lane; j < g.W; j += 32) {
      int b0, b1;
      float l;
      axis_taps(j + g.w1, g.sw, g.wp, g.pad_w, &b0, &b1, &l);
      const float w = (b0 == b ? 1.0f - l : 0.f) + (b1 == b ? l : 0.f);
      acc += w * T[j];
    }
    acc = warp_sum(acc);
    if (lane == 0) gmotion[(c * g.hp + a) * g.wp + b] = -acc;
  }
}

static int make_geom(const char* fn, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W, TileGeom* out) {
  CMAX_REQUIRE(hp >= 1 && wp >= 1 && pad_h >= 0 && pad_w >= 0 && sh >= 1 && sw >= 1 && H >= 1 && W >= 1,
               "%s: bad geometry (grid %dx%d, pad %d,%d, window %dx%d, image %dx%d)", fn, hp, wp, pad_h, pad_w, sh, sw, H, W);
  TileGeom g;
  g.hp = hp; g.wp = wp; g.pad_h = pad_h; g.pad_w = pad_w; g.sh = sh; g.sw = sw; g.H = H; g.W = W;
  const int full_h = (hp + 2 * pad_h) * sh, full_w = (wp + 2 * pad_w) * sw;
  g.h1 = full_h / 2 - H / 2;
  g.w1 = full_w / 2 - W / 2;
  CMAX_REQUIRE(g.h1 >= 0 && g.w1 >= 0 && g.h1 + H <= full_h && g.w1 + W <= full_w,
               "%s: the up-sampled padded grid (%dx%d) does not cover the %dx%d image", fn, full_h, full_w, H, W);
  *out = g;
  return CMAX_OK;
}

}  // namespace cmax

using namespace cmax;

extern "C" {

int cmax_tile_flow_upsample(const float* motion, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W, float* dense,
                            cmax_stream_t stream) {
  CMAX_REQUIRE(motion != nullptr && dense != nullptr, "cmax_tile_flow_upsample: NULL pointer");
  TileGeom g;
  const int rc = make_geom("cmax_tile_flow_upsample", hp, wp, pad_h, pad_w, sh, sw, H, W, &g);
  if (rc) return rc;
  const int64_t HW = (int64_t)H * W;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((HW + 255) / 256, (int64_t)kNumSMs * 4));
  tile_flow_upsample_kernel<<<grid, 256, 0, as_stream(stream)>>>(motion, g, dense);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

int cmax_tile_flow_upsample_backward(const float* grad_dense, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W,
                                     float* grad_motion, cmax_stream_t stream) {
  CMAX_REQUIRE(grad_dense != nullptr && grad_motion != nullptr, "cmax_tile_flow_upsample_backward: NULL pointer");
  TileGeom g;
  const int rc = make_geom("cmax_tile_flow_upsample_backward", hp, wp, pad_h, pad_w, sh, sw, H, W, &g);
  if (rc) return rc;
  CMAX_REQUIRE((size_t)W * sizeof(float) <= 48 * 1024, "cmax_tile_flow_upsample_backward: image wider than %d pixels", 48 * 1024 / 4);
  dim3 grid(hp, 2);
  tile_flow_upsample_backward_kernel<<<grid, 256, (size_t)W * sizeof(float), as_stream(stream)>>>(grad_dense, g, grad_motion);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // extern "C"
