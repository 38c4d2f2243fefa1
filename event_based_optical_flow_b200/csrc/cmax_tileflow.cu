// Tile (patch-grid) flow -> dense flow and its adjoint: the per-iteration map the reference applies in front of the
// warp (src/solver/patch_contrast_base.py:462-506): negate, replicate-pad the grid, bilinear resize by the integer
// sliding window (align_corners = false), central crop to the image.  SURVEY.md section 8(f) row 1.
//
// The map is separable: dense[c,i,j] = - sum_a sum_b Wr[i,a] Wc[j,b] m[c,a,b] with at most two non-zeros per row of Wr /
// Wc (after the replicate padding is folded onto the border nodes).  Forward = one thread per pixel (4 taps per
// channel, the grid is <= a few hundred floats and lives in L1); adjoint = a gather per grid node over its support --
// no atomics, so the gradient of the 512-parameter motion is deterministic.
#include "cmax_tile.cuh"

namespace cmax {

__global__ void __launch_bounds__(256) tile_flow_upsample_kernel(const float* __restrict__ motion, TileGeom g, float* __restrict__ dense) {
  const int64_t HW = (int64_t)g.H * g.W;
  const int np = g.hp * g.wp;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < HW; p += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(p / g.W), j = (int)(p % g.W);
    const TileTaps t = tile_taps(g, i, j);
#pragma unroll
    for (int c = 0; c < 2; ++c) dense[c * HW + p] = tile_value(motion + c * np, g, t);
  }
}

// Adjoint: one CTA per grid node (c, a, b): gm[c,a,b] = - sum_i sum_j Wr[i,a] Wc[j,b] G[c,i,j] over the node's support
// rectangle (~2 sh x 2 sw pixels; the whole padding band for border nodes).  The separable weights of the rectangle are
// staged in shared memory once, the products are summed by a block reduction in a fixed order: hp*wp*2 CTAs (512 for a
// 16x16 grid), independent loads, no atomics.  (The first version used one CTA per grid ROW with a row loop of dependent
// loads behind a branch: 85 us at 480x640 against ~8 us for the forward.)
__device__ __forceinline__ void node_support(int a, int n, int pad, int s, int crop0, int extent, int* lo, int* hi) {
  // output indices (before the crop) whose two taps can touch node a: padded nodes [a+pad-1, a+pad+1] scaled by s; the
  // border nodes also own the whole replicate-padding band
  const int lo_full = (a == 0) ? 0 : (a + pad - 1) * s;
  const int hi_full = (a == n - 1) ? (n + 2 * pad) * s - 1 : (a + pad + 2) * s - 1;
  *lo = max(lo_full - crop0, 0);
  *hi = min(hi_full - crop0, extent - 1);
}

__global__ void __launch_bounds__(256) tile_flow_upsample_backward_kernel(const float* __restrict__ gdense, TileGeom g,
                                                                          float* __restrict__ gmotion) {
  extern __shared__ float wts[];  // [rows of the support][cols of the support]
  __shared__ float red[8];
  const int b = blockIdx.x, a = blockIdx.y, c = blockIdx.z;
  int i_lo, i_hi, j_lo, j_hi;
  node_support(a, g.hp, g.pad_h, g.sh, g.h1, g.H, &i_lo, &i_hi);
  node_support(b, g.wp, g.pad_w, g.sw, g.w1, g.W, &j_lo, &j_hi);
  const int nr = max(i_hi - i_lo + 1, 0), nc = max(j_hi - j_lo + 1, 0);
  float* wr = wts;
  float* wc = wts + nr;
  for (int t = threadIdx.x; t < nr + nc; t += blockDim.x) {
    int t0, t1;
    float l;
    if (t < nr) {
      axis_taps(i_lo + t + g.h1, g.inv_sh, g.hp, g.pad_h, &t0, &t1, &l);
      wr[t] = (t0 == a ? 1.0f - l : 0.f) + (t1 == a ? l : 0.f);
    } else {
      axis_taps(j_lo + (t - nr) + g.w1, g.inv_sw, g.wp, g.pad_w, &t0, &t1, &l);
      wc[t - nr] = (t0 == b ? 1.0f - l : 0.f) + (t1 == b ? l : 0.f);
    }
  }
  __syncthreads();
  const float* G = gdense + (int64_t)c * g.H * g.W;
  float acc = 0.f;
  for (int t = threadIdx.x; t < nr * nc; t += blockDim.x) {
    const int ii = t / nc, jj = t - ii * nc;
    acc += (wr[ii] * wc[jj]) * __ldg(G + (int64_t)(i_lo + ii) * g.W + (j_lo + jj));
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += red[w];
    gmotion[(c * g.hp + a) * g.wp + b] = -tot;
  }
}

int make_tile_geom(const char* fn, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W, TileGeom* out) {
  CMAX_REQUIRE(hp >= 1 && wp >= 1 && pad_h >= 0 && pad_w >= 0 && sh >= 1 && sw >= 1 && H >= 1 && W >= 1,
               "%s: bad geometry (grid %dx%d, pad %d,%d, window %dx%d, image %dx%d)", fn, hp, wp, pad_h, pad_w, sh, sw, H, W);
  TileGeom g;
  g.hp = hp; g.wp = wp; g.pad_h = pad_h; g.pad_w = pad_w; g.sh = sh; g.sw = sw; g.H = H; g.W = W;
  const int full_h = (hp + 2 * pad_h) * sh, full_w = (wp + 2 * pad_w) * sw;
  g.inv_sh = 1.0f / (float)sh;
  g.inv_sw = 1.0f / (float)sw;
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
  const int rc = make_tile_geom("cmax_tile_flow_upsample", hp, wp, pad_h, pad_w, sh, sw, H, W, &g);
  if (rc) return rc;
  const int64_t HW = (int64_t)H * W;
  const int grid = (int)std::max<int64_t>(1, std::min<int64_t>((HW + 255) / 256, (int64_t)num_sms() * 4));
  tile_flow_upsample_kernel<<<grid, 256, 0, as_stream(stream)>>>(motion, g, dense);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

int cmax_tile_flow_upsample_backward(const float* grad_dense, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W,
                                     float* grad_motion, cmax_stream_t stream) {
  CMAX_REQUIRE(grad_dense != nullptr && grad_motion != nullptr, "cmax_tile_flow_upsample_backward: NULL pointer");
  TileGeom g;
  const int rc = make_tile_geom("cmax_tile_flow_upsample_backward", hp, wp, pad_h, pad_w, sh, sw, H, W, &g);
  if (rc) return rc;
  CMAX_REQUIRE((size_t)(H + W) * sizeof(float) <= 48 * 1024, "cmax_tile_flow_upsample_backward: image larger than %d pixels in H + W", 48 * 1024 / 4);
  dim3 grid(wp, hp, 2);
  tile_flow_upsample_backward_kernel<<<grid, 256, (size_t)(H + W) * sizeof(float), as_stream(stream)>>>(grad_dense, g, grad_motion);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // extern "C"
