// Tile (patch-grid) flow geometry shared by the stand-alone up-sampling kernels (cmax_tileflow.cu) and the event kernels
// that evaluate the tile flow themselves (cmax_lean.cu, motion model CMAX_MOTION_TILE).
#pragma once
#include "cmax_common.cuh"

namespace cmax {

struct TileGeom {
  int hp, wp;        // patch grid
  int pad_h, pad_w;  // replicate padding of the grid
  int sh, sw;        // integer up-sampling factors (the sliding window)
  int H, W;          // image
  int h1, w1;        // crop offsets inside the up-sampled padded grid
  float inv_sh, inv_sw;  // 1.0f / sh, 1.0f / sw (IEEE fp32 division, done once on the host: the same value the device would compute)
};

// Source taps of output index `full` (coordinates of the up-sampled padded grid) along one axis, PyTorch bilinear
// align_corners=false semantics: src = max(0, (full + 0.5) / s - 0.5); taps i0 = floor(src), i1 = min(i0 + 1, n_pad - 1)
// with weights (1 - l, l); padded index p -> grid node clamp(p - pad, 0, n - 1) (replicate padding).
__device__ __forceinline__ void axis_taps(int full, float inv_s, int n, int pad, int* a0, int* a1, float* l1) {
  const int n_pad = n + 2 * pad;
  float src = ((float)full + 0.5f) * inv_s - 0.5f;
  src = fmaxf(src, 0.0f);
  const int i0 = (int)src;
  const int i1 = min(i0 + 1, n_pad - 1);
  *l1 = src - (float)i0;
  *a0 = min(max(i0 - pad, 0), n - 1);
  *a1 = min(max(i1 - pad, 0), n - 1);
}

// dense[c] at pixel (i, j) = -((1 - lr) * ((1 - lc) v00 + lc v01) + lr * ((1 - lc) v10 + lc v11)), every operation rounded on
// its own (the library is built -fmad=false): the ONE expression both the up-sampling kernel and the event kernels evaluate,
// so a fused evaluation sees bit-identical flow vectors.                    src/solver/patch_contrast_base.py:462-506
struct TileTaps {
  int a0, a1, b0, b1;
  float lr, lc;
};
__device__ __forceinline__ TileTaps tile_taps(const TileGeom& g, int i, int j) {
  TileTaps t;
  axis_taps(i + g.h1, g.inv_sh, g.hp, g.pad_h, &t.a0, &t.a1, &t.lr);
  axis_taps(j + g.w1, g.inv_sw, g.wp, g.pad_w, &t.b0, &t.b1, &t.lc);
  return t;
}
__device__ __forceinline__ float tile_value(const float* __restrict__ m, const TileGeom& g, const TileTaps& t) {
  const float v00 = __ldg(m + t.a0 * g.wp + t.b0), v01 = __ldg(m + t.a0 * g.wp + t.b1);
  const float v10 = __ldg(m + t.a1 * g.wp + t.b0), v11 = __ldg(m + t.a1 * g.wp + t.b1);
  const float top = (1.0f - t.lc) * v00 + t.lc * v01, bot = (1.0f - t.lc) * v10 + t.lc * v11;
  return -((1.0f - t.lr) * top + t.lr * bot);
}

int make_tile_geom(const char* fn, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W, TileGeom* out);

}  // namespace cmax
