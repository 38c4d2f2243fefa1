// Modular operators behind the reference's Warp / EventImageConverter objects: each is one kernel, usable on any
// event array (no plan).  The fused per-iteration path lives in cmax_fused.cu; these exist so that the reference's
// unchanged solver can hold drop-in `warper` / `imager` objects (src/solver/base.py:139-147).
#include "cmax_common.cuh"

namespace cmax {

// ------------------------------------------------------------------------------------------------ warp
struct WarpArgs {
  const float* ev;
  int64_t n;
  int stride, H, W, model, ref;
  const float* motion;
  const cmax_time_params_t* tp;
};

// Motion-independent part of one event's warp: dt, source pixel, time bin.  Returns false when the source pixel is
// outside the image.                                        src/warp.py:254-258, 305, 346-352
__device__ __forceinline__ bool warp_site(float x, float y, float t, int H, int W, int model,
                                          const cmax_time_params_t* __restrict__ tp, int r, float* dt_out, int* src_out,
                                          int* bin_out) {
  const float dt = normalised_dt(t, tp->ref[r], tp->period[r], tp->normalize_t);
  *dt_out = dt;
  *bin_out = 0;
  *src_out = 0;
  if (model == CMAX_MOTION_2DOF) return true;
  const int row = __float2int_rz(x), col = __float2int_rz(y);
  *src_out = row * W + col;
  if (!(row >= 0 && row < H && col >= 0 && col < W)) return false;
  if (model == CMAX_MOTION_VOXEL) {
    const int T = tp->n_bins;
    const float inv = (float)T / (tp->dt_max[r] - tp->dt_min[r]);
    *bin_out = time_bin(dt, tp->edges[r], T, tp->dt_min[r], inv);  // -1: in no bin, left un-warped like the reference
  }
  return true;
}

// src/warp.py:306-307, 352-357, 507-514
__device__ __forceinline__ bool warp_one(float x, float y, float t, int H, int W, int model, const float* __restrict__ motion,
                                         const cmax_time_params_t* __restrict__ tp, int r, float* xw, float* yw, float* dt_out,
                                         int* src_out, int* bin_out) {
  const bool ok = warp_site(x, y, t, H, W, model, tp, r, dt_out, src_out, bin_out);
  *xw = x;
  *yw = y;
  if (model == CMAX_MOTION_2DOF) {
    *xw = warp_plus(x, *dt_out, __ldg(motion + 0));
    *yw = warp_plus(y, *dt_out, __ldg(motion + 1));
  } else if (ok && *bin_out >= 0) {
    const int HW = H * W;
    const float* f = motion + (int64_t)(*bin_out) * 2 * HW;
    *xw = warp_minus(x, *dt_out, __ldg(f + *src_out));
    *yw = warp_minus(y, *dt_out, __ldg(f + HW + *src_out));
  }
  return ok;
}

__global__ void __launch_bounds__(256) warp_events_kernel(WarpArgs a, float* __restrict__ out, int32_t* __restrict__ status) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += step) {
    const float* e = a.ev + i * a.stride;
    const float x = __ldg(e), y = __ldg(e + 1), t = __ldg(e + 2);
    float xw, yw, dt;
    int src, bin;
    bad |= !warp_one(x, y, t, a.H, a.W, a.model, a.motion, a.tp, a.ref, &xw, &yw, &dt, &src, &bin);
    float* o = out + i * a.stride;
    o[0] = xw;
    o[1] = yw;
    o[2] = dt;
    for (int c = 3; c < a.stride; ++c) o[c] = __ldg(e + c);
  }
  if (status != nullptr && bad) atomicOr(status, 1);
}

// d motion = sum_e  J_e^T grad_out_e;  x' = x - dt f[src]  =>  d f0[src] += -dt * g_x   (2-dof: d theta0 += dt * g_x)
__global__ void __launch_bounds__(256) warp_events_backward_kernel(WarpArgs a, const float* __restrict__ gout,
                                                                   float* __restrict__ gmotion) {
  __shared__ double red[2][8];
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int HW = a.H * a.W;
  double s0 = 0.0, s1 = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += step) {
    const float* e = a.ev + i * a.stride;
    const float x = __ldg(e), y = __ldg(e + 1), t = __ldg(e + 2);
    float dt;
    int src, bin;
    const bool ok = warp_site(x, y, t, a.H, a.W, a.model, a.tp, a.ref, &dt, &src, &bin);
    const float gx = __ldg(gout + i * a.stride), gy = __ldg(gout + i * a.stride + 1);
    if (a.model == CMAX_MOTION_2DOF) {
      s0 += (double)(dt * gx);
      s1 += (double)(dt * gy);
    } else if (ok && bin >= 0) {
      float* g = gmotion + (int64_t)bin * 2 * HW;
      atomicAdd(g + src, -(dt * gx));
      atomicAdd(g + HW + src, -(dt * gy));
    }
  }
  if (a.model == CMAX_MOTION_2DOF) {  // fp64 partial sums per CTA, one fp32 atomic per CTA
    s0 = warp_sum(s0);
    s1 = warp_sum(s1);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) { red[0][wid] = s0; red[1][wid] = s1; }
    __syncthreads();
    if (threadIdx.x < 2) {
      double tot = 0.0;
      for (int w = 0; w < 8; ++w) tot += red[threadIdx.x][w];
      atomicAdd(gmotion + threadIdx.x, (float)tot);
    }
  }
}

// ------------------------------------------------------------------------------------------------ vote
// 4 masked atomics per event; masks are per corner.  src/event_image_converter.py:346-373 (count: :226-254)
__global__ void __launch_bounds__(256) vote_kernel(const float* __restrict__ xy, int64_t n, int stride, const float* __restrict__ weight,
                                                   int Hp, int Wp, int ph, int pw, int count_mode, float* __restrict__ img) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const float xw = __ldg(xy + i * stride), yw = __ldg(xy + i * stride + 1);
    const Vote v = vote_geometry(xw, yw, ph, pw);
    float w[4];
    if (count_mode) {
      w[0] = w[1] = w[2] = w[3] = 1.0f;
    } else {
      vote_weights(v, w);
      if (weight != nullptr) {
        const float s = __ldg(weight + i);
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = __fmul_rn(w[c], s);
      }
    }
    const bool r0 = v.row >= 0 && v.row < Hp, r1 = v.row + 1 >= 0 && v.row + 1 < Hp;
    const bool c0 = v.col >= 0 && v.col < Wp, c1 = v.col + 1 >= 0 && v.col + 1 < Wp;
    float* p = img + (int64_t)v.row * Wp + v.col;
    if (r0 && c0) atomicAdd(p, w[0]);
    if (r1 && c0) atomicAdd(p + Wp, w[1]);
    if (r0 && c1) atomicAdd(p + 1, w[2]);
    if (r1 && c1) atomicAdd(p + Wp + 1, w[3]);
  }
}

// d L / d x' = sum_c m_c G[c] d w_c / d x'  with  d w/d x' = (-(1-fy), (1-fy), -fy, fy),  d w/d y' = (-(1-fx), -fx, (1-fx), fx)
__global__ void __launch_bounds__(256) vote_backward_kernel(const float* __restrict__ xy, int64_t n, int stride,
                                                            const float* __restrict__ weight, int Hp, int Wp, int ph, int pw,
                                                            const float* __restrict__ G, float* __restrict__ gxy,
                                                            float* __restrict__ gweight) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const float xw = __ldg(xy + i * stride), yw = __ldg(xy + i * stride + 1);
    const Vote v = vote_geometry(xw, yw, ph, pw);
    const bool r0 = v.row >= 0 && v.row < Hp, r1 = v.row + 1 >= 0 && v.row + 1 < Hp;
    const bool c0 = v.col >= 0 && v.col < Wp, c1 = v.col + 1 >= 0 && v.col + 1 < Wp;
    const float* p = G + (int64_t)v.row * Wp + v.col;
    const float g00 = (r0 && c0) ? __ldg(p) : 0.f;
    const float g10 = (r1 && c0) ? __ldg(p + Wp) : 0.f;
    const float g01 = (r0 && c1) ? __ldg(p + 1) : 0.f;
    const float g11 = (r1 && c1) ? __ldg(p + Wp + 1) : 0.f;
    const float s = weight ? __ldg(weight + i) : 1.0f;
    const float dx = (1.0f - v.fy) * (g10 - g00) + v.fy * (g11 - g01);
    const float dy = (1.0f - v.fx) * (g01 - g00) + v.fx * (g11 - g10);
    gxy[2 * i] = s * dx;
    gxy[2 * i + 1] = s * dy;
    if (gweight != nullptr) {
      float w[4];
      vote_weights(v, w);
      gweight[i] = w[0] * g00 + w[1] * g10 + w[2] * g01 + w[3] * g11;
    }
  }
}

// Second order of the bilinear vote (Hessian-vector products).  vote_backward is
//   grad_xy = wt * (dx, dy),  dx = (1-fy)(g10-g00) + fy(g11-g01),  dy = (1-fx)(g01-g00) + fx(g11-g10)
// a function of (xy, G) that is bilinear in G and piecewise bilinear in xy.  For a cotangent u = (ux, uy) on grad_xy:
//   d<grad_xy, u>/dG[c]  = wt * (ux dw_c/dx' + uy dw_c/dy')       -> scattered into out_gimage (a "tangent vote")
//   d<grad_xy, u>/d(x',y') = wt * d_r * (uy, ux),  d_r = (g11-g01) - (g10-g00)   (d2w/dx'2 = d2w/dy'2 = 0 inside a cell)
__global__ void __launch_bounds__(256) vote_backward2_kernel(const float* __restrict__ xy, int64_t n, int stride,
                                                             const float* __restrict__ weight, int Hp, int Wp, int ph, int pw,
                                                             const float* __restrict__ G, const float* __restrict__ u, int u_stride,
                                                             float* __restrict__ out_gimage, float* __restrict__ out_gxy) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const float xw = __ldg(xy + i * stride), yw = __ldg(xy + i * stride + 1);
    const Vote v = vote_geometry(xw, yw, ph, pw);
    const bool r0 = v.row >= 0 && v.row < Hp, r1 = v.row + 1 >= 0 && v.row + 1 < Hp;
    const bool c0 = v.col >= 0 && v.col < Wp, c1 = v.col + 1 >= 0 && v.col + 1 < Wp;
    const int64_t base = (int64_t)v.row * Wp + v.col;
    const float s = weight ? __ldg(weight + i) : 1.0f;
    const float ux = s * __ldg(u + i * u_stride), uy = s * __ldg(u + i * u_stride + 1);
    if (out_gimage != nullptr) {
      float* p = out_gimage + base;
      if (r0 && c0) atomicAdd(p, -(1.0f - v.fy) * ux - (1.0f - v.fx) * uy);
      if (r1 && c0) atomicAdd(p + Wp, (1.0f - v.fy) * ux - v.fx * uy);
      if (r0 && c1) atomicAdd(p + 1, -v.fy * ux + (1.0f - v.fx) * uy);
      if (r1 && c1) atomicAdd(p + Wp + 1, v.fy * ux + v.fx * uy);
    }
    if (out_gxy != nullptr) {
      const float* p = G + base;
      const float g00 = (r0 && c0) ? __ldg(p) : 0.f;
      const float g10 = (r1 && c0) ? __ldg(p + Wp) : 0.f;
      const float g01 = (r0 && c1) ? __ldg(p + 1) : 0.f;
      const float g11 = (r1 && c1) ? __ldg(p + Wp + 1) : 0.f;
      const float d_r = (g11 - g01) - (g10 - g00);
      out_gxy[2 * i] = d_r * uy;
      out_gxy[2 * i + 1] = d_r * ux;
    }
  }
}

// Tangent of the warp w.r.t. the motion (= the adjoint of warp_events_backward w.r.t. its incoming gradient):
// out[e] = d(x', y')_e / d motion . tangent  =  -dt (T0[src], T1[src])   (2-dof: +dt (T[0], T[1]))
__global__ void __launch_bounds__(256) warp_events_tangent_kernel(WarpArgs a, const float* __restrict__ tangent, float* __restrict__ out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  const int HW = a.H * a.W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += step) {
    const float* e = a.ev + i * a.stride;
    const float x = __ldg(e), y = __ldg(e + 1), t = __ldg(e + 2);
    float dt;
    int src, bin;
    const bool ok = warp_site(x, y, t, a.H, a.W, a.model, a.tp, a.ref, &dt, &src, &bin);
    float ox = 0.f, oy = 0.f;
    if (a.model == CMAX_MOTION_2DOF) {
      ox = dt * __ldg(tangent);
      oy = dt * __ldg(tangent + 1);
    } else if (ok && bin >= 0) {
      const float* f = tangent + (int64_t)bin * 2 * HW;
      ox = -(dt * __ldg(f + src));
      oy = -(dt * __ldg(f + HW + src));
    }
    out[2 * i] = ox;
    out[2 * i + 1] = oy;
  }
}

// ------------------------------------------------------------------------------------------------ blur
__device__ __forceinline__ int reflect(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

// 3x3 Gaussian with reflect padding = one correlation with the outer-product kernel.
__global__ void __launch_bounds__(256) blur3_kernel(const float* __restrict__ in, float* __restrict__ out, int Hp, int Wp,
                                                    float k0, float k1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  const float* img = in + (int64_t)blockIdx.z * Hp * Wp;
  if (c >= Wp) return;
  const float k[3] = {k0, k1, k0};
  float acc = 0.f;
#pragma unroll
  for (int dr = -1; dr <= 1; ++dr) {
    const int rr = reflect(r + dr, Hp);
#pragma unroll
    for (int dc = -1; dc <= 1; ++dc) acc += (k[dr + 1] * k[dc + 1]) * __ldg(img + (int64_t)rr * Wp + reflect(c + dc, Wp));
  }
  out[(int64_t)blockIdx.z * Hp * Wp + (int64_t)r * Wp + c] = acc;
}

// Transpose of blur3_kernel as a gather: out[p] = sum over q whose (reflected) 3x3 footprint contains p.
__global__ void __launch_bounds__(256) blur3_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int Hp, int Wp,
                                                              float k0, float k1) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const int r = blockIdx.y;
  const float* img = in + (int64_t)blockIdx.z * Hp * Wp;
  if (c >= Wp) return;
  const float k[3] = {k0, k1, k0};
  float acc = 0.f;
  // q ranges over the 5x5 neighbourhood; tap (dr,dc) of q lands on reflect(q+d); count it when that equals p
  for (int qr = max(0, r - 2); qr <= min(Hp - 1, r + 2); ++qr) {
    float wr = 0.f;
#pragma unroll
    for (int dr = -1; dr <= 1; ++dr)
      if (reflect(qr + dr, Hp) == r) wr += k[dr + 1];
    if (wr == 0.f) continue;
    for (int qc = max(0, c - 2); qc <= min(Wp - 1, c + 2); ++qc) {
      float wc = 0.f;
#pragma unroll
      for (int dc = -1; dc <= 1; ++dc)
        if (reflect(qc + dc, Wp) == c) wc += k[dc + 1];
      if (wc != 0.f) acc += wr * wc * __ldg(img + (int64_t)qr * Wp + qc);
    }
  }
  out[(int64_t)blockIdx.z * Hp * Wp + (int64_t)r * Wp + c] = acc;
}

static inline int grid_for(int64_t n, int block, int per_sm) {
  const int64_t want = (n + block - 1) / block;
  return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)num_sms() * per_sm));
}

}  // namespace cmax

using namespace cmax;

static int check_warp_args(const char* fn, const float* events, int64_t n, int ev_stride, int H, int W, int model,
                           const cmax_time_params_t* tp, int ref_index) {
  CMAX_REQUIRE(n >= 0 && ev_stride >= 3, "%s: need n >= 0 and ev_stride >= 3 (got %lld, %d)", fn, (long long)n, ev_stride);
  CMAX_REQUIRE(n == 0 || events != nullptr, "%s: events is NULL", fn);
  CMAX_REQUIRE(tp != nullptr, "%s: time params is NULL", fn);
  CMAX_REQUIRE(ref_index >= 0 && ref_index < CMAX_MAX_REFS, "%s: ref_index %d out of range", fn, ref_index);
  CMAX_REQUIRE(model == CMAX_MOTION_DENSE || model == CMAX_MOTION_VOXEL || model == CMAX_MOTION_2DOF,
               "%s: motion model %d not supported", fn, model);
  CMAX_REQUIRE(model == CMAX_MOTION_2DOF || (H > 0 && W > 0 && (int64_t)H * W < ((int64_t)1 << 30)), "%s: bad image size %dx%d", fn, H, W);
  return CMAX_OK;
}

extern "C" {

int cmax_warp_events(const float* events, int64_t n, int ev_stride, int H, int W, int motion_model, const float* motion,
                     const cmax_time_params_t* d_params, int ref_index, float* out, int32_t* d_status, cmax_stream_t stream) {
  const int rc = check_warp_args("cmax_warp_events", events, n, ev_stride, H, W, motion_model, d_params, ref_index);
  if (rc) return rc;
  CMAX_REQUIRE(motion != nullptr && (n == 0 || out != nullptr), "cmax_warp_events: NULL motion/out");
  cudaStream_t s = as_stream(stream);
  if (d_status) CMAX_CUDA_CHECK(cudaMemsetAsync(d_status, 0, sizeof(int32_t), s));
  if (n > 0) {
    WarpArgs a{events, n, ev_stride, H, W, motion_model, ref_index, motion, d_params};
    warp_events_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(a, out, d_status);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

int cmax_warp_events_backward(const float* events, int64_t n, int ev_stride, int H, int W, int motion_model, int n_bins,
                              const cmax_time_params_t* d_params, int ref_index, const float* grad_out, float* grad_motion,
                              cmax_stream_t stream) {
  const int rc = check_warp_args("cmax_warp_events_backward", events, n, ev_stride, H, W, motion_model, d_params, ref_index);
  if (rc) return rc;
  CMAX_REQUIRE(grad_motion != nullptr && (n == 0 || grad_out != nullptr), "cmax_warp_events_backward: NULL gradient buffer");
  CMAX_REQUIRE(motion_model != CMAX_MOTION_VOXEL || (n_bins >= 1 && n_bins <= CMAX_MAX_BINS),
               "cmax_warp_events_backward: n_bins must be in [1,%d], got %d", CMAX_MAX_BINS, n_bins);
  cudaStream_t s = as_stream(stream);
  size_t bytes = 2 * sizeof(float);
  if (motion_model == CMAX_MOTION_DENSE) bytes = 2 * (size_t)H * W * sizeof(float);
  if (motion_model == CMAX_MOTION_VOXEL) bytes = 2 * (size_t)n_bins * H * W * sizeof(float);
  CMAX_CUDA_CHECK(cudaMemsetAsync(grad_motion, 0, bytes, s));
  if (n > 0) {
    // the warp is affine in the motion, so its adjoint never reads motion values
    WarpArgs a{events, n, ev_stride, H, W, motion_model, ref_index, nullptr, d_params};
    warp_events_backward_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(a, grad_out, grad_motion);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

int cmax_vote(const float* xy, int64_t n, int xy_stride, const float* weight, int Hp, int Wp, int pad_h, int pad_w, int vote,
              float* image, cmax_stream_t stream) {
  CMAX_REQUIRE(n >= 0 && xy_stride >= 2, "cmax_vote: need n >= 0 and xy_stride >= 2");
  CMAX_REQUIRE(Hp > 0 && Wp > 0 && (int64_t)Hp * Wp < ((int64_t)1 << 30) && Hp < (1 << 22) && Wp < (1 << 22), "cmax_vote: bad image size %dx%d", Hp, Wp);
  CMAX_REQUIRE(image != nullptr && (n == 0 || xy != nullptr), "cmax_vote: NULL pointer");
  CMAX_REQUIRE(vote == CMAX_VOTE_BILINEAR || vote == CMAX_VOTE_COUNT, "cmax_vote: method %d is not implemented", vote);
  cudaStream_t s = as_stream(stream);
  CMAX_CUDA_CHECK(cudaMemsetAsync(image, 0, (size_t)Hp * Wp * sizeof(float), s));
  if (n > 0) {
    vote_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(xy, n, xy_stride, weight, Hp, Wp, pad_h, pad_w, vote == CMAX_VOTE_COUNT, image);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

int cmax_vote_backward(const float* xy, int64_t n, int xy_stride, const float* weight, int Hp, int Wp, int pad_h, int pad_w,
                       const float* grad_image, float* grad_xy, float* grad_weight, cmax_stream_t stream) {
  CMAX_REQUIRE(n >= 0 && xy_stride >= 2, "cmax_vote_backward: need n >= 0 and xy_stride >= 2");
  CMAX_REQUIRE(Hp > 0 && Wp > 0, "cmax_vote_backward: bad image size %dx%d", Hp, Wp);
  CMAX_REQUIRE(grad_image != nullptr && (n == 0 || (xy != nullptr && grad_xy != nullptr)), "cmax_vote_backward: NULL pointer");
  if (n > 0) {
    vote_backward_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(xy, n, xy_stride, weight, Hp, Wp, pad_h, pad_w,
                                                                              grad_image, grad_xy, grad_weight);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

int cmax_vote_backward2(const float* xy, int64_t n, int xy_stride, const float* weight, int Hp, int Wp, int pad_h, int pad_w,
                        const float* grad_image, const float* u, int u_stride, float* out_grad_image, float* out_grad_xy,
                        cmax_stream_t stream) {
  CMAX_REQUIRE(n >= 0 && xy_stride >= 2 && u_stride >= 2, "cmax_vote_backward2: need n >= 0 and strides >= 2");
  CMAX_REQUIRE(Hp > 0 && Wp > 0, "cmax_vote_backward2: bad image size %dx%d", Hp, Wp);
  CMAX_REQUIRE(out_grad_xy == nullptr || grad_image != nullptr, "cmax_vote_backward2: out_grad_xy needs grad_image");
  CMAX_REQUIRE(n == 0 || (xy != nullptr && u != nullptr), "cmax_vote_backward2: NULL pointer");
  cudaStream_t s = as_stream(stream);
  if (out_grad_image != nullptr) CMAX_CUDA_CHECK(cudaMemsetAsync(out_grad_image, 0, (size_t)Hp * Wp * sizeof(float), s));
  if (n > 0) {
    vote_backward2_kernel<<<grid_for(n, 256, 8), 256, 0, s>>>(xy, n, xy_stride, weight, Hp, Wp, pad_h, pad_w, grad_image, u, u_stride,
                                                              out_grad_image, out_grad_xy);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

int cmax_warp_events_tangent(const float* events, int64_t n, int ev_stride, int H, int W, int motion_model, const cmax_time_params_t* d_params,
                             int ref_index, const float* tangent_motion, float* out, cmax_stream_t stream) {
  CMAX_REQUIRE(n >= 0 && ev_stride >= 3, "cmax_warp_events_tangent: need n >= 0 and ev_stride >= 3");
  CMAX_REQUIRE(motion_model == CMAX_MOTION_DENSE || motion_model == CMAX_MOTION_VOXEL || motion_model == CMAX_MOTION_2DOF,
               "cmax_warp_events_tangent: motion model %d not supported", motion_model);
  CMAX_REQUIRE(d_params != nullptr && tangent_motion != nullptr && (n == 0 || (events != nullptr && out != nullptr)),
               "cmax_warp_events_tangent: NULL pointer");
  CMAX_REQUIRE(ref_index >= 0 && ref_index < CMAX_MAX_REFS, "cmax_warp_events_tangent: ref_index out of range");
  if (n > 0) {
    WarpArgs a{events, n, ev_stride, H, W, motion_model, ref_index, nullptr, d_params};
    warp_events_tangent_kernel<<<grid_for(n, 256, 8), 256, 0, as_stream(stream)>>>(a, tangent_motion, out);
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

int cmax_blur3(const float* in, float* out, int n_img, int Hp, int Wp, float sigma, int transpose, cmax_stream_t stream) {
  CMAX_REQUIRE(in != nullptr && out != nullptr && in != out, "cmax_blur3: in/out must be distinct non-NULL buffers");
  CMAX_REQUIRE(n_img >= 1 && Hp >= 2 && Wp >= 2, "cmax_blur3: reflect padding needs images of at least 2x2 (got %d x %dx%d)", n_img, Hp, Wp);
  CMAX_REQUIRE(sigma > 0.f, "cmax_blur3: sigma must be > 0");
  // taps exp(-x^2 / 2 sigma^2) at x in {-1,0,1}, normalised -- computed in fp32 like torchvision does
  const float e = expf(-0.5f * (1.0f / sigma) * (1.0f / sigma));
  const float sum = e + 1.0f + e;
  const float k0 = e / sum, k1 = 1.0f / sum;
  dim3 grid((Wp + 255) / 256, Hp, n_img);
  if (transpose) blur3_transpose_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, out, Hp, Wp, k0, k1);
  else blur3_kernel<<<grid, 256, 0, as_stream(stream)>>>(in, out, Hp, Wp, k0, k1);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // extern "C"
