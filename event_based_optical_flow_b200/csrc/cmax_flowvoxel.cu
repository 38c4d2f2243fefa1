// Dense flow at t0 -> flow voxel [T,2,H,W] (the time-aware map the reference applies in front of the voxel warp) and its
// adjoint: first-order upwind / inviscid-Burgers propagation, one explicit step per time bin, forward in time on one
// side of t0 and backward (sign-flipped flow) on the other.  SURVEY.md section 8(f) row 2.
//   src/utils/flow_utils.py:99-161 (construct_dense_flow_voxel_torch), :439-493 (upwind step), :567-639 (Burgers step)
//
// The reference runs T-1 dependent stencil steps, each a dozen full-image torch kernels (and as many again in autograd).
// Here the whole propagation of one side is ONE launch: a CTA owns a 16x32 tile, loads it with a halo of k pixels
// (k = number of steps, <= 8 per launch) into shared memory and advances the k levels in place (ping-pong), the valid
// region shrinking by one pixel per level (temporal blocking; the 5-point stencil has reach 1).  Image borders are real
// boundary conditions (zero one-sided differences / replicated neighbours), not halo.  Longer chains are cut into
// chunks of 8 levels.  The adjoint is the same scheme run from the far level towards t0, in gather form (every pixel
// collects from its 4 neighbours), so there are no atomics and the gradient is deterministic.
//
// Arithmetic: fp32, every operation rounded separately in the reference's order (the library is built -fmad=false), so
// the voxel is bit-identical to the reference's fp32 torch result.
#include <vector>

#include "cmax_common.cuh"

namespace cmax {

constexpr int kFvTH = 16, kFvTW = 32;  // output tile of one CTA
constexpr int kFvK = 8;                // levels per launch (= halo)
constexpr int kFvRH = kFvTH + 2 * kFvK, kFvRW = kFvTW + 2 * kFvK;
constexpr int kFvCells = kFvRH * kFvRW;
constexpr int kFvThreads = 256;

struct FvFwdJob {
  const float* src;     // level the chunk starts from, [2,H,W]
  float* copy_dst;      // if non-NULL the start level is also written here (voxel[t0] = dense)
  float* out[kFvK];     // levels produced by steps 1..k
  int k;
  float sgn;            // +1: forward in time, -1: backward (the reference flips the flow sign, flow_utils.py:459-462)
};
struct FvFwdArgs {
  FvFwdJob job[2];
  int H, W, scheme;
  float dt;
};

struct FvAdjJob {
  const float* w_init;   // cotangent of the far level, [2,H,W]
  const float* f[kFvK];  // input level of step s (the level whose cotangent step s produces)
  const float* gv[kFvK]; // cotangent arriving directly at that level (or NULL)
  float* out;            // cotangent of the near level
  int accumulate;        // 0: out = ; 1: atomicAdd into a pre-zeroed buffer (the two sides of t0 finish in any order: a + b == b + a)
  int k;
  float sgn;
};
struct FvAdjArgs {
  FvAdjJob job[2];       // blockIdx.z: the two sides of t0 run concurrently
  int H, W, scheme;
  float dt;
};

__device__ __forceinline__ float fsign(float x) { return (float)((x > 0.f) - (x < 0.f)); }
// torch.maximum(x, 0) / torch.minimum(x, 0) split the gradient evenly at a tie
__device__ __forceinline__ float dmax0(float x) { return x > 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f); }
__device__ __forceinline__ float dmin0(float x) { return x < 0.f ? 1.f : (x == 0.f ? 0.5f : 0.f); }

// One explicit step at one pixel.  U, V: shared-memory level (already multiplied by the direction sign), c: cell,
// P: row pitch; up/dn/lf/rt: the neighbour exists inside the image.
template <int SCHEME>
__device__ __forceinline__ void fv_step(const float* __restrict__ U, const float* __restrict__ V, int c, int P, bool up, bool dn, bool lf,
                                        bool rt, float dt, float& nu, float& nv) {
  const float u = U[c], v = V[c];
  const float up_u = fmaxf(u, 0.f), um_u = fminf(u, 0.f), up_v = fmaxf(v, 0.f), um_v = fminf(v, 0.f);
  if (SCHEME == 0) {  // upwind: flow - dt * (max(u,0) d-x + min(u,0) d+x + max(v,0) d-y + min(v,0) d+y)   flow_utils.py:481-492
    const float* C[2] = {U, V};
    float o[2];
#pragma unroll
    for (int ch = 0; ch < 2; ++ch) {
      const float* A = C[ch];
      const float a = A[c];
      const float dxb = up ? __fsub_rn(a, A[c - P]) : 0.f, dxf = dn ? __fsub_rn(A[c + P], a) : 0.f;
      const float dyb = lf ? __fsub_rn(a, A[c - 1]) : 0.f, dyf = rt ? __fsub_rn(A[c + 1], a) : 0.f;
      float s = __fadd_rn(__fmul_rn(up_u, dxb), __fmul_rn(um_u, dxf));
      s = __fadd_rn(s, __fmul_rn(up_v, dyb));
      s = __fadd_rn(s, __fmul_rn(um_v, dyf));
      o[ch] = __fsub_rn(a, __fmul_rn(dt, s));
    }
    nu = o[0];
    nv = o[1];
  } else {  // inviscid Burgers: conservative form for dFx/dx, dFy/dy, upwind for the cross terms   flow_utils.py:596-638
    const float ub = up ? U[c - P] : u, uf = dn ? U[c + P] : u;  // replicate padding
    const float vb = lf ? V[c - 1] : v, vf = rt ? V[c + 1] : v;
    float bu = __fadd_rn(__fmul_rn(__fmul_rn(u, u), fsign(u)), __fmul_rn(fmaxf(fsign(ub), 0.f), __fmul_rn(-ub, ub)));
    bu = __fmul_rn(__fsub_rn(bu, __fmul_rn(fminf(fsign(uf), 0.f), __fmul_rn(uf, uf))), 0.5f);
    float bv = __fadd_rn(__fmul_rn(__fmul_rn(v, v), fsign(v)), __fmul_rn(fmaxf(fsign(vb), 0.f), __fmul_rn(-vb, vb)));
    bv = __fmul_rn(__fsub_rn(bv, __fmul_rn(fminf(fsign(vf), 0.f), __fmul_rn(vf, vf))), 0.5f);
    const float u_dyb = lf ? __fsub_rn(u, U[c - 1]) : 0.f, u_dyf = rt ? __fsub_rn(U[c + 1], u) : 0.f;
    const float v_dxb = up ? __fsub_rn(v, V[c - P]) : 0.f, v_dxf = dn ? __fsub_rn(V[c + P], v) : 0.f;
    const float su = __fadd_rn(__fadd_rn(__fmul_rn(up_v, u_dyb), __fmul_rn(um_v, u_dyf)), bu);
    const float sv = __fadd_rn(__fadd_rn(__fmul_rn(up_u, v_dxb), __fmul_rn(um_u, v_dxf)), bv);
    nu = __fsub_rn(u, __fmul_rn(dt, su));
    nv = __fsub_rn(v, __fmul_rn(dt, sv));
  }
}

// Transposed Jacobian of one step applied to the cotangent (WU, WV), gathered at one pixel:
//   g = w - dt * (dS/df)^T w      (the direction sign cancels: out = sgn * step(sgn * in))
template <int SCHEME>
__device__ __forceinline__ void fv_adj(const float* __restrict__ U, const float* __restrict__ V, const float* __restrict__ WU,
                                       const float* __restrict__ WV, int c, int P, bool up, bool dn, bool lf, bool rt, float dt, float& gu,
                                       float& gv) {
  const float u = U[c], v = V[c], wu = WU[c], wv = WV[c];
  // coefficients of the one-sided differences at this pixel and at the neighbours that difference against it
  const float A1 = up ? fmaxf(u, 0.f) : 0.f, A2 = dn ? fminf(u, 0.f) : 0.f;
  const float B1 = lf ? fmaxf(v, 0.f) : 0.f, B2 = rt ? fminf(v, 0.f) : 0.f;
  const float A1dn = dn ? fmaxf(U[c + P], 0.f) : 0.f, A2up = up ? fminf(U[c - P], 0.f) : 0.f;
  const float B1rt = rt ? fmaxf(V[c + 1], 0.f) : 0.f, B2lf = lf ? fminf(V[c - 1], 0.f) : 0.f;
  const float wu_dn = dn ? WU[c + P] : 0.f, wu_up = up ? WU[c - P] : 0.f, wu_rt = rt ? WU[c + 1] : 0.f, wu_lf = lf ? WU[c - 1] : 0.f;
  const float wv_dn = dn ? WV[c + P] : 0.f, wv_up = up ? WV[c - P] : 0.f, wv_rt = rt ? WV[c + 1] : 0.f, wv_lf = lf ? WV[c - 1] : 0.f;
  float tu, tv;
  if (SCHEME == 0) {
    const float diag = A1 - A2 + B1 - B2;
    tu = wu * diag - A1dn * wu_dn + A2up * wu_up - B1rt * wu_rt + B2lf * wu_lf;
    tv = wv * diag - A1dn * wv_dn + A2up * wv_up - B1rt * wv_rt + B2lf * wv_lf;
    const float u_dxb = up ? u - U[c - P] : 0.f, u_dxf = dn ? U[c + P] - u : 0.f, u_dyb = lf ? u - U[c - 1] : 0.f, u_dyf = rt ? U[c + 1] - u : 0.f;
    const float v_dxb = up ? v - V[c - P] : 0.f, v_dxf = dn ? V[c + P] - v : 0.f, v_dyb = lf ? v - V[c - 1] : 0.f, v_dyf = rt ? V[c + 1] - v : 0.f;
    tu += wu * (dmax0(u) * u_dxb + dmin0(u) * u_dxf) + wv * (dmax0(u) * v_dxb + dmin0(u) * v_dxf);
    tv += wu * (dmax0(v) * u_dyb + dmin0(v) * u_dyf) + wv * (dmax0(v) * v_dyb + dmin0(v) * v_dyf);
  } else {
    const float u_dyb = lf ? u - U[c - 1] : 0.f, u_dyf = rt ? U[c + 1] - u : 0.f;
    const float v_dxb = up ? v - V[c - P] : 0.f, v_dxf = dn ? V[c + P] - v : 0.f;
    tu = wu * (B1 - B2) - B1rt * wu_rt + B2lf * wu_lf + wv * (dmax0(u) * v_dxb + dmin0(u) * v_dxf);
    tv = wv * (A1 - A2) - A1dn * wv_dn + A2up * wv_up + wu * (dmax0(v) * u_dyb + dmin0(v) * u_dyf);
    // conservative term: d/du (u^2 sign u)/2 = |u|; the replicated neighbour of a border pixel is the pixel itself
    tu += wu * fabsf(u);
    tv += wv * fabsf(v);
    if (u > 0.f) tu -= u * (wu_dn + (up ? 0.f : wu));
    if (u < 0.f) tu += u * (wu_up + (dn ? 0.f : wu));
    if (v > 0.f) tv -= v * (wv_rt + (lf ? 0.f : wv));
    if (v < 0.f) tv += v * (wv_lf + (rt ? 0.f : wv));
  }
  gu = wu - dt * tu;
  gv = wv - dt * tv;
}

// c / w for 0 <= c < 2^16 and 1 <= w < 2^8 without the ~20-instruction integer division: the exact quotient's distance to
// the next integer is >= 0.5 / w, far above the fp32 rounding error of the product
__device__ __forceinline__ int fast_div(int c, float inv_w) { return (int)(((float)c + 0.5f) * inv_w); }

// region geometry shared by both kernels: the CTA's tile with a halo of k pixels
struct FvRegion {
  int gi0, gj0, rh, rw;
};
__device__ __forceinline__ FvRegion fv_region(int k) {
  FvRegion r;
  r.gi0 = blockIdx.y * kFvTH - k;
  r.gj0 = blockIdx.x * kFvTW - k;
  r.rh = kFvTH + 2 * k;
  r.rw = kFvTW + 2 * k;
  return r;
}

template <int SCHEME>
__global__ void __launch_bounds__(kFvThreads) flow_voxel_kernel(FvFwdArgs a) {
  __shared__ float lev[2][2][kFvCells];  // [ping-pong][channel][cell]
  const FvFwdJob& job = a.job[blockIdx.z];
  const int k = job.k, H = a.H, W = a.W;
  const int64_t HW = (int64_t)H * W;
  const FvRegion R = fv_region(k);
  const int n_cells = R.rh * R.rw;
  const float inv_rw = 1.0f / (float)R.rw;
  for (int c = threadIdx.x; c < n_cells; c += kFvThreads) {
    const int li = fast_div(c, inv_rw), lj = c - li * R.rw, gi = R.gi0 + li, gj = R.gj0 + lj;
    float u = 0.f, v = 0.f;
    if (gi >= 0 && gi < H && gj >= 0 && gj < W) {
      const int64_t p = (int64_t)gi * W + gj;
      u = __ldg(job.src + p);
      v = __ldg(job.src + HW + p);
      if (job.copy_dst != nullptr && li >= k && li < k + kFvTH && lj >= k && lj < k + kFvTW) {
        job.copy_dst[p] = u;
        job.copy_dst[HW + p] = v;
      }
      u *= job.sgn;
      v *= job.sgn;
    }
    lev[0][0][c] = u;
    lev[0][1][c] = v;
  }
  for (int s = 1; s <= k; ++s) {
    __syncthreads();
    const float* U = lev[(s - 1) & 1][0];
    const float* V = lev[(s - 1) & 1][1];
    float* NU = lev[s & 1][0];
    float* NV = lev[s & 1][1];
    float* out = job.out[s - 1];
    const int ih = R.rh - 2 * s, iw = R.rw - 2 * s;  // cells still valid at this level
    const float inv_iw = 1.0f / (float)iw;
    for (int q = threadIdx.x; q < ih * iw; q += kFvThreads) {
      const int qi = fast_div(q, inv_iw);
      const int li = s + qi, lj = s + (q - qi * iw), gi = R.gi0 + li, gj = R.gj0 + lj;
      if (gi < 0 || gi >= H || gj < 0 || gj >= W) continue;
      const int c = li * R.rw + lj;
      float nu, nv;
      fv_step<SCHEME>(U, V, c, R.rw, gi > 0, gi < H - 1, gj > 0, gj < W - 1, a.dt, nu, nv);
      NU[c] = nu;
      NV[c] = nv;
      if (li >= k && li < k + kFvTH && lj >= k && lj < k + kFvTW) {
        const int64_t p = (int64_t)gi * W + gj;
        out[p] = nu * job.sgn;
        out[HW + p] = nv * job.sgn;
      }
    }
  }
}

template <int SCHEME>
__global__ void __launch_bounds__(kFvThreads) flow_voxel_adjoint_kernel(FvAdjArgs a) {
  __shared__ float wbuf[2][2][kFvCells];  // cotangent, ping-pong
  __shared__ float fbuf[2][kFvCells];     // input level of the current step (times the direction sign)
  const FvAdjJob& job = a.job[blockIdx.z];
  const int k = job.k, H = a.H, W = a.W;
  const int64_t HW = (int64_t)H * W;
  const FvRegion R = fv_region(k);
  const int n_cells = R.rh * R.rw;
  const float inv_rw = 1.0f / (float)R.rw;
  for (int c = threadIdx.x; c < n_cells; c += kFvThreads) {
    const int ci = fast_div(c, inv_rw);
    const int gi = R.gi0 + ci, gj = R.gj0 + (c - ci * R.rw);
    float u = 0.f, v = 0.f;
    if (gi >= 0 && gi < H && gj >= 0 && gj < W) {
      const int64_t p = (int64_t)gi * W + gj;
      u = __ldg(job.w_init + p);
      v = __ldg(job.w_init + HW + p);
    }
    wbuf[0][0][c] = u;
    wbuf[0][1][c] = v;
  }
  for (int s = 1; s <= k; ++s) {
    __syncthreads();  // the previous level is complete, and fbuf is free again
    const float* F = job.f[s - 1];
    for (int c = threadIdx.x; c < n_cells; c += kFvThreads) {
      const int ci = fast_div(c, inv_rw);
      const int gi = R.gi0 + ci, gj = R.gj0 + (c - ci * R.rw);
      float u = 0.f, v = 0.f;
      if (gi >= 0 && gi < H && gj >= 0 && gj < W) {
        const int64_t p = (int64_t)gi * W + gj;
        u = __ldg(F + p) * job.sgn;
        v = __ldg(F + HW + p) * job.sgn;
      }
      fbuf[0][c] = u;
      fbuf[1][c] = v;
    }
    __syncthreads();
    const float* WU = wbuf[(s - 1) & 1][0];
    const float* WV = wbuf[(s - 1) & 1][1];
    float* NU = wbuf[s & 1][0];
    float* NV = wbuf[s & 1][1];
    const float* GV = job.gv[s - 1];
    const int ih = R.rh - 2 * s, iw = R.rw - 2 * s;
    const float inv_iw = 1.0f / (float)iw;
    for (int q = threadIdx.x; q < ih * iw; q += kFvThreads) {
      const int qi = fast_div(q, inv_iw);
      const int li = s + qi, lj = s + (q - qi * iw), gi = R.gi0 + li, gj = R.gj0 + lj;
      if (gi < 0 || gi >= H || gj < 0 || gj >= W) continue;
      const int c = li * R.rw + lj;
      float gu, gv;
      fv_adj<SCHEME>(fbuf[0], fbuf[1], WU, WV, c, R.rw, gi > 0, gi < H - 1, gj > 0, gj < W - 1, a.dt, gu, gv);
      const int64_t p = (int64_t)gi * W + gj;
      if (GV != nullptr) {
        gu += __ldg(GV + p);
        gv += __ldg(GV + HW + p);
      }
      NU[c] = gu;
      NV[c] = gv;
      if (s == k && li >= k && li < k + kFvTH && lj >= k && lj < k + kFvTW) {
        if (job.accumulate) {
          atomicAdd(job.out + p, gu);
          atomicAdd(job.out + HW + p, gv);
        } else {
          job.out[p] = gu;
          job.out[HW + p] = gv;
        }
      }
    }
  }
}

// The propagation itself has no structural limit on the number of levels (the reference's own tests use 60 and 100,
// tests/utils/test_flow_utils.py:52-106); the voxel WARP is limited to CMAX_MAX_BINS by its edge table.
constexpr int kFvMaxLevels = 4096;

struct FvGeom {
  int H, W, T, scheme, t0;
  float dt;
  // The reference's Burgers backward loop also runs for i = 0 and writes level -1 = T-1 (flow_utils.py:140-141); the
  // forward loop overwrites that level again unless t0 == T-1 (T == 1, or T == 2 with t0 in the middle), in which case
  // level t0 itself ends up t0+1 backward steps away from the input and no level holds the input any more.
  bool wrap;
};

static int fv_geom(const char* fn, int H, int W, int T, int scheme, int t0_middle, FvGeom* g) {
  CMAX_REQUIRE(H >= 1 && W >= 1, "%s: bad image size %dx%d", fn, H, W);
  CMAX_REQUIRE(T >= 1 && T <= kFvMaxLevels, "%s: time_bin must be in [1,%d], got %d", fn, kFvMaxLevels, T);
  CMAX_REQUIRE(scheme == CMAX_SCHEME_UPWIND || scheme == CMAX_SCHEME_BURGERS, "%s: unknown scheme %d", fn, scheme);
  g->H = H; g->W = W; g->T = T; g->scheme = scheme;
  g->t0 = t0_middle ? T / 2 : 0;
  g->dt = (float)(1.0 / (double)T);  // the reference's Python float 1.0 / time_bin, narrowed by torch to the tensor dtype
  g->wrap = (scheme == CMAX_SCHEME_BURGERS && g->t0 == T - 1);
  return CMAX_OK;
}

// The levels each side of t0 produces, in the order they are stepped: forward side t0+1 .. T-1; backward side
// t0-1 .. 0 (followed by T-1 when the backward loop wraps).
static std::vector<int> fv_side(const FvGeom& g, bool forward) {
  std::vector<int> levels;
  if (forward) {
    for (int l = g.t0 + 1; l <= g.T - 1; ++l) levels.push_back(l);
  } else {
    for (int l = g.t0 - 1; l >= 0; --l) levels.push_back(l);
    if (g.wrap) levels.push_back(g.T - 1);
  }
  return levels;
}

static dim3 fv_grid(const FvGeom& g, int z) { return dim3((g.W + kFvTW - 1) / kFvTW, (g.H + kFvTH - 1) / kFvTH, z); }

static void launch_fwd(const FvGeom& g, const FvFwdArgs& a, int njobs, cudaStream_t s) {
  if (g.scheme == CMAX_SCHEME_UPWIND) flow_voxel_kernel<0><<<fv_grid(g, njobs), kFvThreads, 0, s>>>(a);
  else flow_voxel_kernel<1><<<fv_grid(g, njobs), kFvThreads, 0, s>>>(a);
}
static void launch_adj(const FvGeom& g, const FvAdjArgs& a, int njobs, cudaStream_t s) {
  if (g.scheme == CMAX_SCHEME_UPWIND) flow_voxel_adjoint_kernel<0><<<fv_grid(g, njobs), kFvThreads, 0, s>>>(a);
  else flow_voxel_adjoint_kernel<1><<<fv_grid(g, njobs), kFvThreads, 0, s>>>(a);
}

}  // namespace cmax

using namespace cmax;

extern "C" {

size_t cmax_flow_voxel_workspace_bytes(int H, int W) { return (H < 1 || W < 1) ? 0 : (size_t)4 * 2 * (size_t)H * W * sizeof(float); }

int cmax_flow_voxel(const float* dense, int H, int W, int time_bin, int scheme, int t0_middle, float* voxel, cmax_stream_t stream) {
  CMAX_REQUIRE(dense != nullptr && voxel != nullptr, "cmax_flow_voxel: NULL pointer");
  FvGeom g;
  const int rc = fv_geom("cmax_flow_voxel", H, W, time_bin, scheme, t0_middle, &g);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const int64_t L = 2 * (int64_t)H * W;  // floats per level
  const std::vector<int> lv[2] = {fv_side(g, true), fv_side(g, false)};
  const int n[2] = {(int)lv[0].size(), (int)lv[1].size()};
  if (n[0] == 0 && n[1] == 0) {  // T == 1, upwind
    CMAX_CUDA_CHECK(cudaMemcpyAsync(voxel, dense, (size_t)L * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return CMAX_OK;
  }
  FvFwdArgs a;
  a.H = H; a.W = W; a.scheme = scheme; a.dt = g.dt;
  bool copied = g.wrap;  // (when the backward loop wraps, no level keeps the input)
  int done[2] = {0, 0};
  while (done[0] < n[0] || done[1] < n[1]) {  // wave w = chunk w of both sides in one launch
    int nj = 0;
    for (int side = 0; side < 2; ++side) {
      if (done[side] >= n[side]) continue;
      FvFwdJob& j = a.job[nj++];
      memset(&j, 0, sizeof(j));
      j.k = std::min(kFvK, n[side] - done[side]);
      j.src = (done[side] == 0) ? dense : voxel + (int64_t)lv[side][done[side] - 1] * L;
      if (!copied) {
        j.copy_dst = voxel + (int64_t)g.t0 * L;
        copied = true;
      }
      for (int q = 0; q < j.k; ++q) j.out[q] = voxel + (int64_t)lv[side][done[side] + q] * L;
      j.sgn = side == 0 ? 1.f : -1.f;
      done[side] += j.k;
    }
    launch_fwd(g, a, nj, s);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

int cmax_flow_voxel_backward(const float* dense, const float* voxel, const float* grad_voxel, int H, int W, int time_bin, int scheme,
                             int t0_middle, float* grad_dense, void* workspace, cmax_stream_t stream) {
  CMAX_REQUIRE(dense != nullptr && voxel != nullptr && grad_voxel != nullptr && grad_dense != nullptr, "cmax_flow_voxel_backward: NULL pointer");
  FvGeom g;
  const int rc = fv_geom("cmax_flow_voxel_backward", H, W, time_bin, scheme, t0_middle, &g);
  if (rc) return rc;
  cudaStream_t s = as_stream(stream);
  const int64_t L = 2 * (int64_t)H * W;
  const std::vector<int> lv[2] = {fv_side(g, true), fv_side(g, false)};
  const int n[2] = {(int)lv[0].size(), (int)lv[1].size()};
  if (n[0] == 0 && n[1] == 0) {
    CMAX_CUDA_CHECK(cudaMemcpyAsync(grad_dense, grad_voxel, (size_t)L * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return CMAX_OK;
  }
  CMAX_REQUIRE((n[0] <= kFvK && n[1] <= kFvK) || workspace != nullptr,
               "cmax_flow_voxel_backward: more than %d levels on one side of t0 needs the workspace", kFvK);
  // carry buffers between the chunks of one side (only chains of more than 8 levels need them): two per side, ping-pong
  float* carry[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
  if (workspace != nullptr) {
    float* wsf = static_cast<float*>(workspace);
    carry[0][0] = wsf; carry[0][1] = wsf + L; carry[1][0] = wsf + 2 * L; carry[1][1] = wsf + 3 * L;
  }
  const bool both = n[0] > 0 && n[1] > 0;
  if (both) CMAX_CUDA_CHECK(cudaMemsetAsync(grad_dense, 0, (size_t)L * sizeof(float), s));  // the two sides add into it
  FvAdjArgs a;
  a.H = H; a.W = W; a.scheme = scheme; a.dt = g.dt;
  bool t0_direct = !g.wrap;  // the cotangent of level t0 reaches the input directly (voxel[t0] = dense): add it exactly once
  int done[2] = {0, 0}, pp[2] = {0, 0};
  const float* w[2] = {n[0] ? grad_voxel + (int64_t)lv[0][n[0] - 1] * L : nullptr, n[1] ? grad_voxel + (int64_t)lv[1][n[1] - 1] * L : nullptr};
  // each side walks from its far level back towards t0 (step q undoes the forward step that produced lv[side][n-1-q]; its
  // input level is lv[side][n-2-q], or the dense input itself for the first forward step of the side); wave = chunk of both
  // sides in ONE launch
  while (done[0] < n[0] || done[1] < n[1]) {
    int nj = 0;
    for (int side = 0; side < 2; ++side) {
      if (done[side] >= n[side]) continue;
      FvAdjJob& j = a.job[nj++];
      memset(&j, 0, sizeof(j));
      j.k = std::min(kFvK, n[side] - done[side]);
      j.w_init = w[side];
      j.sgn = side == 0 ? 1.f : -1.f;
      for (int q = 0; q < j.k; ++q) {
        const int in = n[side] - 2 - (done[side] + q);  // index into lv of the step's input level; -1 = the dense input
        if (in >= 0) {
          j.f[q] = voxel + (int64_t)lv[side][in] * L;
          j.gv[q] = grad_voxel + (int64_t)lv[side][in] * L;
        } else {
          j.f[q] = dense;
          j.gv[q] = t0_direct ? grad_voxel + (int64_t)g.t0 * L : nullptr;
          t0_direct = false;
        }
      }
      done[side] += j.k;
      if (done[side] == n[side]) {
        j.out = grad_dense;
        j.accumulate = both ? 1 : 0;
      } else {
        j.out = carry[side][pp[side]];
        j.accumulate = 0;
        w[side] = carry[side][pp[side]];
        pp[side] ^= 1;
      }
    }
    launch_adj(g, a, nj, s);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

}  // extern "C"
