// Event-batch preparation: error plumbing, time range / reference-time parameters, and the resident plan
// (validation + stable tile sort).  Everything here runs once per solver.optimize(), not per CM iteration.
#include <stdarg.h>
#include <string.h>

#include <new>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "cmax_plan.cuh"

namespace cmax {

static thread_local std::string g_last_error;

void set_error(const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return CMAX_ERR_CUDA;
}

// ------------------------------------------------------------------------------------------------ time range
__device__ __forceinline__ void atomic_min_float(float* addr, float v) {
  if (v >= 0.0f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void init_minmax_kernel(float* mm) {
  mm[0] = INFINITY;
  mm[1] = -INFINITY;
}

// min/max over column 2.                                     src/warp.py:217-224, 256-257 (nt_min / nt_max)
__global__ void __launch_bounds__(256) time_range_kernel(const float* __restrict__ ev, int64_t n, int stride,
                                                         float* __restrict__ mm) {
  float lo = INFINITY, hi = -INFINITY;
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    const float t = (stride == 4) ? ld_event(ev, i).z : __ldg(ev + i * stride + 2);
    lo = fminf(lo, t);
    hi = fmaxf(hi, t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0 && lo <= hi) {
    atomic_min_float(mm + 0, lo);
    atomic_max_float(mm + 1, hi);
  }
}

// One thread restates the reference's scalar arithmetic: fp32 for ref/period (0-dim tensor ops), float64 for the
// bin edges (numpy on the host), then the edge is rounded to fp32 because torch compares `scalar <= fp32 tensor`
// in fp32.                                                  src/warp.py:201-233, 254-258, 342-345
__global__ void time_params_kernel(const float* __restrict__ mm, cmax_time_params_t* __restrict__ out, int n_ref,
                                   int n_bins, int normalize_t, cmax_ref r0, cmax_ref r1, cmax_ref r2, cmax_ref r3) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const cmax_ref refs[CMAX_MAX_REFS] = {r0, r1, r2, r3};
  const float tmin = mm[0], tmax = mm[1];
  out->n_ref = n_ref;
  out->n_bins = n_bins;
  out->normalize_t = normalize_t;
  out->pad_ = 0;
  for (int r = 0; r < n_ref; ++r) {
    float ref;
    if (refs[r].mode == 0) ref = tmin;
    else if (refs[r].mode == 1) ref = tmax;
    else ref = __fadd_rn(tmin, __fmul_rn(__fsub_rn(tmax, tmin), refs[r].fraction));
    const float dlo = __fsub_rn(tmin, ref), dhi = __fsub_rn(tmax, ref);
    const float period = __fsub_rn(dhi, dlo);
    const float lo = normalize_t ? __fdiv_rn(dlo, period) : dlo;
    const float hi = normalize_t ? __fdiv_rn(dhi, period) : dhi;
    out->ref[r] = ref;
    out->period[r] = period;
    out->dt_min[r] = lo;
    out->dt_max[r] = hi;
    const double span = __dsub_rn((double)hi, (double)lo);
    for (int b = 0; b < n_bins; ++b) {
      const double e = __dadd_rn(__dmul_rn(__ddiv_rn((double)b, (double)n_bins), span), (double)lo);
      out->edges[r][b] = (float)e;
    }
    out->edges[r][n_bins] = (float)__dadd_rn((double)hi, 1000.0);
    for (int b = n_bins + 1; b <= CMAX_MAX_BINS; ++b) out->edges[r][b] = INFINITY;
  }
}

// ------------------------------------------------------------------------------------------------ plan kernels
// Source pixel of an event: .long() truncation of the un-warped coordinates, src/warp.py:305
__device__ __forceinline__ bool source_pixel(float x, float y, int H, int W, int* r, int* c) {
  *r = __float2int_rz(x);
  *c = __float2int_rz(y);
  return (*r >= 0) && (*r < H) && (*c >= 0) && (*c < W) && (x == x) && (y == y);
}

// One pass over the caller's events: validates the source pixel (status bit 0: outside the image or NaN), notices fractional
// coordinates (status bit 1: no compact / strip packing), and writes the sort key -- tile id (TILE order: events of a tile
// keep their time order) or tile id * 1024 + pixel inside the tile (PIXEL order) -- with the identity permutation.  No
// atomics per event: the per-key counts come from the SORTED keys afterwards (key_first_kernel).  An invalid event gets
// key 0, so everything enqueued behind this kernel stays in bounds until the host has looked at the status word.
__global__ void __launch_bounds__(256) keys_kernel(const float* __restrict__ ev, int64_t n, int stride, int H, int W, int tiles_x,
                                                   int by_pixel, uint32_t* __restrict__ keys, uint32_t* __restrict__ idx,
                                                   int32_t* __restrict__ status) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  bool bad = false, frac = false;
  int hi1 = 0, ilo = 0;  // max(row + 1), max(H - row) over the valid events: 0 = none (status[2], status[3])
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += step) {
    int r, c;
    const float x = __ldg(ev + i * stride), y = __ldg(ev + i * stride + 1);
    const bool ok = source_pixel(x, y, H, W, &r, &c);
    bad |= !ok;
    if (ok) {
      hi1 = max(hi1, r + 1);
      ilo = max(ilo, H - r);
    }
    frac |= (x != truncf(x)) || (y != truncf(y));
    if (keys != nullptr) {
      uint32_t key = 0u;
      if (ok) {
        const uint32_t tile = (uint32_t)((r / kTile) * tiles_x + (c / kTile));
        key = by_pixel ? (tile * (uint32_t)(kTile * kTile) + (uint32_t)((r % kTile) * kTile + (c % kTile))) : tile;
      }
      keys[i] = key;
      idx[i] = (uint32_t)i;
    }
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(status, 1);
  if (__any_sync(0xffffffffu, frac) && (threadIdx.x & 31) == 0) atomicOr(status, 2);
  hi1 = __reduce_max_sync(0xffffffffu, hi1);
  ilo = __reduce_max_sync(0xffffffffu, ilo);
  if ((threadIdx.x & 31) == 0 && hi1 > 0) {
    atomicMax(status + 2, hi1);
    atomicMax(status + 3, ilo);
  }
}

// key_first[k] = index of the first event whose key is >= k (k in [0, n_keys]); events of key k are
// key_first[k] .. key_first[k+1].  One thread per boundary of the sorted key array fills the (possibly empty) keys in between.
__global__ void __launch_bounds__(256) key_first_kernel(const uint32_t* __restrict__ sorted_keys, int64_t n, int64_t n_keys,
                                                        uint32_t* __restrict__ key_first) {
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += (int64_t)gridDim.x * blockDim.x) {
    const int64_t k = j < n ? (int64_t)sorted_keys[j] : n_keys;
    const int64_t kprev = j > 0 ? (int64_t)sorted_keys[j - 1] : -1;
    for (int64_t q = kprev + 1; q <= k; ++q) key_first[q] = (uint32_t)j;
  }
}

__global__ void __launch_bounds__(256) gather_events_kernel(const float* __restrict__ ev, int64_t n, int stride,
                                                            const uint32_t* __restrict__ idx, float4* __restrict__ out) {
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += step) {
    const float* e = ev + (int64_t)idx[j] * stride;
    out[j] = make_float4(__ldg(e), __ldg(e + 1), __ldg(e + 2), stride >= 4 ? __ldg(e + 3) : 0.f);
  }
}

// (x, y, t, p) -> (x, y, tz, bits(src)), written in WARP-TILE order: logical event j = tile*256 + lane*8 + k is
// stored at slot tile*256 + k*32 + lane, so that a coalesced 16-byte load by lane `lane` at step k returns that lane's
// k-th consecutive event (see cmax_plan.cuh).  dt exactly as the kernels compute it (src/warp.py:254-258).
// Slots past n (the last tile's padding) are zero-filled.
template <bool COMPACT>
__global__ void __launch_bounds__(256) repack_kernel(const float4* __restrict__ ev, int64_t n, int64_t slots, int H, int W,
                                                     const cmax_time_params_t* __restrict__ tp, int with_dt, void* __restrict__ out) {
  const float ref = tp->ref[0], period = tp->period[0];
  const int64_t step = (int64_t)gridDim.x * blockDim.x;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < slots; j += step) {
    float4 o = make_float4(0.f, 0.f, 0.f, __int_as_float(-1));
    uint2 oc = make_uint2(0u, 0u);
    if (j < n) {
      const float4 e = ev[j];
      int r, c;
      source_pixel(e.x, e.y, H, W, &r, &c);
      const float tz = with_dt ? normalised_dt(e.z, ref, period, 1) : e.z;
      o = make_float4(e.x, e.y, tz, __int_as_float(r * W + c));
      oc = make_uint2(__float_as_uint(tz), ((unsigned)r << 16) | (unsigned)c);
    }
    const int64_t tile = j / kWarpTile;
    const int within = (int)(j % kWarpTile), lane = within / kRunE, k = within % kRunE;
    const int64_t slot = tile * kWarpTile + k * 32 + lane;
    if (COMPACT) reinterpret_cast<uint2*>(out)[slot] = oc;
    else reinterpret_cast<float4*>(out)[slot] = o;
  }
}

// ---- strips: the packed format of the strip kernels (cmax_lean.cu).  Events are in source-pixel order, so the events of
// one pixel are a run; every run is cut into STRIPS of kRunE events and the last strip of a run is padded, so that a
// strip never mixes two source pixels.  One thread of a strip kernel walks one strip: the source pixel -- hence its float
// coordinates, flat index and flow vector -- is a per-strip constant read from a 4-byte header instead of a per-event
// key that has to be compared, and an event is just its 4-byte (normalised) time.  A warp-tile is 32 strips:
//   [32 headers: row << 17 | col << 4 | count][32 x float4: times 0..3 of lane's strip][32 x float4: times 4..7]
// = 1152 bytes, one TMA bulk copy; both float4 blocks are read with conflict-free LDS.128.  4.5 bytes per event instead
// of 8, plus the padding (half a strip per occupied pixel on average: ~6 % at 55 events per pixel).
__global__ void __launch_bounds__(256) strip_counts_kernel(const uint32_t* __restrict__ key_first, int64_t n_keys,
                                                           uint32_t* __restrict__ key_strips) {
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n_keys; k += (int64_t)gridDim.x * blockDim.x)
    key_strips[k] = k < n_keys ? (key_first[k + 1] - key_first[k] + kRunE - 1) / kRunE : 0u;
}

__global__ void __launch_bounds__(256) pack_strips_kernel(const float4* __restrict__ ev, const uint32_t* __restrict__ sorted_keys, int64_t n,
                                                          const uint32_t* __restrict__ key_first,
                                                          const uint32_t* __restrict__ key_strip0, int H, int W,
                                                          const cmax_time_params_t* __restrict__ tp, int with_dt, int tile_bytes,
                                                          unsigned char* __restrict__ out) {
  const float ref = tp->ref[0], period = tp->period[0];
  const int n_bins = tp->n_bins, n_ref = tp->n_ref;
  for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t key = sorted_keys[j];
    const uint32_t rank = (uint32_t)j - key_first[key];
    const uint32_t strip = key_strip0[key] + rank / kRunE, slot = rank % kRunE;
    const float4 e = ev[j];
    unsigned char* tile = out + (size_t)(strip / 32) * tile_bytes;
    const uint32_t lane = strip % 32;
    const float tz = with_dt ? normalised_dt(e.z, ref, period, 1) : e.z;
    reinterpret_cast<float*>(tile + 128 + (slot / 4) * 512 + lane * 16)[slot % 4] = tz;
    if (n_bins > 0) {
      // time-aware plans: the time bin of every event for every reference time is iteration-invariant too
      // (src/warp.py:342-352) -- one byte each, 255 = in no bin
      for (int r = 0; r < n_ref; ++r) {
        const float dt = normalised_dt(e.z, tp->ref[r], tp->period[r], 1);
        const float inv = (float)n_bins / (tp->dt_max[r] - tp->dt_min[r]);
        const int b = time_bin(dt, tp->edges[r], n_bins, tp->dt_min[r], inv);
        tile[kStripTileBytes + r * kStripBinBytes + lane * kRunE + slot] = (unsigned char)(b < 0 ? 255 : b);
      }
    }
    if (slot == 0) {
      int r, c;
      source_pixel(e.x, e.y, H, W, &r, &c);
      const uint32_t left = key_first[key + 1] - (uint32_t)j;
      reinterpret_cast<uint32_t*>(tile)[lane] = ((uint32_t)r << 17) | ((uint32_t)c << 4) | (left < (uint32_t)kRunE ? left : (uint32_t)kRunE);
    }
  }
}

// slots of the packed copy: whole warp-tiles of kWarpTile events
static inline int64_t packed_slots(int64_t n) { return std::max<int64_t>(1, (n + kWarpTile - 1) / kWarpTile) * kWarpTile; }

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int bits_for(uint64_t v) {
  int b = 1;
  while (b < 32 && (1ull << b) < v) ++b;
  return b;
}

struct PlanLayout {
  size_t off_packed;
  size_t off_params, off_minmax, off_status, off_events, off_keys[2], off_idx[2], off_temp, temp_bytes, total;
  size_t off_kstr, off_kfirst, off_kstrip0, off_strips, off_scan_temp, scan_temp_bytes;
  int64_t n_keys, strip_capacity;  // strips the region holds (whole warp-tiles)
  int tiles_x, tiles_y, n_tiles, key_bits;
};

// Returns false (message set) when the CUB temp-storage query fails, e.g. without a device.
static bool plan_layout(int64_t n, int H, int W, int order, PlanLayout* out) {
  PlanLayout L;
  memset(&L, 0, sizeof(L));
  L.tiles_x = (W + kTile - 1) / kTile;
  L.tiles_y = (H + kTile - 1) / kTile;
  L.n_tiles = L.tiles_x * L.tiles_y;
  size_t off = 0;
  L.off_params = off; off = align_up(off + sizeof(cmax_time_params_t), 256);
  L.off_minmax = off; off = align_up(off + 2 * sizeof(float), 256);
  L.off_status = off; off = align_up(off + 4 * sizeof(int32_t), 256);  // [status bits, number of strips, max(row+1), max(H-row)]
  L.off_packed = off; off = align_up(off + (size_t)packed_slots(n) * sizeof(float4), 256);
  if (order != CMAX_ORDER_ASIS) {
    L.key_bits = bits_for((uint64_t)L.n_tiles * (order == CMAX_ORDER_PIXEL ? kTile * kTile : 1));
    L.off_events = off; off = align_up(off + (size_t)n * sizeof(float4), 256);
    for (int k = 0; k < 2; ++k) {
      L.off_keys[k] = off; off = align_up(off + (size_t)n * sizeof(uint32_t), 256);
      L.off_idx[k] = off; off = align_up(off + (size_t)n * sizeof(uint32_t), 256);
    }
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    size_t temp = 0;
    const cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp, dk, dv, (int)n, 0, L.key_bits, (cudaStream_t)0);
    if (e != cudaSuccess) {
      cuda_fail(e, "cub::DeviceRadixSort::SortPairs (temp-storage query)");
      return false;
    }
    L.temp_bytes = temp;
    L.off_temp = off; off = align_up(off + temp + 256, 256);
    if (order == CMAX_ORDER_PIXEL) {  // strips (see pack_strips_kernel): allowed to be up to 1.5x the events
      L.n_keys = (int64_t)L.n_tiles * kTile * kTile;
      L.strip_capacity = ((n + n / 2) / kRunE + 31) / 32 * 32 + 32;
      for (size_t* o : {&L.off_kstr, &L.off_kfirst, &L.off_kstrip0}) {
        *o = off; off = align_up(off + (size_t)(L.n_keys + 1) * sizeof(uint32_t), 256);
      }
      size_t st = 0;
      const cudaError_t e2 = cub::DeviceScan::ExclusiveSum(nullptr, st, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)(L.n_keys + 1), (cudaStream_t)0);
      if (e2 != cudaSuccess) {
        cuda_fail(e2, "cub::DeviceScan::ExclusiveSum (temp-storage query)");
        return false;
      }
      L.scan_temp_bytes = st;
      L.off_scan_temp = off; off = align_up(off + st + 256, 256);
      L.off_strips = off; off = align_up(off + (size_t)(L.strip_capacity / 32) * (kStripTileBytes + CMAX_MAX_REFS * kStripBinBytes), 256);
    }
  }
  L.total = off;
  *out = L;
  return true;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;  // B200
    cached[dev] = v;
  }
  return cached[dev];
}

// (Re)build the packed copy the RUN kernels read (the strip kernels never touch it, so a plan with strips only pays for it
// when a run variant is selected).
int ensure_packed(const cmax_plan* plan, cudaStream_t s) {
  cmax_plan* p = const_cast<cmax_plan*>(plan);
  if (p->packed_valid || p->n == 0) return CMAX_OK;
  const int64_t slots = packed_slots(p->n);
  const int grid = (int)std::min<int64_t>(num_sms() * 8, (slots + 255) / 256);
  if (p->compact)
    repack_kernel<true><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(p->events), p->n, slots, p->H, p->W, p->d_params, p->packed_has_dt, p->packed);
  else
    repack_kernel<false><<<grid, 256, 0, s>>>(reinterpret_cast<const float4*>(p->events), p->n, slots, p->H, p->W, p->d_params, p->packed_has_dt, p->packed);
  CMAX_CUDA_CHECK(cudaGetLastError());
  p->packed_valid = 1;
  return CMAX_OK;
}

}  // namespace cmax

using namespace cmax;

extern "C" {

int cmax_abi_version(void) { return CMAX_ABI_VERSION; }
const char* cmax_last_error(void) { return g_last_error.c_str(); }
const char* cmax_build_arch(void) { return "sm_100a"; }

int cmax_time_range(const float* events, int64_t n, int ev_stride, float* d_minmax, cmax_stream_t stream) {
  CMAX_REQUIRE(d_minmax != nullptr, "cmax_time_range: d_minmax is NULL");
  CMAX_REQUIRE(n >= 0 && ev_stride >= 3, "cmax_time_range: need n >= 0 and ev_stride >= 3 (got %lld, %d)", (long long)n, ev_stride);
  CMAX_REQUIRE(n == 0 || events != nullptr, "cmax_time_range: events is NULL");
  cudaStream_t s = as_stream(stream);
  init_minmax_kernel<<<1, 1, 0, s>>>(d_minmax);
  if (n > 0) {
    const int grid = (int)std::min<int64_t>(num_sms() * 8, (n + 255) / 256);
    time_range_kernel<<<grid, 256, 0, s>>>(events, n, ev_stride, d_minmax);
  }
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

int cmax_time_params(const float* d_minmax, const cmax_ref* h_refs, int n_ref, int n_bins, int normalize_t,
                     cmax_time_params_t* d_params, cmax_stream_t stream) {
  CMAX_REQUIRE(d_minmax && h_refs && d_params, "cmax_time_params: NULL argument");
  CMAX_REQUIRE(n_ref >= 1 && n_ref <= CMAX_MAX_REFS, "cmax_time_params: n_ref must be in [1,%d], got %d", CMAX_MAX_REFS, n_ref);
  CMAX_REQUIRE(n_bins >= 0 && n_bins <= CMAX_MAX_BINS, "cmax_time_params: n_bins must be in [0,%d], got %d", CMAX_MAX_BINS, n_bins);
  cmax_ref r[CMAX_MAX_REFS];
  for (int i = 0; i < CMAX_MAX_REFS; ++i) r[i] = h_refs[i < n_ref ? i : 0];
  for (int i = 0; i < n_ref; ++i)
    CMAX_REQUIRE(r[i].mode >= 0 && r[i].mode <= 2, "cmax_time_params: ref %d has mode %d (0 first, 1 last, 2 fraction)", i, r[i].mode);
  time_params_kernel<<<1, 1, 0, as_stream(stream)>>>(d_minmax, d_params, n_ref, n_bins, normalize_t ? 1 : 0, r[0], r[1], r[2], r[3]);
  CMAX_CUDA_CHECK(cudaGetLastError());
  return CMAX_OK;
}

size_t cmax_plan_workspace_bytes(int64_t n, int H, int W, int order) {
  if (n < 0 || n >= ((int64_t)1 << 31) || H <= 0 || W <= 0 || order < CMAX_ORDER_ASIS || order > CMAX_ORDER_PIXEL) {
    set_error("cmax_plan_workspace_bytes: bad argument (n=%lld, %dx%d, order %d)", (long long)n, H, W, order);
    return 0;
  }
  PlanLayout L;
  if (!plan_layout(n, H, W, order, &L)) return 0;
  return L.total;
}

int cmax_plan_create(cmax_plan_t** plan, const float* events, int64_t n, int ev_stride, int H, int W, int pad_h,
                     int pad_w, float t_min, float t_max, int order, void* workspace, size_t workspace_bytes,
                     cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_create: plan is NULL");
  *plan = nullptr;
  CMAX_REQUIRE(n >= 0 && n < (int64_t)1 << 31, "cmax_plan_create: n must be in [0, 2^31), got %lld", (long long)n);
  CMAX_REQUIRE(H > 0 && W > 0 && pad_h >= 0 && pad_w >= 0, "cmax_plan_create: bad image size %dx%d pad %d,%d", H, W, pad_h, pad_w);
  CMAX_REQUIRE((int64_t)(H + 2 * pad_h + 1) * (W + 2 * pad_w + 1) < (int64_t)1 << 28, "cmax_plan_create: image too large");
  CMAX_REQUIRE(H + 2 * pad_h < (1 << 22) && W + 2 * pad_w < (1 << 22), "cmax_plan_create: image sides must be < 2^22 (exact-floor range)");
  CMAX_REQUIRE(ev_stride >= 3, "cmax_plan_create: ev_stride must be >= 3");
  CMAX_REQUIRE(n == 0 || events != nullptr, "cmax_plan_create: events is NULL");
  CMAX_REQUIRE(order >= CMAX_ORDER_ASIS && order <= CMAX_ORDER_PIXEL, "cmax_plan_create: unknown event order %d", order);
  const bool sort = order != CMAX_ORDER_ASIS;
  CMAX_REQUIRE(sort || ev_stride == 4, "cmax_plan_create: un-sorted plans borrow the caller's array and need ev_stride == 4");
  CMAX_REQUIRE(sort || (reinterpret_cast<uintptr_t>(events) & 15) == 0, "cmax_plan_create: events must be 16-byte aligned");
  PlanLayout L;
  if (!plan_layout(n, H, W, order, &L)) return CMAX_ERR_CUDA;
  if (workspace == nullptr || workspace_bytes < L.total) {
    set_error("cmax_plan_create: workspace of %zu bytes needed, %zu given", L.total, workspace_bytes);
    return CMAX_ERR_WORKSPACE;
  }
  CMAX_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "cmax_plan_create: workspace must be 256-byte aligned");
  cudaStream_t s = as_stream(stream);
  char* ws = static_cast<char*>(workspace);

  cmax_plan* p = new (std::nothrow) cmax_plan();
  CMAX_REQUIRE(p != nullptr, "cmax_plan_create: out of host memory");
  memset(p, 0, sizeof(*p));
  p->n = n; p->H = H; p->W = W; p->pad_h = pad_h; p->pad_w = pad_w;
  p->Hp = H + 2 * pad_h; p->Wp = W + 2 * pad_w;
  p->tiles_x = L.tiles_x; p->tiles_y = L.tiles_y; p->n_tiles = L.n_tiles;
  p->d_params = reinterpret_cast<cmax_time_params_t*>(ws + L.off_params);
  p->d_minmax = reinterpret_cast<float*>(ws + L.off_minmax);
  p->d_status = reinterpret_cast<int32_t*>(ws + L.off_status);
  p->events = events;
  p->packed = ws + L.off_packed;
  p->order = CMAX_ORDER_ASIS;
  p->stage_mask = 7;
  p->vote_variant = 2;
  p->grad_variant = 2;

  // Everything below is ENQUEUED first and looked at once: one stream synchronisation per plan.  The kernels behind the
  // validation are safe on invalid input (an invalid event sorts under key 0), so the host can afford to learn about it last.
  int rc = CMAX_OK;
  float h_mm[2] = {t_min, t_max};
  int32_t h_status[4] = {0, 0, 0, 0};  // [status bits, number of strips, max(source row + 1), max(H - source row)]
  const bool try_strips = sort && n > 0 && order == CMAX_ORDER_PIXEL && H < (1 << 13) && W < (1 << 13);
  uint32_t *kfirst = nullptr, *kstrip0 = nullptr;
  const uint32_t* sorted_keys = nullptr;
#define PLAN_CHECK(call)                                      \
  do {                                                        \
    cudaError_t e__ = (call);                                 \
    if (e__ != cudaSuccess) { rc = cuda_fail(e__, #call); goto fail; } \
  } while (0)

  PLAN_CHECK(cudaMemsetAsync(p->d_status, 0, 4 * sizeof(int32_t), s));
  {
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(num_sms() * 8, (n + 255) / 256));
    cub::DoubleBuffer<uint32_t> dk(reinterpret_cast<uint32_t*>(ws + L.off_keys[0]), reinterpret_cast<uint32_t*>(ws + L.off_keys[1]));
    cub::DoubleBuffer<uint32_t> dv(reinterpret_cast<uint32_t*>(ws + L.off_idx[0]), reinterpret_cast<uint32_t*>(ws + L.off_idx[1]));
    if (n > 0)
      keys_kernel<<<grid, 256, 0, s>>>(events, n, ev_stride, H, W, L.tiles_x, order == CMAX_ORDER_PIXEL, sort ? dk.Current() : nullptr,
                                       sort ? dv.Current() : nullptr, p->d_status);
    if (t_min != t_min || t_max != t_max) {  // NaN -> compute from this batch
      rc = cmax_time_range(events, n, ev_stride, p->d_minmax, stream);
      if (rc != CMAX_OK) goto fail;
      PLAN_CHECK(cudaMemcpyAsync(h_mm, p->d_minmax, sizeof(h_mm), cudaMemcpyDeviceToHost, s));
    } else {
      PLAN_CHECK(cudaMemcpyAsync(p->d_minmax, h_mm, sizeof(h_mm), cudaMemcpyHostToDevice, s));
    }
    if (sort && n > 0) {
      float4* sorted = reinterpret_cast<float4*>(ws + L.off_events);
      size_t temp = L.temp_bytes + 256;
      PLAN_CHECK(cub::DeviceRadixSort::SortPairs(ws + L.off_temp, temp, dk, dv, (int)n, 0, L.key_bits, s));  // stable
      gather_events_kernel<<<grid, 256, 0, s>>>(events, n, ev_stride, dv.Current(), sorted);
      PLAN_CHECK(cudaGetLastError());
      p->events = reinterpret_cast<const float*>(sorted);
      p->order = order;
      sorted_keys = dk.Current();
      if (try_strips) {
        uint32_t* kstr = reinterpret_cast<uint32_t*>(ws + L.off_kstr);
        kfirst = reinterpret_cast<uint32_t*>(ws + L.off_kfirst);
        kstrip0 = reinterpret_cast<uint32_t*>(ws + L.off_kstrip0);
        key_first_kernel<<<grid, 256, 0, s>>>(sorted_keys, n, L.n_keys, kfirst);
        strip_counts_kernel<<<(int)std::min<int64_t>(num_sms() * 4, (L.n_keys + 256) / 256), 256, 0, s>>>(kfirst, L.n_keys, kstr);
        size_t st = L.scan_temp_bytes + 256;
        PLAN_CHECK(cub::DeviceScan::ExclusiveSum(ws + L.off_scan_temp, st, kstr, kstrip0, (int)(L.n_keys + 1), s));
        PLAN_CHECK(cudaMemcpyAsync(p->d_status + 1, kstrip0 + L.n_keys, sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
      }
    }
  }
  PLAN_CHECK(cudaMemcpyAsync(h_status, p->d_status, sizeof(h_status), cudaMemcpyDeviceToHost, s));
  PLAN_CHECK(cudaStreamSynchronize(s));
  if (h_status[0] & 1) {
    set_error("cmax_plan_create: an event's pixel (x=row, y=col) lies outside the %dx%d image (or is NaN); "
              "the reference's torch.gather raises here (src/warp.py:305-307)", H, W);
    rc = CMAX_ERR_SOURCE_OOB;
    goto fail;
  }
  p->compact_ok = (!(h_status[0] & 2) && H <= 65535 && W <= 65535) ? 1 : 0;
  p->compact = p->compact_ok;
  p->t_min = h_mm[0];
  p->t_max = h_mm[1];
  p->src_row_lo = h_status[2] > 0 ? H - h_status[3] : H;  // rows that hold a source pixel of this batch (empty: lo > hi)
  p->src_row_hi = h_status[2] - 1;
  // strips: integer coordinates, and dense enough that padding every pixel's run to whole strips stays below 50 %
  if (try_strips && p->compact_ok && (int64_t)(uint32_t)h_status[1] <= L.strip_capacity - 32) {
    p->strips = ws + L.off_strips;
    p->n_strips = (int64_t)(uint32_t)h_status[1];
    p->sorted_keys = sorted_keys;
    p->key_first = kfirst;
    p->key_strip0 = kstrip0;
    p->vote_variant = 5;
    p->grad_variant = 5;
  }
#undef PLAN_CHECK
  {
    const cmax_ref first = {0, 0.0f};
    rc = cmax_plan_set_refs(p, &first, 1, 0, stream);
    if (rc != CMAX_OK) goto fail;
  }
  *plan = p;
  return CMAX_OK;
fail:
  delete p;
  return rc;
}

void cmax_plan_destroy(cmax_plan_t* plan) { delete plan; }

int cmax_plan_info(const cmax_plan_t* plan, float* h_tmin, float* h_tmax, int64_t* h_n, int32_t* h_order) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_info: plan is NULL");
  if (h_tmin) *h_tmin = plan->t_min;
  if (h_tmax) *h_tmax = plan->t_max;
  if (h_n) *h_n = plan->n;
  if (h_order) *h_order = plan->order;
  return CMAX_OK;
}

int cmax_plan_strips(const cmax_plan_t* plan, int64_t* h_n_strips) {
  CMAX_REQUIRE(plan != nullptr && h_n_strips != nullptr, "cmax_plan_strips: NULL argument");
  *h_n_strips = plan->strips != nullptr ? plan->n_strips : 0;
  return CMAX_OK;
}

int cmax_plan_set_tile_flow(cmax_plan_t* plan, int hp, int wp, int pad_h, int pad_w, int sh, int sw, float t_scale) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_set_tile_flow: plan is NULL");
  CMAX_REQUIRE(hp >= 1 && wp >= 1 && (int64_t)hp * wp <= 1024, "cmax_plan_set_tile_flow: the patch grid must have between 1 and 1024 nodes (got %dx%d)", hp, wp);
  TileGeom g;
  const int rc = make_tile_geom("cmax_plan_set_tile_flow", hp, wp, pad_h, pad_w, sh, sw, plan->H, plan->W, &g);
  if (rc) return rc;
  plan->tile = g;
  plan->t_scale = t_scale;
  return CMAX_OK;
}

int cmax_plan_set_variant(cmax_plan_t* plan, int vote_variant, int grad_variant) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_set_variant: plan is NULL");
  CMAX_REQUIRE(vote_variant >= 0 && vote_variant <= 5, "cmax_plan_set_variant: vote_variant must be in [0,5]");
  CMAX_REQUIRE(grad_variant >= 0 && grad_variant <= 5, "cmax_plan_set_variant: grad_variant must be in [0,5]");
  plan->vote_variant = vote_variant;
  plan->grad_variant = grad_variant;
  return CMAX_OK;
}

int cmax_plan_set_compact(cmax_plan_t* plan, int enable, int32_t* h_compact, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_set_compact: plan is NULL");
  (void)stream;
  const int want = (enable && plan->compact_ok) ? 1 : 0;
  if (want != plan->compact) {
    plan->compact = want;
    plan->packed_valid = 0;  // re-packed on demand (ensure_packed)
  }
  if (h_compact) *h_compact = plan->compact;
  return CMAX_OK;
}

int cmax_plan_set_stage_mask(cmax_plan_t* plan, int mask) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_set_stage_mask: plan is NULL");
  CMAX_REQUIRE(mask >= 1 && mask <= 7, "cmax_plan_set_stage_mask: mask must be in [1,7]");
#ifndef CMAX_MEASURE
  CMAX_REQUIRE(mask == 7, "cmax_plan_set_stage_mask: this is a release build; partial stage masks (timing-only results) need a library "
                          "built with -DCMAX_MEASURE");
#endif
  plan->stage_mask = mask;
  return CMAX_OK;
}

int cmax_plan_set_refs(cmax_plan_t* plan, const cmax_ref* h_refs, int n_ref, int n_bins, cmax_stream_t stream) {
  CMAX_REQUIRE(plan != nullptr, "cmax_plan_set_refs: plan is NULL");
  const int rc = cmax_time_params(plan->d_minmax, h_refs, n_ref, n_bins, 1, plan->d_params, stream);
  if (rc != CMAX_OK) return rc;
  plan->n_ref = n_ref;
  plan->n_bins = n_bins;
  plan->packed_has_dt = (n_ref == 1) ? 1 : 0;
  plan->packed_valid = 0;
  if (plan->n > 0 && plan->strips != nullptr) {
    const int grid = (int)std::min<int64_t>(num_sms() * 8, (plan->n + 255) / 256);
    const int64_t tiles = (plan->n_strips + 31) / 32;
    plan->strip_tile_bytes = kStripTileBytes + (n_bins > 0 ? n_ref * kStripBinBytes : 0);
    CMAX_CUDA_CHECK(cudaMemsetAsync(plan->strips, 0, (size_t)tiles * plan->strip_tile_bytes, as_stream(stream)));
    pack_strips_kernel<<<grid, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(plan->events), plan->sorted_keys, plan->n,
                                                             plan->key_first, plan->key_strip0, plan->H, plan->W, plan->d_params,
                                                             plan->packed_has_dt, plan->strip_tile_bytes,
                                                             static_cast<unsigned char*>(plan->strips));
    CMAX_CUDA_CHECK(cudaGetLastError());
  }
  return CMAX_OK;
}

}  // extern "C"
