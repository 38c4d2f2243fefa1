// Shared pieces of the event kernels (cmax_fused.cu, cmax_lean.cu): kernel arguments, the packed-event formats, the TMA
// tile pipeline, programmatic-dependent-launch helpers, persistent-grid sizing.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "cmax_plan.cuh"
#include "cmax_tile.cuh"

namespace cmax {

// ------------------------------------------------------------------------------------------------ per-event math
struct FusedArgs {
  const float4* ev;
  const void* packed;    // the plan's packed copy (16-byte or compact 8-byte events), read by the run kernels
  int compact;
  const void* strips;    // the plan's strips (cmax_plan.cuh), read by the strip kernels; NULL when the batch has none
  int64_t n_strips;
  int strip_tile_bytes;
  int64_t n;
  int H, W, Hp, Wp, pad_h, pad_w;
  const float* motion;
  const cmax_time_params_t* tp;
  int64_t cells;  // (Hp+1)*(Wp+1)
  unsigned int* zero256;  // 64 words K1's first CTA clears (statistics block + grid-barrier counters), or NULL
  int seg_reduce;         // strip K3: sum the flow gradient over the strips of one pixel inside a warp before the reductions (dense batches)
  TileGeom tile;          // CMAX_MOTION_TILE: `motion` is the patch grid [2,hp,wp], evaluated per source pixel by the kernels
  float t_scale;
};

// Time parameters one CTA needs, staged in shared memory once per CTA.
struct TimeSmem {
  float ref[CMAX_MAX_REFS], period[CMAX_MAX_REFS], dt_min[CMAX_MAX_REFS], inv_width[CMAX_MAX_REFS];
  float edges[CMAX_MAX_REFS][CMAX_MAX_BINS + 1];
  int n_bins, normalize_t;
};

template <int NREF, bool VOXEL>
__device__ __forceinline__ void stage_time(const cmax_time_params_t* __restrict__ tp, TimeSmem& s) {
  if (threadIdx.x < NREF) {
    const int r = threadIdx.x;
    s.ref[r] = tp->ref[r];
    s.period[r] = tp->period[r];
    s.dt_min[r] = tp->dt_min[r];
    s.inv_width[r] = (float)tp->n_bins / (tp->dt_max[r] - tp->dt_min[r]);
  }
  if (threadIdx.x == 0) {
    s.n_bins = tp->n_bins;
    s.normalize_t = tp->normalize_t;
  }
  if (VOXEL) {
    for (int k = threadIdx.x; k < NREF * (CMAX_MAX_BINS + 1); k += blockDim.x)
      s.edges[k / (CMAX_MAX_BINS + 1)][k % (CMAX_MAX_BINS + 1)] = tp->edges[k / (CMAX_MAX_BINS + 1)][k % (CMAX_MAX_BINS + 1)];
  }
  __syncthreads();
}

// Programmatic dependent launch (PDL): a kernel launched with the programmatic-serialization attribute may start while
// its predecessor in the stream is still running; it must not touch anything the predecessor produces (or still reads)
// before pdl_wait(), which returns once the predecessor grid has completed and its writes are visible.  A predecessor
// calls pdl_trigger() to allow the early start.  Both are no-ops for kernels launched normally.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ void red_add_v4(float4* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// predicated form: no branch, so the walk of consecutive events stays one basic block the scheduler can interleave
__device__ __forceinline__ void red_add_v4_if(bool pred, float4* addr, float a, float b, float c, float d) {
  asm volatile(
      "{\n .reg .pred p;\n setp.ne.u32 p, %5, 0;\n @p red.global.add.v4.f32 [%0], {%1, %2, %3, %4};\n}" ::"l"(addr), "f"(a), "f"(b),
      "f"(c), "f"(d), "r"((unsigned)pred)
      : "memory");
}

constexpr int kRunThreads = 128;
constexpr int kRunWarps = kRunThreads / 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!done);
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// The two packed-event formats (cmax_plan.cuh).  `key` identifies the source pixel for change detection; `src()` turns
// it into the flat pixel index (only needed when the key changes).
template <bool COMPACT>
struct PackedEv;
template <>
struct PackedEv<false> {  // (x, y, dt|t, bits(src)) : 16 bytes, any coordinates
  static constexpr uint32_t kTileBytes = kWarpTile * 16;
  __device__ static __forceinline__ void get(const void* buf, int k, int lane, float& x, float& y, float& tz, int& key) {
    const float4 e = reinterpret_cast<const float4*>(buf)[k * 32 + lane];
    x = e.x; y = e.y; tz = e.z; key = __float_as_int(e.w);
  }
  __device__ static __forceinline__ int key_of(const void* buf, int k, int lane) {
    return __float_as_int(reinterpret_cast<const float4*>(buf)[k * 32 + lane].w);
  }
  __device__ static __forceinline__ int src(int key, int W) { return key; }
};
template <>
struct PackedEv<true> {  // (dt|t, row<<16|col) : 8 bytes, integer pixel coordinates (what an event camera delivers)
  static constexpr uint32_t kTileBytes = kWarpTile * 8;
  __device__ static __forceinline__ void get(const void* buf, int k, int lane, float& x, float& y, float& tz, int& key) {
    const uint2 u = reinterpret_cast<const uint2*>(buf)[k * 32 + lane];
    tz = __uint_as_float(u.x);
    key = (int)u.y;
    // exact small-int -> float without the conversion pipe: 2^23 + v has v in its mantissa
    x = __fsub_rn(__uint_as_float(0x4B000000u | (u.y >> 16)), 8388608.0f);
    y = __fsub_rn(__uint_as_float(0x4B000000u | (u.y & 0xFFFFu)), 8388608.0f);
  }
  __device__ static __forceinline__ int key_of(const void* buf, int k, int lane) {
    return (int)reinterpret_cast<const uint2*>(buf)[k * 32 + lane].y;
  }
  __device__ static __forceinline__ int src(int key, int W) { return (int)((unsigned)key >> 16) * W + (key & 0xFFFF); }
};

template <uint32_t BYTES, int NS = 2>
struct alignas(128) TilePipe {  // one per warp: NS-deep landing zone of the TMA bulk copies
  unsigned char buf[NS][BYTES];
  uint64_t bar[NS];
};

template <uint32_t BYTES, int NS>
__device__ __forceinline__ void pipe_init(TilePipe<BYTES, NS>& p, int lane) {
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NS; ++i) mbar_init(&p.bar[i], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncwarp();
}
template <uint32_t BYTES, int NS>
__device__ __forceinline__ void pipe_issue(TilePipe<BYTES, NS>& p, int stage, const void* __restrict__ packed, int64_t tile, int lane) {
  if (lane == 0) {
    mbar_expect_tx(&p.bar[stage], BYTES);
    bulk_g2s(p.buf[stage], static_cast<const unsigned char*>(packed) + tile * BYTES, BYTES, &p.bar[stage]);
  }
}

// the same with the tile size taken at run time from a template-sized landing zone
template <uint32_t BYTES, int NS>
__device__ __forceinline__ void pipe_issue_n(TilePipe<BYTES, NS>& p, int stage, const void* __restrict__ packed, int64_t tile, uint32_t bytes, int lane) {
  if (lane == 0) {
    mbar_expect_tx(&p.bar[stage], bytes);
    bulk_g2s(p.buf[stage], static_cast<const unsigned char*>(packed) + tile * bytes, bytes, &p.bar[stage]);
  }
}

// Per-reference-time scalars kept in registers (dense / 2-dof); the voxel model also needs the bin edges (shared).
template <int NREF>
struct RefRegs {
  float ref[NREF], period[NREF];
};

template <int NREF>
__device__ __forceinline__ RefRegs<NREF> load_refs(const cmax_time_params_t* __restrict__ tp) {
  RefRegs<NREF> rr;
#pragma unroll
  for (int r = 0; r < NREF; ++r) {
    rr.ref[r] = __ldg(&tp->ref[r]);
    rr.period[r] = __ldg(&tp->period[r]);
  }
  return rr;
}

// accumulator cell of a vote, or -1 when the event touches no pixel
__device__ __forceinline__ int vote_cell(const Vote& v, int Hp, int Wp) {
  const bool inside = ((unsigned)(v.row + 1) <= (unsigned)Hp) & ((unsigned)(v.col + 1) <= (unsigned)Wp);
  return inside ? (v.row + 1) * (Wp + 1) + (v.col + 1) : -1;
}

// Launch with the programmatic-stream-serialization attribute (see pdl_wait); `pdl == false` is an ordinary launch.
static inline bool pdl_enabled() {
#ifdef CMAX_MEASURE  // measurement builds only: CMAX_PDL=0 turns programmatic dependent launch off
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("CMAX_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
#else
  return true;
#endif
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t s, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// Persistent-style grid for the run kernels: exactly the number of CTAs that are resident at once (occupancy x SMs),
// each warp striding over the warp-tiles; the shared-memory carveout is raised to the maximum first (the kernels want
// up to 6 CTAs x 33 KB per SM).  Cached per kernel instantiation (one process drives one GPU).
template <typename K>
static inline int run_grid(K kernel, int64_t n) {
  static int per_sm = 0;
  if (per_sm == 0) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kRunThreads, 0) != cudaSuccess || v < 1) v = 4;
    per_sm = v;
  }
  const int64_t ctas = ((n + kWarpTile - 1) / kWarpTile + kRunWarps - 1) / kRunWarps;
  return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)num_sms() * per_sm));
}


// (re)build the packed copy the run kernels read, if it is stale (cmax_events.cu)
int ensure_packed(const cmax_plan* plan, cudaStream_t s);

// strip kernels (cmax_lean.cu)
int strips_tile_bytes_for(int motion_model, int n_ref);
void launch_vote_strips(int motion_model, int n_ref, cudaStream_t s, const FusedArgs& a, float4* acc);
void launch_grad_strips(int motion_model, int n_ref, bool pdl, cudaStream_t s, const FusedArgs& a, const float4* gq, float* gmotion);

}  // namespace cmax
