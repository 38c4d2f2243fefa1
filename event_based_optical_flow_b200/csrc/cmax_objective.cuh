// Workspace layout of one objective evaluation and the seams between the translation units of the fused path:
//   cmax_fused.cu  the event kernels that read the packed copy (per-event baselines, run kernels) + dispatch
//   cmax_lean.cu   the strip kernels
//   cmax_mid.cu    everything image-sized between K1 and K3 (fold, statistics, cost, gradient quads, the cross-GPU
//                  exchange), the stages, and the C ABI of the fused path
#pragma once
#include "cmax_runs.cuh"
#include "cmax_stats.cuh"

namespace cmax {

struct ObjLayout {
  size_t off_acc, off_iwe, off_iwe_full, off_blur, off_statacc, off_stats, off_affine, off_misc, off_gxy, off_g, off_g2, off_gq, off_slots,
      off_sync, total;
  int64_t cells, HW;
};

constexpr int kMidThreads = 512;
constexpr int kMidMaxCtas = 512;  // statistics slots (one per CTA of the image kernel and reference time)

static inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

static inline ObjLayout obj_layout(int Hp, int Wp) {
  ObjLayout L;
  const int R = CMAX_MAX_REFS;
  L.cells = (int64_t)(Hp + 1) * (Wp + 1) + 1;  // + one cell no pixel reads: its gradient quad is all zero (the strip K3's 'outside' cell)
  L.HW = (int64_t)Hp * Wp;
  size_t off = 0;
  L.off_acc = off;     off = align256(off + (size_t)R * L.cells * sizeof(float4));
  L.off_iwe = off;     off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_iwe_full = off; off = align256(off + (size_t)R * L.HW * sizeof(float));  // sharded: the summed IWE (off_iwe stays the partial peers read)
  L.off_blur = off;    off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_stats = off;   off = align256(off + (size_t)R * 4 * sizeof(double));
  L.off_affine = off;  off = align256(off + (size_t)R * 2 * sizeof(float));
  L.off_misc = off;    off = align256(off + 2 * sizeof(double));  // 2-dof fp64 staging
  // [StatAcc block][per-CTA slots][Sobel pair] is the workspace cmax_image_stats carves for itself (cmax_cost.cu)
  L.off_statacc = off; off = align256(off + (size_t)R * sizeof(StatAcc));
  L.off_gxy = off;     off = align256(off + (size_t)R * 2 * kStatMaxCtas * sizeof(double) + (size_t)R * 2 * L.HW * sizeof(float));
  L.off_g = off;       off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_g2 = off;      off = align256(off + (size_t)R * L.HW * sizeof(float));
  L.off_gq = off;      off = align256(off + (size_t)R * L.cells * sizeof(float4));
  L.off_slots = off;   off = align256(off + (size_t)kMidMaxCtas * R * 2 * sizeof(double));
  L.off_sync = off;    off = align256(off + 256);  // [0] epoch of the cross-GPU exchange (persistent), see image_kernel
  L.total = off;
  return L;
}

struct Ws {
  float4* acc; float* iwe; float* iwe_full; float* blur; StatAcc* sacc; double* stats; float* affine; unsigned int* bar;
  float* G; float* G2; float4* gq; char* stats_ws; double* acc2; double* slots; uint32_t* epoch;
};

static inline Ws carve(void* workspace, const ObjLayout& L) {
  char* ws = static_cast<char*>(workspace);
  Ws w;
  w.acc = reinterpret_cast<float4*>(ws + L.off_acc);
  w.iwe = reinterpret_cast<float*>(ws + L.off_iwe);
  w.iwe_full = reinterpret_cast<float*>(ws + L.off_iwe_full);
  w.blur = reinterpret_cast<float*>(ws + L.off_blur);
  w.sacc = reinterpret_cast<StatAcc*>(ws + L.off_statacc);
  w.stats = reinterpret_cast<double*>(ws + L.off_stats);
  w.affine = reinterpret_cast<float*>(ws + L.off_affine);
  // the grid-barrier counters of the image kernel live in the last 64 bytes of the 256-byte StatAcc block, which K1's
  // first CTA clears on every evaluation (FusedArgs::zero256)
  w.bar = reinterpret_cast<unsigned int*>(ws + L.off_statacc + 192);
  w.G = reinterpret_cast<float*>(ws + L.off_g);
  w.G2 = reinterpret_cast<float*>(ws + L.off_g2);
  w.gq = reinterpret_cast<float4*>(ws + L.off_gq);
  w.stats_ws = ws + L.off_statacc;  // [StatAcc block][Sobel pair], the layout cmax_image_stats expects
  w.acc2 = reinterpret_cast<double*>(ws + L.off_misc);  // 2-dof fp64 staging
  w.slots = reinterpret_cast<double*>(ws + L.off_slots);
  w.epoch = reinterpret_cast<uint32_t*>(ws + L.off_sync);
  return w;
}
static_assert(sizeof(StatAcc) * CMAX_MAX_REFS <= 192, "StatAcc block and the barrier counters share 256 bytes");

FusedArgs fused_args(const cmax_plan* p, const float* motion);
// K1 / K3 of any variant (cmax_fused.cu); `iwe` is only written by vote variant 1 (scalar reds straight into the image)
void launch_vote_any(const cmax_plan* p, int motion_model, cudaStream_t s, const FusedArgs& a, float4* acc, float* iwe);
void launch_grad_any(const cmax_plan* p, int motion_model, cudaStream_t s, const FusedArgs& a, const float4* gq, float* target);

}  // namespace cmax
