/*
 * cmax_b200.h -- C ABI of the B200-native contrast-maximization inner loop.
 *
 * The reference (tub-rip/event_based_optical_flow) is pure Python: it has NO FFI for this path.  The seams a
 * replacement plugs into are the duck-typed Python objects the solver holds (src/solver/base.py:139-147,
 * :185-204) and the composing method src/solver/patch_contrast_base.py:273-352.  This header is therefore the
 * boundary the reference's maintainers would bind (ctypes): every entry point names the
 * reference function (file:line under /root/reference) whose arithmetic it replaces.
 *
 * Conventions
 *   - All pointers are DEVICE pointers (fp32 unless stated) unless the name starts with `h_`.
 *   - An event is 4 consecutive floats (x, y, t, p); x = ROW (height), y = COLUMN (width)
 *     (src/utils/event_utils.py:38, src/event_image_converter.py:344-345).  `ev_stride` = floats per event (>=3).
 *   - Flow is [2,H,W] (channel 0 = row component), a flow voxel is [T,2,H,W]; flat pixel = x*W + y (src/warp.py:305).
 *   - Images are [Hp,Wp] row-major with Hp = H + 2*pad_h, Wp = W + 2*pad_w (src/event_image_converter.py:23-28).
 *   - Every function returns 0 on success, a cmax_status otherwise, and never throws or exits;
 *     cmax_last_error() gives the message for the calling thread.
 *   - Kernels are enqueued on `stream` (a cudaStream_t passed as void*); nothing synchronises unless stated.
 *   - Buffers are BORROWED: the library never frees caller memory and allocates no device memory of its own;
 *     scratch comes from the caller through the *_workspace_bytes()/workspace arguments.
 */
#ifndef CMAX_B200_H
#define CMAX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMAX_ABI_VERSION 2
#define CMAX_MAX_REFS 4   /* reference times fused into one event pass (first / middle / last / one more) */
#define CMAX_MAX_BINS 64  /* time bins of a flow voxel */
#define CMAX_MAX_PEERS 8  /* GPUs of one NVLink domain whose partial images one kernel sums */

typedef void* cmax_stream_t;

typedef enum {
  CMAX_OK = 0,
  CMAX_ERR_ARG = 1,        /* bad argument (null pointer, size, enum) -> Python ValueError */
  CMAX_ERR_CUDA = 2,       /* a CUDA runtime call failed; message holds cudaGetErrorString */
  CMAX_ERR_SOURCE_OOB = 3, /* an event's un-warped pixel lies outside [0,H)x[0,W): the reference's torch.gather raises (src/warp.py:305-307) */
  CMAX_ERR_WORKSPACE = 4   /* caller workspace too small */
} cmax_status;

typedef enum {
  CMAX_MOTION_DENSE = 0, /* "dense-flow"        src/warp.py:263-313 */
  CMAX_MOTION_VOXEL = 1, /* "dense-flow-voxel"  src/warp.py:315-365 */
  CMAX_MOTION_2DOF = 2,  /* "2d-translation" / "rigid-optical-flow"  src/warp.py:483-522 */
  CMAX_MOTION_TILE = 3   /* fused path only: the motion is the patch grid [2,hp,wp] the solver optimises; the event kernels
                            evaluate the tile-flow -> dense-flow map (src/solver/patch_contrast_base.py:462-506, times t_scale)
                            at every source pixel themselves and return dL/d(patch grid).  See cmax_plan_set_tile_flow. */
} cmax_motion;

typedef enum {
  CMAX_VOTE_BILINEAR = 0, /* src/event_image_converter.py:316-374 */
  CMAX_VOTE_COUNT = 1     /* src/event_image_converter.py:209-255 */
} cmax_vote_method;

typedef enum {
  CMAX_STAT_VARIANCE = 0, /* unbiased variance          src/costs/image_variance.py:47-58 */
  CMAX_STAT_GRADMAG = 1   /* mean (Sobel/8)^2 magnitude  src/costs/gradient_magnitude.py:60-76 */
} cmax_stat;

typedef enum {
  CMAX_COST_PLAIN = 0,      /* -stat(iwe)                               image_variance.py:56-58, gradient_magnitude.py:73-76 */
  CMAX_COST_NORMALIZED = 1, /* stat(orig)/stat(iwe)                     normalized_*.py:59-64 / 74-79 */
  CMAX_COST_MULTIFOCAL = 2  /* sum_r w_r stat(orig)/stat(iwe_r)         multi_focal_normalized_*.py:73-101 */
} cmax_cost_form;

/* A reference time: ref = t_min + fraction*(t_max - t_min) in fp32; mode 0 = "first" (ref = t_min exactly),
 * 1 = "last" (ref = t_max exactly), 2 = fraction ("middle" = 0.5, "before" = -1, "after" = 2).  src/warp.py:201-233 */
typedef struct {
  int32_t mode;
  float fraction;
} cmax_ref;

/* Device-resident time parameters for up to CMAX_MAX_REFS reference times (written by cmax_time_params). */
typedef struct {
  float ref[CMAX_MAX_REFS];    /* reference time                               src/warp.py:217-224 */
  float period[CMAX_MAX_REFS]; /* max(dt) - min(dt) of the batch, fp32          src/warp.py:256-257 */
  float dt_min[CMAX_MAX_REFS]; /* normalised dt of the earliest event                                   */
  float dt_max[CMAX_MAX_REFS];
  float edges[CMAX_MAX_REFS][CMAX_MAX_BINS + 1]; /* fp32-rounded float64 bin edges, last = dt_max+1000  src/warp.py:342-345 */
  int32_t n_ref;
  int32_t n_bins;
  int32_t normalize_t;
  int32_t pad_;
} cmax_time_params_t;

typedef struct cmax_plan cmax_plan_t; /* host-side handle; see cmax_plan_create */

/* ------------------------------------------------------------------ library */
int cmax_abi_version(void);
const char* cmax_last_error(void);
/* Name of the first CUDA kernel image compiled into the library ("sm_100a"); lets a loader assert the build arch. */
const char* cmax_build_arch(void);

/* ------------------------------------------------------------------ time  (src/warp.py:201-259, 342-345) */
/* min/max of column 2 of `events` into d_minmax[0..1].  Replaces the 4 full reductions per warp call. */
int cmax_time_range(const float* events, int64_t n, int ev_stride, float* d_minmax, cmax_stream_t stream);
/* (t_min,t_max) on device -> cmax_time_params_t on device, bit-exact with the reference's fp32/fp64 host arithmetic. */
int cmax_time_params(const float* d_minmax, const cmax_ref* h_refs, int n_ref, int n_bins, int normalize_t,
                     cmax_time_params_t* d_params, cmax_stream_t stream);

/* ------------------------------------------------------------------ modular operators */
/* Warp.warp_event (src/warp.py:156-199): out[n,ev_stride] = (x', y', dt, p) for reference time `ref_index`.
 * motion: [2,H,W] | [n_bins,2,H,W] | [2].  d_status (int32, device, may be NULL) is set non-zero on a source pixel
 * outside the image (such events are passed through un-warped). */
int cmax_warp_events(const float* events, int64_t n, int ev_stride, int H, int W, int motion_model,
                     const float* motion, const cmax_time_params_t* d_params, int ref_index, float* out,
                     int32_t* d_status, cmax_stream_t stream);
/* Adjoint of cmax_warp_events w.r.t. motion: grad_out[n,ev_stride] (columns 0,1 used) -> grad_motion (same shape as
 * motion, ZEROED BY THE CALLEE, accumulated with atomics; n_bins = voxel depth, ignored for the other models).
 * (autograd of src/warp.py:306-307 / 352-357 / 507-508) */
int cmax_warp_events_backward(const float* events, int64_t n, int ev_stride, int H, int W, int motion_model,
                              int n_bins, const cmax_time_params_t* d_params, int ref_index, const float* grad_out,
                              float* grad_motion, cmax_stream_t stream);
/* EventImageConverter.bilinear_vote_tensor / count_event_tensor (src/event_image_converter.py:316-374, 209-255).
 * xy: [n,xy_stride] (columns 0,1 used); weight: [n] or NULL (=1); image [Hp,Wp] is ZEROED BY THE CALLEE. */
int cmax_vote(const float* xy, int64_t n, int xy_stride, const float* weight, int Hp, int Wp, int pad_h, int pad_w,
              int vote, float* image, cmax_stream_t stream);
/* Adjoint of the bilinear vote: grad_image [Hp,Wp] -> grad_xy [n,2] (and grad_weight [n] if non-NULL). */
int cmax_vote_backward(const float* xy, int64_t n, int xy_stride, const float* weight, int Hp, int Wp, int pad_h,
                       int pad_w, const float* grad_image, float* grad_xy, float* grad_weight, cmax_stream_t stream);
/* ---- second order (Hessian-vector products for Newton-CG / trust-*: scipy_autograd/torch_wrapper.py:51-73 runs
 * torch.autograd.functional.vhp through warp -> vote -> cost).
 * cmax_vote_backward2: the adjoint of cmax_vote_backward as a function of (xy, grad_image).  u [n,u_stride] (columns 0,1)
 * is the cotangent of grad_xy.  out_grad_image [Hp,Wp] (ZEROED BY THE CALLEE; may be NULL) = d<grad_xy,u>/d grad_image, a
 * vote with the weight derivatives; out_grad_xy [n,2] (may be NULL) = d<grad_xy,u>/d xy = weight * d_r * (u_y, u_x), the
 * only non-zero second derivative of a bilinear weight being the mixed one. */
int cmax_vote_backward2(const float* xy, int64_t n, int xy_stride, const float* weight, int Hp, int Wp, int pad_h, int pad_w,
                        const float* grad_image, const float* u, int u_stride, float* out_grad_image, float* out_grad_xy,
                        cmax_stream_t stream);
/* Tangent of cmax_warp_events w.r.t. the motion = the adjoint of cmax_warp_events_backward w.r.t. grad_out:
 * out [n,2] = d(x',y')/d motion . tangent_motion (same shape as the motion). */
int cmax_warp_events_tangent(const float* events, int64_t n, int ev_stride, int H, int W, int motion_model, const cmax_time_params_t* d_params,
                             int ref_index, const float* tangent_motion, float* out, cmax_stream_t stream);
/* 3x3 Gaussian, reflect padding (torchvision gaussian_blur(kernel_size=3) at src/event_image_converter.py:153-158)
 * and its transpose.  n_img images of [Hp,Wp]; in != out. */
int cmax_blur3(const float* in, float* out, int n_img, int Hp, int Wp, float sigma, int transpose, cmax_stream_t stream);
/* Contrast statistics of n_img images: d_stats[i*4 + {0,1,2,3}] (float64, device) = {value, mean, M, unused}.
 * value = unbiased variance of the crop (VARIANCE) or mean (gx^2+gy^2) (GRADMAG); if grad != NULL also writes
 * d value / d image [n_img,Hp,Wp].  omit_boundary crops [1:-1,1:-1] (src/costs/image_variance.py:37-38,
 * src/costs/gradient_magnitude.py:67-72).  workspace: cmax_stats_workspace_bytes(n_img,Hp,Wp) bytes. */
size_t cmax_stats_workspace_bytes(int n_img, int Hp, int Wp);
int cmax_image_stats(const float* images, int n_img, int Hp, int Wp, int stat, int omit_boundary, double* d_stats,
                     float* grad, void* workspace, cmax_stream_t stream);

/* ------------------------------------------------------------------ tile flow   */
/* PatchContrastMaximization.interpolate_dense_flow_from_patch_tensor (src/solver/patch_contrast_base.py:462-506):
 * motion [2,hp,wp] (patch grid) -> dense [2,H,W] = crop(resize_bilinear(replicate_pad(-motion, pad_h, pad_w), x(sh,sw))),
 * align_corners = false, central crop (offsets full/2 - H/2).  The backward is the exact adjoint (a gather per grid
 * node, no atomics): grad_dense [2,H,W] -> grad_motion [2,hp,wp]. */
int cmax_tile_flow_upsample(const float* motion, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W, float* dense,
                            cmax_stream_t stream);
int cmax_tile_flow_upsample_backward(const float* grad_dense, int hp, int wp, int pad_h, int pad_w, int sh, int sw, int H, int W,
                                     float* grad_motion, cmax_stream_t stream);

/* ------------------------------------------------------------------ per-patch initialiser   */
/* K translation candidates for each of P patches in one call -- replaces P*K calls of
 * PyramidalPatchContrastMaximization.calculate_cost_for_small_patch(events, theta, "2d-translation")
 * (src/solver/patch_contrast_pyramid.py:379-415, driven per patch and per TPE trial by :320-377), the reference's numpy path:
 * 2-dof warp to the middle of the patch's time span (src/warp.py:483-522), numpy bilinear vote
 * (src/event_image_converter.py:257-312), scipy.ndimage.gaussian_filter(sigma) ('reflect' border, radius int(4 sigma + 0.5)),
 * cv2.Sobel(ksize 3) / 8 (BORDER_REFLECT_101), mean(gx^2 + gy^2) over the whole (padded) patch image
 * (src/costs/gradient_magnitude.py:78-95, omit_boundary = False).
 *   patch_events [m,3] f32   the events of all patches, grouped by patch: (x - x_min, y - y_min, dt) with dt already
 *                            (t - t_ref) / period of the PATCH's events (the host side prepares it once per frame and level)
 *   patch_offsets [P+1] i64  device array; patch p owns rows [offsets[p], offsets[p+1]); max_patch_events = the largest count
 *   candidates [P,K,2] f64   (trans_x, trans_y) as the sampler suggests them
 *   theta_scale [P] f64      the patch's time span the reference multiplies a candidate by (pyramid.py:366-371); NULL = 1
 *   orig_energy [P] f64      NULL: out [P,K] = mean squared gradient magnitude of the blurred image of warped events;
 *                            else: out = orig_energy[p] / that (the reference's loss,
 *                            src/costs/normalized_gradient_magnitude.py:90-94 with direction 'minimize'; NaN -> 0 as
 *                            pyramid.py:374-375).  orig_energy is this entry point's own output for a zero candidate.
 *   flags                    CMAX_PATCH_GLOBAL_IMAGES: never use the shared-memory kernel (it serves patch images of up to
 *                            25 600 padded pixels); CMAX_PATCH_KEEP_IMAGES: leave the blurred images [P,K,Hp,Wp] f32 at the
 *                            start of the workspace (always the case with global images).
 * workspace: cmax_patch_candidates_workspace_bytes() bytes; may be NULL when the shared-memory kernel runs without KEEP_IMAGES. */
enum { CMAX_PATCH_GLOBAL_IMAGES = 1, CMAX_PATCH_KEEP_IMAGES = 2 };
size_t cmax_patch_candidates_workspace_bytes(int n_patches, int n_candidates, int h, int w, int pad_h, int pad_w);
int cmax_patch_candidates(const float* patch_events, const int64_t* patch_offsets, int64_t max_patch_events, int n_patches,
                          const double* candidates, const double* theta_scale, int n_candidates, int h, int w, int pad_h, int pad_w,
                          float sigma, const double* orig_energy, int flags, void* workspace, size_t workspace_bytes, double* out,
                          cmax_stream_t stream);

/* ------------------------------------------------------------------ time-aware flow voxel   */
typedef enum {
  CMAX_SCHEME_UPWIND = 0,  /* "upwind"   src/utils/flow_utils.py:439-493 */
  CMAX_SCHEME_BURGERS = 1  /* "burgers"  src/utils/flow_utils.py:567-639 */
} cmax_voxel_scheme;
/* construct_dense_flow_voxel_torch (src/utils/flow_utils.py:99-161): dense [2,H,W] (the flow at t0) -> voxel
 * [time_bin,2,H,W] by explicit upwind / inviscid-Burgers steps of dt = 1/time_bin, forward in time from level t0 up and
 * backward (sign-flipped flow) from t0 down; t0 = time_bin/2 when t0_middle, else 0; 1 <= time_bin <= 4096 (the reference's
 * own tests use 60 and 100 levels; only the voxel WARP is limited to CMAX_MAX_BINS).  fp32, bit-identical to the
 * reference's torch fp32 result.  One launch per 8 levels (both sides of t0 in the same launch).  The backward is the
 * exact adjoint (torch's even split of d max(x,0)/dx at x == 0 included), gather form, no atomics: grad_voxel
 * [time_bin,2,H,W] -> grad_dense [2,H,W].  `dense` and `voxel` are the forward's input and output.  workspace:
 * cmax_flow_voxel_workspace_bytes(H,W) bytes (only touched when one side of t0 has more than 8 levels; may be NULL
 * otherwise). */
size_t cmax_flow_voxel_workspace_bytes(int H, int W);
int cmax_flow_voxel(const float* dense, int H, int W, int time_bin, int scheme, int t0_middle, float* voxel, cmax_stream_t stream);
int cmax_flow_voxel_backward(const float* dense, const float* voxel, const float* grad_voxel, int H, int W, int time_bin, int scheme,
                             int t0_middle, float* grad_dense, void* workspace, cmax_stream_t stream);

/* ------------------------------------------------------------------ fused hot path */
/* Event order inside a plan.  The order never changes results beyond fp32 summation order; it changes locality. */
typedef enum {
  CMAX_ORDER_ASIS = 0,  /* borrow the caller's float4 array (time order)                                       */
  CMAX_ORDER_TILE = 1,  /* stable sort by 32x32 source tile: flow / gradient accesses of a CTA stay in one tile */
  CMAX_ORDER_PIXEL = 2  /* stable sort by source pixel (tile-major): runs of events share one flow vector       */
} cmax_order;

/* Resident, pre-processed event batch (events are constant over one solver.optimize()).  Validates source pixels,
 * takes (t_min,t_max) (pass NaN to compute from this batch; multi-GPU callers pass the GLOBAL range), and -- for
 * order != ASIS -- stably re-orders the events into the workspace as float4 (radix sort by source pixel; the per-pixel
 * counts come from the sorted keys, no atomics).  Everything is enqueued first and `stream` is synchronised ONCE.
 * `events` must stay alive and unchanged while the plan is used when order == CMAX_ORDER_ASIS. */
size_t cmax_plan_workspace_bytes(int64_t n, int H, int W, int order);
int cmax_plan_create(cmax_plan_t** plan, const float* events, int64_t n, int ev_stride, int H, int W, int pad_h,
                     int pad_w, float t_min, float t_max, int order, void* workspace, size_t workspace_bytes,
                     cmax_stream_t stream);
void cmax_plan_destroy(cmax_plan_t* plan);
/* (t_min, t_max, n, order) the plan uses. */
int cmax_plan_info(const cmax_plan_t* plan, float* h_tmin, float* h_tmax, int64_t* h_n, int32_t* h_order);
/* Number of 8-event strips the plan cut the batch into (0 = the batch did not qualify: not pixel-ordered, fractional
 * coordinates, image side >= 8192, or so sparse that padding every pixel's run to whole strips would cost more than 50 %).
 * Variant 5 (the default when there are strips) runs the strip kernels, see cmax_plan_set_variant. */
int cmax_plan_strips(const cmax_plan_t* plan, int64_t* h_n_strips);
/* Select reference times / voxel bins for subsequent calls (enqueues one tiny kernel). */
int cmax_plan_set_refs(cmax_plan_t* plan, const cmax_ref* h_refs, int n_ref, int n_bins, cmax_stream_t stream);
/* CMAX_MOTION_TILE: geometry of the tile-flow map (arguments as in cmax_tile_flow_upsample: patch grid hp x wp <= 1024 nodes,
 * replicate padding, integer sliding window) and the factor t_scale the reference multiplies the up-sampled flow by
 * (src/solver/patch_contrast_pyramid.py:452-453).  The dense flow and its gradient are never materialised: no
 * [2,H,W] round trip through L2, the gradient that leaves the GPU (or crosses NVLink) is 2 * hp * wp floats.  Needs a plan
 * with strips (cmax_plan_strips > 0); other batches compose cmax_tile_flow_upsample with the dense model. */
int cmax_plan_set_tile_flow(cmax_plan_t* plan, int hp, int wp, int pad_h, int pad_w, int sh, int sw, float t_scale);
/* Kernel variants (all give the same results up to fp32 summation order; kept selectable so that profiles/ can show
 * each design choice measured against the others):
 *   vote_variant / grad_variant 5 (default when the plan has strips, else they fall back to 2) = the strip kernels: one
 *                  thread per 8-event strip of ONE source pixel (per-strip coordinates / flow / time bins, 4.5 B per event);
 *   vote_variant 2 = each thread walks a run of consecutive events and sums the weights of events that
 *                  fall into the same accumulator cell in registers: one red.v4 per cell change;
 *                3 = the same with the flow loads of a warp-tile batched (16 more registers);
 *                0 = one red.v4 per event into the per-corner accumulators; 1 = four scalar red.f32 per event
 *                  straight into the image (the textbook scatter).
 *   grad_variant 2 = run walk batched by 4 events (flow loads, then quad gathers, then accumulation), 3 = batched by 8,
 *                  4 = sequential run walk: corner gradients re-gathered only on a cell change, flow gradient summed in
 *                  registers per source-pixel run; 1 = per-event gather + warp-segmented shuffle reduction over runs
 *                  of equal source pixel (needs CMAX_ORDER_PIXEL, else falls back to 0); 0 = scalar red per event. */
int cmax_plan_set_variant(cmax_plan_t* plan, int vote_variant, int grad_variant);
/* Packed-event format.  When every event of the batch has integer pixel coordinates (what a sensor delivers; checked at
 * plan creation) the plan's private copy uses 8 bytes per event -- (dt|t, row<<16|col) -- instead of 16, halving the
 * event stream the kernels read; fractional coordinates (e.g. undistorted events) keep the 16-byte format.  Results are
 * identical.  enable = 0 forces the 16-byte format (measurement / tests); *h_compact (may be NULL) returns the format in
 * use afterwards. */
int cmax_plan_set_compact(cmax_plan_t* plan, int enable, int32_t* h_compact, cmax_stream_t stream);
/* Measurement aid, MEASUREMENT BUILDS ONLY (-DCMAX_MEASURE; the release library rejects every mask but 7 with
 * CMAX_ERR_ARG, so that no switch can silently change the results of a release build): which launches the stages enqueue.
 * Bit 0 = the memsets, bit 1 = the event kernels, bit 2 = the image-sized kernels. */
int cmax_plan_set_stage_mask(cmax_plan_t* plan, int mask);

/* What one CM evaluation computes: cost = form(stat(blur(iwe_r)))   src/costs/<cost>.py, src/solver/patch_contrast_base.py:289-352 */
typedef struct {
  int32_t stat;           /* cmax_stat */
  int32_t form;           /* cmax_cost_form */
  int32_t direction_sign; /* +1 minimize, -1 maximize                     src/costs/base.py:20-25 */
  int32_t omit_boundary;  /* crop [1:-1,1:-1]                              src/solver/patch_contrast_base.py:290 */
  float sigma;            /* 3x3 Gaussian blur of every IWE, 0 = none      src/event_image_converter.py:153-158 */
  float weights[CMAX_MAX_REFS]; /* multi-focal weights per reference time  multi_focal_normalized_*.py */
} cmax_cost_spec;

/* Scratch one evaluation needs (per-corner accumulators, IWEs, gradient pictures, statistics). */
size_t cmax_objective_workspace_bytes(const cmax_plan_t* plan, const cmax_cost_spec* spec);
/* Must be called once on a freshly allocated workspace (and again after any failed call): the per-corner accumulators
 * are kept clean BETWEEN evaluations by the kernels themselves (the fold zeroes what it reads), so no evaluation ever
 * enqueues a memset for them. */
int cmax_objective_workspace_init(const cmax_plan_t* plan, void* workspace, cmax_stream_t stream);

/* One CM iteration on one GPU = THREE launches: K1 (event pass: warp + vote), the image kernel (fold + statistics + scalar
 * cost + per-corner gradient quads, grid-wide barriers between its phases), K3 (event pass: re-warp, gather, chain to
 * dL/dmotion).  cost -> d_cost[0] (float64, device); grad_motion (same shape as motion; NULL = value only) is written in
 * full.  d_orig_stat: device float64 statistic of the un-warped IWE (normalised / multi-focal forms), else NULL.
 * Gradient-magnitude and blurred costs run the operator kernels (blur, Sobel statistics, combine) between a fold-only and
 * a quads-only launch of the image kernel.
 * (src/warp.py:301-313|339-365|506-520 -> src/event_image_converter.py:316-374 -> src/costs/<cost>.py, composed as in
 * src/solver/patch_contrast_base.py:289-352; backward = autograd of all of it) */
int cmax_objective(const cmax_plan_t* plan, int motion_model, const float* motion, const cmax_cost_spec* spec,
                   const double* d_orig_stat, void* workspace, double* d_cost, float* grad_motion /* NULL = value only */,
                   cmax_stream_t stream);
/* The same evaluation stage by stage (what cmax_objective enqueues for a variance cost without blur is exactly
 * vote + cost(zero_grad = grad_motion) + grad(pre_zeroed = 1)); callers that all-reduce with NCCL put their collectives
 * between the stages: vote, fold, [all-reduce *iwe_out in place], cost, grad, [all-reduce the gradient].
 *   vote  K1: warp by `motion` and bilinear-vote into the per-corner accumulators of the n_ref images.
 *   fold  accumulators -> IWE stack [n_ref,Hp,Wp] fp32 inside the workspace (*iwe_out, may be NULL, receives its address).
 *   cost  statistics + scalar cost of the folded IWEs and -- when want_grad -- the gradient quads; zero_grad (may be NULL)
 *         = n_zero floats cleared on the way (the buffer `grad` accumulates into).
 *   grad  K3 into grad_motion; pre_zeroed != 0 promises the buffer was cleared by `cost`. */
int cmax_objective_vote(const cmax_plan_t* plan, int motion_model, const float* motion, void* workspace, cmax_stream_t stream);
int cmax_objective_fold(const cmax_plan_t* plan, void* workspace, float** iwe_out, cmax_stream_t stream);
int cmax_objective_cost(const cmax_plan_t* plan, const cmax_cost_spec* spec, const double* d_orig_stat, void* workspace, int want_grad,
                        double* d_cost, float* zero_grad, int64_t n_zero, cmax_stream_t stream);
int cmax_objective_grad(const cmax_plan_t* plan, int motion_model, const float* motion, void* workspace, float* grad_motion,
                        int pre_zeroed, cmax_stream_t stream);

/* ---- multi-GPU: events sharded over the ranks of one NVLink domain, one process per GPU ----
 * The objective workspace, the partial-gradient buffer and a flag block of every rank live in SYMMETRIC (peer-mapped)
 * memory.  cmax_objective_sharded is cmax_objective with the two sums of the sharded path fused into its kernels, no
 * collective launch and no barrier kernel:
 *   K1 -> image kernel: fold this rank's partial IWE; its last CTA raises this rank's flag on every peer (one fence + one
 *         posted store per peer); every CTA then waits for all ranks' flags and sums the partial IWEs of ALL ranks
 *         through NVLink peer loads, in rank order (bit-identical on every rank), accumulating the variance sums on the
 *         way; cost; gradient quads  -- the all-reduce of the IWE and the cost are ONE launch;
 *   K3 into this rank's partial gradient -> gradient exchange kernel: raise the gradient flag, wait, sum all ranks'
 *         partial gradients into grad_motion (value only: flags only, which keeps the ranks in step).
 * A flag is a 16-byte record {evaluation counter, row_lo, row_hi, -}: next to "complete" it says which image rows of the
 * partial result can be non-zero, and a reader skips the peers whose rows miss its pixels.  With spatially compact shards
 * (distributed.reshard_events_by_pixel: a contiguous slice of the pixel-ordered stream per rank) the exchanges then move the
 * overlaps only instead of N whole images per rank; with time-sliced shards the ranges are the whole image.
 * Flags are evaluation counters (monotonic), the two flag arrays alternate per evaluation, and that alternation is what
 * makes re-use of the partial buffers safe without any further synchronisation.  Every rank must call with the same
 * arguments in the same order; a rank that never arrives makes the waiting kernels trap after ~4 s. */
typedef struct {
  int32_t n_peers, rank;
  const float* iwe[CMAX_MAX_PEERS];  /* rank r's partial IWE stack = its workspace + cmax_objective_iwe_offset */
  const float* grad[CMAX_MAX_PEERS]; /* rank r's partial motion gradient (as many floats as the motion) */
  uint32_t* flags[CMAX_MAX_PEERS];   /* rank r's flag block: 256 bytes (2 x CMAX_MAX_PEERS records of 16 bytes), 16-byte aligned,
                                        zeroed once by the caller */
} cmax_peers;
size_t cmax_objective_iwe_offset(const cmax_plan_t* plan);
/* Measurement builds (-DCMAX_MEASURE) only: where in the workspace the image / exchange kernels leave their per-CTA
 * globaltimer stamps (scripts/image_probe.py, scripts/exchange_probe.py); the release library writes nothing there. */
size_t cmax_objective_probe_offset(const cmax_plan_t* plan);
size_t cmax_objective_full_iwe_offset(const cmax_plan_t* plan); /* where the summed IWE stack of the last evaluation sits */
int cmax_objective_sharded(const cmax_plan_t* plan, int motion_model, const float* motion, const cmax_cost_spec* spec,
                           const double* d_orig_stat, void* workspace, const cmax_peers* peers, double* d_cost,
                           float* grad_motion /* NULL = value only */, cmax_stream_t stream);

/* Scalar combination of per-image statistics (exposed for the modular cost plugins).
 * d_stats: n_ref x 4 doubles from cmax_image_stats; h_weights: n_ref multi-focal weights (NULL = 1).
 * explicit_grad != 0: the affine pair is (d cost/d stat_r, 0); otherwise (VARIANCE) it is
 * (d cost/d stat_r * 2/(M-1), mean_r).  Writes d_cost[0] (float64) and d_affine[2*n_ref] (fp32). */
int cmax_combine_cost(const double* d_stats, int n_ref, int stat, int cost_form, const double* d_orig_stat,
                      const float* h_weights, int direction_sign, int explicit_grad, double* d_cost, float* d_affine,
                      cmax_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CMAX_B200_H */
