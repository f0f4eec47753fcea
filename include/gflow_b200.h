/*
 * gflow_b200.h -- C ABI of the B200-native splat rasteriser (libgflow_b200.so).
 *
 * Drop-in boundary for the `msplat` operator surface GFlow calls
 * (/root/reference/gflow/utils/render.py:21-154, /root/reference/gflow/trainer.py:955).
 * `msplat` itself is a third-party CUDA extension that is NOT part of
 * /root/reference; each entry point below names the msplat
 * Python-level operator it implements and the GFlow call site that pins its
 * contract.  INTEGRATION.md shows the ctypes binding a maintainer adds.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current CUDA device unless noted;
 *  - float = IEEE binary32, row-major contiguous; int32 for integer tensors;
 *    `visible` is uint8 (0/1), may be NULL (= all visible);
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *  - return value: 0 on success, otherwise the cudaError_t of the failing call
 *    (gfb_error_string() decodes it), or a negative GFB_E_* code for bad arguments;
 *  - no entry point allocates device memory or synchronises the stream; work
 *    buffers are passed in by the caller and sized with the *_bytes() helpers;
 *  - N = number of Gaussians, K = number of (tile, Gaussian) intersections,
 *    T = ceil(W/16) * ceil(H/16) tiles, row-major (tile = ty * ceil(W/16) + tx).
 */
#ifndef GFLOW_B200_H
#define GFLOW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFB_TILE 16
#define GFB_E_BADARG (-1)     /* NULL pointer / negative size / unsupported channel count */
#define GFB_E_UNSUPPORTED (-2)
#define GFB_E_CAPACITY (-3)    /* K exceeded the caller's capacity: outputs are truncated, retry larger */
#define GFB_E_STALE (-4)       /* K ticket expired: its hand-off slot has been reused (256 later hand-offs on the device) */
#define GFB_E_NOTREADY (-5)    /* gfb_query_k_ticket: the producing kernel has not finished yet */

/* library / build information */
int gfb_version(void);                 /* 100 * major + minor */
const char *gfb_build_arch(void);      /* "sm_100a" */
const char *gfb_error_string(int code);
/* number of CUDA kernels this library has launched in this process (host-side counter) */
int64_t gfb_kernel_launch_count(void);

/* ------------------------------------------------------------------ msplat.project_point
 * call sites: render.py:21-24,116-119; trainer.py:955.
 * uv (N,2), depth (N,1); culled points give uv = depth = 0. */
int gfb_project_point_fwd(const float *xyz, const float *intr, const float *extr, int N, int W, int H,
                          float nearest, float extent, float *uv, float *depth, void *stream);
/* g_depth may be NULL.  d_cam = 16 floats {d_extr (3x4 row-major), d_intr (fx fy cx cy)},
 * overwritten (zeroed, then reduced over N). */
int gfb_project_point_bwd(const float *xyz, const float *intr, const float *extr, int N, int W, int H,
                          float nearest, float extent, const float *g_uv, const float *g_depth, float *d_xyz,
                          float *d_cam, void *stream);

/* ------------------------------------------------------------------ msplat.compute_cov3d
 * call sites: render.py:37-41,123-127.  cov3d (N,6) = upper triangle of R diag(s^2) R^T. */
int gfb_compute_cov3d_fwd(const float *scale, const float *rotate, const uint8_t *visible, int N, float *cov3d,
                          void *stream);
int gfb_compute_cov3d_bwd(const float *scale, const float *rotate, const uint8_t *visible, int N,
                          const float *g_cov3d, float *d_scale, float *d_rotate, void *stream);

/* ------------------------------------------------------------------ msplat.ewa_project
 * call sites: render.py:44-49,130-135.  conic (N,3), radius (N,1) int32, tiles_touched (N,1) int32. */
int gfb_ewa_project_fwd(const float *xyz, const float *cov3d, const float *intr, const float *extr, const float *uv,
                        int N, int W, int H, const uint8_t *visible, float *conic, int32_t *radius,
                        int32_t *tiles_touched, void *stream);
int gfb_ewa_project_bwd(const float *xyz, const float *cov3d, const float *intr, const float *extr, const float *uv,
                        int N, int W, int H, const uint8_t *visible, const float *g_conic, float *d_xyz,
                        float *d_cov3d, float *d_cam /* as in gfb_project_point_bwd */, void *stream);

/* ------------------------------------------------------------------ msplat.compute_sh
 * (no GFlow call site; north_star surface).  shs (N,C,K), K in {1,4,9,16}; dirs (N,3). */
int gfb_compute_sh_fwd(const float *shs, const float *dirs, const uint8_t *visible, int N, int C, int K, float *out,
                       void *stream);
int gfb_compute_sh_bwd(const float *shs, const float *dirs, const uint8_t *visible, int N, int C, int K,
                       const float *g_out, float *d_shs, float *d_dirs, void *stream);

/* ------------------------------------------------------------------ msplat.sort_gaussian
 * call sites: render.py:52-54,138-140.
 * Per-tile intersection counts -> exclusive scan (K) -> scatter of (depth bits, id) keys into every
 * tile's segment -> one warp per tile sorts its segment.  The result equals a stable ascending sort
 * of (tile << 32 | float bits of depth) over a Gaussian-major emission.
 * K has to reach the host (the caller sizes gaussian_ids_sorted (K,)); to hide that read-back the
 * caller passes a `capacity` (e.g. its previous K plus slack) for keys_ws (gfb_sort_workspace_bytes)
 * and gaussian_ids_sorted, scatter + sort are enqueued speculatively, and the host waits only for
 * count + scan on an internal event.  *K_host (HOST pointer) receives K.  Returns GFB_E_CAPACITY when
 * K > capacity (outputs truncated; call again with capacity >= *K_host).
 * tile_ws: gfb_sort_tile_workspace_bytes(W, H) bytes of scratch; tile_range (T,2) int32. */
size_t gfb_sort_workspace_bytes(int64_t K);
size_t gfb_sort_tile_workspace_bytes(int W, int H);
int gfb_sort_gaussian(const float *uv, const float *depth, const int32_t *radius, const int32_t *tiles_touched, int N,
                      int W, int H, void *tile_ws, int64_t capacity, void *keys_ws, int32_t *gaussian_ids_sorted,
                      int32_t *tile_range, int64_t *K_host, void *stream);
/* Same, for a caller that keeps tile_ws between calls on one stream: zero the block once; the part the kernels rely
 * on (tile counters + ticket) is zero again when the call's kernels have run, so no memset is enqueued per call.
 * After a failed call (other than GFB_E_CAPACITY) the block may be dirty: zero it again. */
int gfb_sort_gaussian_keep(const float *uv, const float *depth, const int32_t *radius, const int32_t *tiles_touched,
                           int N, int W, int H, void *tile_ws_keep, int64_t capacity, void *keys_ws,
                           int32_t *gaussian_ids_sorted, int32_t *tile_range, int64_t *K_host, void *stream);

/* ------------------------------------------------------------------ msplat.alpha_blending
 * call sites: render.py:58-64,68-74,84-90,99-105,148-154; backward via trainer.py:533.
 *
 * The blend kernels read two packed, tile-contiguous streams that are staged into
 * shared memory with 1-D bulk TMA copies:
 *   geometry stream: 2 K records of 16 B: K x {u, v, half-extent x, y} then K x {conic a, b, c, opacity}
 *   feature  stream: K records of 16 B (up to 4 channels, zero padded)
 * gfb_blend_pack_geometry / gfb_blend_pack_feature build them from the msplat-level
 * tensors; a geometry stream can be shared by every blend that uses the same
 * (uv, conic, opacity, gaussian_ids_sorted).  C > 4 is handled by the caller as
 * ceil(C/4) channel groups (c0 = first channel of the group, 1 <= Cg <= 4). */
size_t gfb_blend_geometry_stream_bytes(int64_t K);
size_t gfb_blend_feature_stream_bytes(int64_t K);
int gfb_blend_pack_geometry(const float *uv, const float *conic, const float *opacity,
                            const int32_t *gaussian_ids_sorted, int64_t K, void *geom_stream, void *stream);
int gfb_blend_pack_feature(const float *feature, int C, int c0, int Cg, const int32_t *gaussian_ids_sorted, int64_t K,
                           void *feat_stream, void *stream);
/* Both streams in one launch (the first blend of a geometry; bit-identical to the two calls above). */
int gfb_blend_pack_geometry_feature(const float *uv, const float *conic, const float *opacity, const float *feature,
                                    int C, int c0, int Cg, const int32_t *gaussian_ids_sorted, int64_t K,
                                    void *geom_stream, void *feat_stream, void *stream);
/* out is (C,H,W); this call writes channels [c0, c0+Cg).  final_T (H,W) float and
 * n_contrib (H,W) int32 are saved for the backward pass. */
int gfb_alpha_blending_fwd(const void *geom_stream, const void *feat_stream, int64_t K, const int32_t *tile_range,
                           int C, int c0, int Cg, float bg, int W, int H, float *out, float *final_T,
                           int32_t *n_contrib, void *stream);
/* Accumulates into grad_pack (N records of 12 floats {d_u, d_v, d_a, d_b, d_c, d_opacity,
 * d_f[c0..c0+3], pad, pad}); the caller zeroes grad_pack before each call. */
size_t gfb_blend_grad_pack_bytes(int N);
int gfb_alpha_blending_bwd(const void *geom_stream, const void *feat_stream, int64_t K,
                           const int32_t *gaussian_ids_sorted, const int32_t *tile_range, int C, int c0, int Cg,
                           float bg, int W, int H,
                           const float *final_T, const int32_t *n_contrib, const float *g_out, float *grad_pack,
                           void *stream);
/* Scatter grad_pack into msplat-shaped gradients.  Feature slots go to
 * d_feature[:, c0:c0+Cg] (row stride C).  flags: GFB_UNPACK_ACCUMULATE adds to d_uv / d_conic /
 * d_opacity instead of overwriting them (second and later channel groups); GFB_UNPACK_CLEAR writes
 * zeros back into grad_pack after reading it, so a pack the caller keeps between calls needs no
 * memset of its own. */
#define GFB_UNPACK_ACCUMULATE 1
#define GFB_UNPACK_CLEAR 2
int gfb_blend_unpack_grads(float *grad_pack, int N, int C, int c0, int Cg, float *d_uv, float *d_conic,
                           float *d_opacity, float *d_feature, int flags, void *stream);

/* ------------------------------------------------------------------ msplat.rasterization (fused)
 * The whole chain of render.py:21-64 for callers that only need the image: 4 forward kernels
 * (preprocess+count+scan, scatter, per-tile sort+pack, blend) and 2 backward kernels (blend
 * backward, fused geometry backward).  Per-Gaussian results are bit-identical to the separate
 * operators.  1 <= C <= 4.  Buffers:
 *   uv (N,2) depth (N,1) conic (N,3) radius (N,1): per-Gaussian outputs (also msplat-visible)
 *   rect_ws   : N x 8 bytes        control_ws : gfb_render_control_bytes(W,H), any contents
 *   tile_range (T,2) int32
 *   capacity  : number of intersections the K-sized buffers can hold:
 *               keys_ws (8 B), gaussian_ids_sorted (4 B), geom_stream (32 B), feat_stream (16 B) each
 *   out (C,H,W), final_T (H,W), n_contrib (H,W)
 * K is delivered to *K_host (HOST pointer); GFB_E_CAPACITY as in gfb_sort_gaussian.  With K_host == NULL
 * the call returns without waiting; the caller overlaps host work and then calls gfb_wait_k(&K), which
 * blocks until the K of the calling thread's most recent hand-off has landed (compare it with capacity).
 *
 * Every hand-off (gfb_sort_gaussian, gfb_render_forward) owns a slot of a per-device ring of pinned words +
 * events and is named by a ticket, so calls interleaved from several streams or host threads never see each
 * other's K, and a caller may validate its speculative capacity late (msplat.rasterization does so in its
 * backward, which keeps the host out of the forward path):
 *   gfb_k_ticket()                 ticket of the calling thread's most recent hand-off (-1: none yet)
 *   gfb_wait_k_ticket(t, &K)       blocks on that hand-off only; GFB_E_STALE if the slot has been reused
 *   gfb_query_k_ticket(t, &K)      same without blocking; GFB_E_NOTREADY while the kernel is still running */
size_t gfb_render_control_bytes(int W, int H);
/* Byte offset, inside control_ws, of the int32 word that receives K (device memory).  A forward captured into a CUDA
 * graph makes no K hand-off (gfb_k_ticket() is unchanged, K_host must be NULL); the owner of the graph reads this
 * word after a replay and compares it with the capacity the graph was captured with. */
size_t gfb_render_control_k_offset(int W, int H);
int gfb_wait_k(int64_t *K_host);
int64_t gfb_k_ticket(void);
int gfb_wait_k_ticket(int64_t ticket, int64_t *K_host);
int gfb_query_k_ticket(int64_t ticket, int64_t *K_host);
int gfb_render_forward(const float *xyz, const float *scale, const float *rotate, const float *opacity,
                       const float *feature, int C, const float *intr, const float *extr, int N, int W, int H,
                       float bg, float nearest, float extent, float *uv, float *depth, float *conic, int32_t *radius,
                       void *rect_ws, void *control_ws, int32_t *tile_range, int64_t capacity,
                       void *keys_ws, int32_t *gaussian_ids_sorted, void *geom_stream, void *feat_stream, float *out,
                       float *final_T, int32_t *n_contrib, int64_t *K_host, void *stream);
/* grad_ws: gfb_render_grad_bytes(N) bytes = N x 12 floats of packed per-Gaussian gradients followed by
 * d_cam (16 floats: d_extr 3x4, d_intr 4); zeroed inside.  capacity as passed to the forward call. */
size_t gfb_render_grad_bytes(int N);
int gfb_render_backward(const float *xyz, const float *scale, const float *rotate, const float *intr,
                        const float *extr, int N, int W, int H, int C, float bg, float nearest, float extent,
                        const int32_t *gaussian_ids_sorted, const int32_t *tile_range, int64_t capacity,
                        const void *geom_stream, const void *feat_stream, const float *final_T,
                        const int32_t *n_contrib, const float *g_out, void *grad_ws, float *d_xyz, float *d_scale,
                        float *d_rotate, float *d_opacity, float *d_feature, void *stream);

/* Variants for callers that keep their workspaces ALIVE across calls on one stream, which saves the two memset
 * launches of a render step (the blocks clean themselves):
 *   control_ws      zero before the first call; every call leaves it zero again (scatter hands the tile counters
 *                   back, the scan resets its ticket; only the K word keeps its value)
 *   grad_pack_keep  N x 12 floats, zero before the first call; geometry_bwd clears each row after reading it
 *   d_cam           16 floats (d_extr 3x4, d_intr 4), any contents: cleared by the blend backward before use
 * After a call that returned an error the caller must zero the kept blocks again before reusing them.  Everything
 * else as gfb_render_forward / gfb_render_backward. */
int gfb_render_forward_keep(const float *xyz, const float *scale, const float *rotate, const float *opacity,
                            const float *feature, int C, const float *intr, const float *extr, int N, int W, int H,
                            float bg, float nearest, float extent, float *uv, float *depth, float *conic, int32_t *radius,
                            void *rect_ws, void *control_ws, int32_t *tile_range, int64_t capacity,
                            void *keys_ws, int32_t *gaussian_ids_sorted, void *geom_stream, void *feat_stream, float *out,
                            float *final_T, int32_t *n_contrib, int64_t *K_host, void *stream);
int gfb_render_backward_keep(const float *xyz, const float *scale, const float *rotate, const float *intr,
                             const float *extr, int N, int W, int H, int C, float bg, float nearest, float extent,
                             const int32_t *gaussian_ids_sorted, const int32_t *tile_range, int64_t capacity,
                             const void *geom_stream, const void *feat_stream, const float *final_T,
                             const int32_t *n_contrib, const float *g_out, void *grad_pack_keep, float *d_cam, float *d_xyz,
                             float *d_scale, float *d_rotate, float *d_opacity, float *d_feature, void *stream);

/* ------------------------------------------------------------------ native per-frame optimisation loop
 * The inner loop GFlow runs per frame, /root/reference/gflow/trainer.py:387-558 (driven by
 * /root/reference/gflow/fit_video.py:119-142,256-315), as ONE stream of kernels per iteration with no host
 * synchronisation inside it:
 *   raw parameters -> activations (trainer.py:62-69) + project_point + compute_cov3d + ewa_project + tile
 *   binning (one kernel) -> scatter -> per-tile sort + pack -> ONE 4-channel blend that renders rgb and the
 *   depth map together (render.py:58-74 issues two blends over the same sort) -> photometric loss
 *   (mean squared error [+ 1 - SSIM, pytorch_ssim.py:17-37]) + scale/shift invariant depth loss
 *   (trainer.py:476-488) and dL/d(image) -> blend backward -> geometry backward through the activations,
 *   gradient masks (trainer.py:535-551) and the Adam update (torch.optim.Adam defaults, LinearLR
 *   1.0 -> 0.1, trainer.py:123-153,383-384,554-555) of every attribute in the same kernel -> pose
 *   (roma xyzw quaternion + translation, trainer.py:115-121) and depth_a/depth_b update.
 * All pointers are device pointers.  Raw parameters, pose and depth_ab are updated in place. */
typedef struct gfb_fit_problem {
    float *xyz, *scale, *rotate, *opacity, *rgb; /* raw attributes (N,3) (N,3) (N,4) (N,1) (N,3), trainer.py:81-88 */
    float *pose;                 /* 7: qx qy qz qw tx ty tz (raw; normalised inside, trainer.py:119) */
    float *depth_ab;             /* 2: depth_a, depth_b (trainer.py:146-149) */
    const float *intr;           /* 4: fx fy cx cy */
    const float *gt_image;       /* (H,W,3) in [0,1] */
    const float *gt_depth;       /* (H,W,1) or NULL (no depth term) */
    const uint8_t *pixel_mask;   /* (H,W) 1 = pixel takes part in the losses, or NULL (trainer.py:452-455,484) */
    const uint8_t *still_mask;   /* (n_still) 1 = xyz gradient zeroed (trainer.py:542-546), or NULL */
    const uint8_t *scale_sel;    /* (N) 1 = Gaussian takes part in loss_scale, or NULL = all: trainer.py:467-471 narrows
                                    the in-image index to the still (camera-only) / moving (full stage) set in place, and
                                    trainer.py:495-501 reads it through the alias self.within_index */
    const float *still_ref;      /* (n_still_ref,3) last frame's xyz, or NULL: loss_still (trainer.py:504-508) */
    const uint8_t *still_sel;    /* (n_still_ref) 1 = Gaussian takes part in loss_still (last_still_mask) */
    const float *flow_target;    /* (n_flow,2) last_uv + gt_flow[last_uv], or NULL: loss_flow (trainer.py:510-530) */
    const uint8_t *flow_sel;     /* (n_flow) 1 = Gaussian takes part in loss_flow (and_mask) */
    /* camera-only stage on frames >= 1 (trainer.py:427-451): every iteration the MOVING Gaussians (attributes are
     * frozen in this stage, so the caller passes a compact raw copy) are rendered under the current pose and every
     * pixel they touch (grey > 0) is removed from the losses, cumulatively.  sub_N = 0 switches this off. */
    const float *sub_xyz, *sub_scale, *sub_rotate, *sub_opacity, *sub_rgb; /* (sub_N, 3|3|4|1|3) raw */
    uint8_t *dyn_mask;           /* (H,W) 1 = pixel still counts; read and updated by the kernels; replaces pixel_mask */
    void *sub_workspace;         /* gfb_fit_sub_workspace_bytes(sub_N, W, H, sub_capacity) */
    float *dbg_grads;            /* NULL, or (N,14) raw-attribute gradients of the last iteration before masking */
    float *dbg_act;              /* NULL, or (N,14) activated attributes of the last iteration */
    int32_t N, W, H, n_still;
    int32_t n_still_ref, still_count; /* still_count = number of 1s in still_sel (the mean's denominator) */
    int32_t n_flow, flow_count;       /* flow_count = number of 1s in flow_sel */
    int32_t sub_N, sub_capacity;      /* moving subset: Gaussians, intersection capacity of sub_workspace */
    int32_t total_iters;         /* LinearLR horizon (`iterations` of trainer.train) */
    int32_t camera_only;         /* attribute gradients zeroed, pose still optimised (trainer.py:548-551) */
    int32_t freeze_rgb;          /* rgb gradient zeroed (frames >= 1, trainer.py:537-540) */
    int32_t use_ssim;            /* loss_rgb = mse + (1 - SSIM) as in trainer.py:459-462; 0 = mse only */
    /* state of the optimiser after a densification: the reference re-creates Adam over the attributes only,
     * with the initial lr and no scheduler (trainer.py:941-951), so from then on the Adam step count restarts
     * at adam_t0, the lr stays constant and pose / depth_a / depth_b are no longer updated */
    int32_t adam_t0;             /* iteration index at which the Adam state was last zeroed (0 initially) */
    int32_t constant_lr;         /* 1 = LinearLR factor fixed at 1 */
    int32_t freeze_camera;       /* 1 = pose and depth_a / depth_b not updated */
    float bg, nearest, extent;
    float lr, lr_camera, lambda_rgb, lambda_depth, lambda_var, lambda_scale, lambda_still, lambda_flow;
    float beta1, beta2, eps;     /* Adam; torch defaults 0.9, 0.999, 1e-8 */
    float depth_den_min;         /* lower clamp of the depth-loss denominator (0 = reference behaviour) */
} gfb_fit_problem;

/* byte offsets of the pieces of the fit workspace a caller may want to read back */
typedef struct gfb_fit_layout {
    size_t status;      /* int32[16]: [0] iterations done, [1] K of the last iteration, [2] max K seen,
                           [3] max K of the moving-subset render seen,
                           [8..14] float bits of dL/d(pose) of the last iteration (diagnostics) */
    size_t loss_hist;   /* float[max_iters][8]: total, mse, ssim, depth, var, scale, still, flow per iteration */
    size_t cam;         /* float[16]: extr (3x4) + intr used by the NEXT iteration */
    size_t adam_m;      /* float[14 N + 12]: xyz | scale | rotate | opacity | rgb | pose(7) | depth_ab(2) */
    size_t adam_v;
    size_t uv;          /* float (N,2) of the last iteration */
    size_t depth;       /* float (N,1) */
    size_t conic;       /* float (N,3) */
    size_t radius;      /* int32 (N,1) */
    size_t tile_range;  /* int32 (T,2) */
    size_t ids;         /* int32 (capacity) gaussian_ids_sorted */
    size_t out;         /* float (C,H,W), C = 4 with a depth term (rgb + depth map) else 3 */
    size_t g_out;       /* float (C,H,W) dL/d(out) */
    size_t total;       /* workspace size in bytes */
} gfb_fit_layout;

int gfb_fit_get_layout(int N, int W, int H, int64_t capacity, int max_iters, gfb_fit_layout *layout);
size_t gfb_fit_sub_workspace_bytes(int sub_N, int W, int H, int64_t sub_capacity);
/* zeroes the Adam state / status / loss history and derives the camera of iteration 0 from the pose */
int gfb_fit_init(const gfb_fit_problem *problem, void *workspace, int64_t capacity, int max_iters, void *stream);
/* enqueues iterations [first_iter, first_iter + n_iters) and returns without synchronising.  K is not
 * known on the host: the kernels clamp at `capacity`; status[2] (max K) > capacity after the fact means the
 * iterations since the last check rendered truncated tiles and must be redone with a larger workspace. */
int gfb_fit_iterate(const gfb_fit_problem *problem, void *workspace, int64_t capacity, int max_iters, int first_iter,
                    int n_iters, void *stream);

/* ------------------------------------------------------------------ error-driven densification
 * SimpleGaussian.densify_by_pixels, /root/reference/gflow/trainer.py:878-939, without the host round trip.
 *   gfb_rgb_error_map   : loss_rgb_pixel of trainer.py:457 from the rendered (3,H,W) image and the target
 *   gfb_densify_prepare : weights = (error + min positive error) * mask (mask given, or weights > threshold),
 *                         block sums + their scan.  The workspace starts with int32 stats[8]:
 *                         [1] = number of mask pixels (the caller derives densify_num = int(num_points *
 *                         mask_ratio * percent) from it), [3] = float bits of the total weight.
 *   gfb_densify_sample  : draws `count` pixels with replacement, proportional to the weights (counter-based
 *                         generator, `seed`), and writes the new RAW attributes of trainer.py:908-933 for
 *                         them: xyz = pix2world(pixel, gt_depth) (geometry.py:104-116), scale =
 *                         gt_depth / (min sampled depth * num_points), rotate = (1,0,0,0),
 *                         opacity = logit(0.99)/10, rgb = logit(target colour); sampled_pixels (count) int32. */
int gfb_rgb_error_map(const float *rendered, const float *gt_image, const uint8_t *pixel_mask, int W, int H,
                      float *error_map, void *stream);
size_t gfb_densify_workspace_bytes(int W, int H);
int gfb_densify_prepare(const float *error_map, const uint8_t *mask, int W, int H, float error_threshold,
                        void *workspace, void *stream);
int gfb_densify_sample(const void *workspace, const float *gt_image, const float *gt_depth, const float *intr,
                       const float *extr, int W, int H, int count, int num_points, uint64_t seed, float *new_xyz,
                       float *new_scale, float *new_rotate, float *new_opacity, float *new_rgb,
                       int32_t *sampled_pixels, void *stream);

/* ------------------------------------------------------------------ host pipe
 * Stream plumbing of a render step whose inputs and results live in HOST memory (what a caller of the reference
 * pays around msplat when its Gaussians are numpy / CPU tensors: .cuda() before, .cpu() after).  A pipe owns two copy
 * streams and `depth` slots; gfb_hostpipe_submit enqueues, for one slot,
 *   H2D  host_in (pinned) -> dev_in          on the pipe's upload stream,
 *   the caller's captured CUDA graph         (cudaGraphExec_t, e.g. of gfb_render_forward_keep + gfb_render_backward_keep
 *                                             over dev_in / dev_out) on compute_stream,
 *   D2H  dev_out -> host_out (pinned)        on the pipe's download stream,
 * ordered by events, so with depth >= 2 the copies of neighbouring steps overlap the kernels.  A slot may be
 * re-submitted at once (the call orders it behind the slot's previous step); host_out is valid after
 * gfb_hostpipe_wait.  One pipe per device; not thread safe. */
typedef struct gfb_hostpipe gfb_hostpipe;
int gfb_hostpipe_create(int depth, gfb_hostpipe **pipe);
int gfb_hostpipe_submit(gfb_hostpipe *pipe, int slot, void *dev_in, const void *host_in, size_t in_bytes,
                        void *graph_exec, void *compute_stream, void *host_out, const void *dev_out, size_t out_bytes);
int gfb_hostpipe_wait(gfb_hostpipe *pipe);
int gfb_hostpipe_destroy(gfb_hostpipe *pipe);

#ifdef __cplusplus
}
#endif
#endif /* GFLOW_B200_H */
