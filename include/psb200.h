/*
 * psb200 — C-ABI of the B200-native Gaussian-splatting hot path (libpsb200.so).
 *
 * Plain C, plain pointers and sizes, no torch / C++ types. Every pointer named *device* below is a
 * CUDA device pointer to contiguous float32 / int32 data on the current device; `stream` is a
 * cudaStream_t passed as void* (NULL = the legacy default stream the reference launches on).
 * All functions return >= 0 on success and a negative code on failure; psb_last_error() then returns a
 * human-readable reason (thread-local). CUDA launch errors ARE checked (the reference never checks,
 * SURVEY.md §2.2 quirk 13).
 *
 * Each entry point names the reference interface it replaces (paths relative to the Photo-SLAM tree).
 */
#ifndef PSB200_H_INCLUDED
#define PSB200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_ERR_ARG   (-1)
#define PSB_ERR_CUDA  (-2)
#define PSB_ERR_SIZE  (-3)

/* Scratch allocator callback: must return a device pointer to at least `bytes` bytes that stays valid
 * until the matching backward call. Replaces the std::function<char*(size_t)> resize callbacks of
 * cuda_rasterizer/rasterizer.h:32-34 (created by src/rasterize_points.cu:28-34). */
typedef char* (*psb_alloc_fn)(size_t bytes, void* user);

int psb_version(void);
const char* psb_last_error(void);

/* Scratch sizes; pure functions of P / num_rendered / W*H like `required<State>(n)`,
 * cuda_rasterizer/rasterizer_impl.h:66-72. */
size_t psb_geometry_bytes(int P);
size_t psb_binning_bytes(int num_rendered);
size_t psb_image_bytes(int num_pixels);

/*
 * Forward rasterization. Replaces CudaRasterizer::Rasterizer::forward
 * (cuda_rasterizer/rasterizer.h:31-52, rasterizer_impl.cu:198-336); same argument meaning and order.
 * Optional inputs are NULL ("None" convention, src/gaussian_rasterizer.cpp:209-219): exactly one of
 * {shs, colors_precomp} and one of {scales+rotations, cov3D_precomp}.
 * viewmatrix / projmatrix: 4x4 column-major in memory (m[4*c + r]), device. out_color: [3,H,W] device.
 * radii: [P] int32 device or NULL. Returns num_rendered (number of (Gaussian, tile) instances).
 * Like the reference it blocks the host once (to size the binning scratch through the callback).
 */
int psb_rasterize_forward(psb_alloc_fn geometry_buffer, void* geometry_user,
                          psb_alloc_fn binning_buffer, void* binning_user,
                          psb_alloc_fn image_buffer, void* image_user,
                          int P, int D, int M,
                          const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp,
                          const float* opacities, const float* scales, float scale_modifier,
                          const float* rotations, const float* cov3D_precomp,
                          const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                          float tan_fovx, float tan_fovy, int prefiltered,
                          float* out_color, int* radii, void* stream);

/*
 * Backward rasterization. Replaces CudaRasterizer::Rasterizer::backward
 * (cuda_rasterizer/rasterizer.h:54-82, rasterizer_impl.cu:340-432). R = value returned by forward;
 * the three scratch pointers are the chunks handed out by the forward callbacks. Gradient outputs are
 * caller-allocated, ZERO-INITIALISED device arrays with the reference's shapes
 * (src/rasterize_points.cu:148-157): dL_dmean2D [P,3] (x,y written), dL_dconic [P,2,2] (x,y,w written),
 * dL_dopacity [P], dL_dcolor [P,3], dL_dmean3D [P,3], dL_dcov3D [P,6], dL_dsh [P,M,3],
 * dL_dscale [P,3], dL_drot [P,4]. Rows of invisible Gaussians are left untouched.
 */
int psb_rasterize_backward(int P, int D, int M, int R,
                           const float* background, int width, int height,
                           const float* means3D, const float* shs, const float* colors_precomp,
                           const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                           const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                           char* geom_buffer, char* binning_buffer, char* image_buffer,
                           const float* dL_dpix,
                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor,
                           float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh, float* dL_dscale, float* dL_drot,
                           void* stream);

/* Near-plane visibility (view z > 0.2). Replaces CudaRasterizer::Rasterizer::markVisible
 * (cuda_rasterizer/rasterizer.h:24-29, rasterizer_impl.cu:141-153). present: [P] bytes (bool), device. */
int psb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     unsigned char* present, void* stream);

/*
 * Parity/debug export (tests only): expands this library's private scratch into the reference's
 * intermediate arrays so they can be compared element by element with the reference's GeometryState /
 * BinningState / ImageState (cuda_rasterizer/rasterizer_impl.h:30-64). Any output may be NULL.
 *   depths [P] f32, means2D [P,2] f32, conic_opacity [P,4] f32, rgb [P,3] f32, clamped [P,3] u8,
 *   tiles_touched [P] u32, keys_sorted [R] u64 = (tile << 32 | depth bits), values_sorted [R] u32,
 *   ranges [tiles,2] u32, n_contrib [W*H] u32, final_T [W*H] f32.
 * Entries of culled Gaussians (tiles_touched == 0) are written as zero.
 */
int psb_debug_export(int P, int R, int width, int height,
                     char* geom_buffer, char* binning_buffer, char* image_buffer,
                     float* depths, float* means2D, float* conic_opacity, float* rgb, unsigned char* clamped,
                     uint32_t* tiles_touched, uint64_t* keys_sorted, uint32_t* values_sorted,
                     uint32_t* ranges, uint32_t* n_contrib, float* final_T, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Trainer-level surface: one fused training iteration on the RAW model tensors.
 *
 * Replaces the body of GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp:677-772) /
 * GaussianTrainer::trainingOnce (src/gaussian_trainer.cpp:31-135) between "pick a keyframe" and
 * "densify": GaussianRenderer::render (src/gaussian_renderer.cpp:23-149) incl. the activations of
 * src/gaussian_model.cpp:48-71, loss_utils::l1_loss / ssim (include/loss_utils.h:28-124), loss.backward(),
 * max_radii2D / addDensificationStats (gaussian_mapper.cpp:714-719, gaussian_model.cpp:817-831) and
 * gaussians_->optimizer_->step() + zero_grad (torch::optim::Adam over the 6 groups of
 * gaussian_model.cpp:477-503). Tensor layouts are the reference's (include/gaussian_model.h):
 *   index 0 xyz [P,3] | 1 features_dc [P,1,3] | 2 features_rest [P,15,3] | 3 opacity [P,1] (logit) |
 *   4 scaling [P,3] (log) | 5 rotation [P,4] (w,x,y,z, unnormalised);   all device float32, contiguous.
 * ------------------------------------------------------------------------------------------------ */
#define PSB_ERR_RETRY (-4)

typedef struct psb_trainer psb_trainer; /* opaque: scratch arena of the iteration */

typedef struct psb_model {
	float* param[6];          /* parameters, updated in place by psb_trainer_step / psb_adam_update */
	float* exp_avg[6];        /* Adam first moments  (same shapes) */
	float* exp_avg_sq[6];     /* Adam second moments (same shapes) */
	float* max_radii2D;       /* [P]   or NULL */
	float* xyz_gradient_accum;/* [P,1] or NULL */
	float* denom;             /* [P,1] or NULL */
	int* exist_since_iter;    /* [P] int32 or NULL (GaussianModel::exist_since_iter_; only the densify / prune / insert entry points move it) */
} psb_model;

typedef struct psb_camera {
	const float* viewmatrix;  /* device, 16 floats, column-major (world_view_transform_ memory) */
	const float* projmatrix;  /* device, 16 floats (full_proj_transform_ memory) */
	const float* campos;      /* device, 3 floats */
	float tan_fovx, tan_fovy;
	int width, height;
} psb_camera;

typedef struct psb_step {
	float lr[6];              /* per-group learning rates, order as psb_model.param */
	float beta1, beta2, eps;  /* reference: 0.9, 0.999, 1e-15 (gaussian_model.cpp:483-485) */
	int step;                 /* 1-based Adam step count of THIS update */
	float lambda_dssim;       /* reference default 0.2 */
	int sh_degree;            /* active SH degree 0..3 */
	int update_densify_stats; /* != 0: fold max_radii2D / xyz_gradient_accum / denom updates in */
} psb_step;

int psb_trainer_create(psb_trainer** out);
int psb_trainer_destroy(psb_trainer* t);

/* Forward only, from raw parameters (GaussianMapper::renderFromPose, gaussian_mapper.cpp:1521-1569).
 * out_color [3,H,W] device; radii [P] int32 device or NULL. M must be 16. Asynchronous on `stream`. */
int psb_trainer_render(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera,
                       const float* background, int sh_degree, float* out_color, int* radii, void* stream);

/* One fused iteration: render -> loss -> backward -> Adam (parameters and moments updated in place).
 * gt_image [3,H,W] device; mask [3,H,W] device or NULL (undistortion mask, gaussian_mapper.cpp:692);
 * out_color / radii optional outputs. Asynchronous: no host synchronisation happens inside. */
int psb_trainer_step(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera,
                     const float* background, const float* gt_image, const float* mask, const psb_step* step,
                     float* out_color, int* radii, void* stream);

/* Data-parallel halves of the same iteration: psb_trainer_backward writes the gradients w.r.t. the RAW
 * parameters into grads[0..5] (shapes of psb_model.param; every row written, zeros for invisible Gaussians)
 * and updates the densification statistics but NOT the parameters; after the all-reduce of `grads`,
 * psb_adam_update applies torch::optim::Adam semantics with g = grads * grad_scale. */
int psb_trainer_backward(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera,
                         const float* background, const float* gt_image, const float* mask, const psb_step* step,
                         float* out_color, int* radii, float* const* grads, void* stream);
int psb_adam_update(int P, int M, const psb_model* model, float* const* grads, const psb_step* step,
                    float grad_scale, void* stream);
/* Pipelined form of psb_trainer_backward: _begin runs render, loss and the tile backward (everything up to the 9
 * screen-space sums per Gaussian); _slab then produces the raw-parameter gradients of Gaussians [first, first+count)
 * (first a multiple of 128). grads[i] must point at the address where row 0 of tensor i WOULD be, i.e.
 * slab_buffer_i - first * floats_per_row_i: the kernel writes row g of tensor i at grads[i] + g * floats_per_row_i,
 * which lets the caller keep each slab's six gradient blocks contiguous and all-reduce slab k while slab k+1 is
 * still being computed. */
int psb_trainer_backward_begin(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera,
                               const float* background, const float* gt_image, const float* mask, const psb_step* step,
                               float* out_color, int* radii, void* stream);
int psb_trainer_backward_slab(psb_trainer* t, int P, int M, const psb_model* model, const psb_camera* camera,
                              const psb_step* step, int first, int count, float* const* grads, void* stream);
/* Same update on one flat range of n floats (any slice of a parameter tensor and the matching slices of its moments and
 * gradient) with an explicit learning rate: lets the caller pipeline chunked all-reduces with the optimizer. */
int psb_adam_flat(size_t n, float* param, float* exp_avg, float* exp_avg_sq, const float* grad, float lr,
                  const psb_step* step, float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Keyframe-sharded data parallelism over NVLink peer memory: one process per GPU, replicated
 * Gaussians, rank r trains on its own view, K views per optimizer step (mean gradient). Replaces the
 * optimizer step of the reference iteration (src/gaussian_mapper.cpp:769-772) at world > 1 by a fused
 *   reduce-scatter (80-byte gradient records pushed by the per-Gaussian backward straight into the owner
 *   rank's inbox) -> Adam on the owned rows only (moments live on the owner) -> all-gather (the owner's Adam
 *   kernels store the updated rows into every rank's parameter tensors)
 * with no materialised reduced gradient and no collective call. Ownership: chunks of 128 Gaussians, chunk c
 * belongs to rank c % world. The six parameter tensors must live in this context's arena (psb_dp_params), which
 * is mapped into the peers with CUDA IPC: create on every rank -> exchange the psb_dp_ipc_handle blobs with any
 * host-side transport (torch.distributed, MPI, a socket) -> psb_dp_connect. Moments (psb_model.exp_avg*) stay
 * ordinary full-size tensors; only the rows this rank owns are read or written.
 * ------------------------------------------------------------------------------------------------ */
typedef struct psb_dp psb_dp;
int psb_dp_create(psb_dp** out, int rank, int world, int P);   /* world <= 8; every rank must pass the same P */
int psb_dp_handle_bytes(void);                                 /* size of one IPC handle blob (64) */
int psb_dp_ipc_handle(psb_dp* dp, void* out_handle);           /* this rank's blob */
int psb_dp_connect(psb_dp* dp, const void* handles);           /* world blobs, rank-major; no-op for world == 1 */
int psb_dp_params(psb_dp* dp, float** out6);                   /* device addresses of the six parameter tensors in the arena (order of psb_model.param) */
/* One data-parallel iteration: render -> loss -> tile backward -> per-Gaussian backward with records pushed to the
 * owners -> (device-side wait for every rank's records) -> Adam of the owned rows, updated rows stored to every rank.
 * Asynchronous: no host synchronisation, every cross-rank dependency is a device-side wait on an epoch flag. A view
 * whose binning arena overflows contributes a zero gradient to the step (psb_trainer_result still reports
 * PSB_ERR_RETRY and grows the arena): the replicas stay identical and no rank waits forever. */
int psb_dp_step(psb_trainer* t, psb_dp* dp, int P, int M, const psb_model* model, const psb_camera* camera,
                const float* background, const float* gt_image, const float* mask, const psb_step* step,
                float* out_color, int* radii, void* stream);
/* Enqueues a device-side wait until every rank's rows of the last psb_dp_step have landed in this rank's tensors. */
int psb_dp_sync(psb_dp* dp, void* stream);
/* 0 = healthy, 1 = a cross-rank wait timed out (PSB_DP_TIMEOUT_MS, default 20 s). Synchronises the stream. */
int psb_dp_status(psb_dp* dp, void* stream);
int psb_dp_destroy(psb_dp* dp);

/* Returns the results of the last step / backward / render: out3 = {loss, l1, ssim} (host, may be NULL),
 * *num_rendered (may be NULL). Blocks only until those scalars have reached pinned host memory (an event recorded
 * right behind the loss kernel), NOT until the step has finished: the caller can read the loss of iteration i and
 * enqueue iteration i+1 while the backward half of i is still running. Use a stream/device synchronize to wait for
 * the step itself.
 * Returns PSB_ERR_RETRY when the binning arena was too small for a view SINCE THE LAST CALL of this function (the
 * record is sticky on the device: a queued step that overflowed is reported even if later steps fitted): that step
 * was a no-op on the model, the arena has been grown, call the step again. psb_trainer_overflow_info tells which
 * call it was: sequence numbers count the forward passes (render / step / backward) of this context from 1;
 * *first_seq = the first overflowing one, *count = how many overflowed, *current_seq = the latest pass enqueued. */
int psb_trainer_result(psb_trainer* t, float* out3, int* num_rendered, void* stream);
int psb_trainer_overflow_info(psb_trainer* t, unsigned* first_seq, unsigned* count, unsigned* current_seq);

/* Parity/debug export (tests only) of the last step / render of this context: per pixel the final transmittance
 * (final_T [H*W] f32, device, may be NULL) and the Gaussian index of the last blended splat (last_gauss [H*W] i32,
 * device, -1 = none, may be NULL) — the trainer-path counterpart of psb_debug_export's n_contrib / final_T, comparable
 * with the reference's point_list[ranges.x + n_contrib - 1] (cuda_rasterizer/forward.cu:352-365) even though the
 * tight instance lists number their entries differently. *num_rendered (host, may be NULL) = instance count. */
int psb_trainer_debug_state(psb_trainer* t, int width, int height, int* last_gauss, float* final_T, int* num_rendered, void* stream);

/* Optional per-stage timing with CUDA events recorded on the step's stream (bench.py uses it for the roofline numbers; off by
 * default). psb_trainer_stage_times fills ms[0..n) (n >= 7) and returns how many stages it filled:
 *   psb_trainer_step : 8 = preprocess, depth sort, binning (emit + tile sort + ranges), render forward, loss fwd+bwd, render
 *                      backward, per-Gaussian backward kernel, f_rest Adam kernel;
 *   psb_dp_step      : 10 = the first six, then push backward, wait for every rank's records, owner-side Adam, and ms[9] = the
 *                      wait at the start of the step for the previous step's rows. */
int psb_trainer_set_profiling(psb_trainer* t, int enable);
int psb_trainer_stage_times(psb_trainer* t, float* ms, int n);

/* Stand-alone fused loss (tests / evaluation): L = (1-lambda) L1 + lambda (1 - SSIM), reference
 * include/loss_utils.h:28-124. image/gt/mask/dL_dimage device [3,H,W] (mask, dL_dimage may be NULL);
 * out3_host = {loss, l1, ssim}. Synchronises the stream. */
int psb_loss(int height, int width, const float* image, const float* gt_image, const float* mask,
             float lambda_dssim, float* dL_dimage, float* out3_host, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Fused densification. Replaces GaussianModel::densifyAndPrune and everything under it (src/gaussian_model.cpp:795-815,
 * densifyAndClone :763-793, densifyAndSplit :716-761, prunePoints :588-642, densificationPostfix :644-714): the reference
 * re-materialises all 59 parameters + 118 Adam moments per Gaussian four times (cat, cat, mask-index, mask-index) with a
 * blocking .item(); here every surviving row is read once and every output row written once.
 *   plan : evaluates the clone / split / prune predicates on the P source rows, scans them, returns counts_host[5] =
 *          {P_new, surviving originals, clones, surviving split children PER COPY, split-selected rows}. This is the one host
 *          round trip (the caller sizes the output tensors from P_new).
 *   apply: writes the P_new output rows of all six parameter tensors and their moments into `dst`, in the reference's row
 *          order [surviving originals | clones | split children copy 0 | copy 1]; new rows get zero moments; the three
 *          statistics arrays of `dst` are zeroed (reference :709-711). Split children: xyz = R(q)(z * exp(s)) + xyz,
 *          scaling = log(exp(s)/1.6); z = normal_samples [2 * n_split, 3] (row c * n_split + rank of the parent among the
 *          split-selected rows, the reference's repeat({N,1}) order) or, when NULL, Philox4x32-10 + Box-Muller keyed by
 *          (cfg.seed, cfg.offset, sample row): counter-based, identical on every data-parallel replica.
 * workspace: psb_densify_workspace_bytes(P) device bytes, the same buffer for plan and apply.
 * ------------------------------------------------------------------------------------------------ */
typedef struct psb_densify_cfg {
	float max_grad;          /* densify_grad_threshold */
	float min_opacity;       /* prune below (reference call sites: 0.005) */
	float extent;            /* scene extent (cameras_extent_) */
	float percent_dense;     /* GaussianModel::percentDense() */
	int max_screen_size;     /* != 0 enables the world-size prune (max exp(scaling) > 0.1 extent), like the reference */
	unsigned long long seed, offset;
} psb_densify_cfg;
size_t psb_densify_workspace_bytes(int P);
int psb_densify_plan(int P, const psb_model* src, const psb_densify_cfg* cfg, void* workspace, int* counts_host, void* stream);
int psb_densify_apply(int P, const psb_model* src, const psb_model* dst, int P_new, const psb_densify_cfg* cfg, const void* workspace,
                      const float* normal_samples, void* stream);
/* GaussianModel::prunePoints (src/gaussian_model.cpp:588-642) through the same machinery: plan from a byte mask (!= 0 = remove),
 * counts_host as above (only [0] and [1] non-zero), then psb_densify_apply with the same workspace compacts parameters, moments
 * and exist_since_iter; NOTE: like every densify apply it zeroes dst's statistics — pass keep_stats != 0 to compact them instead
 * (prunePoints keeps them, :637-641). */
int psb_prune_plan(int P, const unsigned char* mask, void* workspace, int* counts_host, void* stream);
int psb_prune_apply(int P, const psb_model* src, const psb_model* dst, int P_new, const void* workspace, int keep_stats, void* stream);
/* GaussianModel::increasePcd (src/gaussian_model.cpp:193-377, both overloads) minus the k-NN call: dst (P + n rows) = src rows
 * followed by n new Gaussians: xyz = points, features_dc = RGB2SH(colors) (include/sh_utils.h:138-141), features_rest = 0,
 * opacity = inverse_sigmoid(0.1), scaling = log(sqrt(max(dist2, 1e-7))) on all three axes (dist2 = psb_dist_cuda2 of the NEW
 * points), rotation = (1,0,0,0), zero moments, exist_since_iter = iteration; statistics of all rows reset to zero
 * (densificationPostfix :709-711). points/colors [n,3], dist2 [n]: device. */
int psb_insert_points(int P, const psb_model* src, const psb_model* dst, int n, const float* points, const float* colors,
                      const float* dist2, int iteration, void* stream);
/* GaussianModel::resetOpacity (src/gaussian_model.cpp:556-565, incl. its no-op clamp) + zeroed opacity moments, in place. */
int psb_reset_opacity(int P, float* opacity, float* exp_avg, float* exp_avg_sq, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Point-cloud helpers exported by the same shared objects in the reference.
 * ------------------------------------------------------------------------------------------------ */

/* Mean squared distance of every point to its 3 nearest neighbours. Replaces SimpleKNN::knn /
 * distCUDA2 (third_party/simple-knn/simple_knn.h:15-19, spatial.cu:15-26). points [P,3], mean_dists [P], device.
 * Asynchronous on `stream` (scratch from the stream-ordered allocator; no host round trips). */
int psb_dist_cuda2(int P, const float* points, float* mean_dists, void* stream);

/* p' = M p for a 4x4 column-major M (rows 0..2 used). Replaces the transform_points kernel behind
 * transformPoints (src/operate_points.cu:38-50, 73-93). Out-of-place: out_points [P,3]. */
int psb_transform_points(int P, const float* points, const float* transform, float* out_points, void* stream);

/* For rows with mask != 0: p' = M (scale * p), q' = quaternion(M3x3 * R(q)) (w,x,y,z). Replaces the
 * scale_and_transform_points kernel behind scaleAndTransformThenMarkVisiblePoints (src/operate_points.cu:52-71,
 * 95-143). Unmasked rows of the outputs are left untouched. fix_quaternion_write = 0 reproduces the reference's
 * insert_rot_to_rots exactly (cuda_rasterizer/operate_points.h:170-178 writes z into slot +2 and never writes
 * slot +3 — SURVEY.md §2.2 quirk 10); != 0 writes the intended (w,x,y,z). */
int psb_scale_transform_points(int P, float scale, const float* points, const float* rots, const float* transform,
                               const unsigned char* mask, float* out_points, float* out_rots,
                               int fix_quaternion_write, void* stream);

/* Depth image -> camera-space points for masked pixels (idx = v * width + u). Replaces the reproject_depths_pinhole
 * kernel behind reprojectDepthPinhole (src/stereo_vision.cu:39-61, 138-166). points [P,3] must be zero-initialised. */
int psb_reproject_depth_pinhole(int P, int width, float fx, float fy, float cx, float cy, const float* depths,
                                const unsigned char* mask, float* points, void* stream);

/* Keypoints with a 3-D point keep it; the others borrow the depth of the nearest keypoint (squared pixel distance
 * <= max_pixel_dist) that has one and are re-projected, or get z = -1. Replaces the kernel behind
 * monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints (src/stereo_vision.cu:63-136, 168-215), including
 * its colour indexing `colors[(int)(v * width + u) + c]`. Outputs [N,3], zero-initialised by the caller. */
int psb_neighbour_depth_pinhole(int N, int width, float fx, float fy, float cx, float cy, float max_pixel_dist,
                                const float* pixels, const unsigned char* has3D, const float* points_local,
                                const float* colors, float* out_points, float* out_colors, void* stream);

/* Standalone sort primitive (tests): stable LSD radix sort of (u32 key, u32 value) pairs on key bits
 * [0, nbits). keys/vals are device arrays of n elements, sorted in place. */
int psb_debug_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, int nbits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_H_INCLUDED */
