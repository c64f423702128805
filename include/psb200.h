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

/* Standalone sort primitive (tests): stable LSD radix sort of (u32 key, u32 value) pairs on key bits
 * [0, nbits). keys/vals are device arrays of n elements, sorted in place. */
int psb_debug_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, int nbits, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_H_INCLUDED */
