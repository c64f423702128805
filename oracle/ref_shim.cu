// TEST INFRASTRUCTURE ONLY (oracle/): a flat C façade over the UNMODIFIED reference rasterizer.
//
// This translation unit is compiled TOGETHER WITH the reference's own sources, taken where they lie
// under /root/reference (cuda_rasterizer/{forward,backward,rasterizer_impl}.cu and
// third_party/simple-knn/simple_knn.cu), into oracle/_ref/libref_rasterizer.so by oracle/Makefile.
// It contains no rasterization code of its own: it only forwards to
//   CudaRasterizer::Rasterizer::{forward,backward,markVisible}   (reference cuda_rasterizer/rasterizer.h:24-82)
//   SimpleKNN::knn                                               (reference third_party/simple-knn/simple_knn.h:15-19)
// and exposes the private scratch layout (GeometryState / BinningState / ImageState,
// reference cuda_rasterizer/rasterizer_impl.h:30-64) so tests can read intermediates
// (radii, tiles_touched, sorted keys/values, tile ranges, n_contrib ...).
//
// Only tests/, __graft_entry__.smoke() and bench.py (--impl reference) load the resulting library.
#include <cstdint>
#include <cstddef>
#include <functional>
#include <cuda_runtime.h>
#include "cuda_rasterizer/rasterizer_impl.h"
#include "simple_knn.h"

extern "C" {

typedef char* (*ref_alloc_fn)(size_t bytes, void* user);

int ref_forward(ref_alloc_fn geom_fn, void* geom_user,
                ref_alloc_fn bin_fn, void* bin_user,
                ref_alloc_fn img_fn, void* img_user,
                int P, int D, int M, const float* background, int width, int height,
                const float* means3D, const float* shs, const float* colors_precomp,
                const float* opacities, const float* scales, float scale_modifier,
                const float* rotations, const float* cov3D_precomp,
                const float* viewmatrix, const float* projmatrix, const float* cam_pos,
                float tan_fovx, float tan_fovy, int prefiltered, float* out_color, int* radii)
{
	std::function<char*(size_t)> g = [=](size_t n) { return geom_fn(n, geom_user); };
	std::function<char*(size_t)> b = [=](size_t n) { return bin_fn(n, bin_user); };
	std::function<char*(size_t)> i = [=](size_t n) { return img_fn(n, img_user); };
	return CudaRasterizer::Rasterizer::forward(g, b, i, P, D, M, background, width, height,
		means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp,
		viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy, prefiltered != 0, out_color, radii);
}

void ref_backward(int P, int D, int M, int R, const float* background, int width, int height,
                  const float* means3D, const float* shs, const float* colors_precomp,
                  const float* scales, float scale_modifier, const float* rotations,
                  const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix,
                  const float* campos, float tan_fovx, float tan_fovy, const int* radii,
                  char* geom_buffer, char* binning_buffer, char* image_buffer,
                  const float* dL_dpix, float* dL_dmean2D, float* dL_dconic, float* dL_dopacity,
                  float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                  float* dL_dscale, float* dL_drot)
{
	CudaRasterizer::Rasterizer::backward(P, D, M, R, background, width, height, means3D, shs,
		colors_precomp, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix,
		campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix,
		dL_dmean2D, dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot);
}

void ref_mark_visible(int P, float* means3D, float* viewmatrix, float* projmatrix, unsigned char* present)
{
	CudaRasterizer::Rasterizer::markVisible(P, means3D, viewmatrix, projmatrix, (bool*)present);
}

void ref_simple_knn(int P, float* points, float* mean_dists)
{
	SimpleKNN::knn(P, (float3*)points, mean_dists);
}

// out[0..9] = depths, clamped, internal_radii, means2D, cov3D, conic_opacity, rgb, point_offsets, tiles_touched, scanning_space
void ref_geom_pointers(char* chunk, int P, void** out)
{
	CudaRasterizer::GeometryState g = CudaRasterizer::GeometryState::fromChunk(chunk, (size_t)P);
	out[0] = g.depths; out[1] = g.clamped; out[2] = g.internal_radii; out[3] = g.means2D; out[4] = g.cov3D;
	out[5] = g.conic_opacity; out[6] = g.rgb; out[7] = g.point_offsets; out[8] = g.tiles_touched; out[9] = g.scanning_space;
}

// out[0..3] = point_list_keys_unsorted, point_list_keys, point_list_unsorted, point_list
void ref_binning_pointers(char* chunk, int R, void** out)
{
	CudaRasterizer::BinningState b = CudaRasterizer::BinningState::fromChunk(chunk, (size_t)R);
	out[0] = b.point_list_keys_unsorted; out[1] = b.point_list_keys; out[2] = b.point_list_unsorted; out[3] = b.point_list;
}

// out[0..2] = ranges (uint2 per tile), n_contrib, accum_alpha (final T)
void ref_image_pointers(char* chunk, int N, void** out)
{
	CudaRasterizer::ImageState s = CudaRasterizer::ImageState::fromChunk(chunk, (size_t)N);
	out[0] = s.ranges; out[1] = s.n_contrib; out[2] = s.accum_alpha;
}

int ref_memcpy_d2d(void* dst, const void* src, size_t n) { return (int)cudaMemcpy(dst, src, n, cudaMemcpyDeviceToDevice); }
int ref_device_synchronize() { return (int)cudaDeviceSynchronize(); }
int ref_last_error() { return (int)cudaGetLastError(); }

}  // extern "C"
