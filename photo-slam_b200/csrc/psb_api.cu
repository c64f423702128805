// C-ABI entry points of libpsb200 (declared in include/psb200.h) and the host orchestration of one
// rasterization: scratch carving through the caller's allocator callbacks, kernel sequence, error checks.
#include <cstdio>
#include <cstring>
#include <string>
#include "psb_kernels.h"
#include "../../include/psb200.h"

namespace psb {

static thread_local std::string g_last_error;

void set_error(const char* what, cudaError_t e, const char* file, int line)
{
	char buf[512];
	snprintf(buf, sizeof(buf), "%s failed: %s (%s:%d)", what, cudaGetErrorString(e), file, line);
	g_last_error = buf;
}
void set_error_msg(const char* what) { g_last_error = what; }

static Camera make_camera(int width, int height, const float* view, const float* proj, const float* campos, float tan_fovx, float tan_fovy)
{
	Camera cam;
	cam.view = view; cam.proj = proj; cam.campos = campos;
	cam.tan_fovx = tan_fovx; cam.tan_fovy = tan_fovy;
	cam.focal_y = height / (2.0f * tan_fovy);
	cam.focal_x = width / (2.0f * tan_fovx);
	cam.W = width; cam.H = height;
	cam.grid_x = (width + PSB_TILE_X - 1) / PSB_TILE_X;
	cam.grid_y = (height + PSB_TILE_Y - 1) / PSB_TILE_Y;
	return cam;
}

static GaussIn make_input(int P, int D, int M, const float* means3D, const float* shs, const float* colors_precomp, const float* opacities,
                          const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp)
{
	GaussIn in;
	memset(&in, 0, sizeof(in));
	in.P = P; in.D = D; in.M = M;
	in.means3D = means3D; in.scales = scales; in.rotations = rotations; in.opacities = opacities;
	in.shs = shs; in.cov3D_precomp = cov3D_precomp; in.colors_precomp = colors_precomp;
	in.scale_modifier = scale_modifier;
	in.sh_vec4 = (shs != nullptr && M == 16 && (reinterpret_cast<uintptr_t>(shs) & 15) == 0) ? 1 : 0;
	return in;
}

// ---- debug export kernels ------------------------------------------------------------------------
__global__ void export_geom_kernel(int P, GeomState geom, float* depths, float* means2D, float* conic_opacity, float* rgb, uint8_t* clamped,
                                   uint32_t* tiles_touched)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	const uint32_t tt = geom.tile_info[i].z;
	GaussRec r;
	r.q0 = r.q1 = r.q2 = make_float4(0, 0, 0, 0);
	if (tt) r = geom.rec[i];
	if (tiles_touched) tiles_touched[i] = tt;
	if (depths) depths[i] = r.q1.w;
	if (means2D) { means2D[2 * i] = r.q0.x; means2D[2 * i + 1] = r.q0.y; }
	if (conic_opacity) { conic_opacity[4 * i] = r.q0.z; conic_opacity[4 * i + 1] = r.q0.w; conic_opacity[4 * i + 2] = r.q1.x; conic_opacity[4 * i + 3] = r.q1.y; }
	if (rgb) { rgb[3 * i] = r.q2.x; rgb[3 * i + 1] = r.q2.y; rgb[3 * i + 2] = r.q2.z; }
	if (clamped) {
		const uint32_t cb = rec_clamp_bits(__float_as_uint(r.q2.w));
		clamped[3 * i] = cb & 1u; clamped[3 * i + 1] = (cb >> 1) & 1u; clamped[3 * i + 2] = (cb >> 2) & 1u;
	}
}
__global__ void export_keys_kernel(int R, const uint32_t* tile_key, const uint32_t* inst, const GaussRec* rec, uint64_t* keys, uint32_t* values)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= R) return;
	const uint32_t g = inst[i];
	if (keys) keys[i] = ((uint64_t)tile_key[i] << 32) | (uint64_t)__float_as_uint(rec[g].q1.w);
	if (values) values[i] = g;
}

}  // namespace psb

using namespace psb;

extern "C" {

int psb_version(void) { return 100; }
const char* psb_last_error(void) { return g_last_error.c_str(); }

size_t psb_geometry_bytes(int P) { return required_bytes<GeomState>((size_t)P); }
size_t psb_binning_bytes(int R) { return required_bytes<BinState>((size_t)R); }
size_t psb_image_bytes(int N) { return required_bytes<ImgState>((size_t)N); }

int psb_rasterize_forward(psb_alloc_fn geometry_buffer, void* geometry_user, psb_alloc_fn binning_buffer, void* binning_user,
                          psb_alloc_fn image_buffer, void* image_user, int P, int D, int M, const float* background, int width, int height,
                          const float* means3D, const float* shs, const float* colors_precomp, const float* opacities, const float* scales,
                          float scale_modifier, const float* rotations, const float* cov3D_precomp, const float* viewmatrix,
                          const float* projmatrix, const float* cam_pos, float tan_fovx, float tan_fovy, int prefiltered, float* out_color,
                          int* radii, void* stream_)
{
	(void)prefiltered;  // the reference only uses it for a device-side assert (auxiliary.h:156-160)
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || width <= 0 || height <= 0 || !geometry_buffer || !binning_buffer || !image_buffer) { set_error_msg("psb_rasterize_forward: bad argument"); return PSB_ERR_ARG; }
	if (P > 0 && (!means3D || !opacities || !viewmatrix || !projmatrix || !background || !out_color)) { set_error_msg("psb_rasterize_forward: null required pointer"); return PSB_ERR_ARG; }
	if (P > 0 && ((shs == nullptr) == (colors_precomp == nullptr))) { set_error_msg("psb_rasterize_forward: provide exactly one of shs / colors_precomp"); return PSB_ERR_ARG; }
	if (P > 0 && (((scales == nullptr) || (rotations == nullptr)) == (cov3D_precomp == nullptr))) { set_error_msg("psb_rasterize_forward: provide exactly one of scales+rotations / cov3D_precomp"); return PSB_ERR_ARG; }
	if (P > 0 && shs && (M < (D + 1) * (D + 1) || D > 3 || D < 0)) { set_error_msg("psb_rasterize_forward: SH degree / coefficient count mismatch"); return PSB_ERR_ARG; }
	if (P > 0 && shs && !cam_pos) { set_error_msg("psb_rasterize_forward: cam_pos required with shs"); return PSB_ERR_ARG; }
	if (width > 65535 * 16 || height > 65535 * 16) { set_error_msg("psb_rasterize_forward: image too large"); return PSB_ERR_ARG; }

	const Camera cam = make_camera(width, height, viewmatrix, projmatrix, cam_pos, tan_fovx, tan_fovy);
	if ((size_t)cam.grid_x * cam.grid_y >= (1u << 24)) { set_error_msg("psb_rasterize_forward: more than 2^24 tiles"); return PSB_ERR_ARG; }

	char* geom_chunk = geometry_buffer(required_bytes<GeomState>((size_t)P), geometry_user);
	char* img_chunk = image_buffer(required_bytes<ImgState>((size_t)width * height), image_user);
	if (!geom_chunk || !img_chunk) { set_error_msg("psb_rasterize_forward: allocator returned null"); return PSB_ERR_ARG; }
	GeomState geom = GeomState::from_chunk(geom_chunk, (size_t)P);
	ImgState img = ImgState::from_chunk(img_chunk, (size_t)width * height);

	const GaussIn in = make_input(P, D, M, means3D, shs, colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp);
	int rc;
	if ((rc = launch_preprocess(in, cam, radii, geom, /*raw=*/false, /*tight=*/false, stream))) return rc;
	if ((rc = launch_depth_sort_and_scan(P, geom, /*scan=*/true, stream))) return rc;

	uint32_t num_rendered = 0;
	if (P > 0) {
		PSB_CUDA_OK(cudaMemcpyAsync(&num_rendered, geom.counters, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
		PSB_CUDA_OK(cudaStreamSynchronize(stream));
	}
	if (num_rendered >= (1u << 30)) { set_error_msg("psb_rasterize_forward: more than 2^30 instances"); return PSB_ERR_SIZE; }

	char* bin_chunk = binning_buffer(required_bytes<BinState>((size_t)num_rendered), binning_user);
	if (!bin_chunk) { set_error_msg("psb_rasterize_forward: allocator returned null"); return PSB_ERR_ARG; }
	BinState bin = BinState::from_chunk(bin_chunk, (size_t)num_rendered);
	if ((rc = launch_binning(P, cam, geom, bin, img, (size_t)num_rendered, stream))) return rc;

	const int res = make_sort_plan(tile_id_bits(cam.grid_x * cam.grid_y)).npass & 1;
	if ((rc = launch_render_forward(cam, img.ranges, bin.inst[res], geom.rec, background, out_color, img.final_T, img.n_contrib, stream))) return rc;
	return (int)num_rendered;
}

int psb_rasterize_backward(int P, int D, int M, int R, const float* background, int width, int height, const float* means3D, const float* shs,
                           const float* colors_precomp, const float* scales, float scale_modifier, const float* rotations,
                           const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos, float tan_fovx,
                           float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix,
                           float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D,
                           float* dL_dsh, float* dL_dscale, float* dL_drot, void* stream_)
{
	(void)radii;  // visibility is re-derived from the private geometry state (tiles_touched), identical to radii > 0
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P == 0) return 0;
	if (P < 0 || R < 0 || !geom_buffer || !binning_buffer || !image_buffer || !dL_dpix || !dL_dmean2D || !dL_dconic || !dL_dopacity || !dL_dcolor) {
		set_error_msg("psb_rasterize_backward: bad argument");
		return PSB_ERR_ARG;
	}
	const Camera cam = make_camera(width, height, viewmatrix, projmatrix, campos, tan_fovx, tan_fovy);
	GeomState geom = GeomState::from_chunk(geom_buffer, (size_t)P);
	BinState bin = BinState::from_chunk(binning_buffer, (size_t)R);
	ImgState img = ImgState::from_chunk(image_buffer, (size_t)width * height);
	const int res = make_sort_plan(tile_id_bits(cam.grid_x * cam.grid_y)).npass & 1;

	GradSink sink;
	sink.mean2D = dL_dmean2D; sink.mean2D_stride = 3;
	sink.conic = dL_dconic; sink.conic_stride = 4;
	sink.opacity = dL_dopacity; sink.opacity_stride = 1;
	sink.color = dL_dcolor; sink.color_stride = 3;
	sink.packed = 0;
	int rc;
	if (R > 0)
		if ((rc = launch_render_backward(cam, img.ranges, bin.inst[res], geom.rec, background, img.final_T, img.n_contrib, dL_dpix, sink, stream))) return rc;

	const GaussIn in = make_input(P, D, M, means3D, shs, colors_precomp, nullptr, scales, scale_modifier, rotations, cov3D_precomp);
	GaussGradOut out;
	out.dL_dmeans3D = dL_dmean3D; out.dL_dcov3D = dL_dcov3D; out.dL_dsh = dL_dsh; out.dL_dscales = dL_dscale; out.dL_drots = dL_drot;
	if ((rc = launch_preprocess_backward(in, cam, geom, dL_dmean2D, 3, dL_dconic, 4, dL_dcolor, 3, out, stream))) return rc;
	return 0;
}

int psb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix, unsigned char* present, void* stream_)
{
	(void)projmatrix;
	if (P < 0 || (P > 0 && (!means3D || !viewmatrix || !present))) { set_error_msg("psb_mark_visible: bad argument"); return PSB_ERR_ARG; }
	return launch_mark_visible(P, means3D, viewmatrix, present, (cudaStream_t)stream_);
}

int psb_debug_export(int P, int R, int width, int height, char* geom_buffer, char* binning_buffer, char* image_buffer, float* depths,
                     float* means2D, float* conic_opacity, float* rgb, unsigned char* clamped, uint32_t* tiles_touched, uint64_t* keys_sorted,
                     uint32_t* values_sorted, uint32_t* ranges, uint32_t* n_contrib, float* final_T, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P <= 0) return 0;
	GeomState geom = GeomState::from_chunk(geom_buffer, (size_t)P);
	export_geom_kernel<<<cdiv(P, 256), 256, 0, stream>>>(P, geom, depths, means2D, conic_opacity, rgb, clamped, tiles_touched);
	PSB_LAUNCH_OK();
	const int gx = (width + PSB_TILE_X - 1) / PSB_TILE_X, gy = (height + PSB_TILE_Y - 1) / PSB_TILE_Y;
	if (binning_buffer && R > 0 && (keys_sorted || values_sorted)) {
		BinState bin = BinState::from_chunk(binning_buffer, (size_t)R);
		const int res = make_sort_plan(tile_id_bits(gx * gy)).npass & 1;
		export_keys_kernel<<<cdiv(R, 256), 256, 0, stream>>>(R, bin.tile_key[res], bin.inst[res], geom.rec, keys_sorted, values_sorted);
		PSB_LAUNCH_OK();
	}
	if (image_buffer) {
		ImgState img = ImgState::from_chunk(image_buffer, (size_t)width * height);
		if (ranges) PSB_CUDA_OK(cudaMemcpyAsync(ranges, img.ranges, (size_t)gx * gy * sizeof(uint2), cudaMemcpyDeviceToDevice, stream));
		if (n_contrib) PSB_CUDA_OK(cudaMemcpyAsync(n_contrib, img.n_contrib, (size_t)width * height * sizeof(uint32_t), cudaMemcpyDeviceToDevice, stream));
		if (final_T) PSB_CUDA_OK(cudaMemcpyAsync(final_T, img.final_T, (size_t)width * height * sizeof(float), cudaMemcpyDeviceToDevice, stream));
	}
	return 0;
}

int psb_debug_sort_pairs(uint32_t* keys, uint32_t* vals, size_t n, int nbits, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (n == 0) return 0;
	const SortPlan plan = make_sort_plan(nbits);
	uint32_t *k1 = nullptr, *v1 = nullptr;
	void* scratch = nullptr;
	const size_t sb = sort_scratch_bytes(n, plan.npass);
	PSB_CUDA_OK(cudaMalloc(&k1, n * 4));
	PSB_CUDA_OK(cudaMalloc(&v1, n * 4));
	PSB_CUDA_OK(cudaMalloc(&scratch, sb));
	uint32_t* kk[2] = {keys, k1};
	uint32_t* vv[2] = {vals, v1};
	int rc = radix_sort_pairs(kk, vv, false, nullptr, n, plan, scratch, sb, stream);
	if (rc == 0 && (plan.npass & 1)) {
		cudaMemcpyAsync(keys, k1, n * 4, cudaMemcpyDeviceToDevice, stream);
		cudaMemcpyAsync(vals, v1, n * 4, cudaMemcpyDeviceToDevice, stream);
	}
	cudaStreamSynchronize(stream);
	cudaFree(k1); cudaFree(v1); cudaFree(scratch);
	return rc;
}

}  // extern "C"
