// libcuda_rasterizer drop-in: the reference's torch-typed glue (B1) and raw-pointer class (B2) re-exported with
// the reference's exact C++ signatures, implemented on the psb200 C-ABI.
//
//   RasterizeGaussiansCUDA / RasterizeGaussiansBackwardCUDA / markVisible
//        reference include/rasterize_points.h:18-65 (definitions src/rasterize_points.cu:36-214)
//   CudaRasterizer::Rasterizer::{forward, backward, markVisible}
//        reference cuda_rasterizer/rasterizer.h:24-82
//
// libgaussian_mapper.so links against exactly these symbols; building this file into
// libcuda_rasterizer.so (+ libpsb200.so) swaps the rasterizer without touching any caller. Pure C++/LibTorch host
// code: no kernels here. The TORCH_LIBRARY block at the end only exists so the Python test-suite can call the
// C++ entry points (torch.ops.psb200.*).
#include <functional>
#include <stdexcept>
#include <tuple>
#include <vector>
#include <torch/torch.h>
#include <torch/library.h>
#include <c10/cuda/CUDAStream.h>
#include "../../include/psb200.h"

namespace {

struct TensorSlot { torch::Tensor* t; };

char* resize_cb(size_t n, void* user)
{
	// reference resizeFunctional (src/rasterize_points.cu:28-34)
	auto* t = static_cast<torch::Tensor*>(user);
	t->resize_({(long long)n});
	return reinterpret_cast<char*>(t->contiguous().data_ptr());
}

const float* fptr(const torch::Tensor& t) { return t.numel() ? t.contiguous().data_ptr<float>() : nullptr; }

void* current_stream() { return (void*)c10::cuda::getCurrentCUDAStream().stream(); }

char* fn_cb(size_t n, void* user) { return (*static_cast<std::function<char*(size_t)>*>(user))(n); }

}  // namespace

std::tuple<int, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& colors, const torch::Tensor& opacity,
                       const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier, const torch::Tensor& cov3D_precomp,
                       const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix, const float tan_fovx, const float tan_fovy,
                       const int image_height, const int image_width, const torch::Tensor& sh, const int degree, const torch::Tensor& campos,
                       const bool prefiltered)
{
	if (means3D.ndimension() != 2 || means3D.size(1) != 3) { AT_ERROR("means3D must have dimensions (num_points, 3)"); }
	const int P = means3D.size(0);
	const int H = image_height, W = image_width;
	auto float_opts = means3D.options().dtype(torch::kFloat32);
	torch::Tensor out_color = torch::full({3, H, W}, 0.0, float_opts);
	torch::Tensor radii = torch::full({P}, 0, means3D.options().dtype(torch::kInt32));
	torch::TensorOptions options(torch::kByte);
	torch::Tensor geomBuffer = torch::empty({0}, options.device(means3D.device()));
	torch::Tensor binningBuffer = torch::empty({0}, options.device(means3D.device()));
	torch::Tensor imgBuffer = torch::empty({0}, options.device(means3D.device()));
	int rendered = 0;
	if (P != 0) {
		int M = 0;
		if (sh.size(0) != 0) M = sh.size(1);
		// keep contiguous copies alive for the duration of the call
		const auto bg = background.contiguous(), m3 = means3D.contiguous(), shc = sh.contiguous(), col = colors.contiguous(), op = opacity.contiguous(),
		           sc = scales.contiguous(), rot = rotations.contiguous(), cov = cov3D_precomp.contiguous(), vm = viewmatrix.contiguous(),
		           pm = projmatrix.contiguous(), cp = campos.contiguous();
		rendered = psb_rasterize_forward(resize_cb, &geomBuffer, resize_cb, &binningBuffer, resize_cb, &imgBuffer, P, degree, M, fptr(bg), W, H,
		                                 fptr(m3), fptr(shc), fptr(col), fptr(op), fptr(sc), scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(pm),
		                                 fptr(cp), tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color.data_ptr<float>(), radii.data_ptr<int>(),
		                                 current_stream());
		if (rendered < 0) { AT_ERROR("psb_rasterize_forward: ", psb_last_error()); }
	}
	return std::make_tuple(rendered, out_color, radii, geomBuffer, binningBuffer, imgBuffer);
}

std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor, torch::Tensor>
RasterizeGaussiansBackwardCUDA(const torch::Tensor& background, const torch::Tensor& means3D, const torch::Tensor& radii, const torch::Tensor& colors,
                               const torch::Tensor& scales, const torch::Tensor& rotations, const float scale_modifier,
                               const torch::Tensor& cov3D_precomp, const torch::Tensor& viewmatrix, const torch::Tensor& projmatrix,
                               const float tan_fovx, const float tan_fovy, const torch::Tensor& dL_dout_color, const torch::Tensor& sh,
                               const int degree, const torch::Tensor& campos, const torch::Tensor& geomBuffer, const int R,
                               const torch::Tensor& binningBuffer, const torch::Tensor& imageBuffer)
{
	const int P = means3D.size(0);
	const int H = dL_dout_color.size(1), W = dL_dout_color.size(2);
	int M = 0;
	if (sh.size(0) != 0) M = sh.size(1);
	torch::Tensor dL_dmeans3D = torch::zeros({P, 3}, means3D.options());
	torch::Tensor dL_dmeans2D = torch::zeros({P, 3}, means3D.options());
	torch::Tensor dL_dcolors = torch::zeros({P, 3}, means3D.options());
	torch::Tensor dL_dconic = torch::zeros({P, 2, 2}, means3D.options());
	torch::Tensor dL_dopacity = torch::zeros({P, 1}, means3D.options());
	torch::Tensor dL_dcov3D = torch::zeros({P, 6}, means3D.options());
	torch::Tensor dL_dsh = torch::zeros({P, M, 3}, means3D.options());
	torch::Tensor dL_dscales = torch::zeros({P, 3}, means3D.options());
	torch::Tensor dL_drotations = torch::zeros({P, 4}, means3D.options());
	if (P != 0) {
		const auto bg = background.contiguous(), m3 = means3D.contiguous(), shc = sh.contiguous(), col = colors.contiguous(), sc = scales.contiguous(),
		           rot = rotations.contiguous(), cov = cov3D_precomp.contiguous(), vm = viewmatrix.contiguous(), pm = projmatrix.contiguous(),
		           cp = campos.contiguous(), dpix = dL_dout_color.contiguous(), rad = radii.contiguous();
		const int rc = psb_rasterize_backward(
			P, degree, M, R, fptr(bg), W, H, fptr(m3), fptr(shc), fptr(col), fptr(sc), scale_modifier, fptr(rot), fptr(cov), fptr(vm), fptr(pm), fptr(cp),
			tan_fovx, tan_fovy, rad.data_ptr<int>(), reinterpret_cast<char*>(geomBuffer.contiguous().data_ptr()),
			reinterpret_cast<char*>(binningBuffer.contiguous().data_ptr()), reinterpret_cast<char*>(imageBuffer.contiguous().data_ptr()), fptr(dpix),
			dL_dmeans2D.data_ptr<float>(), dL_dconic.data_ptr<float>(), dL_dopacity.data_ptr<float>(), dL_dcolors.data_ptr<float>(),
			dL_dmeans3D.data_ptr<float>(), dL_dcov3D.data_ptr<float>(), M ? dL_dsh.data_ptr<float>() : nullptr, dL_dscales.data_ptr<float>(),
			dL_drotations.data_ptr<float>(), current_stream());
		if (rc < 0) { AT_ERROR("psb_rasterize_backward: ", psb_last_error()); }
	}
	return std::make_tuple(dL_dmeans2D, dL_dcolors, dL_dopacity, dL_dmeans3D, dL_dcov3D, dL_dsh, dL_dscales, dL_drotations);
}

torch::Tensor markVisible(torch::Tensor& means3D, torch::Tensor& viewmatrix, torch::Tensor& projmatrix)
{
	const int P = means3D.size(0);
	torch::Tensor present = torch::full({P}, false, means3D.options().dtype(at::kBool));
	if (P != 0) {
		const int rc = psb_mark_visible(P, means3D.contiguous().data_ptr<float>(), viewmatrix.contiguous().data_ptr<float>(),
		                                projmatrix.contiguous().data_ptr<float>(), reinterpret_cast<unsigned char*>(present.data_ptr<bool>()),
		                                current_stream());
		if (rc < 0) { AT_ERROR("psb_mark_visible: ", psb_last_error()); }
	}
	return present;
}

// ---- the other torch-typed exports of the reference's libcuda_rasterizer.so / libsimple_knn.so ----
//   transformPoints, scaleAndTransformThenMarkVisiblePoints   include/operate_points.h:27-40, src/operate_points.cu:73-143
//   reprojectDepthPinhole, monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints
//                                                              include/stereo_vision.h:26-40, src/stereo_vision.cu:138-215
//   distCUDA2                                                  third_party/simple-knn/spatial.h:14, spatial.cu:15-26
void transformPoints(torch::Tensor& points, torch::Tensor& transformmatrix)
{
	if (points.ndimension() != 2 || points.size(1) != 3) { AT_ERROR("points must have dimensions (num_points, 3)"); }
	const int P = points.size(0);
	torch::Tensor transformed_points = torch::zeros_like(points);
	if (P != 0) {
		if (psb_transform_points(P, points.contiguous().data_ptr<float>(), transformmatrix.contiguous().data_ptr<float>(),
		                         transformed_points.data_ptr<float>(), current_stream()) < 0) { AT_ERROR(psb_last_error()); }
		points = transformed_points;
	}
}

void scaleAndTransformThenMarkVisiblePoints(torch::Tensor& points, torch::Tensor& rots, torch::Tensor& point_not_transformed_mask,
                                            torch::Tensor& point_unstable_mask, torch::Tensor& transformmatrix, torch::Tensor& viewmatrix,
                                            torch::Tensor& projmatrix, int& num_transformed, const float scale)
{
	if (points.ndimension() != 2 || points.size(1) != 3) { AT_ERROR("points must have dimensions (num_points, 3)"); }
	torch::Tensor present = markVisible(points, viewmatrix, projmatrix);
	auto num_points = present.size(0);
	if (point_not_transformed_mask.size(0) != num_points || point_unstable_mask.size(0) != num_points) { AT_ERROR("points_mask must have dimensions (num_points)"); }
	torch::Tensor final_mask = torch::logical_and(torch::logical_and(point_not_transformed_mask, point_unstable_mask), present);
	num_transformed += final_mask.sum().item<int>();
	const int P = points.size(0);
	if (P != 0) {
		torch::Tensor transformed_points = torch::zeros_like(points);
		torch::Tensor transformed_rots = torch::zeros_like(rots);
		const auto fm = final_mask.contiguous();
		if (psb_scale_transform_points(P, scale, points.contiguous().data_ptr<float>(), rots.contiguous().data_ptr<float>(),
		                               transformmatrix.contiguous().data_ptr<float>(), reinterpret_cast<const unsigned char*>(fm.data_ptr<bool>()),
		                               transformed_points.data_ptr<float>(), transformed_rots.data_ptr<float>(), /*fix_quaternion_write=*/0,
		                               current_stream()) < 0) { AT_ERROR(psb_last_error()); }
		points.index_put_({final_mask}, transformed_points.index({final_mask}));
		rots.index_put_({final_mask}, transformed_rots.index({final_mask}));
		point_not_transformed_mask.index_put_({final_mask}, torch::full({P}, false, point_not_transformed_mask.options()).index({final_mask}));
	}
}

torch::Tensor reprojectDepthPinhole(torch::Tensor& depth, torch::Tensor& mask, std::vector<float>& intr, int width)
{
	if (depth.ndimension() != 1) { AT_ERROR("points must have dimensions (num_points)"); }
	const int P = depth.size(0);
	torch::Tensor points;
	if (P != 0) {
		points = torch::zeros({P, 3}, depth.options());
		const auto mk = mask.contiguous();
		if (psb_reproject_depth_pinhole(P, width, intr[0], intr[1], intr[2], intr[3], depth.contiguous().data_ptr<float>(),
		                                reinterpret_cast<const unsigned char*>(mk.data_ptr<bool>()), points.data_ptr<float>(), current_stream()) < 0) { AT_ERROR(psb_last_error()); }
	}
	return points;
}

std::tuple<torch::Tensor, torch::Tensor> monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(
	torch::Tensor& kps_pixel, torch::Tensor& kps_has3D, torch::Tensor& kps_point_local, torch::Tensor& colors, float max_pixel_dist,
	std::vector<float>& intr, int width)
{
	if (kps_pixel.ndimension() != 2 || kps_pixel.size(1) != 2) AT_ERROR("kps_pixel must have dimensions (num_points, 2)");
	if (kps_has3D.ndimension() != 1) AT_ERROR("kps_has3D must have dimensions (num_points)");
	if (kps_point_local.ndimension() != 2 || kps_point_local.size(1) != 3) AT_ERROR("kps_point_local must have dimensions (num_points, 3)");
	const int N = kps_pixel.size(0);
	torch::Tensor result_pt, result_color;
	if (N != 0) {
		result_pt = torch::zeros_like(kps_point_local);
		result_color = torch::zeros_like(kps_point_local);
		const auto h3 = kps_has3D.contiguous();
		if (psb_neighbour_depth_pinhole(N, width, intr[0], intr[1], intr[2], intr[3], max_pixel_dist, kps_pixel.contiguous().data_ptr<float>(),
		                                reinterpret_cast<const unsigned char*>(h3.data_ptr<bool>()), kps_point_local.contiguous().data_ptr<float>(),
		                                colors.contiguous().data_ptr<float>(), result_pt.data_ptr<float>(), result_color.data_ptr<float>(),
		                                current_stream()) < 0) { AT_ERROR(psb_last_error()); }
		torch::Tensor depth_valid_flags = torch::where(result_pt.index({torch::indexing::Slice(), 2}) > 0.0f, true, false);
		result_pt = result_pt.index({depth_valid_flags});
		result_color = result_color.index({depth_valid_flags});
	}
	return std::make_tuple(result_pt, result_color);
}

torch::Tensor distCUDA2(const torch::Tensor& points)
{
	const int P = points.size(0);
	auto float_opts = points.options().dtype(torch::kFloat32);
	torch::Tensor means = torch::full({P}, 0.0, float_opts);
	if (P != 0) {
		const auto pts = points.contiguous();
		if (psb_dist_cuda2(P, pts.data_ptr<float>(), means.data_ptr<float>(), current_stream()) < 0) { AT_ERROR(psb_last_error()); }
	}
	return means;
}

// ---- B2: the raw-pointer class of cuda_rasterizer/rasterizer.h, same static members, same argument lists ----
namespace CudaRasterizer {
class Rasterizer {
public:
	static void markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present);
	static int forward(std::function<char*(size_t)> geometryBuffer, std::function<char*(size_t)> binningBuffer, std::function<char*(size_t)> imageBuffer,
	                   const int P, int D, int M, const float* background, const int width, int height, const float* means3D, const float* shs,
	                   const float* colors_precomp, const float* opacities, const float* scales, const float scale_modifier, const float* rotations,
	                   const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos, const float tan_fovx,
	                   float tan_fovy, const bool prefiltered, float* out_color, int* radii = nullptr);
	static void backward(const int P, int D, int M, int R, const float* background, const int width, int height, const float* means3D, const float* shs,
	                     const float* colors_precomp, const float* scales, const float scale_modifier, const float* rotations,
	                     const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos, const float tan_fovx,
	                     float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix,
	                     float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
	                     float* dL_dscale, float* dL_drot);
};

void Rasterizer::markVisible(int P, float* means3D, float* viewmatrix, float* projmatrix, bool* present)
{
	if (psb_mark_visible(P, means3D, viewmatrix, projmatrix, reinterpret_cast<unsigned char*>(present), nullptr) < 0) throw std::runtime_error(psb_last_error());
}

int Rasterizer::forward(std::function<char*(size_t)> geometryBuffer, std::function<char*(size_t)> binningBuffer, std::function<char*(size_t)> imageBuffer,
                        const int P, int D, int M, const float* background, const int width, int height, const float* means3D, const float* shs,
                        const float* colors_precomp, const float* opacities, const float* scales, const float scale_modifier, const float* rotations,
                        const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* cam_pos, const float tan_fovx,
                        float tan_fovy, const bool prefiltered, float* out_color, int* radii)
{
	// the reference launches on the legacy default stream (no stream argument anywhere): keep that here
	const int r = psb_rasterize_forward(fn_cb, &geometryBuffer, fn_cb, &binningBuffer, fn_cb, &imageBuffer, P, D, M, background, width, height, means3D,
	                                    shs, colors_precomp, opacities, scales, scale_modifier, rotations, cov3D_precomp, viewmatrix, projmatrix, cam_pos,
	                                    tan_fovx, tan_fovy, prefiltered ? 1 : 0, out_color, radii, nullptr);
	if (r < 0) throw std::runtime_error(psb_last_error());  // reference throws std::runtime_error too (rasterizer_impl.cu:241-244)
	return r;
}

void Rasterizer::backward(const int P, int D, int M, int R, const float* background, const int width, int height, const float* means3D, const float* shs,
                          const float* colors_precomp, const float* scales, const float scale_modifier, const float* rotations,
                          const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos, const float tan_fovx,
                          float tan_fovy, const int* radii, char* geom_buffer, char* binning_buffer, char* image_buffer, const float* dL_dpix,
                          float* dL_dmean2D, float* dL_dconic, float* dL_dopacity, float* dL_dcolor, float* dL_dmean3D, float* dL_dcov3D, float* dL_dsh,
                          float* dL_dscale, float* dL_drot)
{
	if (psb_rasterize_backward(P, D, M, R, background, width, height, means3D, shs, colors_precomp, scales, scale_modifier, rotations, cov3D_precomp,
	                           viewmatrix, projmatrix, campos, tan_fovx, tan_fovy, radii, geom_buffer, binning_buffer, image_buffer, dL_dpix, dL_dmean2D,
	                           dL_dconic, dL_dopacity, dL_dcolor, dL_dmean3D, dL_dcov3D, dL_dsh, dL_dscale, dL_drot, nullptr) < 0)
		throw std::runtime_error(psb_last_error());
}
}  // namespace CudaRasterizer

// ---- test hook: expose the C++ entry points to Python without pybind ----
namespace {
using T = torch::Tensor;
std::tuple<int64_t, T, T, T, T, T> op_fwd(const T& bg, const T& m3, const T& col, const T& op, const T& sc, const T& rot, double smod, const T& cov,
                                          const T& vm, const T& pm, double tfx, double tfy, int64_t H, int64_t W, const T& sh, int64_t deg, const T& cp,
                                          bool pre)
{
	auto r = RasterizeGaussiansCUDA(bg, m3, col, op, sc, rot, (float)smod, cov, vm, pm, (float)tfx, (float)tfy, (int)H, (int)W, sh, (int)deg, cp, pre);
	return std::make_tuple((int64_t)std::get<0>(r), std::get<1>(r), std::get<2>(r), std::get<3>(r), std::get<4>(r), std::get<5>(r));
}
std::tuple<T, T, T, T, T, T, T, T> op_bwd(const T& bg, const T& m3, const T& radii, const T& col, const T& sc, const T& rot, double smod, const T& cov,
                                          const T& vm, const T& pm, double tfx, double tfy, const T& dpix, const T& sh, int64_t deg, const T& cp,
                                          const T& gb, int64_t R, const T& bb, const T& ib)
{
	return RasterizeGaussiansBackwardCUDA(bg, m3, radii, col, sc, rot, (float)smod, cov, vm, pm, (float)tfx, (float)tfy, dpix, sh, (int)deg, cp, gb, (int)R, bb, ib);
}
T op_vis(T m3, T vm, T pm) { return markVisible(m3, vm, pm); }
T op_dist(const T& p) { return distCUDA2(p); }
T op_xform(T p, T m) { transformPoints(p, m); return p; }
T op_reproject(T d, T mk, double fx, double fy, double cx, double cy, int64_t w)
{
	std::vector<float> intr{(float)fx, (float)fy, (float)cx, (float)cy};
	return reprojectDepthPinhole(d, mk, intr, (int)w);
}
std::tuple<T, T> op_neigh(T px, T h3, T pl, T col, double maxd, double fx, double fy, double cx, double cy, int64_t w)
{
	std::vector<float> intr{(float)fx, (float)fy, (float)cx, (float)cy};
	return monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(px, h3, pl, col, (float)maxd, intr, (int)w);
}
// B2 through std::function allocators, for the test of the raw-pointer class
std::tuple<int64_t, T, T> op_b2_fwd(const T& bg, const T& m3, const T& op, const T& sc, const T& rot, const T& vm, const T& pm, double tfx, double tfy,
                                    int64_t H, int64_t W, const T& sh, int64_t deg, const T& cp)
{
	const int P = m3.size(0);
	T out = torch::zeros({3, H, W}, m3.options()), radii = torch::zeros({P}, m3.options().dtype(torch::kInt32));
	T g = torch::empty({0}, m3.options().dtype(torch::kByte)), b = g.clone(), i = g.clone();
	auto mk = [](T& t) { return std::function<char*(size_t)>([&t](size_t n) { t.resize_({(long long)n}); return reinterpret_cast<char*>(t.data_ptr()); }); };
	const int R = CudaRasterizer::Rasterizer::forward(mk(g), mk(b), mk(i), P, (int)deg, (int)sh.size(1), bg.data_ptr<float>(), (int)W, (int)H,
	                                                  m3.data_ptr<float>(), sh.data_ptr<float>(), nullptr, op.data_ptr<float>(), sc.data_ptr<float>(), 1.0f,
	                                                  rot.data_ptr<float>(), nullptr, vm.data_ptr<float>(), pm.data_ptr<float>(), cp.data_ptr<float>(),
	                                                  (float)tfx, (float)tfy, false, out.data_ptr<float>(), radii.data_ptr<int>());
	return std::make_tuple((int64_t)R, out, radii);
}
}  // namespace

TORCH_LIBRARY(psb200, m)
{
	m.def("rasterize_gaussians", &op_fwd);
	m.def("rasterize_gaussians_backward", &op_bwd);
	m.def("mark_visible", &op_vis);
	m.def("b2_forward", &op_b2_fwd);
	m.def("dist_cuda2", &op_dist);
	m.def("transform_points", &op_xform);
	m.def("reproject_depth_pinhole", &op_reproject);
	m.def("neighbour_depth_pinhole", &op_neigh);
}
