// TEST INFRASTRUCTURE ONLY (oracle/): torch-op façade over the UNMODIFIED reference point operators.
//
// Compiled by oracle/Makefile (target `refpoints`) TOGETHER WITH the reference's own sources, taken where they lie under
// /root/reference: src/operate_points.cu, src/stereo_vision.cu, src/rasterize_points.cu (markVisible, which
// scaleAndTransformThenMarkVisiblePoints calls), linked against oracle/_ref/libref_rasterizer.so (the reference rasterizer core)
// and LibTorch, into oracle/_ref/libref_points.so. It contains no algorithm of its own: every op forwards to
//   transformPoints / scaleAndTransformThenMarkVisiblePoints        reference include/operate_points.h:27-40
//   reprojectDepthPinhole / monocularPinholeInactiveGeoDensify...    reference include/stereo_vision.h:26-40
// so that tests/test_points_gpu.py can pin psb_transform_points / psb_scale_transform_points / psb_reproject_depth_pinhole /
// psb_neighbour_depth_pinhole to the reference's own kernels on the same B200.
#include <torch/library.h>
#include <torch/torch.h>

#include "include/operate_points.h"
#include "include/stereo_vision.h"

static torch::Tensor ref_transform_points(torch::Tensor points, torch::Tensor m)
{
	transformPoints(points, m);
	return points;
}
static std::tuple<torch::Tensor, torch::Tensor, torch::Tensor, int64_t> ref_scale_transform(torch::Tensor points, torch::Tensor rots, torch::Tensor not_transformed,
                                                                                            torch::Tensor unstable, torch::Tensor m, torch::Tensor view,
                                                                                            torch::Tensor proj, int64_t num_transformed, double scale)
{
	int n = (int)num_transformed;
	scaleAndTransformThenMarkVisiblePoints(points, rots, not_transformed, unstable, m, view, proj, n, (float)scale);
	return std::make_tuple(points, rots, not_transformed, (int64_t)n);
}
static torch::Tensor ref_reproject_depth(torch::Tensor depth, torch::Tensor mask, std::vector<double> intr, int64_t width)
{
	std::vector<float> k(intr.begin(), intr.end());
	return reprojectDepthPinhole(depth, mask, k, (int)width);
}
static std::tuple<torch::Tensor, torch::Tensor> ref_neighbour_depth(torch::Tensor kps_pixel, torch::Tensor has3D, torch::Tensor pts_local, torch::Tensor colors,
                                                                   double max_pixel_dist, std::vector<double> intr, int64_t width)
{
	std::vector<float> k(intr.begin(), intr.end());
	return monocularPinholeInactiveGeoDensifyBySearchingNeighborhoodKeypoints(kps_pixel, has3D, pts_local, colors, (float)max_pixel_dist, k, (int)width);
}

TORCH_LIBRARY(psbref, m)
{
	m.def("transform_points", &ref_transform_points);
	m.def("scale_transform_mark_visible", &ref_scale_transform);
	m.def("reproject_depth_pinhole", &ref_reproject_depth);
	m.def("neighbour_depth_pinhole", &ref_neighbour_depth);
}
