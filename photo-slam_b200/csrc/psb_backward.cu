// Per-Gaussian backward: gradients of the 9 screen-space sums produced by the tile kernel
// (dL/dmean2D.xy, dL/dconic.xyw, dL/dopacity, dL/drgb) with respect to the 3D parameters.
// ONE kernel for what the reference runs as computeCov2DCUDA + preprocessCUDA
// (reference cuda_rasterizer/backward.cu:144-274, 346-396, SH part :20-139, cov3D part :278-341):
// the 3D covariance is recomputed in registers instead of being written in forward and re-read, and
// dL/dcov3D never round-trips through memory unless the caller asks for it.
#include "psb_backward.cuh"

namespace psb {

__global__ void __launch_bounds__(128) preprocess_bwd_kernel(GaussIn in, Camera cam, GeomState geom, const float* __restrict__ dL_dmean2D,
                                                             int mean2D_stride, const float* __restrict__ dL_dconic, int conic_stride,
                                                             const float* __restrict__ dL_dcolor, int color_stride, GaussGradOut out)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= in.P || geom.tile_info[idx].z == 0) return;

	const float3 mean = make_float3(in.means3D[3 * idx], in.means3D[3 * idx + 1], in.means3D[3 * idx + 2]);
	float cov3D[6];
	float3 scale = make_float3(0, 0, 0);
	float4 rot = make_float4(0, 0, 0, 0);
	if (in.cov3D_precomp != nullptr) {
#pragma unroll
		for (int i = 0; i < 6; i++) cov3D[i] = in.cov3D_precomp[6 * idx + i];
	} else {
		scale = make_float3(in.scales[3 * idx], in.scales[3 * idx + 1], in.scales[3 * idx + 2]);
		rot = reinterpret_cast<const float4*>(in.rotations)[idx];
		cov3d_from_scale_rot(scale, in.scale_modifier, rot, cov3D);
	}
	const uint32_t clamp_bits = rec_clamp_bits(__float_as_uint(geom.rec[idx].q2.w));
	const float2 g2 = make_float2(dL_dmean2D[(size_t)idx * mean2D_stride], dL_dmean2D[(size_t)idx * mean2D_stride + 1]);
	const float3 gc = make_float3(dL_dconic[(size_t)idx * conic_stride], dL_dconic[(size_t)idx * conic_stride + 1],
	                              dL_dconic[(size_t)idx * conic_stride + 3]);
	const float3 gcol = make_float3(dL_dcolor[(size_t)idx * color_stride], dL_dcolor[(size_t)idx * color_stride + 1],
	                                dL_dcolor[(size_t)idx * color_stride + 2]);
	const float* sh_row = in.shs ? in.shs + (size_t)idx * in.M * 3 : nullptr;
	float* dsh_row = (in.shs && out.dL_dsh) ? out.dL_dsh + (size_t)idx * in.M * 3 : nullptr;

	float3 dL_dmean, dL_dscale = make_float3(0, 0, 0);
	float4 dL_drot = make_float4(0, 0, 0, 0);
	float dL_dcov[6];
	gaussian_backward(in, cam, idx, mean, cov3D, scale, rot, sh_row, clamp_bits, g2, gc, gcol, dsh_row, dL_dmean, dL_dcov, dL_dscale, dL_drot);

	if (out.dL_dmeans3D) { out.dL_dmeans3D[3 * idx] = dL_dmean.x; out.dL_dmeans3D[3 * idx + 1] = dL_dmean.y; out.dL_dmeans3D[3 * idx + 2] = dL_dmean.z; }
	if (out.dL_dcov3D) {
#pragma unroll
		for (int i = 0; i < 6; i++) out.dL_dcov3D[6 * idx + i] = dL_dcov[i];
	}
	if (in.scales) {
		if (out.dL_dscales) { out.dL_dscales[3 * idx] = dL_dscale.x; out.dL_dscales[3 * idx + 1] = dL_dscale.y; out.dL_dscales[3 * idx + 2] = dL_dscale.z; }
		if (out.dL_drots) reinterpret_cast<float4*>(out.dL_drots)[idx] = dL_drot;
	}
}

int launch_preprocess_backward(const GaussIn& in, const Camera& cam, const GeomState& geom, const float* dL_dmean2D, int mean2D_stride,
                               const float* dL_dconic, int conic_stride, const float* dL_dcolor, int color_stride,
                               const GaussGradOut& out, cudaStream_t stream)
{
	if (in.P == 0) return 0;
	preprocess_bwd_kernel<<<cdiv(in.P, 128), 128, 0, stream>>>(in, cam, geom, dL_dmean2D, mean2D_stride, dL_dconic, conic_stride, dL_dcolor,
	                                                           color_stride, out);
	PSB_LAUNCH_OK();
	return 0;
}

}  // namespace psb
