// Per-Gaussian backward: gradients of the 9 screen-space sums produced by the tile kernel
// (dL/dmean2D.xy, dL/dconic.xyw, dL/dopacity, dL/drgb) with respect to the 3D parameters.
// ONE kernel for what the reference runs as computeCov2DCUDA + preprocessCUDA
// (reference cuda_rasterizer/backward.cu:144-274, 346-396, SH part :20-139, cov3D part :278-341):
// the 3D covariance is recomputed in registers instead of being written in forward and re-read, and
// dL/dcov3D never round-trips through memory unless the caller asks for it.
#include "psb_geom.cuh"
#include "psb_kernels.h"

namespace psb {

namespace {

// d(normalize(v))/dv applied to dv  (reference auxiliary.h:107-117)
__device__ __forceinline__ float3 dnormvdv(float3 v, float3 dv)
{
	const float sum2 = v.x * v.x + v.y * v.y + v.z * v.z;
	const float invsum32 = 1.0f / sqrt(sum2 * sum2 * sum2);
	float3 o;
	o.x = ((+sum2 - v.x * v.x) * dv.x - v.y * v.x * dv.y - v.z * v.x * dv.z) * invsum32;
	o.y = (-v.x * v.y * dv.x + (sum2 - v.y * v.y) * dv.y - v.z * v.y * dv.z) * invsum32;
	o.z = (-v.x * v.z * dv.x - v.y * v.z * dv.y + (sum2 - v.z * v.z) * dv.z) * invsum32;
	return o;
}

// SH backward for one Gaussian: writes dL/dsh (coefficients of the active degree only) and returns the
// view-direction contribution to dL/dmean.
__device__ __forceinline__ float3 sh_backward(int deg, const float3 pos, const float3 campos, const float* __restrict__ sh,
                                              uint32_t clamp_bits, const float3 dL_dcolor, float* __restrict__ dL_dsh)
{
	const float3 dir_orig = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
	const float len = sqrtf(dir_orig.x * dir_orig.x + dir_orig.y * dir_orig.y + dir_orig.z * dir_orig.z);
	const float x = dir_orig.x / len, y = dir_orig.y / len, z = dir_orig.z / len;
	float g[3] = {dL_dcolor.x, dL_dcolor.y, dL_dcolor.z};
#pragma unroll
	for (int ch = 0; ch < 3; ch++) g[ch] *= ((clamp_bits >> ch) & 1u) ? 0 : 1;
	float dx[3] = {0, 0, 0}, dy[3] = {0, 0, 0}, dz[3] = {0, 0, 0};
#define SHC(k) sh[3 * (k) + ch]
#define DSH(k, w) if (dL_dsh) dL_dsh[3 * (k) + ch] = (w) * g[ch]
#pragma unroll
	for (int ch = 0; ch < 3; ch++) {
		DSH(0, kSH_C0);
		if (deg > 0) {
			DSH(1, -kSH_C1 * y); DSH(2, kSH_C1 * z); DSH(3, -kSH_C1 * x);
			dx[ch] = -kSH_C1 * SHC(3); dy[ch] = -kSH_C1 * SHC(1); dz[ch] = kSH_C1 * SHC(2);
			if (deg > 1) {
				const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
				DSH(4, kSH_C2_0 * xy); DSH(5, kSH_C2_1 * yz); DSH(6, kSH_C2_2 * (2.f * zz - xx - yy));
				DSH(7, kSH_C2_3 * xz); DSH(8, kSH_C2_4 * (xx - yy));
				dx[ch] += kSH_C2_0 * y * SHC(4) + kSH_C2_2 * 2.f * -x * SHC(6) + kSH_C2_3 * z * SHC(7) + kSH_C2_4 * 2.f * x * SHC(8);
				dy[ch] += kSH_C2_0 * x * SHC(4) + kSH_C2_1 * z * SHC(5) + kSH_C2_2 * 2.f * -y * SHC(6) + kSH_C2_4 * 2.f * -y * SHC(8);
				dz[ch] += kSH_C2_1 * y * SHC(5) + kSH_C2_2 * 2.f * 2.f * z * SHC(6) + kSH_C2_3 * x * SHC(7);
				if (deg > 2) {
					DSH(9, kSH_C3_0 * y * (3.f * xx - yy)); DSH(10, kSH_C3_1 * xy * z);
					DSH(11, kSH_C3_2 * y * (4.f * zz - xx - yy)); DSH(12, kSH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy));
					DSH(13, kSH_C3_4 * x * (4.f * zz - xx - yy)); DSH(14, kSH_C3_5 * z * (xx - yy));
					DSH(15, kSH_C3_6 * x * (xx - 3.f * yy));
					dx[ch] += (kSH_C3_0 * SHC(9) * 3.f * 2.f * xy + kSH_C3_1 * SHC(10) * yz + kSH_C3_2 * SHC(11) * -2.f * xy +
					           kSH_C3_3 * SHC(12) * -3.f * 2.f * xz + kSH_C3_4 * SHC(13) * (-3.f * xx + 4.f * zz - yy) +
					           kSH_C3_5 * SHC(14) * 2.f * xz + kSH_C3_6 * SHC(15) * 3.f * (xx - yy));
					dy[ch] += (kSH_C3_0 * SHC(9) * 3.f * (xx - yy) + kSH_C3_1 * SHC(10) * xz +
					           kSH_C3_2 * SHC(11) * (-3.f * yy + 4.f * zz - xx) + kSH_C3_3 * SHC(12) * -3.f * 2.f * yz +
					           kSH_C3_4 * SHC(13) * -2.f * xy + kSH_C3_5 * SHC(14) * -2.f * yz + kSH_C3_6 * SHC(15) * -3.f * 2.f * xy);
					dz[ch] += (kSH_C3_1 * SHC(10) * xy + kSH_C3_2 * SHC(11) * 4.f * 2.f * yz +
					           kSH_C3_3 * SHC(12) * 3.f * (2.f * zz - xx - yy) + kSH_C3_4 * SHC(13) * 4.f * 2.f * xz +
					           kSH_C3_5 * SHC(14) * (xx - yy));
				}
			}
		}
	}
#undef SHC
#undef DSH
	const float3 dL_ddir = make_float3(dx[0] * g[0] + dx[1] * g[1] + dx[2] * g[2], dy[0] * g[0] + dy[1] * g[1] + dy[2] * g[2],
	                                   dz[0] * g[0] + dz[1] * g[1] + dz[2] * g[2]);
	return dnormvdv(dir_orig, dL_ddir);
}

// (scale, quaternion) <- dL/dSigma  (reference backward.cu:278-341)
__device__ __forceinline__ void cov3d_backward(const float3 scale, float mod, const float4 rot, const float* dL_dcov3D, float3& dL_dscale,
                                               float4& dL_drot)
{
	const float r = rot.x, x = rot.y, y = rot.z, z = rot.w;
	const Mat3 R = quat_to_mat3(rot);
	const float3 s = make_float3(mod * scale.x, mod * scale.y, mod * scale.z);
	const Mat3 S = diag3(s.x, s.y, s.z);
	const Mat3 M = mat3_mul(S, R);
	Mat3 dL_dSigma;
	dL_dSigma(0, 0) = dL_dcov3D[0]; dL_dSigma(0, 1) = 0.5f * dL_dcov3D[1]; dL_dSigma(0, 2) = 0.5f * dL_dcov3D[2];
	dL_dSigma(1, 0) = 0.5f * dL_dcov3D[1]; dL_dSigma(1, 1) = dL_dcov3D[3]; dL_dSigma(1, 2) = 0.5f * dL_dcov3D[4];
	dL_dSigma(2, 0) = 0.5f * dL_dcov3D[2]; dL_dSigma(2, 1) = 0.5f * dL_dcov3D[4]; dL_dSigma(2, 2) = dL_dcov3D[5];
	Mat3 M2;
#pragma unroll
	for (int i = 0; i < 9; i++) M2.m[i] = M.m[i] * 2.0f;
	const Mat3 dL_dM = mat3_mul(M2, dL_dSigma);
	const Mat3 Rt = mat3_transpose(R);
	Mat3 dL_dMt = mat3_transpose(dL_dM);
	dL_dscale.x = Rt(0, 0) * dL_dMt(0, 0) + Rt(0, 1) * dL_dMt(0, 1) + Rt(0, 2) * dL_dMt(0, 2);
	dL_dscale.y = Rt(1, 0) * dL_dMt(1, 0) + Rt(1, 1) * dL_dMt(1, 1) + Rt(1, 2) * dL_dMt(1, 2);
	dL_dscale.z = Rt(2, 0) * dL_dMt(2, 0) + Rt(2, 1) * dL_dMt(2, 1) + Rt(2, 2) * dL_dMt(2, 2);
#pragma unroll
	for (int rr = 0; rr < 3; rr++) { dL_dMt(0, rr) *= s.x; dL_dMt(1, rr) *= s.y; dL_dMt(2, rr) *= s.z; }
	dL_drot.x = 2 * z * (dL_dMt(0, 1) - dL_dMt(1, 0)) + 2 * y * (dL_dMt(2, 0) - dL_dMt(0, 2)) + 2 * x * (dL_dMt(1, 2) - dL_dMt(2, 1));
	dL_drot.y = 2 * y * (dL_dMt(1, 0) + dL_dMt(0, 1)) + 2 * z * (dL_dMt(2, 0) + dL_dMt(0, 2)) + 2 * r * (dL_dMt(1, 2) - dL_dMt(2, 1)) -
	            4 * x * (dL_dMt(2, 2) + dL_dMt(1, 1));
	dL_drot.z = 2 * x * (dL_dMt(1, 0) + dL_dMt(0, 1)) + 2 * r * (dL_dMt(2, 0) - dL_dMt(0, 2)) + 2 * z * (dL_dMt(1, 2) + dL_dMt(2, 1)) -
	            4 * y * (dL_dMt(2, 2) + dL_dMt(0, 0));
	dL_drot.w = 2 * r * (dL_dMt(0, 1) - dL_dMt(1, 0)) + 2 * x * (dL_dMt(2, 0) + dL_dMt(0, 2)) + 2 * y * (dL_dMt(1, 2) + dL_dMt(2, 1)) -
	            4 * z * (dL_dMt(1, 1) + dL_dMt(0, 0));
}

}  // namespace

// Everything between the 9 screen-space sums and the activated 3D parameters, for one visible Gaussian.
// Returns dL/dmean3D, dL/dcov3D[6]; dL/dscale, dL/drot if scale/rot given; writes dL/dsh through the pointer.
__device__ __forceinline__ void gaussian_backward(const GaussIn& in, const Camera& cam, int idx, const float3 mean, const float* cov3D,
                                                  const float3 scale, const float4 rot, const float* __restrict__ sh_row,
                                                  uint32_t clamp_bits, const float2 dL_dmean2D, const float3 dL_dconic,
                                                  const float3 dL_dcolor, float* __restrict__ dL_dsh_row, float3& dL_dmean, float* dL_dcov,
                                                  float3& dL_dscale, float4& dL_drot)
{
	Cov2DTerms ct;
	cov2d_terms(mean, cam.focal_x, cam.focal_y, cam.tan_fovx, cam.tan_fovy, cov3D, cam.view, ct);
	const Mat3& T = ct.T;
	const Mat3& Vrk = ct.Vrk;
	const float3 t = ct.t;
	const float h_x = cam.focal_x, h_y = cam.focal_y;
	const float a = ct.cov(0, 0) + 0.3f;
	const float b = ct.cov(0, 1);
	const float c = ct.cov(1, 1) + 0.3f;
	const float denom = a * c - b * b;
	float dL_da = 0, dL_db = 0, dL_dc = 0;
	const float denom2inv = 1.0f / ((denom * denom) + 0.0000001f);
	if (denom2inv != 0) {
		dL_da = denom2inv * (-c * c * dL_dconic.x + 2 * b * c * dL_dconic.y + (denom - a * c) * dL_dconic.z);
		dL_dc = denom2inv * (-a * a * dL_dconic.z + 2 * a * b * dL_dconic.y + (denom - a * c) * dL_dconic.x);
		dL_db = denom2inv * 2 * (b * c * dL_dconic.x - (denom + 2 * b * b) * dL_dconic.y + a * b * dL_dconic.z);
		dL_dcov[0] = (T(0, 0) * T(0, 0) * dL_da + T(0, 0) * T(1, 0) * dL_db + T(1, 0) * T(1, 0) * dL_dc);
		dL_dcov[3] = (T(0, 1) * T(0, 1) * dL_da + T(0, 1) * T(1, 1) * dL_db + T(1, 1) * T(1, 1) * dL_dc);
		dL_dcov[5] = (T(0, 2) * T(0, 2) * dL_da + T(0, 2) * T(1, 2) * dL_db + T(1, 2) * T(1, 2) * dL_dc);
		dL_dcov[1] = 2 * T(0, 0) * T(0, 1) * dL_da + (T(0, 0) * T(1, 1) + T(0, 1) * T(1, 0)) * dL_db + 2 * T(1, 0) * T(1, 1) * dL_dc;
		dL_dcov[2] = 2 * T(0, 0) * T(0, 2) * dL_da + (T(0, 0) * T(1, 2) + T(0, 2) * T(1, 0)) * dL_db + 2 * T(1, 0) * T(1, 2) * dL_dc;
		dL_dcov[4] = 2 * T(0, 2) * T(0, 1) * dL_da + (T(0, 1) * T(1, 2) + T(0, 2) * T(1, 1)) * dL_db + 2 * T(1, 1) * T(1, 2) * dL_dc;
	} else {
#pragma unroll
		for (int i = 0; i < 6; i++) dL_dcov[i] = 0;
	}
	const float dL_dT00 = 2 * (T(0, 0) * Vrk(0, 0) + T(0, 1) * Vrk(0, 1) + T(0, 2) * Vrk(0, 2)) * dL_da +
	                      (T(1, 0) * Vrk(0, 0) + T(1, 1) * Vrk(0, 1) + T(1, 2) * Vrk(0, 2)) * dL_db;
	const float dL_dT01 = 2 * (T(0, 0) * Vrk(1, 0) + T(0, 1) * Vrk(1, 1) + T(0, 2) * Vrk(1, 2)) * dL_da +
	                      (T(1, 0) * Vrk(1, 0) + T(1, 1) * Vrk(1, 1) + T(1, 2) * Vrk(1, 2)) * dL_db;
	const float dL_dT02 = 2 * (T(0, 0) * Vrk(2, 0) + T(0, 1) * Vrk(2, 1) + T(0, 2) * Vrk(2, 2)) * dL_da +
	                      (T(1, 0) * Vrk(2, 0) + T(1, 1) * Vrk(2, 1) + T(1, 2) * Vrk(2, 2)) * dL_db;
	const float dL_dT10 = 2 * (T(1, 0) * Vrk(0, 0) + T(1, 1) * Vrk(0, 1) + T(1, 2) * Vrk(0, 2)) * dL_dc +
	                      (T(0, 0) * Vrk(0, 0) + T(0, 1) * Vrk(0, 1) + T(0, 2) * Vrk(0, 2)) * dL_db;
	const float dL_dT11 = 2 * (T(1, 0) * Vrk(1, 0) + T(1, 1) * Vrk(1, 1) + T(1, 2) * Vrk(1, 2)) * dL_dc +
	                      (T(0, 0) * Vrk(1, 0) + T(0, 1) * Vrk(1, 1) + T(0, 2) * Vrk(1, 2)) * dL_db;
	const float dL_dT12 = 2 * (T(1, 0) * Vrk(2, 0) + T(1, 1) * Vrk(2, 1) + T(1, 2) * Vrk(2, 2)) * dL_dc +
	                      (T(0, 0) * Vrk(2, 0) + T(0, 1) * Vrk(2, 1) + T(0, 2) * Vrk(2, 2)) * dL_db;
	// W(c, r): c-th column of the view rotation as filled in cov2d_terms
	const float* vm = cam.view;
	const float W00 = vm[0], W01 = vm[4], W02 = vm[8], W10 = vm[1], W11 = vm[5], W12 = vm[9], W20 = vm[2], W21 = vm[6], W22 = vm[10];
	const float dL_dJ00 = W00 * dL_dT00 + W01 * dL_dT01 + W02 * dL_dT02;
	const float dL_dJ02 = W20 * dL_dT00 + W21 * dL_dT01 + W22 * dL_dT02;
	const float dL_dJ11 = W10 * dL_dT10 + W11 * dL_dT11 + W12 * dL_dT12;
	const float dL_dJ12 = W20 * dL_dT10 + W21 * dL_dT11 + W22 * dL_dT12;
	const float tz = 1.f / t.z;
	const float tz2 = tz * tz;
	const float tz3 = tz2 * tz;
	const float dL_dtx = ct.x_grad_mul * -h_x * tz2 * dL_dJ02;
	const float dL_dty = ct.y_grad_mul * -h_y * tz2 * dL_dJ12;
	const float dL_dtz = -h_x * tz2 * dL_dJ00 - h_y * tz2 * dL_dJ11 + (2 * h_x * t.x) * tz3 * dL_dJ02 + (2 * h_y * t.y) * tz3 * dL_dJ12;
	dL_dmean = make_float3(vm[0] * dL_dtx + vm[1] * dL_dty + vm[2] * dL_dtz, vm[4] * dL_dtx + vm[5] * dL_dty + vm[6] * dL_dtz,
	                       vm[8] * dL_dtx + vm[9] * dL_dty + vm[10] * dL_dtz);

	const float* proj = cam.proj;
	const float3 m = mean;
	const float4 m_hom = xform4x4(m, proj);
	const float m_w = 1.0f / (m_hom.w + 0.0000001f);
	const float mul1 = (proj[0] * m.x + proj[4] * m.y + proj[8] * m.z + proj[12]) * m_w * m_w;
	const float mul2 = (proj[1] * m.x + proj[5] * m.y + proj[9] * m.z + proj[13]) * m_w * m_w;
	float3 dm2;
	dm2.x = (proj[0] * m_w - proj[3] * mul1) * dL_dmean2D.x + (proj[1] * m_w - proj[3] * mul2) * dL_dmean2D.y;
	dm2.y = (proj[4] * m_w - proj[7] * mul1) * dL_dmean2D.x + (proj[5] * m_w - proj[7] * mul2) * dL_dmean2D.y;
	dm2.z = (proj[8] * m_w - proj[11] * mul1) * dL_dmean2D.x + (proj[9] * m_w - proj[11] * mul2) * dL_dmean2D.y;
	dL_dmean.x += dm2.x; dL_dmean.y += dm2.y; dL_dmean.z += dm2.z;

	if (sh_row) {
		const float3 campos = make_float3(cam.campos[0], cam.campos[1], cam.campos[2]);
		const float3 dsh = sh_backward(in.D, m, campos, sh_row, clamp_bits, dL_dcolor, dL_dsh_row);
		dL_dmean.x += dsh.x; dL_dmean.y += dsh.y; dL_dmean.z += dsh.z;
	}
	if (in.scales) cov3d_backward(scale, in.scale_modifier, rot, dL_dcov, dL_dscale, dL_drot);
}

__global__ void __launch_bounds__(128) preprocess_bwd_kernel(GaussIn in, Camera cam, GeomState geom, const float* __restrict__ dL_dmean2D,
                                                             int mean2D_stride, const float* __restrict__ dL_dconic, int conic_stride,
                                                             const float* __restrict__ dL_dcolor, int color_stride, GaussGradOut out)
{
	const int idx = blockIdx.x * blockDim.x + threadIdx.x;
	if (idx >= in.P || geom.tiles_touched[idx] == 0) return;

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
	const uint32_t clamp_bits = __float_as_uint(geom.rec[idx].q2.w);
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
