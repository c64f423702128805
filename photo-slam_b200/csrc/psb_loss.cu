// Fused photometric loss of the training step:  L = (1 - lambda) * mean|I - G| + lambda * (1 - mean SSIM(I, G))
// with I = rendered * mask, SSIM over an 11x11 Gaussian window (sigma 1.5), zero padding, per channel —
// the loss of reference gaussian_mapper.cpp:692-698 / include/loss_utils.h:28-124, which the reference runs
// as five grouped cuDNN convolutions plus ~20 elementwise kernels forward and the autograd mirror backward.
// Here: two kernels. Forward computes the five windowed moments with a separable filter in shared memory,
// the SSIM map, the two loss sums, and stores the three partial derivatives of the map that the backward
// needs; backward filters those three maps (the window is symmetric, so the adjoint of the zero-padded
// convolution is the same convolution) and writes dL/d(rendered) directly.
#include "psb_common.cuh"
#include "psb_train.h"

namespace psb {

namespace {

constexpr int LT = 16;          // output tile edge
constexpr int LH = 5;           // window half width
constexpr int LS = LT + 2 * LH; // staged tile edge (26)

struct Win { float w[11]; };

__device__ __forceinline__ float block_sum_256(float v, float* s_red)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	__syncthreads();
	if (lane == 0) s_red[warp] = v;
	__syncthreads();
	float t = 0.f;
#pragma unroll
	for (int i = 0; i < 8; i++) t += s_red[i];
	return t;
}

// grid (tiles_x, tiles_y, 3), block 256 (16x16).
__global__ void __launch_bounds__(256) loss_fwd_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                                                       const float* __restrict__ mask, Win win, float g_ssim,
                                                       float* __restrict__ dmap /* [3 maps][3][H][W] */, double* __restrict__ sums /* [2] */)
{
	__shared__ float s_x[LS][LS + 1], s_y[LS][LS + 1];
	__shared__ float s_h[5][LS][LT + 1];
	__shared__ float s_red[8];
	const int ch = blockIdx.z;
	const size_t HW = (size_t)H * W;
	const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
	const int tid = threadIdx.x;
	for (int i = tid; i < LS * LS; i += 256) {
		const int ly = i / LS, lx = i % LS;
		const int gy = y0 + ly - LH, gx = x0 + lx - LH;
		float a = 0.f, b = 0.f;
		if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
			const size_t p = ch * HW + (size_t)gy * W + gx;
			a = img[p];
			if (mask) a *= mask[p];
			b = gt[p];
		}
		s_x[ly][lx] = a;
		s_y[ly][lx] = b;
	}
	__syncthreads();
	// horizontal pass: 26 rows x 16 columns x 5 moments
	for (int i = tid; i < LS * LT; i += 256) {
		const int ly = i / LT, lx = i % LT;
		float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
		for (int k = 0; k < 11; k++) {
			const float a = s_x[ly][lx + k], b = s_y[ly][lx + k], wk = win.w[k];
			m1 += wk * a; m2 += wk * b; e11 += wk * a * a; e22 += wk * b * b; e12 += wk * a * b;
		}
		s_h[0][ly][lx] = m1; s_h[1][ly][lx] = m2; s_h[2][ly][lx] = e11; s_h[3][ly][lx] = e22; s_h[4][ly][lx] = e12;
	}
	__syncthreads();
	const int lx = tid % LT, ly = tid / LT;
	const int gx = x0 + lx, gy = y0 + ly;
	float l1 = 0.f, ss = 0.f;
	if (gx < W && gy < H) {
		float m1 = 0.f, m2 = 0.f, e11 = 0.f, e22 = 0.f, e12 = 0.f;
#pragma unroll
		for (int k = 0; k < 11; k++) {
			const float wk = win.w[k];
			m1 += wk * s_h[0][ly + k][lx]; m2 += wk * s_h[1][ly + k][lx]; e11 += wk * s_h[2][ly + k][lx];
			e22 += wk * s_h[3][ly + k][lx]; e12 += wk * s_h[4][ly + k][lx];
		}
		const float C1 = 0.01f * 0.01f, C2 = 0.03f * 0.03f;
		const float s1 = e11 - m1 * m1, s2 = e22 - m2 * m2, s12 = e12 - m1 * m2;
		const float A1 = 2.f * m1 * m2 + C1, A2 = 2.f * s12 + C2, B1 = m1 * m1 + m2 * m2 + C1, B2 = s1 + s2 + C2;
		const float inv = 1.f / (B1 * B2);
		const float map = A1 * A2 * inv;
		// partials of the map w.r.t. the three windowed moments that depend on the rendered image
		const float dm_dm1 = (2.f * m2 * A2 - 2.f * m2 * A1) * inv - map * (2.f * m1 / B1 - 2.f * m1 / B2);
		const float dm_de11 = -map / B2;
		const float dm_de12 = 2.f * A1 * inv;
		const size_t p = ch * HW + (size_t)gy * W + gx;
		dmap[p] = g_ssim * dm_dm1;
		dmap[3 * HW + p] = g_ssim * dm_de11;
		dmap[6 * HW + p] = g_ssim * dm_de12;
		ss = map;
		l1 = fabsf(s_x[ly + LH][lx + LH] - s_y[ly + LH][lx + LH]);
	}
	const float tl1 = block_sum_256(l1, s_red);
	const float tss = block_sum_256(ss, s_red);
	if (tid == 0) { atomicAdd(&sums[0], (double)tl1); atomicAdd(&sums[1], (double)tss); }
}

__global__ void __launch_bounds__(256) loss_bwd_kernel(int H, int W, const float* __restrict__ img, const float* __restrict__ gt,
                                                       const float* __restrict__ mask, Win win, float g_l1,
                                                       const float* __restrict__ dmap, float* __restrict__ dL_dimg)
{
	__shared__ float s_d[3][LS][LS + 1];
	__shared__ float s_h[3][LS][LT + 1];
	const int ch = blockIdx.z;
	const size_t HW = (size_t)H * W;
	const int x0 = blockIdx.x * LT, y0 = blockIdx.y * LT;
	const int tid = threadIdx.x;
	for (int i = tid; i < LS * LS; i += 256) {
		const int ly = i / LS, lx = i % LS;
		const int gy = y0 + ly - LH, gx = x0 + lx - LH;
		float a = 0.f, b = 0.f, c = 0.f;
		if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
			const size_t p = ch * HW + (size_t)gy * W + gx;
			a = dmap[p]; b = dmap[3 * HW + p]; c = dmap[6 * HW + p];
		}
		s_d[0][ly][lx] = a; s_d[1][ly][lx] = b; s_d[2][ly][lx] = c;
	}
	__syncthreads();
	for (int i = tid; i < LS * LT; i += 256) {
		const int ly = i / LT, lx = i % LT;
		float a = 0.f, b = 0.f, c = 0.f;
#pragma unroll
		for (int k = 0; k < 11; k++) {
			const float wk = win.w[k];
			a += wk * s_d[0][ly][lx + k]; b += wk * s_d[1][ly][lx + k]; c += wk * s_d[2][ly][lx + k];
		}
		s_h[0][ly][lx] = a; s_h[1][ly][lx] = b; s_h[2][ly][lx] = c;
	}
	__syncthreads();
	const int lx = tid % LT, ly = tid / LT;
	const int gx = x0 + lx, gy = y0 + ly;
	if (gx < W && gy < H) {
		float c1 = 0.f, c2 = 0.f, c3 = 0.f;
#pragma unroll
		for (int k = 0; k < 11; k++) {
			const float wk = win.w[k];
			c1 += wk * s_h[0][ly + k][lx]; c2 += wk * s_h[1][ly + k][lx]; c3 += wk * s_h[2][ly + k][lx];
		}
		const size_t p = ch * HW + (size_t)gy * W + gx;
		const float mk = mask ? mask[p] : 1.f;
		const float x = img[p] * mk, y = gt[p];
		const float d = x - y;
		const float sgn = (d > 0.f) ? 1.f : ((d < 0.f) ? -1.f : 0.f);
		dL_dimg[p] = (g_l1 * sgn + c1 + 2.f * x * c2 + y * c3) * mk;
	}
}

}  // namespace

// sums: device double[2] (zeroed here); after the launch sums[0] = sum|I-G|, sums[1] = sum ssim_map.
// dmap: device float[9*H*W] scratch. dL_dimg may be null (loss value only).
int launch_loss(int H, int W, const float* img, const float* gt, const float* mask, float lambda_dssim, float* dmap, double* sums,
                float* dL_dimg, cudaStream_t stream)
{
	Win win;
	float s = 0.f;
	for (int x = 0; x < 11; x++) { const int t = x - 5; win.w[x] = expf(-(float)(t * t) / (2.0f * 1.5f * 1.5f)); s += win.w[x]; }
	for (int x = 0; x < 11; x++) win.w[x] /= s;
	const float inv_n = 1.0f / (3.0f * (float)H * (float)W);
	PSB_CUDA_OK(cudaMemsetAsync(sums, 0, 2 * sizeof(double), stream));
	dim3 grid((W + LT - 1) / LT, (H + LT - 1) / LT, 3);
	loss_fwd_kernel<<<grid, 256, 0, stream>>>(H, W, img, gt, mask, win, -lambda_dssim * inv_n, dmap, sums);
	PSB_LAUNCH_OK();
	if (dL_dimg) {
		loss_bwd_kernel<<<grid, 256, 0, stream>>>(H, W, img, gt, mask, win, (1.0f - lambda_dssim) * inv_n, dmap, dL_dimg);
		PSB_LAUNCH_OK();
	}
	return 0;
}

}  // namespace psb
