// Fused densify / clone / split / prune (declared in include/psb200.h).
//
// Replaces reference GaussianModel::densifyAndPrune and everything under it (src/gaussian_model.cpp:795-815 driver,
// :763-793 clone, :716-761 split, :588-642 prunePoints, :644-714 densificationPostfix) — in the reference dozens of
// LibTorch launches, three re-materialisations of all 59 parameters + 118 Adam moments per Gaussian (cat, cat, mask-index,
// mask-index) and a blocking `.item()` — by three launches that read every surviving row ONCE and write every output row ONCE:
//
//   densify_mask_kernel     per Gaussian: the reference's three predicates evaluated on the ORIGINAL rows
//                             clone   |grad| >= tau  &&  max exp(scaling) <= percent_dense * extent              (:769-774)
//                             split   grad  >= tau  &&  max exp(scaling) >  percent_dense * extent              (:726-731)
//                             prune   sigmoid(opacity) < min_opacity || (max_screen_size && max exp(s) > 0.1 extent)  (:806-811)
//                           (the max_radii2D > max_screen_size term of :808 can never fire in the reference: densificationPostfix
//                            has just reset max_radii2D_ to zeros, :711 — reproduced by not evaluating it), composed into what
//                           survives of each row's up to four output rows {itself, its clone, split child 0, split child 1};
//                           per-block counts.
//   densify_scan_kernel     one block: exclusive scan of the block counts -> block offsets + totals (the only value the host
//                           needs: the new row count, to size the output tensors).
//   densify_scatter_kernel  per block of 128 source rows: destination rows from the scan, then every parameter / moment tensor
//                           is streamed through once (coalesced reads, run-coalesced writes). Output order = the reference's:
//                           [surviving originals | clones | split children copy 0 | split children copy 1], each in source
//                           order (cat order of :666-681 after the two prunePoints calls). New rows get zero moments (:675-676);
//                           split children get xyz = R(q) (z * exp(s)) + xyz and scaling = log(exp(s) / 1.6) (:733-738), with z
//                           either injected (tests: same numbers as the oracle) or drawn from Philox4x32-10 + Box-Muller keyed by
//                           (seed, offset, child row) — counter-based, so every data-parallel replica draws identical samples.
#include <cmath>
#include <cstring>
#include "psb_common.cuh"
#include "../../include/psb200.h"

namespace psb {

namespace {

constexpr int DN_TB = 128;
constexpr uint32_t F_ORIG = 1u, F_CLONE = 2u, F_CHILD = 4u, F_SPLIT = 8u;
__host__ __device__ constexpr int dn_row(int ti) { return ti == 2 ? 45 : (ti == 3 ? 1 : (ti == 5 ? 4 : 3)); }  // floats per row of xyz, f_dc, f_rest, opacity, scaling, rotation

struct DensifyCfg {
	float tau, min_opacity, size_thr, ws_thr;
	int ws_prune;
};

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__global__ void __launch_bounds__(DN_TB) densify_mask_kernel(int P, const float* __restrict__ scaling, const float* __restrict__ opacity,
                                                           const float* __restrict__ accum, const float* __restrict__ denom, DensifyCfg cfg,
                                                           uint8_t* __restrict__ flags, uint32_t* __restrict__ block_counts)
{
	__shared__ uint32_t s_cnt[DN_TB / 32][4];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int i = blockIdx.x * DN_TB + tid;
	uint32_t f = 0;
	if (i < P) {
		float g = accum[i] / denom[i];
		if (isnan(g)) g = 0.0f;
		const float s0 = expf(scaling[3 * i]), s1 = expf(scaling[3 * i + 1]), s2 = expf(scaling[3 * i + 2]);
		const float smax = fmaxf(s0, fmaxf(s1, s2));
		const float act = sigmoidf(opacity[i]);
		const bool sel_c = sqrtf(__fmul_rn(g, g)) >= cfg.tau && smax <= cfg.size_thr;
		const bool sel_s = g >= cfg.tau && smax > cfg.size_thr;
		const bool prune_self = act < cfg.min_opacity || (cfg.ws_prune && smax > cfg.ws_thr);
		// children carry scaling = log(exp(s) * (1/1.6)) (ATen divides by a CPU scalar through its reciprocal); the prune test sees exp() of that
		const float inv = 1.0f / 1.6f;
		const float c0 = expf(logf(__fmul_rn(s0, inv))), c1 = expf(logf(__fmul_rn(s1, inv))), c2 = expf(logf(__fmul_rn(s2, inv)));
		const bool prune_child = act < cfg.min_opacity || (cfg.ws_prune && fmaxf(c0, fmaxf(c1, c2)) > cfg.ws_thr);
		if (!sel_s && !prune_self) f |= F_ORIG;
		if (sel_c && !prune_self) f |= F_CLONE;
		if (sel_s) f |= F_SPLIT;
		if (sel_s && !prune_child) f |= F_CHILD;
		flags[i] = (uint8_t)f;
	}
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t b = __ballot_sync(0xffffffffu, (f >> k) & 1u);
		if (lane == 0) s_cnt[warp][k] = __popc(b);
	}
	__syncthreads();
	if (tid < 4) {
		uint32_t t = 0;
#pragma unroll
		for (int w = 0; w < DN_TB / 32; w++) t += s_cnt[w][tid];
		block_counts[4 * blockIdx.x + tid] = t;
	}
}

// exclusive scan of [nblocks][4] counts by ONE block; totals[0..3] = sums
__global__ void __launch_bounds__(1024) densify_scan_kernel(int nblocks, const uint32_t* __restrict__ counts, uint32_t* __restrict__ offsets,
                                                          uint32_t* __restrict__ totals)
{
	__shared__ uint32_t s_w[32][4];
	__shared__ uint32_t s_run[4];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < 4) s_run[tid] = 0;
	__syncthreads();
	for (int b0 = 0; b0 < nblocks; b0 += 1024) {
		const int b = b0 + tid;
		uint32_t v[4], inc[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			v[k] = b < nblocks ? counts[4 * b + k] : 0u;
			inc[k] = v[k];
#pragma unroll
			for (int o = 1; o < 32; o <<= 1) {
				const uint32_t u = __shfl_up_sync(0xffffffffu, inc[k], o);
				if (lane >= o) inc[k] += u;
			}
			if (lane == 31) s_w[warp][k] = inc[k];
		}
		__syncthreads();
		uint32_t base[4];
#pragma unroll
		for (int k = 0; k < 4; k++) {
			uint32_t wb = 0;
			for (int w = 0; w < warp; w++) wb += s_w[w][k];
			base[k] = s_run[k] + wb + inc[k] - v[k];
			if (b < nblocks) offsets[4 * b + k] = base[k];
		}
		__syncthreads();
		if (tid == 1023) {
#pragma unroll
			for (int k = 0; k < 4; k++) s_run[k] = base[k] + v[k];
		}
		__syncthreads();
	}
	if (tid < 4) totals[tid] = s_run[tid];
}

// ---- Philox4x32-10 (Salmon et al. 2011), the counter-based generator cuRAND / ATen use; written out: no cuRAND dependency
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
	constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
	for (int r = 0; r < 10; r++) {
		const uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
		const uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
		ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
		key.x += W0; key.y += W1;
	}
	return ctr;
}
// three standard normals for child row j: Box-Muller on the four 32-bit words of Philox(counter = (j, 0, offset), key = seed)
__device__ __forceinline__ float3 normal3(unsigned long long seed, unsigned long long offset, uint32_t j)
{
	const uint4 r = philox4x32_10(make_uint4(j, 0u, (uint32_t)offset, (uint32_t)(offset >> 32)), make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
	const float k = 2.3283064365386963e-10f;  // 2^-32
	const float u0 = ((float)r.x + 0.5f) * k, u1 = ((float)r.y + 0.5f) * k, u2 = ((float)r.z + 0.5f) * k, u3 = ((float)r.w + 0.5f) * k;
	const float ra = sqrtf(-2.0f * logf(fminf(u0, 0.99999994f))), rb = sqrtf(-2.0f * logf(fminf(u2, 0.99999994f)));
	float sa, ca, sb, cb;
	sincosf(6.283185307179586f * u1, &sa, &ca);
	sincosf(6.283185307179586f * u3, &sb, &cb);
	(void)sb;
	return make_float3(ra * ca, ra * sa, rb * cb);
}

struct DensifyTensors {
	const float* p[6]; const float* m[6]; const float* v[6];   // source [P]
	float* dp[6]; float* dm[6]; float* dv[6];                   // destination [P_new]
	const int* exist; int* dexist;                              // exist_since_iter (optional): children / clones inherit the parent's (:741, :781)
	const float* stat[3]; float* dstat[3];                      // statistics, compacted only by prunePoints (null otherwise)
};

__global__ void __launch_bounds__(DN_TB) densify_scatter_kernel(int P, DensifyTensors t, const uint8_t* __restrict__ flags,
                                                              const uint32_t* __restrict__ block_offsets, const uint32_t* __restrict__ totals,
                                                              const float* __restrict__ samples, unsigned long long seed, unsigned long long offset)
{
	__shared__ int s_dst[4][DN_TB];
	__shared__ float s_cxyz[2][DN_TB][3];
	__shared__ uint32_t s_wcnt[DN_TB / 32][4];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int base = blockIdx.x * DN_TB;
	const int rows = min(DN_TB, P - base);
	const int i = base + tid;
	const uint32_t f = i < P ? flags[i] : 0u;
	uint32_t pre[4];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		const uint32_t b = __ballot_sync(0xffffffffu, (f >> k) & 1u);
		pre[k] = __popc(b & ((1u << lane) - 1u));
		if (lane == 0) s_wcnt[warp][k] = __popc(b);
	}
	__syncthreads();
	const uint32_t K0 = totals[0], K1 = totals[1], K2 = totals[2], NS = totals[3];
#pragma unroll
	for (int k = 0; k < 4; k++) {
		for (int w = 0; w < warp; w++) pre[k] += s_wcnt[w][k];
		pre[k] += block_offsets[4 * blockIdx.x + k];
	}
	s_dst[0][tid] = (f & F_ORIG) ? (int)pre[0] : -1;
	s_dst[1][tid] = (f & F_CLONE) ? (int)(K0 + pre[1]) : -1;
	s_dst[2][tid] = (f & F_CHILD) ? (int)(K0 + K1 + pre[2]) : -1;
	s_dst[3][tid] = (f & F_CHILD) ? (int)(K0 + K1 + K2 + pre[2]) : -1;
	if (f & F_CHILD) {
		// reference :733-736: samples = normal(0, exp(scaling)); new_xyz = build_rotation(rotation) @ samples + xyz
		const float4 qr = reinterpret_cast<const float4*>(t.p[5])[i];
		const float n = sqrtf(qr.x * qr.x + qr.y * qr.y + qr.z * qr.z + qr.w * qr.w);
		const float r = qr.x / n, x = qr.y / n, y = qr.z / n, z = qr.w / n;
		const float R00 = 1.f - 2.f * (y * y + z * z), R01 = 2.f * (x * y - r * z), R02 = 2.f * (x * z + r * y);
		const float R10 = 2.f * (x * y + r * z), R11 = 1.f - 2.f * (x * x + z * z), R12 = 2.f * (y * z - r * x);
		const float R20 = 2.f * (x * z - r * y), R21 = 2.f * (y * z + r * x), R22 = 1.f - 2.f * (x * x + y * y);
		const float s0 = expf(t.p[4][3 * i]), s1 = expf(t.p[4][3 * i + 1]), s2 = expf(t.p[4][3 * i + 2]);
		const float px = t.p[0][3 * i], py = t.p[0][3 * i + 1], pz = t.p[0][3 * i + 2];
#pragma unroll
		for (int c = 0; c < 2; c++) {
			const uint32_t j = (uint32_t)c * NS + pre[3];  // row of the [2 * n_split, 3] sample block (repeat({N,1}) order)
			float3 zz;
			if (samples) zz = make_float3(samples[3 * (size_t)j], samples[3 * (size_t)j + 1], samples[3 * (size_t)j + 2]);
			else zz = normal3(seed, offset, j);
			const float a = zz.x * s0, b = zz.y * s1, d = zz.z * s2;
			s_cxyz[c][tid][0] = R00 * a + R01 * b + R02 * d + px;
			s_cxyz[c][tid][1] = R10 * a + R11 * b + R12 * d + py;
			s_cxyz[c][tid][2] = R20 * a + R21 * b + R22 * d + pz;
		}
	}
	__syncthreads();

	if (i < P) {
		if (t.exist) {
			const int ex = t.exist[i];
#pragma unroll
			for (int o = 0; o < 4; o++) if (s_dst[o][tid] >= 0) t.dexist[s_dst[o][tid]] = ex;
		}
		if (t.stat[0] && s_dst[0][tid] >= 0) {
#pragma unroll
			for (int q = 0; q < 3; q++) t.dstat[q][s_dst[0][tid]] = t.stat[q][i];
		}
	}

	const float inv = 1.0f / 1.6f;
#pragma unroll
	for (int ti = 0; ti < 6; ti++) {
		const int k = dn_row(ti);
		const float* __restrict__ sp = t.p[ti] + (size_t)base * k;
		const float* __restrict__ sm = t.m[ti] + (size_t)base * k;
		const float* __restrict__ sv = t.v[ti] + (size_t)base * k;
		for (int e = tid; e < rows * k; e += DN_TB) {
			const int r = e / k, c = e - r * k;
			const int d0 = s_dst[0][r], d1 = s_dst[1][r], d2 = s_dst[2][r], d3 = s_dst[3][r];
			if (d0 < 0 && d1 < 0 && d2 < 0) continue;  // row vanishes (pruned, or a split parent whose children are pruned)
			const float pv = sp[e];
			if (d0 >= 0) {
				const size_t o = (size_t)d0 * k + c;
				t.dp[ti][o] = pv; t.dm[ti][o] = sm[e]; t.dv[ti][o] = sv[e];
			}
			if (d1 >= 0) {
				const size_t o = (size_t)d1 * k + c;
				t.dp[ti][o] = pv; t.dm[ti][o] = 0.f; t.dv[ti][o] = 0.f;
			}
			if (d2 >= 0) {
				float v0 = pv, v1 = pv;
				if (ti == 0) { v0 = s_cxyz[0][r][c]; v1 = s_cxyz[1][r][c]; }
				else if (ti == 4) { v0 = v1 = logf(__fmul_rn(expf(pv), inv)); }
				const size_t o0 = (size_t)d2 * k + c, o1 = (size_t)d3 * k + c;
				t.dp[ti][o0] = v0; t.dm[ti][o0] = 0.f; t.dv[ti][o0] = 0.f;
				t.dp[ti][o1] = v1; t.dm[ti][o1] = 0.f; t.dv[ti][o1] = 0.f;
			}
		}
	}
}

// reference GaussianModel::resetOpacity (src/gaussian_model.cpp:556-565): inverse_sigmoid(min(sigmoid(o), ones_like(...))) — the
// misplaced parenthesis makes the clamp a no-op — and fresh (zero) Adam moments for the opacity group (:576-578)
__global__ void reset_opacity_kernel(int P, float* __restrict__ opacity, float* __restrict__ m, float* __restrict__ v)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= P) return;
	const float a = fminf(sigmoidf(opacity[i]), 1.0f);
	opacity[i] = logf(a / (1.0f - a));
	m[i] = 0.f; v[i] = 0.f;
}

// prunePoints: flags from a byte mask
__global__ void __launch_bounds__(DN_TB) prune_mask_kernel(int P, const uint8_t* __restrict__ mask, uint8_t* __restrict__ flags, uint32_t* __restrict__ block_counts)
{
	__shared__ uint32_t s_cnt[DN_TB / 32];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int i = blockIdx.x * DN_TB + tid;
	const uint32_t f = (i < P && !mask[i]) ? F_ORIG : 0u;
	if (i < P) flags[i] = (uint8_t)f;
	const uint32_t b = __ballot_sync(0xffffffffu, f & 1u);
	if (lane == 0) s_cnt[warp] = __popc(b);
	__syncthreads();
	if (tid < 4) {
		uint32_t t = 0;
		if (tid == 0) for (int w = 0; w < DN_TB / 32; w++) t += s_cnt[w];
		block_counts[4 * blockIdx.x + tid] = t;
	}
}

// increasePcd: the n appended rows (reference src/gaussian_model.cpp:222-262)
__global__ void insert_fill_kernel(int P, int n, const float* __restrict__ points, const float* __restrict__ colors, const float* __restrict__ dist2,
                                   int iteration, DensifyTensors t)
{
	const int j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const size_t d = (size_t)P + j;
	const float sc = logf(sqrtf(fmaxf(dist2[j], 0.0000001f)));
	const float a = 0.1f * 1.0f;
	const float op = logf(a / (1.0f - a));
#pragma unroll
	for (int c = 0; c < 3; c++) {
		t.dp[0][3 * d + c] = points[3 * (size_t)j + c];
		t.dp[1][3 * d + c] = (colors[3 * (size_t)j + c] - 0.5f) / kSH_C0;   // RGB2SH
		t.dp[4][3 * d + c] = sc;
	}
	for (int c = 0; c < 45; c++) t.dp[2][45 * d + c] = 0.f;
	t.dp[3][d] = op;
	reinterpret_cast<float4*>(t.dp[5])[d] = make_float4(1.f, 0.f, 0.f, 0.f);
	if (t.dexist) t.dexist[d] = iteration;
}

struct Workspace {
	uint8_t* flags; uint32_t* counts; uint32_t* offsets; uint32_t* totals;
};
size_t ws_bytes(int P)
{
	const size_t nb = ((size_t)P + DN_TB - 1) / DN_TB;
	return align_up((size_t)P, 256) + 2 * align_up(nb * 4 * sizeof(uint32_t), 256) + 256;
}
Workspace ws_carve(void* w, int P)
{
	const size_t nb = ((size_t)P + DN_TB - 1) / DN_TB;
	char* c = static_cast<char*>(w);
	Workspace s;
	s.flags = reinterpret_cast<uint8_t*>(c); c += align_up((size_t)P, 256);
	s.counts = reinterpret_cast<uint32_t*>(c); c += align_up(nb * 4 * sizeof(uint32_t), 256);
	s.offsets = reinterpret_cast<uint32_t*>(c); c += align_up(nb * 4 * sizeof(uint32_t), 256);
	s.totals = reinterpret_cast<uint32_t*>(c);
	return s;
}
DensifyCfg to_cfg(const psb_densify_cfg* c)
{
	DensifyCfg d;
	d.tau = c->max_grad; d.min_opacity = c->min_opacity;
	d.size_thr = c->percent_dense * c->extent;   // float * float like percentDense() * scene_extent
	d.ws_thr = 0.1f * c->extent;
	d.ws_prune = c->max_screen_size != 0;
	return d;
}

}  // namespace
}  // namespace psb

using namespace psb;

extern "C" {

size_t psb_densify_workspace_bytes(int P) { return P > 0 ? ws_bytes(P) : 256; }

int psb_densify_plan(int P, const psb_model* src, const psb_densify_cfg* cfg, void* workspace, int* counts_host, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || !cfg || !counts_host || (P > 0 && (!src || !workspace || !src->param[3] || !src->param[4] || !src->xyz_gradient_accum || !src->denom))) {
		set_error_msg("psb_densify_plan: bad argument (model with opacity, scaling, xyz_gradient_accum, denom and a workspace required)");
		return PSB_ERR_ARG;
	}
	for (int k = 0; k < 5; k++) counts_host[k] = 0;
	if (P == 0) return 0;
	const Workspace w = ws_carve(workspace, P);
	const int nb = (P + DN_TB - 1) / DN_TB;
	densify_mask_kernel<<<nb, DN_TB, 0, stream>>>(P, src->param[4], src->param[3], src->xyz_gradient_accum, src->denom, to_cfg(cfg), w.flags, w.counts);
	PSB_LAUNCH_OK();
	densify_scan_kernel<<<1, 1024, 0, stream>>>(nb, w.counts, w.offsets, w.totals);
	PSB_LAUNCH_OK();
	uint32_t tot[4];
	PSB_CUDA_OK(cudaMemcpyAsync(tot, w.totals, sizeof(tot), cudaMemcpyDeviceToHost, stream));
	PSB_CUDA_OK(cudaStreamSynchronize(stream));  // the one host round trip: the caller must size the output tensors
	counts_host[0] = (int)(tot[0] + tot[1] + 2 * tot[2]);
	counts_host[1] = (int)tot[0]; counts_host[2] = (int)tot[1]; counts_host[3] = (int)tot[2]; counts_host[4] = (int)tot[3];
	return 0;
}

int psb_densify_apply(int P, const psb_model* src, const psb_model* dst, int P_new, const psb_densify_cfg* cfg, const void* workspace,
                      const float* normal_samples, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || P_new < 0 || !cfg || (P > 0 && (!src || !workspace)) || (P_new > 0 && !dst)) { set_error_msg("psb_densify_apply: bad argument"); return PSB_ERR_ARG; }
	if (P_new > 0) {
		if (dst->max_radii2D) PSB_CUDA_OK(cudaMemsetAsync(dst->max_radii2D, 0, (size_t)P_new * sizeof(float), stream));          // :709-711
		if (dst->xyz_gradient_accum) PSB_CUDA_OK(cudaMemsetAsync(dst->xyz_gradient_accum, 0, (size_t)P_new * sizeof(float), stream));
		if (dst->denom) PSB_CUDA_OK(cudaMemsetAsync(dst->denom, 0, (size_t)P_new * sizeof(float), stream));
	}
	if (P == 0 || P_new == 0) return 0;
	DensifyTensors t;
	for (int i = 0; i < 6; i++) {
		if (!src->param[i] || !src->exp_avg[i] || !src->exp_avg_sq[i] || !dst->param[i] || !dst->exp_avg[i] || !dst->exp_avg_sq[i]) {
			set_error_msg("psb_densify_apply: null tensor"); return PSB_ERR_ARG;
		}
		t.p[i] = src->param[i]; t.m[i] = src->exp_avg[i]; t.v[i] = src->exp_avg_sq[i];
		t.dp[i] = dst->param[i]; t.dm[i] = dst->exp_avg[i]; t.dv[i] = dst->exp_avg_sq[i];
	}
	t.exist = (src->exist_since_iter && dst->exist_since_iter) ? src->exist_since_iter : nullptr;
	t.dexist = dst->exist_since_iter;
	for (int q = 0; q < 3; q++) { t.stat[q] = nullptr; t.dstat[q] = nullptr; }
	if (reinterpret_cast<uintptr_t>(src->param[5]) & 15) { set_error_msg("psb_densify_apply: rotation tensor must be 16-byte aligned"); return PSB_ERR_ARG; }
	const Workspace w = ws_carve(const_cast<void*>(workspace), P);
	densify_scatter_kernel<<<(P + DN_TB - 1) / DN_TB, DN_TB, 0, stream>>>(P, t, w.flags, w.offsets, w.totals, normal_samples, cfg->seed, cfg->offset);
	PSB_LAUNCH_OK();
	return 0;
}

int psb_prune_plan(int P, const unsigned char* mask, void* workspace, int* counts_host, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || !counts_host || (P > 0 && (!mask || !workspace))) { set_error_msg("psb_prune_plan: bad argument"); return PSB_ERR_ARG; }
	for (int k = 0; k < 5; k++) counts_host[k] = 0;
	if (P == 0) return 0;
	const Workspace w = ws_carve(workspace, P);
	const int nb = (P + DN_TB - 1) / DN_TB;
	prune_mask_kernel<<<nb, DN_TB, 0, stream>>>(P, mask, w.flags, w.counts);
	PSB_LAUNCH_OK();
	densify_scan_kernel<<<1, 1024, 0, stream>>>(nb, w.counts, w.offsets, w.totals);
	PSB_LAUNCH_OK();
	uint32_t tot[4];
	PSB_CUDA_OK(cudaMemcpyAsync(tot, w.totals, sizeof(tot), cudaMemcpyDeviceToHost, stream));
	PSB_CUDA_OK(cudaStreamSynchronize(stream));
	counts_host[0] = counts_host[1] = (int)tot[0];
	return 0;
}

int psb_prune_apply(int P, const psb_model* src, const psb_model* dst, int P_new, const void* workspace, int keep_stats, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || P_new < 0 || (P > 0 && (!src || !workspace)) || (P_new > 0 && !dst)) { set_error_msg("psb_prune_apply: bad argument"); return PSB_ERR_ARG; }
	if (P == 0 || P_new == 0) return 0;
	DensifyTensors t;
	for (int i = 0; i < 6; i++) {
		if (!src->param[i] || !src->exp_avg[i] || !src->exp_avg_sq[i] || !dst->param[i] || !dst->exp_avg[i] || !dst->exp_avg_sq[i]) {
			set_error_msg("psb_prune_apply: null tensor"); return PSB_ERR_ARG;
		}
		t.p[i] = src->param[i]; t.m[i] = src->exp_avg[i]; t.v[i] = src->exp_avg_sq[i];
		t.dp[i] = dst->param[i]; t.dm[i] = dst->exp_avg[i]; t.dv[i] = dst->exp_avg_sq[i];
	}
	t.exist = (src->exist_since_iter && dst->exist_since_iter) ? src->exist_since_iter : nullptr;
	t.dexist = dst->exist_since_iter;
	const bool stats = keep_stats && src->max_radii2D && src->xyz_gradient_accum && src->denom && dst->max_radii2D && dst->xyz_gradient_accum && dst->denom;
	t.stat[0] = stats ? src->max_radii2D : nullptr; t.stat[1] = src->xyz_gradient_accum; t.stat[2] = src->denom;
	t.dstat[0] = dst->max_radii2D; t.dstat[1] = dst->xyz_gradient_accum; t.dstat[2] = dst->denom;
	if (reinterpret_cast<uintptr_t>(src->param[5]) & 15) { set_error_msg("psb_prune_apply: rotation tensor must be 16-byte aligned"); return PSB_ERR_ARG; }
	const Workspace w = ws_carve(const_cast<void*>(workspace), P);
	densify_scatter_kernel<<<(P + DN_TB - 1) / DN_TB, DN_TB, 0, stream>>>(P, t, w.flags, w.offsets, w.totals, nullptr, 0ull, 0ull);
	PSB_LAUNCH_OK();
	return 0;
}

int psb_insert_points(int P, const psb_model* src, const psb_model* dst, int n, const float* points, const float* colors, const float* dist2,
                      int iteration, void* stream_)
{
	cudaStream_t stream = (cudaStream_t)stream_;
	if (P < 0 || n < 0 || !dst || (P > 0 && !src) || (n > 0 && (!points || !colors || !dist2))) { set_error_msg("psb_insert_points: bad argument"); return PSB_ERR_ARG; }
	const size_t Q = (size_t)P + n;
	DensifyTensors t;
	memset(&t, 0, sizeof(t));
	for (int i = 0; i < 6; i++) {
		if (Q > 0 && (!dst->param[i] || !dst->exp_avg[i] || !dst->exp_avg_sq[i])) { set_error_msg("psb_insert_points: null tensor"); return PSB_ERR_ARG; }
		const size_t k = (size_t)dn_row(i);
		if (P > 0) {
			PSB_CUDA_OK(cudaMemcpyAsync(dst->param[i], src->param[i], (size_t)P * k * sizeof(float), cudaMemcpyDeviceToDevice, stream));
			PSB_CUDA_OK(cudaMemcpyAsync(dst->exp_avg[i], src->exp_avg[i], (size_t)P * k * sizeof(float), cudaMemcpyDeviceToDevice, stream));
			PSB_CUDA_OK(cudaMemcpyAsync(dst->exp_avg_sq[i], src->exp_avg_sq[i], (size_t)P * k * sizeof(float), cudaMemcpyDeviceToDevice, stream));
		}
		if (n > 0) {   // moments of the new rows: zeros (cat(..., zeros_like(extension)), :675-676)
			PSB_CUDA_OK(cudaMemsetAsync(dst->exp_avg[i] + (size_t)P * k, 0, (size_t)n * k * sizeof(float), stream));
			PSB_CUDA_OK(cudaMemsetAsync(dst->exp_avg_sq[i] + (size_t)P * k, 0, (size_t)n * k * sizeof(float), stream));
		}
		t.dp[i] = dst->param[i];
	}
	if (P > 0 && src->exist_since_iter && dst->exist_since_iter)
		PSB_CUDA_OK(cudaMemcpyAsync(dst->exist_since_iter, src->exist_since_iter, (size_t)P * sizeof(int), cudaMemcpyDeviceToDevice, stream));
	t.dexist = dst->exist_since_iter;
	if (Q > 0) {
		if (dst->max_radii2D) PSB_CUDA_OK(cudaMemsetAsync(dst->max_radii2D, 0, Q * sizeof(float), stream));
		if (dst->xyz_gradient_accum) PSB_CUDA_OK(cudaMemsetAsync(dst->xyz_gradient_accum, 0, Q * sizeof(float), stream));
		if (dst->denom) PSB_CUDA_OK(cudaMemsetAsync(dst->denom, 0, Q * sizeof(float), stream));
	}
	if (n > 0) {
		if (reinterpret_cast<uintptr_t>(dst->param[5]) & 15) { set_error_msg("psb_insert_points: rotation tensor must be 16-byte aligned"); return PSB_ERR_ARG; }
		insert_fill_kernel<<<(n + 127) / 128, 128, 0, stream>>>(P, n, points, colors, dist2, iteration, t);
		PSB_LAUNCH_OK();
	}
	return 0;
}

int psb_reset_opacity(int P, float* opacity, float* exp_avg, float* exp_avg_sq, void* stream_)
{
	if (P < 0 || (P > 0 && (!opacity || !exp_avg || !exp_avg_sq))) { set_error_msg("psb_reset_opacity: bad argument"); return PSB_ERR_ARG; }
	if (P == 0) return 0;
	reset_opacity_kernel<<<(P + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(P, opacity, exp_avg, exp_avg_sq);
	PSB_LAUNCH_OK();
	return 0;
}

}  // extern "C"
