// Trainer step kernels: the per-Gaussian backward fused with the activations' backward, the Adam update of
// all 59 parameters of the Gaussian and the densification statistics — one read and one write of every
// parameter / moment, no gradient tensor in memory. Two launches per iteration:
//
//   gaussian_backward_kernel  one thread per Gaussian: screen-space sums -> gradients of the 14 small parameters
//                             + their Adam update + densification statistics + the 18-float SH gradient "seed"
//                             (latency-bound: every global load is issued at the top of the kernel, the f_rest
//                             parameter rows it has to read arrive by one TMA bulk copy per block);
//   frest_stream_kernel       Adam over the [P,15,3] SH rows (76 % of the parameter bytes), one float4 per thread,
//                             gradient = seed products; runs at the HBM copy rate.
//
// Replaces, per iteration of reference GaussianMapper::trainForOneIteration (src/gaussian_mapper.cpp:614-774):
//   computeCov2DCUDA + preprocessCUDA backward (cuda_rasterizer/backward.cu:144-396),
//   autograd through sigmoid / exp / normalize / cat (src/gaussian_model.cpp:48-71),
//   9 torch::zeros gradient tensors incl. [P,16,3] (src/rasterize_points.cu:148-157),
//   max_radii2D / addDensificationStats (gaussian_mapper.cpp:714-719, gaussian_model.cpp:817-831),
//   torch::optim::Adam::step over 6 parameter groups + zero_grad (gaussian_mapper.cpp:769-772).
#include "psb_backward.cuh"
#include "psb_train.h"
#include <cstdlib>
#include <cstdint>
#include <cstring>

namespace psb {

namespace {

#ifndef PSB_A_MINBLOCKS
#define PSB_A_MINBLOCKS 5
#endif
constexpr int TB = 128;        // Gaussians per block
constexpr int REST = 45;       // floats per f_rest row (15 coefficients x 3 channels)
constexpr int SEED = 20;       // floats per seed row: w_1..w_15 at [0..14], masked dL/dRGB at [16..18]

struct AdamCoef { float beta1, beta2, eps, inv_bc1, inv_bc2_sqrt; };

// gaussian_backward_kernel modes
constexpr int MODE_GRADS = 0;  // dense gradients of the raw parameters -> `grads` (NCCL data-parallel path / tests)
constexpr int MODE_ADAM = 1;   // fused single-GPU step: parameters and moments updated in place
constexpr int MODE_PUSH = 2;   // NVLink data-parallel step: 80-byte gradient records pushed into the owner rank's inbox

// torch::optim::Adam single-tensor update (LibTorch adam.cpp): exp_avg.mul_(b1).add_(g, 1-b1);
// exp_avg_sq.mul_(b2).addcmul_(g, g, 1-b2); denom = exp_avg_sq.sqrt()/sqrt(bc2) + eps; p.addcdiv_(exp_avg, denom, -lr/bc1).
// sqrt and the division go through the MUFU unit (rsqrt / rcp, ~1-2 ulp): 59 IEEE sqrt+div sequences per
// Gaussian would otherwise make this HBM-streaming kernel issue-bound.
__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float lr_eff, const AdamCoef& c)
{
	m = m * c.beta1 + (1.f - c.beta1) * g;
	v = v * c.beta2 + (1.f - c.beta2) * g * g;
	const float sq = v > 0.f ? v * rsqrtf(v) : 0.f;
	const float denom = sq * c.inv_bc2_sqrt + c.eps;
	p = p - lr_eff * __fdividef(m, denom);
}

// ADAM = true : fused update (parameters and moments of the 14 small parameters updated in place).
// ADAM = false: gradients w.r.t. the RAW parameters are written to `grads` (reference tensor shapes) for the
//               data-parallel path (all-reduce between this kernel and adam_kernel).
//
// Block = 128 Gaussians, one thread each. The SH gradient of coefficient k, channel ch is w_k(dir) * dL/dRGB[ch]
// (clamp-masked): the thread leaves the 15 weights and the 3 masked colour gradients in `seeds` ([P][SEED] floats) and
// frest_stream_kernel expands them. The f_rest parameter rows the SH backward has to READ (view-direction term of
// dL/dxyz) arrive by one TMA bulk copy per block; each thread walks its own row in shared memory (row stride 45 words:
// bank-conflict free).
template <int MODE>
__global__ void __launch_bounds__(TB, PSB_A_MINBLOCKS) gaussian_backward_kernel(int first, int P, TrainTensors t, Camera cam, GeomState geom, float4* __restrict__ sink,
                                                            StepHyper h, GradSegments grads, DensifyStats st,
                                                            const uint32_t* __restrict__ counters, uint32_t capacity, float4* __restrict__ seeds,
                                                            DpPush dp)
{
	constexpr bool ADAM = MODE == MODE_ADAM;
	__shared__ __align__(128) float s_p[TB * REST];  // f_rest parameter rows of this block (TMA destination)
	__shared__ __align__(16) float s_w[TB][SEED];    // seed rows of this block (copied out 128-bit coalesced at the end)
	__shared__ __align__(8) uint64_t s_bar;

	// A binning arena that turned out too small leaves the tile lists incomplete. Fused step: no-op (the host sees the
	// same counter, grows the arena and repeats the step). Data-parallel modes: this view contributes a ZERO gradient to
	// the step (every rank still takes part in the exchange, so the replicas stay identical and nobody waits forever).
	const bool overflow = counters[0] > capacity;
	if (ADAM && overflow) return;

	const int tid = threadIdx.x;
	const int base = first + blockIdx.x * TB;  // this launch covers Gaussians [first, P)
	const int rows = min(TB, P - base);
	const uint32_t row_bytes = (uint32_t)rows * REST * sizeof(float);
	const uint32_t bulk_bytes = row_bytes & ~15u;
	const int rem_floats = (int)(row_bytes & 15u) / 4;
	const size_t goff = (size_t)base * REST;

	if (tid == 0) {
		mbar_init(&s_bar, 1);
		mbar_fence_init();
		mbar_arrive_expect_tx(&s_bar, bulk_bytes);
		bulk_g2s(s_p, t.p[2] + goff, bulk_bytes, &s_bar);
		for (int i = 0; i < rem_floats; i++) s_p[bulk_bytes / 4 + i] = t.p[2][goff + bulk_bytes / 4 + i];  // < 16 trailing bytes (last block)
	}
	__syncthreads();  // the mbarrier must be initialised before any other thread waits on it

	const int idx = base + tid;
	const bool valid = idx < P;
	AdamCoef ac;
	ac.beta1 = h.beta1; ac.beta2 = h.beta2; ac.eps = h.eps; ac.inv_bc1 = h.inv_bc1; ac.inv_bc2_sqrt = 1.0f / h.bc2_sqrt;

	// ---- phase 1a: everything that does not need the staged rows
	bool visible = false;
	float3 g_xyz = make_float3(0, 0, 0), g_scale = make_float3(0, 0, 0);
	float4 g_rot = make_float4(0, 0, 0, 0);
	float g_opac = 0.f;
	float3 mean = make_float3(0, 0, 1), dc = make_float3(0, 0, 0);
	float3 dL_dcolor = make_float3(0, 0, 0);
	uint32_t clamp_bits = 0;
	// the 14 small parameters (every Gaussian needs them for its Adam update): issued up front, with the
	// screen-space sums, as one batch of independent loads — a single memory round trip instead of a chain
	float sp_xyz[3] = {0, 0, 1}, sp_dc[3] = {0, 0, 0}, sp_sc[3] = {0, 0, 0}, sp_op = 0.f;
	float4 sp_rot = make_float4(1, 0, 0, 0);
	// ... and their moments (consumed last): loaded here as well so their latency hides behind the whole backward
	float m3[3][3], v3[3][3], mo = 0.f, vo = 0.f;
	float4 mr = make_float4(0, 0, 0, 0), vr = mr;
	float st_rad = 0.f, st_acc = 0.f, st_den = 0.f;
	if (valid) {
		if (ADAM) {
			const int tsel[3] = {0, 1, 4};
#pragma unroll
			for (int a = 0; a < 3; a++)
#pragma unroll
				for (int c = 0; c < 3; c++) { m3[a][c] = t.m[tsel[a]][3 * idx + c]; v3[a][c] = t.v[tsel[a]][3 * idx + c]; }
			mo = t.m[3][idx]; vo = t.v[3][idx];
			mr = reinterpret_cast<const float4*>(t.m[5])[idx]; vr = reinterpret_cast<const float4*>(t.v[5])[idx];
		}
		if (st.enabled) { st_rad = st.max_radii2D[idx]; st_acc = st.xyz_gradient_accum[idx]; st_den = st.denom[idx]; }
		const uint32_t tt = geom.tile_info[idx].z;
		const float4 s0 = sink[3 * idx], s1 = sink[3 * idx + 1], s2 = sink[3 * idx + 2];
		const uint32_t meta = __float_as_uint(geom.rec[idx].q2.w);
#pragma unroll
		for (int c = 0; c < 3; c++) { sp_xyz[c] = t.p[0][3 * idx + c]; sp_dc[c] = t.p[1][3 * idx + c]; sp_sc[c] = t.p[4][3 * idx + c]; }
		sp_op = t.p[3][idx];
		sp_rot = reinterpret_cast<const float4*>(t.p[5])[idx];
		visible = tt != 0 && !overflow;
		// ready for the next iteration (rows nothing was added to — culled or hidden Gaussians — are zero already: not rewritten)
		const bool sink_dirty = s0.x != 0.f || s0.y != 0.f || s0.w != 0.f || s1.x != 0.f || s1.z != 0.f || s1.w != 0.f || s2.x != 0.f || s2.y != 0.f || s2.z != 0.f;
		if (sink_dirty) { const float4 z = make_float4(0, 0, 0, 0); sink[3 * idx] = z; sink[3 * idx + 1] = z; sink[3 * idx + 2] = z; }
		if (visible) {
			// sink row: [0,1] mean2D.xy | [3,4,6] conic.xyw | [7] opacity | [8,9,10] rgb
			const float2 dL_dmean2D = make_float2(s0.x, s0.y);
			const float3 dL_dconic = make_float3(s0.w, s1.x, s1.z);
			const float dL_dopacity = s1.w;
			dL_dcolor = make_float3(s2.x, s2.y, s2.z);
			mean = make_float3(sp_xyz[0], sp_xyz[1], sp_xyz[2]);
			dc = make_float3(sp_dc[0], sp_dc[1], sp_dc[2]);
			const float3 sraw = make_float3(sp_sc[0], sp_sc[1], sp_sc[2]);
			const float4 qraw = sp_rot;
			const float oraw = sp_op;
			clamp_bits = rec_clamp_bits(meta);
			// activations exactly as in the forward (preprocess_fwd_kernel<RAW>)
			const float3 s = make_float3(expf(sraw.x), expf(sraw.y), expf(sraw.z));
			const float qn = fmaxf(sqrtf(qraw.x * qraw.x + qraw.y * qraw.y + qraw.z * qraw.z + qraw.w * qraw.w), 1e-12f);
			const float4 q = make_float4(qraw.x / qn, qraw.y / qn, qraw.z / qn, qraw.w / qn);
			const float sig = 1.0f / (1.0f + expf(-oraw));
			float cov3D[6], dL_dcov[6];
			cov3d_from_scale_rot(s, 1.0f, q, cov3D);
			gaussian_backward_geom(cam, mean, cov3D, dL_dmean2D, dL_dconic, g_xyz, dL_dcov);
			float3 dL_dscale;
			float4 dL_drot;
			cov3d_backward(s, 1.0f, q, dL_dcov, dL_dscale, dL_drot);
			// autograd of the activations (src/gaussian_model.cpp:48-71): exp, normalize, sigmoid
			g_scale = make_float3(dL_dscale.x * s.x, dL_dscale.y * s.y, dL_dscale.z * s.z);
			const float qd = q.x * dL_drot.x + q.y * dL_drot.y + q.z * dL_drot.z + q.w * dL_drot.w;
			g_rot = make_float4((dL_drot.x - q.x * qd) / qn, (dL_drot.y - q.y * qd) / qn, (dL_drot.z - q.z * qd) / qn, (dL_drot.w - q.w * qd) / qn);
			g_opac = dL_dopacity * sig * (1.0f - sig);
			if (st.enabled) {
				st.max_radii2D[idx] = fmaxf(st_rad, (float)rec_radius(meta));
				st.xyz_gradient_accum[idx] = st_acc + sqrtf(dL_dmean2D.x * dL_dmean2D.x + dL_dmean2D.y * dL_dmean2D.y);
				st.denom[idx] = st_den + 1.0f;
			}
		}
	}

	mbar_wait(&s_bar, 0);  // f_rest rows have landed
	__syncthreads();       // (also publishes thread 0's trailing plain stores)

	// ---- phase 1b: SH backward seeds; the view-direction term completes dL/dxyz
	float3 g_dc = make_float3(0, 0, 0);
	{
		const float* rp = s_p + tid * REST;
		float w[16];
#pragma unroll
		for (int k = 0; k < 16; k++) w[k] = 0.f;
		float gm[3] = {0.f, 0.f, 0.f};
		if (visible) {
			const float3 campos = make_float3(cam.campos[0], cam.campos[1], cam.campos[2]);
			const float3 sh_dmean = sh_backward_t(
				h.D, mean, campos, clamp_bits, dL_dcolor,
				[&](int k, int ch) { return k == 0 ? (ch == 0 ? dc.x : (ch == 1 ? dc.y : dc.z)) : rp[3 * (k - 1) + ch]; },
				[&](int k, int ch, float g) { if (k == 0) { if (ch == 0) g_dc.x = g; else if (ch == 1) g_dc.y = g; else g_dc.z = g; } },
				w);
			g_xyz.x += sh_dmean.x; g_xyz.y += sh_dmean.y; g_xyz.z += sh_dmean.z;
			gm[0] = ((clamp_bits >> 0) & 1u) ? 0.f : dL_dcolor.x;
			gm[1] = ((clamp_bits >> 1) & 1u) ? 0.f : dL_dcolor.y;
			gm[2] = ((clamp_bits >> 2) & 1u) ? 0.f : dL_dcolor.z;
		}
		if (MODE != MODE_PUSH) {
			const int ncoef = (h.D + 1) * (h.D + 1);
#pragma unroll
			for (int k = 1; k < 16; k++) s_w[tid][k - 1] = (k < ncoef) ? w[k] : 0.f;
			s_w[tid][15] = 0.f; s_w[tid][16] = gm[0]; s_w[tid][17] = gm[1]; s_w[tid][18] = gm[2]; s_w[tid][19] = 0.f;
		} else {
			// gradient record of this Gaussian for its owner rank (see DpPush): the SH weights are NOT sent, the owner
			// re-evaluates w_k(dir) from its bit-identical copy of xyz and this rank's camera centre
			s_w[tid][0] = g_xyz.x; s_w[tid][1] = g_xyz.y; s_w[tid][2] = g_xyz.z;
			s_w[tid][3] = g_dc.x; s_w[tid][4] = g_dc.y; s_w[tid][5] = g_dc.z;
			s_w[tid][6] = g_opac;
			s_w[tid][7] = g_scale.x; s_w[tid][8] = g_scale.y; s_w[tid][9] = g_scale.z;
			s_w[tid][10] = g_rot.x; s_w[tid][11] = g_rot.y; s_w[tid][12] = g_rot.z; s_w[tid][13] = g_rot.w;
			s_w[tid][14] = gm[0]; s_w[tid][15] = gm[1]; s_w[tid][16] = gm[2];
			s_w[tid][17] = 0.f; s_w[tid][18] = 0.f;
			// a visible Gaussian hidden behind nearer splats receives no gradient: like an invisible one it sends nothing
			const bool sends = visible && (g_xyz.x != 0.f || g_xyz.y != 0.f || g_xyz.z != 0.f || g_opac != 0.f || gm[0] != 0.f || gm[1] != 0.f || gm[2] != 0.f ||
			                               g_scale.x != 0.f || g_scale.y != 0.f || g_scale.z != 0.f || g_rot.x != 0.f || g_rot.y != 0.f || g_rot.z != 0.f || g_rot.w != 0.f);
			s_w[tid][19] = __uint_as_float(sends ? dp.epoch : 0u);
		}
	}

	// ---- the 14 small parameters of this Gaussian
	if (valid) {
		// a zero gradient on zero moments is an exact no-op of Adam (see frest_stream_kernel): hidden / never-reached Gaussians are not rewritten
		bool idle = false;
		if (ADAM) {
			bool nz = g_xyz.x != 0.f || g_xyz.y != 0.f || g_xyz.z != 0.f || g_dc.x != 0.f || g_dc.y != 0.f || g_dc.z != 0.f || g_scale.x != 0.f ||
			          g_scale.y != 0.f || g_scale.z != 0.f || g_opac != 0.f || g_rot.x != 0.f || g_rot.y != 0.f || g_rot.z != 0.f || g_rot.w != 0.f ||
			          mo != 0.f || vo != 0.f || mr.x != 0.f || mr.y != 0.f || mr.z != 0.f || mr.w != 0.f || vr.x != 0.f || vr.y != 0.f || vr.z != 0.f || vr.w != 0.f;
#pragma unroll
			for (int a = 0; a < 3; a++)
#pragma unroll
				for (int c = 0; c < 3; c++) nz = nz || m3[a][c] != 0.f || v3[a][c] != 0.f;
			idle = !nz;
		}
		if (ADAM && !idle) {
			const float gx[3] = {g_xyz.x, g_xyz.y, g_xyz.z}, gd[3] = {g_dc.x, g_dc.y, g_dc.z}, gs[3] = {g_scale.x, g_scale.y, g_scale.z};
#pragma unroll
			for (int c = 0; c < 3; c++) {
				adam1(sp_xyz[c], m3[0][c], v3[0][c], gx[c], h.lr[0] * ac.inv_bc1, ac);
				adam1(sp_dc[c], m3[1][c], v3[1][c], gd[c], h.lr[1] * ac.inv_bc1, ac);
				adam1(sp_sc[c], m3[2][c], v3[2][c], gs[c], h.lr[4] * ac.inv_bc1, ac);
			}
			adam1(sp_op, mo, vo, g_opac, h.lr[3] * ac.inv_bc1, ac);
			const float lrr = h.lr[5] * ac.inv_bc1;
			adam1(sp_rot.x, mr.x, vr.x, g_rot.x, lrr, ac);
			adam1(sp_rot.y, mr.y, vr.y, g_rot.y, lrr, ac);
			adam1(sp_rot.z, mr.z, vr.z, g_rot.z, lrr, ac);
			adam1(sp_rot.w, mr.w, vr.w, g_rot.w, lrr, ac);
#pragma unroll
			for (int c = 0; c < 3; c++) {
				t.p[0][3 * idx + c] = sp_xyz[c]; t.m[0][3 * idx + c] = m3[0][c]; t.v[0][3 * idx + c] = v3[0][c];
				t.p[1][3 * idx + c] = sp_dc[c]; t.m[1][3 * idx + c] = m3[1][c]; t.v[1][3 * idx + c] = v3[1][c];
				t.p[4][3 * idx + c] = sp_sc[c]; t.m[4][3 * idx + c] = m3[2][c]; t.v[4][3 * idx + c] = v3[2][c];
			}
			t.p[3][idx] = sp_op; t.m[3][idx] = mo; t.v[3][idx] = vo;
			reinterpret_cast<float4*>(t.p[5])[idx] = sp_rot; reinterpret_cast<float4*>(t.m[5])[idx] = mr; reinterpret_cast<float4*>(t.v[5])[idx] = vr;
		} else if (MODE == MODE_GRADS) {
			grads.g[0][3 * idx] = g_xyz.x; grads.g[0][3 * idx + 1] = g_xyz.y; grads.g[0][3 * idx + 2] = g_xyz.z;
			grads.g[1][3 * idx] = g_dc.x; grads.g[1][3 * idx + 1] = g_dc.y; grads.g[1][3 * idx + 2] = g_dc.z;
			grads.g[3][idx] = g_opac;
			grads.g[4][3 * idx] = g_scale.x; grads.g[4][3 * idx + 1] = g_scale.y; grads.g[4][3 * idx + 2] = g_scale.z;
			reinterpret_cast<float4*>(grads.g[5])[idx] = g_rot;
		}
	}
	__syncthreads();

	if (MODE != MODE_PUSH) {
		// seed rows of this block -> global, 128-bit coalesced
		const float4* src = reinterpret_cast<const float4*>(&s_w[0][0]);
		float4* dst = seeds + (size_t)base * (SEED / 4);
		for (int i = tid; i < rows * (SEED / 4); i += TB) dst[i] = src[i];
		return;
	}
	// MODE_PUSH: the block's records go straight into the inbox of the rank that owns this 128-Gaussian chunk (remote
	// stores over NVLink: fire and forget, they overlap the arithmetic of the blocks still running). Rows this view does
	// not see are not sent: their stale epoch word tells the owner to count them as zero.
	{
		const int chunk = base / TB;
		const int owner = chunk % dp.world, lc = chunk / dp.world;
		float4* dst = reinterpret_cast<float4*>(dp.inbox[owner]) + ((size_t)dp.rank * dp.nlocal_max + lc) * (TB * (SEED / 4));
		const float4* src = reinterpret_cast<const float4*>(&s_w[0][0]);
		for (int i = tid; i < rows * (SEED / 4); i += TB) {
			if (__float_as_uint(s_w[i / (SEED / 4)][19]) == dp.epoch) dst[i] = src[i];
		}
		// "all my records of this step have landed". Default: nothing here — dp_signal_kernel, queued right behind this kernel,
		// publishes the camera centre and raises the flags after ONE system-scope fence (the kernel boundary orders it after every
		// store of this grid). PSB_DP_SIGNAL=fence: every thread fences its own remote stores, the last block of the grid publishes
		// (threadFenceReduction pattern) — measured slower: a system-scope fence per block costs microseconds with NVLink stores in flight.
		if (!dp.fence_in_kernel) return;
		__threadfence_system();
		__syncthreads();
		if (tid == 0) {
			const uint32_t prev = atomicAdd(dp.done_counter, 1u);
			if (prev == gridDim.x - 1) {
				*dp.done_counter = 0u;
				__threadfence_system();
				for (int j = 0; j < dp.world; j++) {
					float* mt = dp.meta[j] + 8 * dp.rank;
					mt[0] = cam.campos[0]; mt[1] = cam.campos[1]; mt[2] = cam.campos[2]; mt[3] = (float)h.D;
				}
				__threadfence_system();
				for (int j = 0; j < dp.world; j++) *reinterpret_cast<volatile uint32_t*>(dp.grad_flag[j] + dp.rank) = dp.epoch;
			}
		}
	}
}

// Adam over the f_rest rows of Gaussians [first, P) (ADAM) or their gradient written out
// (!ADAM), one float4 per thread. The gradient of element (row, k, ch) is seed[row][k-1] * seed[row][16+ch].
template <bool ADAM>
__global__ void __launch_bounds__(256) frest_stream_kernel(uint32_t e_first, uint32_t e_end, float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                           float* __restrict__ gout, const float* __restrict__ seeds, float lr_eff, AdamCoef ac,
                                                           const uint32_t* __restrict__ counters, uint32_t capacity)
{
	if (ADAM && counters[0] > capacity) return;  // overflowing fused step: no-op (see gaussian_backward_kernel; !ADAM: the seeds are zero)
	const uint32_t e = e_first + 4u * (blockIdx.x * 256u + threadIdx.x);
	if (e >= e_end) return;
	uint32_t row = e / REST;
	int c = (int)(e - row * REST);
	const float* sd = seeds + (size_t)row * SEED;
	float g[4];
#pragma unroll
	for (int j = 0; j < 4; j++) {
		const int k = c / 3, ch = c - 3 * k;
		g[j] = sd[k] * sd[16 + ch];
		if (++c == REST) { c = 0; sd += SEED; }
	}
	if (e + 4 <= e_end) {
		if (ADAM) {
			float4 mm = __ldcs(reinterpret_cast<const float4*>(m + e)), vv = __ldcs(reinterpret_cast<const float4*>(v + e));
			// a zero gradient on zero moments is an exact no-op of Adam (m' = v' = 0, p' = p - lr * 0 / (0 + eps) = p): such elements
			// (Gaussians no view has reached since their moments were created) are neither re-read nor rewritten
			if (g[0] == 0.f && g[1] == 0.f && g[2] == 0.f && g[3] == 0.f && mm.x == 0.f && mm.y == 0.f && mm.z == 0.f && mm.w == 0.f &&
			    vv.x == 0.f && vv.y == 0.f && vv.z == 0.f && vv.w == 0.f) return;
			float4 pp = __ldcs(reinterpret_cast<const float4*>(p + e));
			adam1(pp.x, mm.x, vv.x, g[0], lr_eff, ac);
			adam1(pp.y, mm.y, vv.y, g[1], lr_eff, ac);
			adam1(pp.z, mm.z, vv.z, g[2], lr_eff, ac);
			adam1(pp.w, mm.w, vv.w, g[3], lr_eff, ac);
			__stcs(reinterpret_cast<float4*>(p + e), pp); __stcs(reinterpret_cast<float4*>(m + e), mm); __stcs(reinterpret_cast<float4*>(v + e), vv);
		} else {
			__stcs(reinterpret_cast<float4*>(gout + e), make_float4(g[0], g[1], g[2], g[3]));
		}
	} else {  // < 4 trailing elements (P not a multiple of 4)
		for (uint32_t j = 0; e + j < e_end; j++) {
			if (ADAM) adam1(p[e + j], m[e + j], v[e + j], g[j], lr_eff, ac);
			else gout[e + j] = g[j];
		}
	}
}

// Plain Adam over one flat tensor (data-parallel path, after the gradient all-reduce). 128-bit accesses.
__global__ void __launch_bounds__(256) adam_kernel(size_t n4, float4* __restrict__ p, float4* __restrict__ m, float4* __restrict__ v,
                                                   const float4* __restrict__ g, float lr, AdamCoef c, float grad_scale)
{
	const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n4) return;
	float4 pp = p[i], mm = m[i], vv = v[i];
	const float4 gg = g[i];
	const float le = lr * c.inv_bc1;
	adam1(pp.x, mm.x, vv.x, gg.x * grad_scale, le, c);
	adam1(pp.y, mm.y, vv.y, gg.y * grad_scale, le, c);
	adam1(pp.z, mm.z, vv.z, gg.z * grad_scale, le, c);
	adam1(pp.w, mm.w, vv.w, gg.w * grad_scale, le, c);
	p[i] = pp; m[i] = mm; v[i] = vv;
}
__global__ void adam_tail_kernel(size_t start, size_t n, float* p, float* m, float* v, const float* g, float lr, AdamCoef c, float grad_scale)
{
	const size_t i = start + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) adam1(p[i], m[i], v[i], g[i] * grad_scale, lr * c.inv_bc1, c);
}


// ------------------------------------------------------------------------------------------------------------------------
// NVLink data-parallel step, owner side. After every rank's records of this epoch have landed (launch_wait_flags), the
// owner of a chunk sums the <= world records of each of its rows, applies Adam to its rows only (moments exist only on the
// owner) and writes the updated parameters into EVERY rank's parameter tensors (remote stores over NVLink): reduce-scatter,
// sharded optimizer and all-gather without a materialised reduced gradient and without a collective call.
// ------------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void sh_weights(int deg, const float3 pos, const float3 campos, float* w /*[16]*/)
{
	// same expressions as sh_backward_t (psb_backward.cuh): the owner reproduces the source rank's weights bit for bit
	const float3 d = make_float3(pos.x - campos.x, pos.y - campos.y, pos.z - campos.z);
	const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
	const float x = d.x / len, y = d.y / len, z = d.z / len;
	const float xx = x * x, yy = y * y, zz = z * z, xy = x * y, yz = y * z, xz = x * z;
	w[0] = kSH_C0;
	w[1] = -kSH_C1 * y; w[2] = kSH_C1 * z; w[3] = -kSH_C1 * x;
	w[4] = kSH_C2_0 * xy; w[5] = kSH_C2_1 * yz; w[6] = kSH_C2_2 * (2.f * zz - xx - yy); w[7] = kSH_C2_3 * xz; w[8] = kSH_C2_4 * (xx - yy);
	w[9] = kSH_C3_0 * y * (3.f * xx - yy); w[10] = kSH_C3_1 * xy * z; w[11] = kSH_C3_2 * y * (4.f * zz - xx - yy);
	w[12] = kSH_C3_3 * z * (2.f * zz - 3.f * xx - 3.f * yy); w[13] = kSH_C3_4 * x * (4.f * zz - xx - yy); w[14] = kSH_C3_5 * z * (xx - yy);
	w[15] = kSH_C3_6 * x * (xx - 3.f * yy);
	const int ncoef = (deg + 1) * (deg + 1);
#pragma unroll
	for (int k = 0; k < 16; k++) if (k >= ncoef) w[k] = 0.f;
}

// One block per owned chunk, one thread per Gaussian: the 14 small parameters (Adam + all-gather) and the summed f_rest gradient row.
// The updated small parameters leave through shared memory: each of the five tensors' 128-row chunk is contiguous in every replica, so
// it is sent to each rank as fully coalesced 16-byte stores (a thread storing its own 3-float rows would emit 4-byte stores at a
// 12-byte stride: a third of every NVLink write packet used — measured 0.55 ms of the 1.04 ms owner stage at N = 8).
__global__ void __launch_bounds__(TB) shard_adam_small_kernel(DpShard d, TrainTensors t, StepHyper h, float grad_scale)
{
	__shared__ __align__(16) float s_g[TB * REST];
	__shared__ __align__(16) float s_out[TB * 14];   // [xyz 384 | f_dc 384 | scaling 384 | opacity 128 | rotation 512] in tensor-chunk layout
	const int tid = threadIdx.x, lc = d.lc_first + blockIdx.x;
	const int chunk = lc * d.world + d.rank;
	const int idx = chunk * TB + tid;
	const bool valid = idx < d.P;
	AdamCoef ac;
	ac.beta1 = h.beta1; ac.beta2 = h.beta2; ac.eps = h.eps; ac.inv_bc1 = h.inv_bc1; ac.inv_bc2_sqrt = 1.0f / h.bc2_sqrt;

	float acc[14];
	float G[REST];
#pragma unroll
	for (int i = 0; i < 14; i++) acc[i] = 0.f;
#pragma unroll
	for (int i = 0; i < REST; i++) G[i] = 0.f;
	float sp_xyz[3] = {0, 0, 1}, sp_dc[3] = {0, 0, 0}, sp_sc[3] = {0, 0, 0}, sp_op = 0.f;
	float4 sp_rot = make_float4(1, 0, 0, 0);
	float m3[3][3], v3[3][3], mo = 0.f, vo = 0.f;
	float4 mr = make_float4(0, 0, 0, 0), vr = mr;
	bool live = false;
	if (valid) {
		// every rank's epoch word of this row first (independent loads: one round trip instead of one per source)
		float4 tail[DP_MAX_WORLD];
#pragma unroll
		for (int s = 0; s < DP_MAX_WORLD; s++) {
			tail[s] = make_float4(0.f, 0.f, 0.f, 0.f);
			if (s < d.world) tail[s] = (reinterpret_cast<const float4*>(d.inbox) + (((size_t)s * d.nlocal_max + lc) * TB + tid) * (DP_REC / 4))[4];
		}
		const int tsel[3] = {0, 1, 4};
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int c = 0; c < 3; c++) { m3[a][c] = t.m[tsel[a]][3 * idx + c]; v3[a][c] = t.v[tsel[a]][3 * idx + c]; }
		mo = t.m[3][idx]; vo = t.v[3][idx];
		mr = reinterpret_cast<const float4*>(t.m[5])[idx]; vr = reinterpret_cast<const float4*>(t.v[5])[idx];
#pragma unroll
		for (int c = 0; c < 3; c++) { sp_xyz[c] = t.p[0][3 * idx + c]; sp_dc[c] = t.p[1][3 * idx + c]; sp_sc[c] = t.p[4][3 * idx + c]; }
		sp_op = t.p[3][idx];
		sp_rot = reinterpret_cast<const float4*>(t.p[5])[idx];
		const float3 pos = make_float3(sp_xyz[0], sp_xyz[1], sp_xyz[2]);
#pragma unroll
		for (int s = 0; s < DP_MAX_WORLD; s++) {
			if (s >= d.world || __float_as_uint(tail[s].w) != d.epoch) continue;  // rank s sent nothing for this Gaussian in this step
			const float4* r = reinterpret_cast<const float4*>(d.inbox) + (((size_t)s * d.nlocal_max + lc) * TB + tid) * (DP_REC / 4);
			const float4 r0 = r[0], r1 = r[1], r2 = r[2], r3 = r[3];
			acc[0] += r0.x; acc[1] += r0.y; acc[2] += r0.z; acc[3] += r0.w;
			acc[4] += r1.x; acc[5] += r1.y; acc[6] += r1.z; acc[7] += r1.w;
			acc[8] += r2.x; acc[9] += r2.y; acc[10] += r2.z; acc[11] += r2.w;
			acc[12] += r3.x; acc[13] += r3.y;
			const float gm[3] = {r3.z, r3.w, tail[s].x};
			const float* mt = d.meta + 8 * s;
			float w[16];
			sh_weights((int)mt[3], pos, make_float3(mt[0], mt[1], mt[2]), w);
#pragma unroll
			for (int k = 1; k < 16; k++)
#pragma unroll
				for (int ch = 0; ch < 3; ch++) G[3 * (k - 1) + ch] += w[k] * gm[ch];
		}
		const float gs = grad_scale;
		const float lx = h.lr[0] * ac.inv_bc1, ld = h.lr[1] * ac.inv_bc1, lo = h.lr[3] * ac.inv_bc1, ls = h.lr[4] * ac.inv_bc1, lrr = h.lr[5] * ac.inv_bc1;
		// zero gradient on zero moments = an exact no-op of Adam: nothing to update (hidden / never-reached Gaussians)
		live = mo != 0.f || vo != 0.f || mr.x != 0.f || mr.y != 0.f || mr.z != 0.f || mr.w != 0.f || vr.x != 0.f || vr.y != 0.f || vr.z != 0.f || vr.w != 0.f;
#pragma unroll
		for (int i = 0; i < 14; i++) live = live || acc[i] != 0.f;
#pragma unroll
		for (int a = 0; a < 3; a++)
#pragma unroll
			for (int c = 0; c < 3; c++) live = live || m3[a][c] != 0.f || v3[a][c] != 0.f;
		if (live) {
#pragma unroll
			for (int c = 0; c < 3; c++) {
				adam1(sp_xyz[c], m3[0][c], v3[0][c], acc[c] * gs, lx, ac);
				adam1(sp_dc[c], m3[1][c], v3[1][c], acc[3 + c] * gs, ld, ac);
				adam1(sp_sc[c], m3[2][c], v3[2][c], acc[7 + c] * gs, ls, ac);
			}
			adam1(sp_op, mo, vo, acc[6] * gs, lo, ac);
			adam1(sp_rot.x, mr.x, vr.x, acc[10] * gs, lrr, ac);
			adam1(sp_rot.y, mr.y, vr.y, acc[11] * gs, lrr, ac);
			adam1(sp_rot.z, mr.z, vr.z, acc[12] * gs, lrr, ac);
			adam1(sp_rot.w, mr.w, vr.w, acc[13] * gs, lrr, ac);
#pragma unroll
			for (int c = 0; c < 3; c++) {
				t.m[0][3 * idx + c] = m3[0][c]; t.v[0][3 * idx + c] = v3[0][c];
				t.m[1][3 * idx + c] = m3[1][c]; t.v[1][3 * idx + c] = v3[1][c];
				t.m[4][3 * idx + c] = m3[2][c]; t.v[4][3 * idx + c] = v3[2][c];
			}
			t.m[3][idx] = mo; t.v[3][idx] = vo;
			reinterpret_cast<float4*>(t.m[5])[idx] = mr; reinterpret_cast<float4*>(t.v[5])[idx] = vr;
		}
	}
	// the block's (updated or unchanged) small parameters in tensor-chunk layout
#pragma unroll
	for (int c = 0; c < 3; c++) { s_out[3 * tid + c] = sp_xyz[c]; s_out[384 + 3 * tid + c] = sp_dc[c]; s_out[768 + 3 * tid + c] = sp_sc[c]; }
	s_out[1152 + tid] = sp_op;
	reinterpret_cast<float4*>(s_out + 1280)[tid] = sp_rot;
	// summed f_rest gradient row -> local scratch (coalesced through shared memory; row stride 45 words is conflict-free)
#pragma unroll
	for (int i = 0; i < REST; i++) s_g[tid * REST + i] = G[i] * grad_scale;
	const bool any_live = __syncthreads_or(live);
	float4* dst = reinterpret_cast<float4*>(d.g_rest + (size_t)lc * (TB * REST));
	const float4* src = reinterpret_cast<const float4*>(s_g);
	for (int i = tid; i < TB * REST / 4; i += TB) dst[i] = src[i];
	// all-gather by remote stores: every replica (this one included) receives the chunk of each tensor as coalesced 16-byte stores.
	// Destinations are visited starting at rank + 1, so ranks running in step aim at different receivers. A chunk in which no row moved
	// (all hidden / never reached) is not sent at all.
	if (any_live) {
		const int rows = min(TB, d.P - chunk * TB);
		const int tsel[5] = {0, 1, 4, 3, 5}, width[5] = {3, 3, 3, 1, 4}, soff[5] = {0, 384, 768, 1152, 1280};
		if (d.bulk && rows == TB) {
			// TMA bulk stores (cp.async.bulk shared -> global, 0.5-2 KB each, 5 tensors x world destinations issued back to back by one
			// thread): the copy engine keeps all of them in flight over NVLink, where per-thread stores are limited by the SM's store queue
			// (measured: the per-block time of this kernel grew 9x from 1 to 7 remote destinations).
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // this thread's s_out writes -> visible to the async proxy
			__syncthreads();
			if (tid == 0) {
				for (int jj = 0; jj < d.world; jj++) {
					const int j = d.rotate ? (d.rank + 1 + jj) % d.world : jj;
#pragma unroll
					for (int q = 0; q < 5; q++) {
						float* gdst = d.param[j][tsel[q]] + (size_t)chunk * TB * width[q];
						asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(s_out + soff[q])),
						             "r"((uint32_t)(TB * width[q] * sizeof(float)))
						             : "memory");
					}
				}
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
				asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // writes complete (not only the shared-memory reads) before the block retires
			}
		} else {
			for (int jj = 0; jj < d.world; jj++) {
				const int j = d.rotate ? (d.rank + 1 + jj) % d.world : jj;
#pragma unroll
				for (int q = 0; q < 5; q++) {
					const int n = rows * width[q];   // floats of this tensor's chunk; the chunk starts 16-byte aligned (128 rows x width x 4 B)
					float* gdst = d.param[j][tsel[q]] + (size_t)chunk * TB * width[q];
					const float* ssrc = s_out + soff[q];
					for (int i = tid; i < n / 4; i += TB) reinterpret_cast<float4*>(gdst)[i] = reinterpret_cast<const float4*>(ssrc)[i];
					if (tid < (n & 3)) gdst[(n & ~3) + tid] = ssrc[(n & ~3) + tid];
				}
			}
		}
	}
	if (d.fence_in_kernel) __threadfence_system();  // this block's remote stores are performed before the kernel can complete
}

// Adam over the f_rest rows of the owned chunks, one block per chunk (5760 floats), 128-bit accesses; the updated values go
// to every rank. Rows nobody has ever seen (g = m = v = 0) are left alone: their Adam update is exactly zero.
__global__ void __launch_bounds__(256) shard_adam_frest_kernel(DpShard d, TrainTensors t, float lr_eff, AdamCoef ac)
{
	const int lc = d.lc_first + blockIdx.x;
	const int chunk = lc * d.world + d.rank;
	const size_t e0 = (size_t)chunk * (TB * REST);
	const size_t e_end = (size_t)d.P * REST;
	float* p = t.p[2];
	float* m = t.m[2];
	float* v = t.v[2];
	const float* g = d.g_rest + (size_t)lc * (TB * REST);
	for (int i = threadIdx.x; i < TB * REST / 4; i += 256) {
		const size_t e = e0 + 4 * (size_t)i;
		if (e >= e_end) break;
		if (e + 4 <= e_end) {
			const float4 gg = __ldcs(reinterpret_cast<const float4*>(g + 4 * i));
			float4 mm = __ldcs(reinterpret_cast<const float4*>(m + e)), vv = __ldcs(reinterpret_cast<const float4*>(v + e));
			const bool idle = gg.x == 0.f && gg.y == 0.f && gg.z == 0.f && gg.w == 0.f && mm.x == 0.f && mm.y == 0.f && mm.z == 0.f && mm.w == 0.f &&
			                  vv.x == 0.f && vv.y == 0.f && vv.z == 0.f && vv.w == 0.f;
			if (idle) continue;
			float4 pp = __ldcs(reinterpret_cast<const float4*>(p + e));
			adam1(pp.x, mm.x, vv.x, gg.x, lr_eff, ac);
			adam1(pp.y, mm.y, vv.y, gg.y, lr_eff, ac);
			adam1(pp.z, mm.z, vv.z, gg.z, lr_eff, ac);
			adam1(pp.w, mm.w, vv.w, gg.w, lr_eff, ac);
			__stcs(reinterpret_cast<float4*>(m + e), mm); __stcs(reinterpret_cast<float4*>(v + e), vv);
			for (int jj = 0; jj < d.world; jj++) {
				const int j = d.rotate ? (d.rank + 1 + jj) % d.world : jj;
				*reinterpret_cast<float4*>(d.param[j][2] + e) = pp;
			}
		} else {  // < 4 trailing elements (P * 45 not a multiple of 4)
			for (size_t k = e; k < e_end; k++) {
				float pp = p[k], mm = m[k], vv = v[k];
				adam1(pp, mm, vv, g[k - e0], lr_eff, ac);
				m[k] = mm; v[k] = vv;
				for (int j = 0; j < d.world; j++) d.param[j][2][k] = pp;
			}
		}
	}
	// "my updated rows have landed everywhere": dp_signal_kernel behind this kernel (default), or the last block of the grid
	if (!d.fence_in_kernel) return;
	__threadfence_system();
	__syncthreads();
	if (threadIdx.x == 0) {
		const uint32_t prev = atomicAdd(d.done_counter, 1u);
		if (prev == gridDim.x - 1) {
			*d.done_counter = 0u;
			__threadfence_system();
			for (int j = 0; j < d.world; j++) *reinterpret_cast<volatile uint32_t*>(d.param_flag[j] + d.rank) = d.epoch;
		}
	}
}

// One warp, queued right behind a data kernel of the data-parallel step: lane j publishes (optionally) this rank's meta row and then
// raises this rank's epoch word on rank j. Stream order puts it after every store of the data kernel; the single system-scope fence
// orders those stores (cumulativity) before the flag for an observer on another GPU.
struct DpSignal {
	uint32_t* flag[DP_MAX_WORLD];
	float* meta[DP_MAX_WORLD];   // null: no meta row
	const float* campos;
	int world, rank, degree;
	uint32_t epoch;
};
__global__ void dp_signal_kernel(DpSignal s)
{
	const int j = threadIdx.x;
	if (j < s.world && s.meta[j]) {
		float* mt = s.meta[j] + 8 * s.rank;
		mt[0] = s.campos[0]; mt[1] = s.campos[1]; mt[2] = s.campos[2]; mt[3] = (float)s.degree;
	}
	__threadfence_system();
	__syncwarp();
	if (j < s.world) *reinterpret_cast<volatile uint32_t*>(s.flag[j] + s.rank) = s.epoch;
}

__device__ __forceinline__ unsigned long long global_timer_ns()
{
	unsigned long long t;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
	return t;
}

// One warp: lane r waits for flags[r] >= epoch (written by rank r over NVLink). Bounded: a peer that never arrives (crashed
// process) must not hang this GPU — after timeout_ns the kernel records the failure and returns.
__global__ void wait_flags_kernel(const uint32_t* flags, int world, uint32_t epoch, uint32_t* status, unsigned long long timeout_ns)
{
	const int lane = threadIdx.x;   // `world` = number of flag words (ranks, or groups x ranks), <= blockDim.x
	if (lane >= world) return;
	const volatile uint32_t* f = flags + lane;
	const unsigned long long t0 = global_timer_ns();
	while ((int32_t)(*f - epoch) < 0) {
		__nanosleep(200);
		if (global_timer_ns() - t0 > timeout_ns) { atomicExch(status, 1u); break; }
	}
	__threadfence_system();
}

}  // namespace

int launch_fused_backward(bool adam, int first, int P, const TrainTensors& t, const Camera& cam, const GeomState& geom, float* sink, float* seeds,
                          const StepHyper& h, const GradSegments& grads, const DensifyStats& st, const uint32_t* counters, uint32_t capacity,
                          cudaStream_t stream, cudaEvent_t between)
{
	if (P - first <= 0) return 0;
	const int grid = cdiv(P - first, TB);
	float4* sd4 = reinterpret_cast<float4*>(seeds);
	float4* sink4 = reinterpret_cast<float4*>(sink);
	if (!seeds) { set_error_msg("launch_fused_backward: seed scratch missing"); return -1; }
	if ((size_t)P * REST >= (size_t)UINT32_MAX - 8) { set_error_msg("psb_trainer: too many Gaussians for 32-bit element indices"); return -1; }
	DpPush nodp;
	memset(&nodp, 0, sizeof(nodp));
	if (adam) gaussian_backward_kernel<MODE_ADAM><<<grid, TB, 0, stream>>>(first, P, t, cam, geom, sink4, h, grads, st, counters, capacity, sd4, nodp);
	else gaussian_backward_kernel<MODE_GRADS><<<grid, TB, 0, stream>>>(first, P, t, cam, geom, sink4, h, grads, st, counters, capacity, sd4, nodp);
	PSB_LAUNCH_OK();
	if (between) cudaEventRecord(between, stream);
	AdamCoef ac;
	ac.beta1 = h.beta1; ac.beta2 = h.beta2; ac.eps = h.eps; ac.inv_bc1 = h.inv_bc1; ac.inv_bc2_sqrt = 1.0f / h.bc2_sqrt;
	const uint32_t e0 = (uint32_t)first * REST, e1 = (uint32_t)P * REST;
	const unsigned gridB = (unsigned)(((size_t)(e1 - e0) + 1023) / 1024);
	if (adam) frest_stream_kernel<true><<<gridB, 256, 0, stream>>>(e0, e1, t.p[2], t.m[2], t.v[2], nullptr, seeds, h.lr[2] * ac.inv_bc1, ac, counters, capacity);
	else frest_stream_kernel<false><<<gridB, 256, 0, stream>>>(e0, e1, nullptr, nullptr, nullptr, grads.g[2], seeds, 0.f, ac, counters, capacity);
	PSB_LAUNCH_OK();
	return 0;
}

int launch_push_backward(int first, int P, const TrainTensors& t, const Camera& cam, const GeomState& geom, float* sink, const StepHyper& h,
                         const DensifyStats& st, const uint32_t* counters, uint32_t capacity, const DpPush& dp, cudaStream_t stream)
{
	GradSegments nog;
	memset(&nog, 0, sizeof(nog));
	if (P - first > 0) {  // Gaussians [first, P): one pipeline group (first is a multiple of 128 * world)
		gaussian_backward_kernel<MODE_PUSH><<<cdiv(P - first, TB), TB, 0, stream>>>(first, P, t, cam, geom, reinterpret_cast<float4*>(sink), h, nog, st, counters,
		                                                                          capacity, nullptr, dp);
		PSB_LAUNCH_OK();
	}
	if (!dp.fence_in_kernel) {
		DpSignal s;
		memset(&s, 0, sizeof(s));
		for (int j = 0; j < dp.world; j++) { s.flag[j] = dp.grad_flag[j]; s.meta[j] = dp.meta[j]; }
		s.campos = cam.campos; s.world = dp.world; s.rank = dp.rank; s.degree = h.D; s.epoch = dp.epoch;
		dp_signal_kernel<<<1, 32, 0, stream>>>(s);
		PSB_LAUNCH_OK();
	}
	return 0;
}

int launch_shard_adam(const DpShard& d, const TrainTensors& t, const StepHyper& h, float grad_scale, cudaStream_t stream, cudaEvent_t between)
{
	AdamCoef ac;
	ac.beta1 = h.beta1; ac.beta2 = h.beta2; ac.eps = h.eps; ac.inv_bc1 = h.inv_bc1; ac.inv_bc2_sqrt = 1.0f / h.bc2_sqrt;
	if (d.nlocal > 0) {  // owned chunks [lc_first, lc_first + nlocal) of this pipeline group
		shard_adam_small_kernel<<<d.nlocal, TB, 0, stream>>>(d, t, h, grad_scale);
		PSB_LAUNCH_OK();
		if (between) cudaEventRecord(between, stream);
		shard_adam_frest_kernel<<<d.nlocal, 256, 0, stream>>>(d, t, h.lr[2] * ac.inv_bc1, ac);
		PSB_LAUNCH_OK();
	}
	if (!d.fence_in_kernel) {
		DpSignal s;
		memset(&s, 0, sizeof(s));
		for (int j = 0; j < d.world; j++) s.flag[j] = d.param_flag[j];
		s.world = d.world; s.rank = d.rank; s.epoch = d.epoch;
		dp_signal_kernel<<<1, 32, 0, stream>>>(s);
		PSB_LAUNCH_OK();
	}
	return 0;
}

int launch_wait_flags(const uint32_t* flags, int world, uint32_t epoch, uint32_t* status, cudaStream_t stream)
{
	static const unsigned long long timeout_ns = (unsigned long long)(getenv("PSB_DP_TIMEOUT_MS") ? atoll(getenv("PSB_DP_TIMEOUT_MS")) : 20000) * 1000000ull;
	wait_flags_kernel<<<1, world <= 32 ? 32 : 64, 0, stream>>>(flags, world, epoch, status, timeout_ns);
	PSB_LAUNCH_OK();
	return 0;
}

int launch_adam(size_t n, float* p, float* m, float* v, const float* g, float lr, const StepHyper& h, float grad_scale, cudaStream_t stream)
{
	if (n == 0) return 0;
	AdamCoef c;
	c.beta1 = h.beta1; c.beta2 = h.beta2; c.eps = h.eps; c.inv_bc1 = h.inv_bc1; c.inv_bc2_sqrt = 1.0f / h.bc2_sqrt;
	const bool aligned = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v) | reinterpret_cast<uintptr_t>(g)) & 15) == 0;
	const size_t n4 = aligned ? n / 4 : 0;
	if (n4) {
		adam_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, stream>>>(n4, reinterpret_cast<float4*>(p), reinterpret_cast<float4*>(m), reinterpret_cast<float4*>(v),
		                                                             reinterpret_cast<const float4*>(g), lr, c, grad_scale);
		PSB_LAUNCH_OK();
	}
	if (n4 * 4 < n) {  // unaligned tensors / < 4 trailing elements: ONE scalar launch over the remainder
		adam_tail_kernel<<<(unsigned)((n - n4 * 4 + 255) / 256), 256, 0, stream>>>(n4 * 4, n, p, m, v, g, lr, c, grad_scale);
		PSB_LAUNCH_OK();
	}
	return 0;
}

}  // namespace psb
