// Per-tile alpha blending, forward and backward.
//
// One 256-thread block per 16x16 tile, one thread per pixel, eight warps each owning an 8x4 pixel
// footprint. Per batch of 256 list entries:
//   * every thread reads one sorted Gaussian index and issues ONE 48-byte TMA bulk copy
//     (cp.async.bulk, completion on an mbarrier) of that Gaussian's packed record into shared memory;
//     two stages, so the gather of batch b+1 overlaps the blending of batch b;
//   * the thread that staged a record tests it against the tile rectangle (exact minimum of the
//     conic quadratic over the rectangle, conservatively padded for rounding) and the survivors are
//     compacted in list order; records that cannot reach alpha >= 1/255 on any pixel of the tile are
//     never evaluated per pixel. The per-pixel tests themselves are unchanged, so results
//     (colour, final T, n_contrib) are identical to evaluating the whole list;
//   * backward: per-pixel terms are reduced over the 32 pixels of a warp with a 12-shuffle butterfly
//     (9 values), accumulated per list entry in shared memory, and leave the block as ONE set of
//     atomics per (Gaussian, tile) instead of 9 atomics per (Gaussian, pixel).
//
// Blending semantics (thresholds, order of operations, n_contrib bookkeeping) follow reference
// cuda_rasterizer/forward.cu:261-374 and backward.cu:399-557.
#include <cstdlib>
#ifndef PSB_FWD_MINBLOCKS
#define PSB_FWD_MINBLOCKS 8
#endif
#ifndef PSB_BWD_MINBLOCKS
#define PSB_BWD_MINBLOCKS 6
#endif
#include "psb_common.cuh"
#include "psb_kernels.h"

namespace psb {

namespace {

constexpr int RB = 256;        // list entries per batch

// Ordered compaction of `keep` flags over the block; returns total, writes kept entry ids to s_cidx.
template <int NWARPS>
__device__ __forceinline__ int compact_block(bool keep, uint8_t* s_cidx, int* s_wcnt, int tid)
{
	const int lane = tid & 31, warp = tid >> 5;
	const uint32_t bal = __ballot_sync(0xffffffffu, keep);
	if (lane == 0) s_wcnt[warp] = __popc(bal);
	__syncthreads();
	int base = 0, total = 0;
#pragma unroll
	for (int w = 0; w < NWARPS; w++) {
		const int c = s_wcnt[w];
		if (w < warp) base += c;
		total += c;
	}
	if (keep) s_cidx[base + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)tid;
	__syncthreads();
	return total;
}

}  // namespace

// Thread/pixel geometry shared by both tile kernels: PPT pixels per thread, 256 / PPT threads per 16x16 tile.
// A warp owns an 8-wide, 4*PPT-tall footprint; lane (lx, ly) = (lane & 7, lane >> 3) owns the pixels of rows
// ly, ly + 4, ... ly + 4 (PPT - 1) of that footprint (same column: dx is shared by a thread's pixels).
template <int PPT>
struct TileGeom {
	static constexpr int THREADS = PSB_TILE_PIX / PPT;
	static constexpr int NWARPS = THREADS / 32;
	static constexpr int FOOT_H = 4 * PPT;
	static constexpr int WARPS_Y = PSB_TILE_Y / FOOT_H;
};

// =================================================================================================
// Forward
// =================================================================================================
template <int PPT>
__global__ void __launch_bounds__(TileGeom<PPT>::THREADS, PPT == 2 ? PSB_FWD_MINBLOCKS : 1) render_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                                            const GaussRec* __restrict__ rec, int W, int H,
                                                                            const float* __restrict__ bg_color, float* __restrict__ out_color,
                                                                            float* __restrict__ final_T, uint32_t* __restrict__ n_contrib, int dbg_arg)
{
#ifdef PSB_FWD_DEBUG_SWITCHES
	const int dbg = dbg_arg;  // bit0: no tile-level cull, bit1: no per-warp cull, bit2: no pmin pre-test
#else
	constexpr int dbg = 0;
	(void)dbg_arg;
#endif
	using G = TileGeom<PPT>;
	constexpr int NT = G::THREADS;  // = list entries staged per batch (one per thread)
	__shared__ __align__(16) GaussRec s_rec[2][NT];
	__shared__ uint8_t s_cidx[NT];
	__shared__ int s_wcnt[G::NWARPS];
	__shared__ __align__(8) uint64_t s_bar[2];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile_x0 = blockIdx.x * PSB_TILE_X, tile_y0 = blockIdx.y * PSB_TILE_Y;
	const int wx = tile_x0 + (warp % 2) * 8, wy = tile_y0 + (warp / 2) * G::FOOT_H;
	const int px = wx + (lane & 7);
	const float pixfx = (float)px;
	int py[PPT];
	float pixfy[PPT];
	bool done[PPT];
	float T[PPT], C[PPT][3];
	uint32_t last_contributor[PPT];
#pragma unroll
	for (int p = 0; p < PPT; p++) {
		py[p] = wy + (lane >> 3) + 4 * p;
		pixfy[p] = (float)py[p];
		done[p] = !(px < W && py[p] < H);
		T[p] = 1.0f;
		C[p][0] = C[p][1] = C[p][2] = 0.f;
		last_contributor[p] = 0;
	}
	const float rx0 = (float)tile_x0, ry0 = (float)tile_y0;
	const float rx1 = (float)(min(tile_x0 + PSB_TILE_X, W) - 1), ry1 = (float)(min(tile_y0 + PSB_TILE_Y, H) - 1);
	// pixel rectangle of this warp's footprint (clamped to the image; degenerate if the warp is outside)
	const float wx0 = (float)wx, wy0 = (float)wy;
	const float wx1 = fmaxf(wx0, fminf(wx0 + 7.f, (float)(W - 1))), wy1 = fmaxf(wy0, fminf(wy0 + (float)(G::FOOT_H - 1), (float)(H - 1)));

	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
	const int n = (int)(range.y - range.x);
	const int nbatch = (n + NT - 1) / NT;

	auto stage = [&](int batch, int st) {
		const int e = batch * NT + tid;
		if (tid == 0) mbar_arrive_expect_tx(&s_bar[st], (uint32_t)min(NT, n - batch * NT) * (uint32_t)sizeof(GaussRec));
		if (e < n) bulk_g2s(&s_rec[st][tid], &rec[point_list[range.x + (uint32_t)e]], (uint32_t)sizeof(GaussRec), &s_bar[st]);
	};
	auto all_done = [&]() {
		bool d = true;
#pragma unroll
		for (int p = 0; p < PPT; p++) d = d && done[p];
		return d;
	};

	if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
	__syncthreads();
	if (nbatch > 0) stage(0, 0);

	for (int b = 0; b < nbatch; b++) {
		const int st = b & 1;
		// (the barrier also orders stage reuse: every thread has left batch b-1)
		if (__syncthreads_count(all_done()) == NT) {
			mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));  // drain the in-flight copy of batch b before exiting
			break;
		}
		if (b + 1 < nbatch) stage(b + 1, st ^ 1);
		mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));

		const int cnt = min(NT, n - b * NT);
		bool keep = false;
		if (tid < cnt) keep = (dbg & 1) || splat_reaches_tile(s_rec[st][tid].q0, s_rec[st][tid].q1, rx0, ry0, rx1, ry1);
		const int ccount = compact_block<G::NWARPS>(keep, s_cidx, s_wcnt, tid);

		// Second, per-warp cull: each lane tests one surviving entry against this warp's pixel footprint
		// (32 entries per ballot), then the warp walks only the entries that can reach one of its pixels.
		for (int c0 = 0; c0 < ccount; c0 += 32) {
			if (__all_sync(0xffffffffu, all_done())) break;
			const int kk = c0 + lane;
			int j = 0;
			bool hit = false;
			if (kk < ccount) {
				j = s_cidx[kk];
				hit = (dbg & 2) || splat_reaches_tile(s_rec[st][j].q0, s_rec[st][j].q1, wx0, wy0, wx1, wy1);
			}
			uint32_t hits = __ballot_sync(0xffffffffu, hit);
			while (hits) {
				const int src = __ffs(hits) - 1;
				hits &= hits - 1;
				const int jj = __shfl_sync(0xffffffffu, j, src);
				const float4 q0 = s_rec[st][jj].q0;
				const float4 q1 = s_rec[st][jj].q1;
				const float2 xy = make_float2(q0.x, q0.y);
				const float4 con_o = make_float4(q0.z, q0.w, q1.x, q1.y);
				const uint32_t contributor = (uint32_t)(b * NT + jj + 1);
#pragma unroll
				for (int p = 0; p < PPT; p++) {
					if (done[p]) continue;
					const float2 d = make_float2(xy.x - pixfx, xy.y - pixfy[p]);
					const float power = splat_power(con_o.x, con_o.y, con_o.z, d.x, d.y);
					if (power > 0.0f) continue;
					if (power < q1.z && !(dbg & 4)) continue;  // below pmin: alpha < 1/255 for certain, skip exp (see GaussRec)
					const float alpha = min(0.99f, __fmul_rn(con_o.w, exp(power)));
					if (alpha < 1.0f / 255.0f) continue;
					const float test_T = __fmul_rn(T[p], __fadd_rn(1.f, -alpha));
					if (test_T < 0.0001f) { done[p] = true; continue; }
					const float4 q2 = s_rec[st][jj].q2;
					C[p][0] = __fmaf_rn(T[p], __fmul_rn(q2.x, alpha), C[p][0]);
					C[p][1] = __fmaf_rn(T[p], __fmul_rn(q2.y, alpha), C[p][1]);
					C[p][2] = __fmaf_rn(T[p], __fmul_rn(q2.z, alpha), C[p][2]);
					T[p] = test_T;
					last_contributor[p] = contributor;
				}
			}
		}
	}

	const size_t HW = (size_t)H * W;
#pragma unroll
	for (int p = 0; p < PPT; p++) {
		if (px < W && py[p] < H) {
			const uint32_t pix_id = (uint32_t)W * py[p] + px;
			final_T[pix_id] = T[p];
			n_contrib[pix_id] = last_contributor[p];
#pragma unroll
			for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[p][ch] + T[p] * bg_color[ch];
		}
	}
}

// =================================================================================================
// Backward
// =================================================================================================
// 9 values per lane -> totals spread over lanes, 12 shuffles. After the call, lane l with
// (l & 1) == 0 holds the warp total of value index `warp_reduce9_index(l)` (or -1: padding).
__device__ __forceinline__ int warp_reduce9_index(int lane)
{
	const int h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
	int l5;
	if (!h8 && !h4) l5 = h2 ? 1 : 0;
	else if (!h8 && h4) l5 = h2 ? -1 : 2;
	else if (h8 && !h4) l5 = h2 ? 4 : 3;
	else l5 = -1;
	if (l5 < 0) return -1;
	const int idx = ((lane & 16) ? 5 : 0) + l5;
	return idx < 9 ? idx : -1;
}
__device__ __forceinline__ float warp_reduce9(const float v[9], int lane)
{
	const bool h16 = lane & 16, h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
	float a[5], b[3], c[2];
#pragma unroll
	for (int i = 0; i < 5; i++) {
		const float hi = (i < 4) ? v[5 + i] : 0.f;
		const float send = h16 ? v[i] : hi;
		const float keep = h16 ? hi : v[i];
		a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
	}
#pragma unroll
	for (int i = 0; i < 3; i++) {
		const float hi = (i < 2) ? a[3 + i] : 0.f;
		const float send = h8 ? a[i] : hi;
		const float keep = h8 ? hi : a[i];
		b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
	}
#pragma unroll
	for (int i = 0; i < 2; i++) {
		const float hi = (i < 1) ? b[2] : 0.f;
		const float send = h4 ? b[i] : hi;
		const float keep = h4 ? hi : b[i];
		c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
	}
	const float send = h2 ? c[0] : c[1];
	const float keep = h2 ? c[1] : c[0];
	float dsum = keep + __shfl_xor_sync(0xffffffffu, send, 2);
	dsum += __shfl_xor_sync(0xffffffffu, dsum, 1);
	return dsum;
}

// Backward batch geometry: RBB list entries per batch, staged by the first RBB threads.
template <int PPT>
struct BwdSmem {
	static constexpr int RBB = TileGeom<PPT>::THREADS < 128 ? TileGeom<PPT>::THREADS : 128;
	// floats per entry: NWARPS x (9 sums + 1 marker: != 0 when that warp wrote its sums for this entry), +1 pad (odd stride: conflict-free row reads)
	static constexpr int ACC_ROW = TileGeom<PPT>::NWARPS * 10 + 1;
	GaussRec rec[2][RBB];
	float acc[RBB * ACC_ROW];
	uint32_t gid[2][RBB];
	uint64_t bar[2];
	int wcnt[TileGeom<PPT>::NWARPS];
	uint8_t cidx[RBB];
};

template <int PPT>
__global__ void __launch_bounds__(TileGeom<PPT>::THREADS, PPT == 2 ? PSB_BWD_MINBLOCKS : 1) render_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                                            const GaussRec* __restrict__ rec, int W, int H,
                                                                            const float* __restrict__ bg_color, const float* __restrict__ final_Ts,
                                                                            const uint32_t* __restrict__ n_contrib,
                                                                            const float* __restrict__ dL_dpixels, GradSink sink)
{
	using G = TileGeom<PPT>;
	using SM = BwdSmem<PPT>;
	constexpr int NT = G::THREADS, RBB = SM::RBB, ACC_ROW = SM::ACC_ROW;
	extern __shared__ __align__(16) unsigned char smem_raw[];
	SM& sm = *reinterpret_cast<SM*>(smem_raw);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile_x0 = blockIdx.x * PSB_TILE_X, tile_y0 = blockIdx.y * PSB_TILE_Y;
	const int wx = tile_x0 + (warp % 2) * 8, wy = tile_y0 + (warp / 2) * G::FOOT_H;
	const int px = wx + (lane & 7);
	const float pixfx = (float)px;
	const float rx0 = (float)tile_x0, ry0 = (float)tile_y0;
	const float rx1 = (float)(min(tile_x0 + PSB_TILE_X, W) - 1), ry1 = (float)(min(tile_y0 + PSB_TILE_Y, H) - 1);
	const float wx0 = (float)wx, wy0 = (float)wy;
	const float wx1 = fmaxf(wx0, fminf(wx0 + 7.f, (float)(W - 1))), wy1 = fmaxf(wy0, fminf(wy0 + (float)(G::FOOT_H - 1), (float)(H - 1)));
	const size_t HW = (size_t)H * W;

	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];

	// Per-pixel state of the back-to-front sweep. The reference keeps accum_rec[3] and last_color[3] per pixel and evaluates
	// dL/dalpha = sum_ch (c_ch - accum_rec_ch) dL/dpix_ch (backward.cu:497-510). dL/dpix is constant along the sweep and the accum_rec
	// recursion is linear, so its DOT with dL/dpix obeys the same recursion: two scalars (acc_dot, last_cdot) replace six.
	float pixfy[PPT], T_final[PPT], T[PPT], acc_dot[PPT], last_cdot[PPT], dL_dpixel[PPT][3], last_alpha[PPT], bg_dot_dpixel[PPT];
	int last_contributor[PPT];
	int my_max = 0;
#pragma unroll
	for (int p = 0; p < PPT; p++) {
		const int py = wy + (lane >> 3) + 4 * p;
		const bool inside = px < W && py < H;
		const uint32_t pix_id = (uint32_t)W * py + px;
		pixfy[p] = (float)py;
		T_final[p] = inside ? final_Ts[pix_id] : 0;
		T[p] = T_final[p];
		last_contributor[p] = inside ? (int)n_contrib[pix_id] : 0;
		my_max = max(my_max, last_contributor[p]);
		last_alpha[p] = 0;
		bg_dot_dpixel[p] = 0;
		acc_dot[p] = 0.f;
		last_cdot[p] = 0.f;
#pragma unroll
		for (int i = 0; i < 3; i++) {
			dL_dpixel[p][i] = inside ? dL_dpixels[i * HW + pix_id] : 0.f;
			bg_dot_dpixel[p] += bg_color[i] * dL_dpixel[p][i];
		}
	}

	const bool has_bg = bg_color[0] != 0.f || bg_color[1] != 0.f || bg_color[2] != 0.f;  // uniform over the grid

	// Nothing behind the deepest last contributor of the tile can receive gradient: start there.
	const int warp_maxc = __reduce_max_sync(0xffffffffu, my_max);
	if (lane == 0) sm.wcnt[warp] = warp_maxc;
	if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); mbar_fence_init(); }
	for (int i = tid; i < RBB * G::NWARPS; i += NT) sm.acc[(i / G::NWARPS) * ACC_ROW + (i % G::NWARPS) * 10 + 9] = 0.f;  // markers
	__syncthreads();
	int maxc = 0;
#pragma unroll
	for (int w = 0; w < G::NWARPS; w++) maxc = max(maxc, sm.wcnt[w]);
	__syncthreads();
	const int n = min(maxc, (int)(range.y - range.x));
	if (n == 0) return;
	const int nbatch = (n + RBB - 1) / RBB;

	auto stage = [&](int batch, int st) {
		const int e = batch * RBB + tid;  // entry number, back to front
		if (tid == 0) mbar_arrive_expect_tx(&sm.bar[st], (uint32_t)min(RBB, n - batch * RBB) * (uint32_t)sizeof(GaussRec));
		if (tid < RBB && e < n) {
			const uint32_t g = point_list[range.x + (uint32_t)(n - 1 - e)];
			sm.gid[st][tid] = g;
			bulk_g2s(&sm.rec[st][tid], &rec[g], (uint32_t)sizeof(GaussRec), &sm.bar[st]);
		}
	};
	stage(0, 0);

	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	// after warp_reduce9 the even lanes hold the nine totals; lane 1 (idle) writes the marker in the SAME store instruction
	const int my_slot = ((lane & 1) == 0) ? warp_reduce9_index(lane) : (lane == 1 ? 9 : -1);

	for (int b = 0; b < nbatch; b++) {
		const int st = b & 1;
		__syncthreads();  // previous batch fully flushed; stage st^1 free
		if (b + 1 < nbatch) stage(b + 1, st ^ 1);
		mbar_wait(&sm.bar[st], (uint32_t)((b >> 1) & 1));

		const int cnt = min(RBB, n - b * RBB);
		bool keep = false;
		if (tid < cnt) keep = splat_reaches_tile(sm.rec[st][tid].q0, sm.rec[st][tid].q1, rx0, ry0, rx1, ry1);
		const int ccount = compact_block<G::NWARPS>(keep, sm.cidx, sm.wcnt, tid);

		for (int c0 = 0; c0 < ccount; c0 += 32) {
			const int kk = c0 + lane;
			int jl = 0;
			bool hit = false;
			if (kk < ccount) {
				jl = sm.cidx[kk];
				// entries behind every pixel's last contributor of this warp cannot receive gradient either
				hit = (n - 1 - (b * RBB + jl)) < warp_maxc && splat_reaches_tile(sm.rec[st][jl].q0, sm.rec[st][jl].q1, wx0, wy0, wx1, wy1);
			}
			uint32_t hits = __ballot_sync(0xffffffffu, hit);
			while (hits) {
				const int src = __ffs(hits) - 1;
				hits &= hits - 1;
				const int j = __shfl_sync(0xffffffffu, jl, src);
				// 0-based list position of this entry; the reference's `contributor` after its decrement
				const int pos = n - 1 - (b * RBB + j);
				const float4 q0 = sm.rec[st][j].q0;
				const float4 q1 = sm.rec[st][j].q1;
				const float2 xy = make_float2(q0.x, q0.y);
				const float4 con_o = make_float4(q0.z, q0.w, q1.x, q1.y);
				float2 d[PPT];
				float G_[PPT], alpha[PPT];
				bool active[PPT];
				bool any = false;
#pragma unroll
				for (int p = 0; p < PPT; p++) {
					d[p] = make_float2(xy.x - pixfx, xy.y - pixfy[p]);
					const float power = splat_power(con_o.x, con_o.y, con_o.z, d[p].x, d[p].y);
					active[p] = pos < last_contributor[p] && !(power > 0.0f) && !(power < q1.z);  // q1.z = pmin, see GaussRec
					G_[p] = 0.f; alpha[p] = 0.f;
					if (active[p]) {
						G_[p] = exp(power);
						alpha[p] = min(0.99f, __fmul_rn(con_o.w, G_[p]));
						active[p] = !(alpha[p] < 1.0f / 255.0f);
					}
					any = any || active[p];
				}
				if (!__any_sync(0xffffffffu, any)) continue;

				// Per pixel only the weight w = G dL/dG is formed; the warp reduces the six MOMENTS of w over the pixel offsets
				// (sum w, w dx, w dy, w dx^2, w dx dy, w dy^2) plus the three colour sums, and the flush turns the moments into the
				// reference's nine sums once per (Gaussian, tile) (backward.cu:518-547):
				//   dL/dmean2D = -(A Mx + B My) ddelx_dx, -(C My + B Mx) ddely_dy;  dL/dconic = -0.5 (Mxx, Mxy, Myy);  dL/dopacity = M0 / opacity
				// (dG/ddel = -G (A dx + B dy): linear in the offsets). A thread's pixels share the column, so dx multiplies their sums once.
				const float4 q2 = sm.rec[st][j].q2;
				float W0 = 0.f, Wy = 0.f, Wyy = 0.f, vc0 = 0.f, vc1 = 0.f, vc2 = 0.f;
#pragma unroll
				for (int p = 0; p < PPT; p++) {
					if (!active[p]) continue;
					T[p] = T[p] / (1.f - alpha[p]);
					const float aT = alpha[p] * T[p];
					const float cdot = q2.x * dL_dpixel[p][0] + q2.y * dL_dpixel[p][1] + q2.z * dL_dpixel[p][2];
					acc_dot[p] = last_alpha[p] * last_cdot[p] + (1.f - last_alpha[p]) * acc_dot[p];
					last_cdot[p] = cdot;
					float dL_dalpha = (cdot - acc_dot[p]) * T[p];
					last_alpha[p] = alpha[p];
					// background term (reference backward.cu:512-516); with a black background it adds (-x) * 0 = -0: skipped
					if (has_bg) dL_dalpha += (-T_final[p] / (1.f - alpha[p])) * bg_dot_dpixel[p];
					vc0 += aT * dL_dpixel[p][0];
					vc1 += aT * dL_dpixel[p][1];
					vc2 += aT * dL_dpixel[p][2];
					const float w = (con_o.w * dL_dalpha) * G_[p];
					const float wy = w * d[p].y;
					W0 += w;
					Wy += wy;
					Wyy += wy * d[p].y;
				}
				const float dx = d[0].x, Wx = W0 * dx;
				const float v[9] = {W0, Wx, Wy, Wx * dx, Wy * dx, Wyy, vc0, vc1, vc2};
				const float tot = warp_reduce9(v, lane);
				// each (warp, entry) pair is visited once per batch: plain stores, no shared-memory atomics
				if (my_slot >= 0) sm.acc[j * ACC_ROW + warp * 10 + my_slot] = (lane == 1) ? 1.f : tot;
			}
		}
		__syncthreads();

		// one set of global reductions per (Gaussian, tile)
		if (tid < cnt) {
			float a[9];
#pragma unroll
			for (int i = 0; i < 9; i++) a[i] = 0.f;
			bool any_w = false;
#pragma unroll
			for (int w = 0; w < G::NWARPS; w++) {
				float* row = &sm.acc[tid * ACC_ROW + w * 10];
				if (row[9] != 0.f) {
					row[9] = 0.f;
					any_w = true;
#pragma unroll
					for (int i = 0; i < 9; i++) a[i] += row[i];
				}
			}
			if (any_w) {
				{   // moments -> the reference's nine sums
					const float4 q0 = sm.rec[st][tid].q0, q1 = sm.rec[st][tid].q1;
					const float A = q0.z, B = q0.w, C = q1.x, M0 = a[0], Mx = a[1], My = a[2], Mxx = a[3], Mxy = a[4], Myy = a[5];
					a[0] = -(A * Mx + B * My) * ddelx_dx;
					a[1] = -(C * My + B * Mx) * ddely_dy;
					a[2] = -0.5f * Mxx;
					a[3] = -0.5f * Mxy;
					a[4] = -0.5f * Myy;
					a[5] = M0 / q1.y;
				}
				const uint32_t g = sm.gid[st][tid];
				if (sink.packed) {
					// trainer path: the 9 sums of a Gaussian live in one 48-byte row -> three 128-bit vector reductions
					float4* row = reinterpret_cast<float4*>(sink.mean2D + (size_t)g * 12);
					atomicAdd(row + 0, make_float4(a[0], a[1], 0.f, a[2]));
					atomicAdd(row + 1, make_float4(a[3], 0.f, a[4], a[5]));
					atomicAdd(row + 2, make_float4(a[6], a[7], a[8], 0.f));
				} else {
					atomicAdd(sink.mean2D + (size_t)g * sink.mean2D_stride + 0, a[0]);
					atomicAdd(sink.mean2D + (size_t)g * sink.mean2D_stride + 1, a[1]);
					atomicAdd(sink.conic + (size_t)g * sink.conic_stride + 0, a[2]);
					atomicAdd(sink.conic + (size_t)g * sink.conic_stride + 1, a[3]);
					atomicAdd(sink.conic + (size_t)g * sink.conic_stride + 3, a[4]);
					atomicAdd(sink.opacity + (size_t)g * sink.opacity_stride, a[5]);
					atomicAdd(sink.color + (size_t)g * sink.color_stride + 0, a[6]);
					atomicAdd(sink.color + (size_t)g * sink.color_stride + 1, a[7]);
					atomicAdd(sink.color + (size_t)g * sink.color_stride + 2, a[8]);
				}
			}
		}
	}
}

// pixels per thread of the two tile kernels (1, 2 or 4); overridable for experiments: PSB_FWD_PPT / PSB_BWD_PPT
static int env_ppt(const char* name, int dflt)
{
	const char* e = getenv(name);
	if (!e) return dflt;
	const int v = atoi(e);
	return (v == 1 || v == 2 || v == 4) ? v : dflt;
}

template <int PPT>
static int launch_bwd_t(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                        const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const GradSink& sink, cudaStream_t stream)
{
	// the attribute is per device (and cheap): set it on every launch rather than cache it in a process-wide flag
	PSB_CUDA_OK(cudaFuncSetAttribute(render_bwd_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem<PPT>)));
	dim3 grid(cam.grid_x, cam.grid_y, 1);
	render_bwd_kernel<PPT><<<grid, TileGeom<PPT>::THREADS, sizeof(BwdSmem<PPT>), stream>>>(ranges, point_list, rec, cam.W, cam.H, bg, final_T, n_contrib,
	                                                                                      dL_dpix, sink);
	PSB_LAUNCH_OK();
	return 0;
}

int launch_render_forward(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                          float* out_color, float* final_T, uint32_t* n_contrib, cudaStream_t stream)
{
	static const int ppt = env_ppt("PSB_FWD_PPT", 2);
	// debug only: PSB_FWD_DEBUG bit0 = no tile-level cull, bit1 = no per-warp cull, bit2 = no pmin pre-test
	static const int dbg = getenv("PSB_FWD_DEBUG") ? atoi(getenv("PSB_FWD_DEBUG")) : 0;
	dim3 grid(cam.grid_x, cam.grid_y, 1);
	if (ppt == 1) render_fwd_kernel<1><<<grid, TileGeom<1>::THREADS, 0, stream>>>(ranges, point_list, rec, cam.W, cam.H, bg, out_color, final_T, n_contrib, dbg);
	else if (ppt == 2) render_fwd_kernel<2><<<grid, TileGeom<2>::THREADS, 0, stream>>>(ranges, point_list, rec, cam.W, cam.H, bg, out_color, final_T, n_contrib, dbg);
	else render_fwd_kernel<4><<<grid, TileGeom<4>::THREADS, 0, stream>>>(ranges, point_list, rec, cam.W, cam.H, bg, out_color, final_T, n_contrib, dbg);
	PSB_LAUNCH_OK();
	return 0;
}

int launch_render_backward(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                           const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const GradSink& sink,
                           cudaStream_t stream)
{
	static const int ppt = env_ppt("PSB_BWD_PPT", 2);
	if (ppt == 1) return launch_bwd_t<1>(cam, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, sink, stream);
	if (ppt == 2) return launch_bwd_t<2>(cam, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, sink, stream);
	return launch_bwd_t<4>(cam, ranges, point_list, rec, bg, final_T, n_contrib, dL_dpix, sink, stream);
}

}  // namespace psb
