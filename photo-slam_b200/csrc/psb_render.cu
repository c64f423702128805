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
#include "psb_common.cuh"
#include "psb_kernels.h"

namespace psb {

namespace {

constexpr int RB = 256;        // list entries per batch

// Can this splat reach alpha >= 1/255 (and power <= 0 is not required here: keeping more is safe)
// on any pixel of the tile whose pixel-coordinate rectangle is [px0,px1] x [py0,py1]?
// q(d) = 0.5 (A dx^2 + C dy^2) + B dx dy, d = mean - pixel; contribution needs q <= ln(255 * opacity).
// Returns false only when q > threshold (+ rounding pad) on the WHOLE rectangle.
__device__ __forceinline__ bool splat_reaches_tile(const float4 q0, const float4 q1, float px0, float py0, float px1, float py1)
{
	const float mx = q0.x, my = q0.y, A = q0.z, B = q0.w, C = q1.x, opac = q1.y;
	if (opac < (1.0f / 255.0f)) return false;  // alpha <= opacity < 1/255 on every pixel
	const float dxlo = mx - px1, dxhi = mx - px0, dylo = my - py1, dyhi = my - py0;
	if (dxlo <= 0.f && dxhi >= 0.f && dylo <= 0.f && dyhi >= 0.f) return true;  // centre inside: q = 0 reachable
	const float thr = __logf(255.0f * opac) + 1e-3f;
	const float dxm = fmaxf(fabsf(dxlo), fabsf(dxhi)), dym = fmaxf(fabsf(dylo), fabsf(dyhi));
	const float S = 0.5f * (fabsf(A) * dxm * dxm + fabsf(C) * dym * dym) + fabsf(B) * dxm * dym;
	const float pad = 1e-5f * S + 1e-4f;
	auto q = [&](float dx, float dy) { return 0.5f * (A * dx * dx + C * dy * dy) + B * dx * dy; };
	// the minimum of a quadratic over a box that does not contain its stationary point lies on the boundary:
	// check the 4 corners and the clamped 1-D stationary point of each edge.
	float qmin = fminf(fminf(q(dxlo, dylo), q(dxhi, dylo)), fminf(q(dxlo, dyhi), q(dxhi, dyhi)));
	if (C > 0.f) {
		const float s0 = fminf(fmaxf(-B * dxlo / C, dylo), dyhi), s1 = fminf(fmaxf(-B * dxhi / C, dylo), dyhi);
		qmin = fminf(qmin, fminf(q(dxlo, s0), q(dxhi, s1)));
	}
	if (A > 0.f) {
		const float s0 = fminf(fmaxf(-B * dylo / A, dxlo), dxhi), s1 = fminf(fmaxf(-B * dyhi / A, dxlo), dxhi);
		qmin = fminf(qmin, fminf(q(s0, dylo), q(s1, dyhi)));
	}
	return !(qmin > thr + pad);
}

// Stage one batch: entry i of the batch <- record of Gaussian list[pos(i)].
template <bool REVERSE>
__device__ __forceinline__ void stage_batch(GaussRec* s_rec, uint32_t* s_gid, uint64_t* bar, const GaussRec* __restrict__ rec,
                                            const uint32_t* __restrict__ point_list, uint32_t list_begin, int n, int batch, int tid)
{
	const int e = batch * RB + tid;  // entry number in traversal order
	const int cnt = min(RB, n - batch * RB);
	if (tid == 0) mbar_arrive_expect_tx(bar, (uint32_t)cnt * (uint32_t)sizeof(GaussRec));
	if (e < n) {
		const uint32_t pos = REVERSE ? (list_begin + (uint32_t)(n - 1 - e)) : (list_begin + (uint32_t)e);
		const uint32_t g = point_list[pos];
		s_gid[tid] = g;
		bulk_g2s(&s_rec[tid], &rec[g], (uint32_t)sizeof(GaussRec), bar);
	}
}

// Ordered compaction of `keep` flags over the block; returns total, writes kept thread ids to s_cidx.
__device__ __forceinline__ int compact_block(bool keep, uint8_t* s_cidx, int* s_wcnt, int tid)
{
	const int lane = tid & 31, warp = tid >> 5;
	const uint32_t bal = __ballot_sync(0xffffffffu, keep);
	if (lane == 0) s_wcnt[warp] = __popc(bal);
	__syncthreads();
	int base = 0, total = 0;
#pragma unroll
	for (int w = 0; w < RB / 32; w++) {
		const int c = s_wcnt[w];
		if (w < warp) base += c;
		total += c;
	}
	if (keep) s_cidx[base + __popc(bal & ((1u << lane) - 1u))] = (uint8_t)tid;
	__syncthreads();
	return total;
}

}  // namespace

// =================================================================================================
// Forward
// =================================================================================================
__global__ void __launch_bounds__(RB) render_fwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                        const GaussRec* __restrict__ rec, int W, int H, const float* __restrict__ bg_color,
                                                        float* __restrict__ out_color, float* __restrict__ final_T,
                                                        uint32_t* __restrict__ n_contrib)
{
	__shared__ __align__(16) GaussRec s_rec[2][RB];
	__shared__ uint32_t s_gid[2][RB];
	__shared__ uint8_t s_cidx[RB];
	__shared__ int s_wcnt[RB / 32];
	__shared__ __align__(8) uint64_t s_bar[2];

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile_x0 = blockIdx.x * PSB_TILE_X, tile_y0 = blockIdx.y * PSB_TILE_Y;
	const int px = tile_x0 + (warp & 1) * 8 + (lane & 7);
	const int py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
	const bool inside = px < W && py < H;
	const uint32_t pix_id = (uint32_t)W * py + px;
	const float2 pixf = make_float2((float)px, (float)py);
	const float rx0 = (float)tile_x0, ry0 = (float)tile_y0;
	const float rx1 = (float)(min(tile_x0 + PSB_TILE_X, W) - 1), ry1 = (float)(min(tile_y0 + PSB_TILE_Y, H) - 1);
	// pixel rectangle of this warp's 8x4 footprint (clamped to the image; degenerate if the warp is outside)
	const float wx0 = (float)(tile_x0 + (warp & 1) * 8), wy0 = (float)(tile_y0 + (warp >> 1) * 4);
	const float wx1 = fmaxf(wx0, fminf(wx0 + 7.f, (float)(W - 1))), wy1 = fmaxf(wy0, fminf(wy0 + 3.f, (float)(H - 1)));

	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];
	const int n = (int)(range.y - range.x);
	const int nbatch = (n + RB - 1) / RB;

	if (tid == 0) { mbar_init(&s_bar[0], 1); mbar_init(&s_bar[1], 1); mbar_fence_init(); }
	__syncthreads();
	if (nbatch > 0) stage_batch<false>(s_rec[0], s_gid[0], &s_bar[0], rec, point_list, range.x, n, 0, tid);

	bool done = !inside;
	float T = 1.0f;
	uint32_t last_contributor = 0;
	float C[3] = {0.f, 0.f, 0.f};

	for (int b = 0; b < nbatch; b++) {
		const int st = b & 1;
		// (the barrier also orders stage reuse: every thread has left batch b-1)
		if (__syncthreads_count(done) == RB) {
			mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));  // drain the in-flight copy of batch b before exiting
			break;
		}
		if (b + 1 < nbatch) stage_batch<false>(s_rec[st ^ 1], s_gid[st ^ 1], &s_bar[st ^ 1], rec, point_list, range.x, n, b + 1, tid);
		mbar_wait(&s_bar[st], (uint32_t)((b >> 1) & 1));

		const int cnt = min(RB, n - b * RB);
		bool keep = false;
		if (tid < cnt) keep = splat_reaches_tile(s_rec[st][tid].q0, s_rec[st][tid].q1, rx0, ry0, rx1, ry1);
		const int ccount = compact_block(keep, s_cidx, s_wcnt, tid);

		// Second, per-warp cull: each lane tests one surviving entry against this warp's 8x4 pixel footprint
		// (32 entries per ballot), then the warp walks only the entries that can reach one of its pixels.
		for (int c0 = 0; c0 < ccount; c0 += 32) {
			if (__all_sync(0xffffffffu, done)) break;
			const int kk = c0 + lane;
			int j = 0;
			bool hit = false;
			if (kk < ccount) {
				j = s_cidx[kk];
				hit = splat_reaches_tile(s_rec[st][j].q0, s_rec[st][j].q1, wx0, wy0, wx1, wy1);
			}
			uint32_t hits = __ballot_sync(0xffffffffu, hit);
			while (hits) {
				const int src = __ffs(hits) - 1;
				hits &= hits - 1;
				const int jj = __shfl_sync(0xffffffffu, j, src);
				if (done) continue;
				const float4 q0 = s_rec[st][jj].q0;
				const float4 q1 = s_rec[st][jj].q1;
				const float2 xy = make_float2(q0.x, q0.y);
				const float2 d = make_float2(xy.x - pixf.x, xy.y - pixf.y);
				const float4 con_o = make_float4(q0.z, q0.w, q1.x, q1.y);
				const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
				if (power > 0.0f) continue;
				if (power < q1.z) continue;  // below pmin: alpha < 1/255 for certain, skip exp (see GaussRec)
				const float alpha = min(0.99f, con_o.w * exp(power));
				if (alpha < 1.0f / 255.0f) continue;
				const float test_T = T * (1 - alpha);
				if (test_T < 0.0001f) { done = true; continue; }
				const float4 q2 = s_rec[st][jj].q2;
				C[0] += q2.x * alpha * T;
				C[1] += q2.y * alpha * T;
				C[2] += q2.z * alpha * T;
				T = test_T;
				last_contributor = (uint32_t)(b * RB + jj + 1);
			}
		}
	}

	if (inside) {
		final_T[pix_id] = T;
		n_contrib[pix_id] = last_contributor;
		const size_t HW = (size_t)H * W;
#pragma unroll
		for (int ch = 0; ch < 3; ch++) out_color[ch * HW + pix_id] = C[ch] + T * bg_color[ch];
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

// Backward batch geometry: 128 list entries per batch (staged by the first 128 threads), 256 pixel threads.
constexpr int RBB = 128;
constexpr int ACC_ROW = 73;  // floats per entry: 8 warps x 9 sums (+1 pad: conflict-free row reads)
struct BwdSmem {
	GaussRec rec[2][RBB];
	float acc[RBB * ACC_ROW];
	unsigned long long dirty[RBB];  // byte w != 0: warp w wrote its 9 sums for this entry
	uint32_t gid[2][RBB];
	uint64_t bar[2];
	int wcnt[RB / 32];
	uint8_t cidx[RBB];
};

__global__ void __launch_bounds__(RB) render_bwd_kernel(const uint2* __restrict__ ranges, const uint32_t* __restrict__ point_list,
                                                        const GaussRec* __restrict__ rec, int W, int H, const float* __restrict__ bg_color,
                                                        const float* __restrict__ final_Ts, const uint32_t* __restrict__ n_contrib,
                                                        const float* __restrict__ dL_dpixels, GradSink sink)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	BwdSmem& sm = *reinterpret_cast<BwdSmem*>(smem_raw);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const int tile_x0 = blockIdx.x * PSB_TILE_X, tile_y0 = blockIdx.y * PSB_TILE_Y;
	const int px = tile_x0 + (warp & 1) * 8 + (lane & 7);
	const int py = tile_y0 + (warp >> 1) * 4 + (lane >> 3);
	const bool inside = px < W && py < H;
	const uint32_t pix_id = (uint32_t)W * py + px;
	const float2 pixf = make_float2((float)px, (float)py);
	const float rx0 = (float)tile_x0, ry0 = (float)tile_y0;
	const float rx1 = (float)(min(tile_x0 + PSB_TILE_X, W) - 1), ry1 = (float)(min(tile_y0 + PSB_TILE_Y, H) - 1);
	// pixel rectangle of this warp's 8x4 footprint (clamped to the image; degenerate if the warp is outside)
	const float wx0 = (float)(tile_x0 + (warp & 1) * 8), wy0 = (float)(tile_y0 + (warp >> 1) * 4);
	const float wx1 = fmaxf(wx0, fminf(wx0 + 7.f, (float)(W - 1))), wy1 = fmaxf(wy0, fminf(wy0 + 3.f, (float)(H - 1)));

	const uint2 range = ranges[blockIdx.y * gridDim.x + blockIdx.x];

	const float T_final = inside ? final_Ts[pix_id] : 0;
	float T = T_final;
	const int last_contributor = inside ? (int)n_contrib[pix_id] : 0;

	// Nothing behind the deepest last contributor of the tile can receive gradient: start there.
	int maxc = __reduce_max_sync(0xffffffffu, last_contributor);
	const int warp_maxc = maxc;
	if (lane == 0) sm.wcnt[warp] = maxc;
	if (tid == 0) { mbar_init(&sm.bar[0], 1); mbar_init(&sm.bar[1], 1); mbar_fence_init(); }
	if (tid < RBB) sm.dirty[tid] = 0ull;
	__syncthreads();
	maxc = 0;
#pragma unroll
	for (int w = 0; w < RB / 32; w++) maxc = max(maxc, sm.wcnt[w]);
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

	float accum_rec[3] = {0.f, 0.f, 0.f};
	float dL_dpixel[3] = {0.f, 0.f, 0.f};
	if (inside) {
		const size_t HW = (size_t)H * W;
#pragma unroll
		for (int i = 0; i < 3; i++) dL_dpixel[i] = dL_dpixels[i * HW + pix_id];
	}
	float last_alpha = 0;
	float last_color[3] = {0.f, 0.f, 0.f};
	const float ddelx_dx = 0.5 * W;
	const float ddely_dy = 0.5 * H;
	float bg_dot_dpixel = 0;
#pragma unroll
	for (int i = 0; i < 3; i++) bg_dot_dpixel += bg_color[i] * dL_dpixel[i];
	const int my_slot = ((lane & 1) == 0) ? warp_reduce9_index(lane) : -1;
	unsigned char* dirty8 = reinterpret_cast<unsigned char*>(sm.dirty);

	for (int b = 0; b < nbatch; b++) {
		const int st = b & 1;
		__syncthreads();  // previous batch fully flushed; stage st^1 free
		if (b + 1 < nbatch) stage(b + 1, st ^ 1);
		mbar_wait(&sm.bar[st], (uint32_t)((b >> 1) & 1));

		const int cnt = min(RBB, n - b * RBB);
		bool keep = false;
		if (tid < cnt) keep = splat_reaches_tile(sm.rec[st][tid].q0, sm.rec[st][tid].q1, rx0, ry0, rx1, ry1);
		const int ccount = compact_block(keep, sm.cidx, sm.wcnt, tid);

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
			bool active = pos < last_contributor;
			const float4 q0 = sm.rec[st][j].q0;
			const float4 q1 = sm.rec[st][j].q1;
			const float2 xy = make_float2(q0.x, q0.y);
			const float2 d = make_float2(xy.x - pixf.x, xy.y - pixf.y);
			const float4 con_o = make_float4(q0.z, q0.w, q1.x, q1.y);
			const float power = -0.5f * (con_o.x * d.x * d.x + con_o.z * d.y * d.y) - con_o.y * d.x * d.y;
			active = active && !(power > 0.0f) && !(power < q1.z);  // q1.z = pmin, see GaussRec
			float G = 0.f, alpha = 0.f;
			if (active) {
				G = exp(power);
				alpha = min(0.99f, con_o.w * G);
				active = !(alpha < 1.0f / 255.0f);
			}
			if (!__any_sync(0xffffffffu, active)) continue;

			float v[9];
#pragma unroll
			for (int i = 0; i < 9; i++) v[i] = 0.f;
			if (active) {
				T = T / (1.f - alpha);
				const float dchannel_dcolor = alpha * T;
				const float4 q2 = sm.rec[st][j].q2;
				const float col[3] = {q2.x, q2.y, q2.z};
				float dL_dalpha = 0.0f;
#pragma unroll
				for (int ch = 0; ch < 3; ch++) {
					const float c = col[ch];
					accum_rec[ch] = last_alpha * last_color[ch] + (1.f - last_alpha) * accum_rec[ch];
					last_color[ch] = c;
					const float dL_dchannel = dL_dpixel[ch];
					dL_dalpha += (c - accum_rec[ch]) * dL_dchannel;
					v[6 + ch] = dchannel_dcolor * dL_dchannel;
				}
				dL_dalpha *= T;
				last_alpha = alpha;
				dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

				const float dL_dG = con_o.w * dL_dalpha;
				const float gdx = G * d.x;
				const float gdy = G * d.y;
				const float dG_ddelx = -gdx * con_o.x - gdy * con_o.y;
				const float dG_ddely = -gdy * con_o.z - gdx * con_o.y;
				v[0] = dL_dG * dG_ddelx * ddelx_dx;
				v[1] = dL_dG * dG_ddely * ddely_dy;
				v[2] = -0.5f * gdx * d.x * dL_dG;
				v[3] = -0.5f * gdx * d.y * dL_dG;
				v[4] = -0.5f * gdy * d.y * dL_dG;
				v[5] = G * dL_dalpha;
			}
			const float tot = warp_reduce9(v, lane);
			// each (warp, entry) pair is visited once per batch: plain stores, no shared-memory atomics
			if (my_slot >= 0) sm.acc[j * ACC_ROW + warp * 9 + my_slot] = tot;
			if (lane == 0) dirty8[j * 8 + warp] = 1;
		}
		}
		__syncthreads();

		// one set of global reductions per (Gaussian, tile)
		if (tid < cnt) {
			const unsigned long long dm = sm.dirty[tid];
			if (dm) {
				sm.dirty[tid] = 0ull;
				float a[9];
#pragma unroll
				for (int i = 0; i < 9; i++) a[i] = 0.f;
#pragma unroll
				for (int w = 0; w < RB / 32; w++) {
					if ((dm >> (8 * w)) & 0xffull) {
#pragma unroll
						for (int i = 0; i < 9; i++) a[i] += sm.acc[tid * ACC_ROW + w * 9 + i];
					}
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

int launch_render_forward(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                          float* out_color, float* final_T, uint32_t* n_contrib, cudaStream_t stream)
{
	dim3 grid(cam.grid_x, cam.grid_y, 1);
	render_fwd_kernel<<<grid, RB, 0, stream>>>(ranges, point_list, rec, cam.W, cam.H, bg, out_color, final_T, n_contrib);
	PSB_LAUNCH_OK();
	return 0;
}

int launch_render_backward(const Camera& cam, const uint2* ranges, const uint32_t* point_list, const GaussRec* rec, const float* bg,
                           const float* final_T, const uint32_t* n_contrib, const float* dL_dpix, const GradSink& sink,
                           cudaStream_t stream)
{
	dim3 grid(cam.grid_x, cam.grid_y, 1);
	static bool attr_set = false;
	if (!attr_set) {
		PSB_CUDA_OK(cudaFuncSetAttribute(render_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BwdSmem)));
		attr_set = true;
	}
	render_bwd_kernel<<<grid, RB, sizeof(BwdSmem), stream>>>(ranges, point_list, rec, cam.W, cam.H, bg, final_T, n_contrib, dL_dpix, sink);
	PSB_LAUNCH_OK();
	return 0;
}

}  // namespace psb
