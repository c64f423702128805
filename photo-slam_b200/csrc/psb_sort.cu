// Onesweep LSD radix sort for (u32 key, u32 value) pairs — see psb_sort.cuh.
//
// Per digit pass every thread block
//   1. takes a ticket (atomic) so tiles are processed in a globally monotonic order,
//   2. loads 4096 keys (16 per thread, warp-striped, fully coalesced) into registers,
//   3. ranks them per warp with one ballot per digit bit + one shared-memory atomic per run (stable inside the warp chunk),
//   4. turns per-warp digit counts into per-block counts, publishes them and resolves the cross-block
//      prefix per digit by decoupled look-back (one 32-bit status word = 2 flag bits + 30-bit count),
//   5. reorders keys/values through shared memory so each digit's run leaves as a contiguous,
//      coalesced store.
// HBM traffic per pass: 8 B read + 8 B written per element (+ 1 KiB of status per 4096 elements).
#include "psb_sort.cuh"

namespace psb {

namespace {

constexpr uint32_t ST_FLAG_AGG = 1u << 30;
constexpr uint32_t ST_FLAG_INC = 2u << 30;
constexpr uint32_t ST_VAL_MASK = (1u << 30) - 1;

__device__ __forceinline__ uint32_t ld_volatile(const uint32_t* p)
{
	uint32_t v;
	asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	return v;
}
__device__ __forceinline__ void st_volatile(uint32_t* p, uint32_t v)
{
	asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// All digit histograms of all passes in one read of the keys.
__global__ void __launch_bounds__(256) rs_histogram_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ n_dev,
                                                           uint32_t n_host, SortPlan plan, uint32_t* __restrict__ hist)
{
	__shared__ uint32_t sh[RS_MAX_PASS][RS_RADIX];
	const uint32_t n = n_dev ? (*n_dev > n_host ? 0u : *n_dev) : n_host;  // overflowed arena: sort nothing
	for (int i = threadIdx.x; i < RS_MAX_PASS * RS_RADIX; i += blockDim.x) (&sh[0][0])[i] = 0;
	__syncthreads();
	const uint32_t stride = gridDim.x * blockDim.x * 4;
	for (uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
		uint32_t k[4];
		if (i + 3 < n) {
			const uint4 v = *reinterpret_cast<const uint4*>(keys + i);
			k[0] = v.x; k[1] = v.y; k[2] = v.z; k[3] = v.w;
		} else {
#pragma unroll
			for (int j = 0; j < 4; j++) k[j] = (i + j < n) ? keys[i + j] : 0xFFFFFFFFu;
		}
#pragma unroll
		for (int j = 0; j < 4; j++) {
			if (i + j < n) {
				for (int p = 0; p < plan.npass; p++)
					atomicAdd(&sh[p][(k[j] >> plan.shift[p]) & ((1u << plan.bits[p]) - 1)], 1u);
			}
		}
	}
	__syncthreads();
	for (int i = threadIdx.x; i < plan.npass * RS_RADIX; i += blockDim.x) {
		const uint32_t c = (&sh[0][0])[i];
		if (c) atomicAdd(&hist[i], c);
	}
}

// In-place exclusive scan of each pass' 256-bin histogram. grid = npass, block = 256.
__global__ void __launch_bounds__(256) rs_scan_hist_kernel(uint32_t* __restrict__ hist)
{
	__shared__ uint32_t warp_tot[8];
	uint32_t* h = hist + blockIdx.x * RS_RADIX;
	const int t = threadIdx.x, lane = t & 31, w = t >> 5;
	const uint32_t v = h[t];
	uint32_t inc = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += u;
	}
	if (lane == 31) warp_tot[w] = inc;
	__syncthreads();
	uint32_t base = 0;
	for (int i = 0; i < w; i++) base += warp_tot[i];
	h[t] = base + inc - v;
}

template <bool IOTA>
__global__ void __launch_bounds__(RS_THREADS, 3) rs_onesweep_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in,
                                                                uint32_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                const uint32_t* __restrict__ n_dev, uint32_t n_host, int shift, int bits,
                                                                const uint32_t* __restrict__ hist_excl, uint32_t* __restrict__ ticket,
                                                                uint32_t* __restrict__ status)
{
	__shared__ uint32_t s_warp_hist[RS_THREADS / 32][RS_RADIX];
	__shared__ uint32_t s_keys[RS_TILE];
	__shared__ uint32_t s_vals[RS_TILE];
	__shared__ uint32_t s_local_start[RS_RADIX];
	__shared__ uint32_t s_digit_base[RS_RADIX];
	__shared__ uint32_t s_warp_tot[RS_THREADS / 32];
	__shared__ uint32_t s_tile;

	const uint32_t n = n_dev ? (*n_dev > n_host ? 0u : *n_dev) : n_host;  // overflowed arena: sort nothing
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0) s_tile = atomicAdd(ticket, 1u);
	for (int i = tid; i < (RS_THREADS / 32) * RS_RADIX; i += RS_THREADS) (&s_warp_hist[0][0])[i] = 0;
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint32_t base = tile * RS_TILE;
	if (base >= n) return;
	const uint32_t tile_count = min((uint32_t)RS_TILE, n - base);
	const uint32_t mask = (1u << bits) - 1u;
	const int radix = 1 << bits;

	// 1. load (warp-striped: warp w owns [w*512, w*512+512), item i of lane l is w*512 + i*32 + l)
	uint32_t key[RS_ITEMS], val[RS_ITEMS], rank[RS_ITEMS];
	const uint32_t wbase = base + warp * (32 * RS_ITEMS);
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const uint32_t idx = wbase + i * 32 + lane;
		const bool ok = idx < n;
		key[i] = ok ? keys_in[idx] : 0xFFFFFFFFu;
		if (IOTA) val[i] = idx;
		else val[i] = ok ? vals_in[idx] : 0u;
	}

	// 2. stable rank inside the warp chunk. Lanes holding the same digit are found with one ballot per digit bit
	//    (independent across the 16 items, unlike MATCH.ANY whose long latency serialises the loop); the lowest
	//    such lane reserves the run in the warp's counter row with ONE shared-memory atomic (same-address atomics
	//    of a warp retire in program order, which keeps the ranking stable across the 16 rounds).
	const uint32_t lanemask_lt = (1u << lane) - 1u;
	constexpr int RG = 8;  // items ranked per group (bounds the live registers, still 8-way ILP)
#pragma unroll
	for (int g0 = 0; g0 < RS_ITEMS; g0 += RG) {
		uint32_t pre[RG];
#pragma unroll
		for (int u = 0; u < RG; u++) {
			const int i = g0 + u;
			const uint32_t d = (key[i] >> shift) & mask;
			uint32_t peers = 0xffffffffu;
			for (int bit = 0; bit < bits; bit++) {
				const bool one = (d >> bit) & 1u;
				const uint32_t bal = __ballot_sync(0xffffffffu, one);
				peers &= one ? bal : ~bal;
			}
			const int leader = __ffs(peers) - 1;
			rank[i] = (uint32_t)__popc(peers & lanemask_lt) | ((uint32_t)leader << 8);
			pre[u] = 0;
			if (lane == leader) pre[u] = atomicAdd(&s_warp_hist[warp][d], (uint32_t)__popc(peers));
		}
#pragma unroll
		for (int u = 0; u < RG; u++) {
			const int i = g0 + u;
			rank[i] = (rank[i] & 0xffu) + __shfl_sync(0xffffffffu, pre[u], (int)(rank[i] >> 8));
		}
	}
	__syncthreads();

	// 3. per-digit: exclusive offsets over warps, block count, cross-block prefix by look-back
	uint32_t count = 0, excl_prev = 0;
	if (tid < radix) {
#pragma unroll
		for (int w = 0; w < RS_THREADS / 32; w++) {
			const uint32_t c = s_warp_hist[w][tid];
			s_warp_hist[w][tid] = count;
			count += c;
		}
		uint32_t* st = status + (size_t)tile * RS_RADIX + tid;
		if (tile == 0) {
			st_volatile(st, ST_FLAG_INC | count);
		} else {
			st_volatile(st, ST_FLAG_AGG | count);
			// Decoupled look-back. Tiles of one wave start together, so the inclusive prefix of the nearest
			// predecessors is usually not published yet and the walk has to add up to a few hundred aggregates:
			// read LB predecessors per round trip (independent loads) instead of one.
			constexpr int LB = 8;
			int p = (int)tile - 1;
			bool found = false;
			while (!found) {
				uint32_t sv[LB];
#pragma unroll
				for (int u = 0; u < LB; u++) sv[u] = (p - u >= 0) ? ld_volatile(status + (size_t)(p - u) * RS_RADIX + tid) : ST_FLAG_INC;
#pragma unroll
				for (int u = 0; u < LB; u++) {
					if (found) break;
					uint32_t s = sv[u];
					while ((s >> 30) == 0u) s = ld_volatile(status + (size_t)(p - u) * RS_RADIX + tid);  // not published yet: poll this one
					excl_prev += s & ST_VAL_MASK;
					if ((s >> 30) == 2u) found = true;
				}
				p -= LB;
			}
			st_volatile(st, ST_FLAG_INC | ((excl_prev + count) & ST_VAL_MASK));
		}
	}
	// block-wide exclusive scan of `count` over digits (threads >= radix contribute 0)
	uint32_t inc = count;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
		if (lane >= o) inc += u;
	}
	if (lane == 31) s_warp_tot[warp] = inc;
	__syncthreads();
	uint32_t wb = 0;
	for (int i = 0; i < warp; i++) wb += s_warp_tot[i];
	const uint32_t local_start = wb + inc - count;
	if (tid < RS_RADIX) {
		s_local_start[tid] = local_start;
		s_digit_base[tid] = (tid < radix ? hist_excl[tid] : 0u) + excl_prev - local_start;
	}
	__syncthreads();

	// 4. reorder through shared memory
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const uint32_t d = (key[i] >> shift) & mask;
		const uint32_t pos = s_local_start[d] + s_warp_hist[warp][d] + rank[i];
		s_keys[pos] = key[i];
		s_vals[pos] = val[i];
	}
	__syncthreads();

	// 5. coalesced scatter: consecutive local positions of one digit map to consecutive addresses
#pragma unroll
	for (int i = 0; i < RS_ITEMS; i++) {
		const uint32_t k = i * RS_THREADS + tid;
		if (k < tile_count) {
			const uint32_t kk = s_keys[k];
			const uint32_t d = (kk >> shift) & mask;
			const uint32_t dst = s_digit_base[d] + k;
			keys_out[dst] = kk;
			vals_out[dst] = s_vals[k];
		}
	}
}

}  // namespace

SortPlan make_sort_plan(int nbits)
{
	SortPlan p{};
	if (nbits < 1) nbits = 1;
	if (nbits > 32) nbits = 32;
	p.npass = (nbits + 7) / 8;
	const int per = (nbits + p.npass - 1) / p.npass;
	int s = 0;
	for (int i = 0; i < p.npass; i++) {
		p.shift[i] = s;
		p.bits[i] = (s + per <= nbits) ? per : (nbits - s);
		if (p.bits[i] < 1) p.bits[i] = 1;
		s += p.bits[i];
	}
	return p;
}

static inline size_t sort_tiles(size_t n) { return (n + RS_TILE - 1) / RS_TILE; }

size_t sort_scratch_bytes(size_t max_n, int npass)
{
	const size_t words = (size_t)RS_MAX_PASS * RS_RADIX + 16 + (size_t)npass * sort_tiles(max_n) * RS_RADIX;
	return align_up(words * sizeof(uint32_t), 128);
}

int radix_sort_pairs(uint32_t* keys[2], uint32_t* vals[2], bool iota_vals, const uint32_t* n_dev, size_t n_host,
                     const SortPlan& plan, void* scratch, size_t scratch_bytes, cudaStream_t stream)
{
	if (n_host == 0) return 0;
	const size_t need = sort_scratch_bytes(n_host, plan.npass);
	if (scratch_bytes < need) { set_error_msg("radix_sort_pairs: scratch too small"); return -3; }
	uint32_t* w = reinterpret_cast<uint32_t*>(scratch);
	uint32_t* hist = w;
	uint32_t* tickets = w + RS_MAX_PASS * RS_RADIX;
	uint32_t* status = tickets + 16;
	const size_t ntiles = sort_tiles(n_host);
	PSB_CUDA_OK(cudaMemsetAsync(scratch, 0, need, stream));
	int hgrid = (int)((n_host + 256 * 4 * 8 - 1) / (256 * 4 * 8));
	if (hgrid > 148 * 8) hgrid = 148 * 8;
	if (hgrid < 1) hgrid = 1;
	rs_histogram_kernel<<<hgrid, 256, 0, stream>>>(keys[0], n_dev, (uint32_t)n_host, plan, hist);
	PSB_LAUNCH_OK();
	rs_scan_hist_kernel<<<plan.npass, 256, 0, stream>>>(hist);
	PSB_LAUNCH_OK();
	for (int p = 0; p < plan.npass; p++) {
		const int src = p & 1, dst = (p + 1) & 1;
		uint32_t* st = status + (size_t)p * ntiles * RS_RADIX;
		if (p == 0 && iota_vals)
			rs_onesweep_kernel<true><<<(unsigned)ntiles, RS_THREADS, 0, stream>>>(keys[src], nullptr, keys[dst], vals[dst], n_dev, (uint32_t)n_host,
			                                                                     plan.shift[p], plan.bits[p], hist + p * RS_RADIX, tickets + p, st);
		else
			rs_onesweep_kernel<false><<<(unsigned)ntiles, RS_THREADS, 0, stream>>>(keys[src], vals[src], keys[dst], vals[dst], n_dev, (uint32_t)n_host,
			                                                                      plan.shift[p], plan.bits[p], hist + p * RS_RADIX, tickets + p, st);
		PSB_LAUNCH_OK();
	}
	return 0;
}

}  // namespace psb
