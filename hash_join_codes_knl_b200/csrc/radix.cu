// radix.cu -- one radix-partitioning pass = four kernels on one stream:
//   k_make_items : cut every parent partition into work items of <= chunk tuples
//   k_hist       : per-item digit histogram in shared memory        (reference: histogram, cpra2.cpp:801-880)
//   k_scan       : single-pass decoupled-look-back prefix sum over   (reference: interleave, phj.cpp:1263-1291)
//                  the (parent, digit, item) ordered counts
//   k_scatter    : tile-wise rank + shared-memory reorder + run-wise (reference: partition / partition_shared +
//                  coalesced stores                                   flush, cpra2.cpp:882-1075, phj.cpp:877-1028)
// The work item plays the role of the reference's thread: it owns a contiguous chunk of the
// input, its counts row is the thread's counts[] and the scan hands it one start offset per
// digit, so items never synchronise while scattering.
//
// Also here:
//   k_hist_small   : k_hist for <= 8 digits (CPRA's GPU-assign pass), register counters
//   k_scatter_bulk : the scatter whose output columns live in other GPUs' memory; digit runs leave the SM as
//                    TMA bulk copies (cp.async.bulk) -- the fused exchange of CPRA (product path for N > 1)
//   k_hist_global  : whole-column histogram behind the public hjb_histogram
// and four schedule experiments for the local scatter that pass every parity test but measured slower
// than k_scatter; they stay selectable (HJB_SCATTER_VARIANT) and out of the default path:
//   k_scatter_bulk<.., PEER=false>  local TMA bulk copies (7) / 16-byte vector stores from a line-aligned tile (8)
//   k_scatter_ov                    stream of tile t-1 overlapped with the plan of tile t (10)
//   k_scatter_fx                    fixed digit regions, rank + placement in one step (11)
//   k_scatter_2x                    two sub-tiles ranked per plan, 16384 tuples placed and streamed per round (12)
#include "hj_device.cuh"
#include "hj_internal.h"
#include <atomic>
#include <stdlib.h>

namespace hjb {

// ------------------------------------------------------------------ work items

__global__ void __launch_bounds__(1024)
k_make_items(const uint32_t *__restrict__ parent_off, uint32_t np, uint64_t n, uint32_t chunk,
             uint32_t *__restrict__ item_prefix, uint32_t *__restrict__ child_off, uint32_t child_total_idx)
{
	__shared__ uint32_t warp_totals[34];
	const uint32_t per = (np + blockDim.x - 1) / blockDim.x;
	const uint32_t q0 = threadIdx.x * per;
	uint32_t local = 0;
	for (uint32_t q = q0; q < q0 + per && q < np; ++q) {
		const uint64_t size = parent_off ? (uint64_t)(parent_off[q + 1] - parent_off[q]) : n;
		const uint32_t items = size ? (uint32_t)((size + chunk - 1) / chunk) : 1u;   // empty parents keep one (empty) item
		local += items;
	}
	uint32_t total;
	uint32_t run = block_exclusive_scan(local, warp_totals, &total);
	for (uint32_t q = q0; q < q0 + per && q < np; ++q) {
		const uint64_t size = parent_off ? (uint64_t)(parent_off[q + 1] - parent_off[q]) : n;
		item_prefix[q] = run;
		run += size ? (uint32_t)((size + chunk - 1) / chunk) : 1u;
	}
	if (threadIdx.x == 0) {
		item_prefix[np] = total;
		child_off[child_total_idx] = (uint32_t)n;       // end sentinel of the child offsets
	}
}

struct ItemRange {
	uint64_t beg, end;
};

__device__ __forceinline__ bool locate_item(const uint32_t *item_prefix, uint32_t np, const uint32_t *parent_off,
                                            uint64_t n, uint32_t chunk, uint32_t item, ItemRange *r)
{
	if (item >= item_prefix[np]) return false;
	const uint32_t q = upper_parent(item_prefix, np, item);
	const uint32_t j = item - item_prefix[q];
	const uint64_t pbeg = parent_off ? parent_off[q] : 0, pend = parent_off ? parent_off[q + 1] : n;
	uint64_t beg = pbeg + (uint64_t)j * chunk;
	if (beg > pend) beg = pend;
	uint64_t end = beg + chunk;
	if (end > pend) end = pend;
	r->beg = beg;
	r->end = end;
	return true;
}

// Fetches the absolutely aligned group of four elements g (indices 4g .. 4g+3) of a column with
// one 128-bit load, so that consecutive threads read consecutive 16-byte words whatever the
// alignment of the range being processed; callers mask the ragged first / last group.
// `n` is the column length: a vector load never crosses it.
__device__ __forceinline__ void load_group4(const uint32_t *col, uint64_t g, uint64_t n, uint32_t (&out)[4])
{
	const uint64_t idx = g << 2;
	if (idx + 3 < n) {
		const uint4 w = ldg_stream_u4(reinterpret_cast<const uint4 *>(col) + g);
		out[0] = w.x; out[1] = w.y; out[2] = w.z; out[3] = w.w;
	} else {
#pragma unroll
		for (int e = 0; e < 4; ++e) out[e] = idx + e < n ? col[idx + e] : 0;
	}
}

// Rank for small fan-outs: the lanes of a warp that hold the same digit reserve their ranks with ONE
// shared-memory atomic (match.any groups them).  With a few digits only, per-lane atomics on the
// same counter serialise 16-fold and bound the whole pass.
__device__ __forceinline__ uint32_t rank_aggregated(uint32_t *cnt, uint32_t d, bool valid)
{
	const unsigned m = __match_any_sync(kFullMask, valid ? d : 0xFFFFFFFFu);
	const int leader = __ffs(m) - 1;
	uint32_t base = 0;
	if (valid && (int)lane_id() == leader) base = atomicAdd(&cnt[d], (uint32_t)__popc(m));
	base = __shfl_sync(kFullMask, base, leader);
	return valid ? (d << 16) | (base + (uint32_t)__popc(m & lanemask_lt())) : 0xFFFFFFFFu;
}

// ------------------------------------------------------------------ histogram

__global__ void __launch_bounds__(kHistThreads)
k_hist(const uint32_t *__restrict__ keys, uint64_t n, uint32_t np, const uint32_t *__restrict__ parent_off,
       const uint32_t *__restrict__ item_prefix, uint32_t chunk, uint32_t factor, int rshift, int bits,
       uint32_t *__restrict__ counts)
{
	extern __shared__ uint32_t s_hist[];
	const uint32_t F = 1u << bits, mask = F - 1;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	for (uint32_t p = threadIdx.x; p < F; p += blockDim.x) s_hist[p] = 0;
	__syncthreads();
	const uint64_t g_end = (r.end + 3) >> 2;
	for (uint64_t g = (r.beg >> 2) + threadIdx.x; g < g_end; g += blockDim.x) {
		if ((g << 2) >= r.beg && (g << 2) + 4 <= r.end) {          // interior group: one 128-bit load, no per-key range test
			const uint4 w = ldg_stream_u4(reinterpret_cast<const uint4 *>(keys) + g);
			atomicAdd(&s_hist[radix_digit(hash_mul(w.x, factor), rshift, mask)], 1u);
			atomicAdd(&s_hist[radix_digit(hash_mul(w.y, factor), rshift, mask)], 1u);
			atomicAdd(&s_hist[radix_digit(hash_mul(w.z, factor), rshift, mask)], 1u);
			atomicAdd(&s_hist[radix_digit(hash_mul(w.w, factor), rshift, mask)], 1u);
			continue;
		}
		uint32_t k[4];
		load_group4(keys, g, n, k);
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			const uint64_t idx = (g << 2) + e;
			if (idx >= r.beg && idx < r.end) atomicAdd(&s_hist[radix_digit(hash_mul(k[e], factor), rshift, mask)], 1u);
		}
	}
	__syncthreads();
	uint32_t *row = counts + (size_t)blockIdx.x * F;
	for (uint32_t p = threadIdx.x; p < F; p += blockDim.x) row[p] = s_hist[p];
}

// k_hist for at most 8 digits (CPRA's GPU-assign pass): per-lane shared-memory atomics on so few
// counters serialise 16-fold; private counters in registers, one atomic per warp and digit instead
__global__ void __launch_bounds__(kHistThreads)
k_hist_small(const uint32_t *__restrict__ keys, uint64_t n, uint32_t np, const uint32_t *__restrict__ parent_off,
             const uint32_t *__restrict__ item_prefix, uint32_t chunk, uint32_t factor, int rshift, int bits,
             uint32_t *__restrict__ counts)
{
	__shared__ uint32_t s_hist[8];
	const uint32_t F = 1u << bits, mask = F - 1;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	if (threadIdx.x < 8) s_hist[threadIdx.x] = 0;
	__syncthreads();
	const uint64_t g_end = (r.end + 3) >> 2;
	uint32_t c[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	for (uint64_t g = (r.beg >> 2) + threadIdx.x; g < g_end; g += blockDim.x) {
		uint32_t k[4];
		load_group4(keys, g, n, k);
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			const uint64_t idx = (g << 2) + e;
			const uint32_t d = (idx >= r.beg && idx < r.end) ? radix_digit(hash_mul(k[e], factor), rshift, mask) : 8u;
#pragma unroll
			for (int j = 0; j < 8; ++j) c[j] += d == (uint32_t)j;
		}
	}
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		const uint32_t t = (uint32_t)warp_sum_u64(c[j]);
		if (lane_id() == 0 && j < (int)F && t) atomicAdd(&s_hist[j], t);
	}
	__syncthreads();
	if (threadIdx.x < F) counts[(size_t)blockIdx.x * F + threadIdx.x] = s_hist[threadIdx.x];
}

// whole-column histogram into one global counts[F] (public hjb_histogram)
__global__ void __launch_bounds__(kHistThreads)
k_hist_global(const uint32_t *__restrict__ keys, uint64_t n, uint32_t factor, int rshift, int bits,
              uint32_t *__restrict__ counts)
{
	extern __shared__ uint32_t s_hist[];
	const uint32_t F = 1u << bits, mask = F - 1;
	for (uint32_t p = threadIdx.x; p < F; p += blockDim.x) s_hist[p] = 0;
	__syncthreads();
	const uint64_t groups = (n + 3) >> 2;
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t k[4];
		load_group4(keys, g, n, k);
#pragma unroll
		for (int e = 0; e < 4; ++e)
			if ((g << 2) + e < n) atomicAdd(&s_hist[radix_digit(hash_mul(k[e], factor), rshift, mask)], 1u);
	}
	__syncthreads();
	for (uint32_t p = threadIdx.x; p < F; p += blockDim.x)
		if (s_hist[p]) atomicAdd(&counts[p], s_hist[p]);
}

// ------------------------------------------------------------------ scan (decoupled look-back)

constexpr uint64_t kFlagAggregate = 1ull << 62, kFlagInclusive = 2ull << 62, kFlagMask = 3ull << 62;

// Exclusive prefix sum over the counts taken in (parent, digit, item) order: the value that
// lands in counts[item][digit] is the absolute output position of that item's first tuple with
// that digit.  One pass: each tile publishes its aggregate, then looks back over its
// predecessors' status words until it meets an inclusive prefix (Merrill & Garland).
__global__ void __launch_bounds__(kScanThreads)
k_scan(const uint32_t *__restrict__ item_prefix, uint32_t np, int bits, uint32_t *__restrict__ counts,
       uint32_t *__restrict__ child_off, uint64_t *__restrict__ status, uint32_t *__restrict__ tile_counter)
{
	__shared__ uint32_t warp_totals[34];
	__shared__ uint32_t s_tile;
	__shared__ uint32_t s_excl;
	const uint32_t F = 1u << bits;
	if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);   // tiles start in look-back order
	__syncthreads();
	const uint32_t tile = s_tile;
	const uint64_t E = (uint64_t)item_prefix[np] * F;
	const uint64_t tile_base = (uint64_t)tile * (kScanThreads * kScanItems);
	if (tile_base >= E) return;

	uint32_t addr[kScanItems], child[kScanItems], v[kScanItems];
	uint64_t i = tile_base + (uint64_t)threadIdx.x * kScanItems;
	uint32_t q = 0, nq = 1, p2 = 0, j = 0;
	if (i < E) {
		q = upper_parent(item_prefix, np, (uint32_t)(i >> bits));
		nq = item_prefix[q + 1] - item_prefix[q];
		const uint32_t local = (uint32_t)(i - (uint64_t)item_prefix[q] * F);
		p2 = local / nq;
		j = local - p2 * nq;
	}
	uint32_t sum = 0;
#pragma unroll
	for (uint32_t k = 0; k < kScanItems; ++k) {
		v[k] = 0;
		addr[k] = 0xFFFFFFFFu;
		child[k] = 0xFFFFFFFFu;
		if (i + k < E) {
			addr[k] = (item_prefix[q] + j) * F + p2;
			v[k] = counts[addr[k]];
			if (j == 0) child[k] = q * F + p2;
			sum += v[k];
			if (++j == nq) {
				j = 0;
				if (++p2 == F) {
					p2 = 0;
					++q;
					if (q < np) nq = item_prefix[q + 1] - item_prefix[q];
				}
			}
		}
	}
	uint32_t block_total;
	const uint32_t thread_excl = block_exclusive_scan(sum, warp_totals, &block_total);

	if (threadIdx.x < 32) {
		const uint32_t lane = threadIdx.x;
		if (tile == 0) {
			if (lane == 0) {
				st_volatile_u64(&status[0], kFlagInclusive | block_total);
				s_excl = 0;
			}
		} else {
			if (lane == 0) st_volatile_u64(&status[tile], kFlagAggregate | block_total);
			uint64_t excl = 0;
			int look = (int)tile - 1;
			while (true) {
				const int idx = look - (int)lane;
				uint64_t w = idx >= 0 ? ld_volatile_u64(&status[idx]) : kFlagInclusive;
				while (__any_sync(kFullMask, (w & kFlagMask) == 0)) {
					if ((w & kFlagMask) == 0) w = ld_volatile_u64(&status[idx]);
				}
				const unsigned incl = __ballot_sync(kFullMask, (w & kFlagMask) == kFlagInclusive);
				uint64_t contrib = w & ~kFlagMask;
				if (incl) {
					const int first = __ffs(incl) - 1;     // nearest predecessor holding an inclusive prefix
					if ((int)lane > first) contrib = 0;
					excl += warp_sum_u64(contrib);
					break;
				}
				excl += warp_sum_u64(contrib);
				look -= 32;
			}
			if (lane == 0) {
				st_volatile_u64(&status[tile], kFlagInclusive | (excl + block_total));
				s_excl = (uint32_t)excl;
			}
		}
	}
	__syncthreads();
	uint32_t run = s_excl + thread_excl;
#pragma unroll
	for (uint32_t k = 0; k < kScanItems; ++k) {
		if (addr[k] != 0xFFFFFFFFu) {
			counts[addr[k]] = run;
			if (child[k] != 0xFFFFFFFFu) child_off[child[k]] = run;
			run += v[k];
		}
	}
}

// ------------------------------------------------------------------ scatter

// One tile = THREADS * 8 tuples = THREADS * 2 absolutely aligned groups; thread t owns groups t
// and t + THREADS of the tile (coalesced 128-bit loads).
template <int THREADS, int G, bool FULL>
__device__ __forceinline__ void load_col8(uint32_t (&x)[4 * G], uint32_t &ok, const uint32_t *col, uint64_t g0, uint64_t g_end,
                                          uint64_t beg, uint64_t end, uint64_t n)
{
	ok = 0;
#pragma unroll
	for (int h = 0; h < G; ++h) {
		const uint64_t g = g0 + threadIdx.x + (uint64_t)h * THREADS;
		if (FULL) {
			const uint4 w = ldg_stream_u4(reinterpret_cast<const uint4 *>(col) + g);
			x[4 * h + 0] = w.x; x[4 * h + 1] = w.y; x[4 * h + 2] = w.z; x[4 * h + 3] = w.w;
			ok |= 0xFu << (4 * h);
		} else {
			uint32_t k4[4] = {0, 0, 0, 0};
			if (g < g_end) load_group4(col, g, n, k4);
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				const uint64_t idx = (g << 2) + e;
				x[4 * h + e] = k4[e];
				if (g < g_end && idx >= beg && idx < end) ok |= 1u << (4 * h + e);
			}
		}
	}
}

// Per tile: (1) every tuple takes a rank inside its digit with a shared-memory atomicAdd,
// (2) warp 0 turns the digit counts into tile offsets and decides, per digit, how far the
// item's output may be flushed, (3) tuples are placed into shared memory grouped by digit,
// (4) the tile is streamed out, neighbouring threads writing neighbouring addresses of one
// partition's run.  The key loads of the NEXT tile are issued before (1) and stay in flight
// through all four steps; the payload loads of this tile are issued before (1) and first
// needed in (3).
//
// Software write-combining (the reference's per-partition staging buffers, cpra2.cpp:976-1008,
// flush cpra2.cpp:711-729): a digit's run is only written up to the last 32-byte sector boundary
// of its output position; the < 8 tuples beyond it wait in a per-digit carry buffer and lead
// the digit's run of the next tile.  Every store but an item's first and last per digit then
// ends on a sector boundary, so L2 rarely has to fetch the rest of a half-written sector from HBM.
// (WC is on for fan-outs <= 256, where the carry buffers fit; wider passes write runs as is.)
// dynamic shared memory: cnt base fpos oldp wpos pend [F] | golim[F] (uint2) | buf[TILE] (uint2) | carry[F*8] (uint2)
constexpr uint32_t kLocalCarry = 8;    // tuples per 32-byte sector of a 4-byte column

// PEER: digit d's run does not go to keys_out / vals_out but to peers.k[d] / peers.v[d] -- the
// receive buffers of GPU d, mapped into this process (CUDA IPC) and pre-offset so that the
// position the scan produced indexes them directly.  The stores then travel over NVLink: the
// GPU-assign pass of CPRA and its all-to-all are one kernel.
// debug (HJB_SCATTER_CLOCKS=1 selects the CLK instantiation): cycles thread 0 of every CTA spent per phase of a tile
__device__ unsigned long long g_scatter_clk[8];

template <int THREADS, int MINB, bool PREFETCH, bool PEER, int G = 2, bool CLK = false>
__global__ void __launch_bounds__(THREADS, MINB)
k_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
          const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
          uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets,
          uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out, const PeerTable peers)
{
	constexpr int IT = 4 * G;                                   // tuples per thread and tile
	constexpr uint32_t kCarry = PEER ? kPeerCarry : kLocalCarry;  // write-combining granule in tuples
	constexpr uint32_t TILE = THREADS * IT, kGroupsPerTile = TILE / 4;
	extern __shared__ __align__(16) uint32_t s_mem[];
	__shared__ uint32_t warp_totals[34];
	__shared__ uint32_t s_tile_n;
	__shared__ uint32_t *s_pk[PEER ? 64 : 1], *s_pv[PEER ? 64 : 1];
	const uint32_t F = 1u << bits, mask = F - 1;
	const bool wc = F <= 256;
	uint32_t *cnt = s_mem, *base = cnt + F, *fpos = base + F, *oldp = fpos + F, *wpos = oldp + F, *pend = wpos + F;
	uint2 *golim = reinterpret_cast<uint2 *>(pend + F);                 // x: global offset of tile index 0, y: flush limit
	uint2 *buf = golim + F;
	uint2 *carry = buf + TILE;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	const uint32_t *row = offsets + (size_t)blockIdx.x * F;
	for (uint32_t p = threadIdx.x; p < F; p += THREADS) {
		wpos[p] = row[p] + (PEER ? peers.bias[p & 63] : 0u);
		pend[p] = 0;
		cnt[p] = 0;
	}
	if (PEER && threadIdx.x < 64) {
		s_pk[threadIdx.x] = peers.k[threadIdx.x];
		s_pv[threadIdx.x] = peers.v[threadIdx.x];
	}
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	auto tile_is_full = [&](uint64_t g0) {
		return (g0 << 2) >= r.beg && ((g0 + kGroupsPerTile) << 2) <= r.end;    // r.end <= n: vector loads stay inside
	};
	uint32_t key[IT], nkey[IT], val[IT], ok = 0, nok = 0;
	if (PREFETCH && g_beg < g_end) {
		if (tile_is_full(g_beg)) load_col8<THREADS, G, true>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
		else load_col8<THREADS, G, false>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
	}
	__syncthreads();
	long long clk_t = 0;
#define SCATTER_MARK(k)                                                              \
	if (CLK && threadIdx.x == 0) {                                                   \
		const long long now_ = clock64();                                            \
		atomicAdd(&g_scatter_clk[k], (unsigned long long)(now_ - clk_t));            \
		clk_t = now_;                                                                \
	}
	if (CLK && threadIdx.x == 0) clk_t = clock64();
	for (uint64_t g0 = g_beg; g0 < g_end; g0 += kGroupsPerTile) {
		const uint64_t g1 = g0 + kGroupsPerTile;
		const bool last = g1 >= g_end;
		const bool full = tile_is_full(g0);
		if (PREFETCH) {
#pragma unroll
			for (int e = 0; e < IT; ++e) key[e] = nkey[e];
			ok = nok;
		} else {
			if (full) load_col8<THREADS, G, true>(key, ok, keys, g0, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(key, ok, keys, g0, g_end, r.beg, r.end, n);
		}
		uint32_t vok;
		if (G < 4) {          // 16 tuples per thread: the payloads are fetched after step (1) to keep its register need down
			if (full) load_col8<THREADS, G, true>(val, vok, vals, g0, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(val, vok, vals, g0, g_end, r.beg, r.end, n);
		}
		if (PREFETCH && !last) {
			if (tile_is_full(g1)) load_col8<THREADS, G, true>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
		}
		// (1) rank: digit << 16 | rank-in-digit (rank < TILE <= 2^16, digit < 2^11).  Interior tiles
		// (all but an item's first and last) have every element valid: no per-element predicate.
		// G >= 4 (16 tuples per thread): only the 16-bit ranks are kept, two per register; the digit is
		// recomputed from the key when the tuple is placed
		constexpr bool PACK = G >= 4;
		uint32_t dr[PACK ? IT / 2 : IT];
		if (PACK) {
#pragma unroll
			for (int e = 0; e < IT; e += 2) {
				const uint32_t d0 = radix_digit(hash_mul(key[e], factor), rshift, mask);
				const uint32_t d1 = radix_digit(hash_mul(key[e + 1], factor), rshift, mask);
				const uint32_t r0 = (full || ((ok >> e) & 1u)) ? atomicAdd(&cnt[d0], 1u) : 0xFFFFu;
				const uint32_t r1 = (full || ((ok >> (e + 1)) & 1u)) ? atomicAdd(&cnt[d1], 1u) : 0xFFFFu;
				dr[e / 2] = r0 | (r1 << 16);
			}
		} else if (bits <= 4) {
#pragma unroll
			for (int e = 0; e < IT; ++e)
				dr[e] = rank_aggregated(cnt, radix_digit(hash_mul(key[e], factor), rshift, mask), full || ((ok >> e) & 1u));
		} else if (full) {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				dr[e] = (d << 16) | atomicAdd(&cnt[d], 1u);
			}
		} else {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				dr[e] = (ok >> e) & 1u ? (d << 16) | atomicAdd(&cnt[d], 1u) : 0xFFFFFFFFu;
			}
		}
		if (G >= 4) {
			if (full) load_col8<THREADS, G, true>(val, vok, vals, g0, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(val, vok, vals, g0, g_end, r.beg, r.end, n);
		}
		SCATTER_MARK(0)        // loads issued, keys arrived, ranks taken
		__syncthreads();
		SCATTER_MARK(1)        // waiting for the slowest warp's ranks
		// (2) per digit: tile offset, global offset, flush limit, what stays pending
		auto plan_digit = [&](uint32_t p, uint32_t c, uint32_t run) {
			const uint32_t w = wpos[p], pe = pend[p], endpos = w + pe + c;
			uint32_t lim = (last || !wc) ? endpos : (endpos & ~(kCarry - 1));
			const bool flush = lim > w;
			if (!flush) lim = w;
			base[p] = run;
			golim[p] = make_uint2(w + pe - run, lim);
			fpos[p] = w;
			oldp[p] = flush ? pe : 0;
			wpos[p] = lim;
			pend[p] = endpos - lim;
			cnt[p] = 0;
		};
		if (F <= THREADS) {
			// one digit per thread (threads >= F idle): warp scans, then the warp totals through shared memory
			const uint32_t p = threadIdx.x;
			const uint32_t nw = (F + 31) >> 5;                 // warps that own digits
			uint32_t c = 0, incl = 0;
			if (p < nw * 32) {
				c = p < F ? cnt[p] : 0;
				incl = warp_inclusive_scan_u32(c);
				if (lane_id() == 31) warp_totals[p >> 5] = incl;
			}
			__syncthreads();
			if (p < nw * 32) {
				const uint32_t t = lane_id() < nw ? warp_totals[lane_id()] : 0;
				const uint32_t tincl = warp_inclusive_scan_u32(t);
				const uint32_t before = __shfl_sync(kFullMask, tincl - t, p >> 5);
				if (p < F) plan_digit(p, c, before + incl - c);
				const uint32_t tile_total = __shfl_sync(kFullMask, tincl, 31);
				if (p == 0) s_tile_n = tile_total;
			}
			__syncthreads();
		} else {
			const uint32_t ept = (F + THREADS - 1) / THREADS;
			const uint32_t p0 = threadIdx.x * ept;
			uint32_t local = 0;
			for (uint32_t p = p0; p < p0 + ept && p < F; ++p) local += cnt[p];
			uint32_t tile_total;
			uint32_t run = block_exclusive_scan(local, warp_totals, &tile_total);
			for (uint32_t p = p0; p < p0 + ept && p < F; ++p) {
				const uint32_t c = cnt[p];
				plan_digit(p, c, run);
				run += c;
			}
			if (threadIdx.x == 0) s_tile_n = tile_total;
			__syncthreads();
		}
		SCATTER_MARK(2)        // plan (two barriers inside)
		// (3) place the tile's tuples; flush the carried tuples of every digit that reached a boundary
		if (PACK) {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t rk_ = (dr[e / 2] >> (16 * (e & 1))) & 0xFFFFu;
				if (rk_ != 0xFFFFu) buf[base[radix_digit(hash_mul(key[e], factor), rshift, mask)] + rk_] = make_uint2(key[e], val[e]);
			}
		} else if (full) {
#pragma unroll
			for (int e = 0; e < IT; ++e) buf[base[dr[e] >> 16] + (dr[e] & 0xFFFFu)] = make_uint2(key[e], val[e]);
		} else {
#pragma unroll
			for (int e = 0; e < IT; ++e)
				if (dr[e] != 0xFFFFFFFFu) buf[base[dr[e] >> 16] + (dr[e] & 0xFFFFu)] = make_uint2(key[e], val[e]);
		}
		if (wc)
			for (uint32_t i = threadIdx.x; i < F * kCarry; i += THREADS) {
				const uint32_t d = i / kCarry, j = i % kCarry;
				if (j < oldp[d]) {
					const uint2 kv = carry[i];
					const uint32_t dst = fpos[d] + j;
					(PEER ? s_pk[d] : keys_out)[dst] = kv.x;
					(PEER ? s_pv[d] : vals_out)[dst] = kv.y;
				}
			}
		SCATTER_MARK(3)        // place + carry flush
		__syncthreads();
		SCATTER_MARK(4)        // waiting for the slowest warp's placement
		// (4) stream the digit-grouped tile: below the digit's limit to global memory, beyond it into
		// the carry buffer.  The next tile's step (1) barrier orders this loop before step (2) rewrites
		// golim and before step (3) reads carry.
		const uint32_t tile_n = s_tile_n;
		uint32_t *const ko = keys_out, *const vo = vals_out;
#pragma unroll
		for (int it = 0; it < IT; ++it) {                 // tile_n <= TILE: at most IT rounds, all in flight together
			const uint32_t i = threadIdx.x + it * THREADS;
			if (i < tile_n) {
				const uint2 kv = buf[i];
				const uint32_t d = radix_digit(hash_mul(kv.x, factor), rshift, mask);
				const uint2 gl = golim[d];
				const uint32_t pos = gl.x + i;
				if (pos < gl.y) {
					(PEER ? s_pk[d] : ko)[pos] = kv.x;
					(PEER ? s_pv[d] : vo)[pos] = kv.y;
				} else {
					carry[d * kCarry + (pos - gl.y)] = kv;
				}
			}
		}
		SCATTER_MARK(5)        // stream
	}
#undef SCATTER_MARK
}

// ------------------------------------------------------------------ local scatter, stream overlapped with the next plan
//
// Same four steps as k_scatter (1024 threads, 8 tuples per thread, fan-out <= 256), but the steps
// of successive tiles are skewed: the plan of tile t -- a latency chain of 256 threads that left the
// other 24 warps idle for ~1250 of a tile's ~10750 cycles (phase clocks, HJB_SCATTER_CLOCKS) -- now runs
// while those warps stream tile t-1 out.  The stream's work is handed out in 128-tuple chunks from a
// shared counter, so the planning warps join in when their plan is done.  golim and the tile size
// are double-buffered; the plan synchronises its 8 warps with a named barrier of its own.
//
//   iteration t:  rank(t) | A | plan(t) by warps 0-7  ||  stream(t-1) by all, planners late | B | place(t), carry flush(t) | C
//
// dynamic shared memory: cnt base fpos oldp wpos pend [F] | golim[2][F] (uint2) | buf[TILE] (uint2) | carry[F*8] (uint2)
template <bool CLK>
__global__ void __launch_bounds__(1024, 1)
k_scatter_ov(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
             const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
             uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets,
             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
	constexpr int THREADS = 1024, G = 2, IT = 4 * G;
	constexpr uint32_t kCarry = kLocalCarry, TILE = THREADS * IT, kGroupsPerTile = TILE / 4, kChunk = 128;
	extern __shared__ __align__(16) uint32_t s_mem[];
	__shared__ uint32_t warp_totals[8];
	__shared__ uint32_t s_tile_n[2], s_next[2];
	const uint32_t F = 1u << bits, mask = F - 1;
	uint32_t *cnt = s_mem, *base = cnt + F, *fpos = base + F, *oldp = fpos + F, *wpos = oldp + F, *pend = wpos + F;
	uint2 *golim = reinterpret_cast<uint2 *>(pend + F);                 // [2][F]  x: global offset of tile index 0, y: flush limit
	uint2 *buf = golim + 2 * F;
	uint2 *carry = buf + TILE;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	const uint32_t *row = offsets + (size_t)blockIdx.x * F;
	for (uint32_t p = threadIdx.x; p < F; p += THREADS) {
		wpos[p] = row[p];
		pend[p] = 0;
		cnt[p] = 0;
	}
	if (threadIdx.x < 2) s_next[threadIdx.x] = 0;
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	auto tile_is_full = [&](uint64_t g0) { return (g0 << 2) >= r.beg && ((g0 + kGroupsPerTile) << 2) <= r.end; };
	// stream one tile: chunks of 128 consecutive slots, handed out by a shared counter
	auto stream = [&](uint32_t par) {
		const uint32_t tile_n = s_tile_n[par];
		const uint2 *gl_tab = golim + par * F;
		while (true) {
			uint32_t c0 = 0;
			if (lane_id() == 0) c0 = atomicAdd(&s_next[par], kChunk);
			c0 = __shfl_sync(kFullMask, c0, 0);
			if (c0 >= tile_n) break;
#pragma unroll
			for (uint32_t j = 0; j < kChunk / 32; ++j) {
				const uint32_t i = c0 + j * 32 + lane_id();
				if (i < tile_n) {
					const uint2 kv = buf[i];
					const uint32_t d = radix_digit(hash_mul(kv.x, factor), rshift, mask);
					const uint2 gl = gl_tab[d];
					const uint32_t pos = gl.x + i;
					if (pos < gl.y) {
						keys_out[pos] = kv.x;
						vals_out[pos] = kv.y;
					} else {
						carry[d * kCarry + (pos - gl.y)] = kv;
					}
				}
			}
		}
	};
	uint32_t key[IT], nkey[IT], val[IT], ok = 0, nok = 0;
	if (g_beg < g_end) {
		if (tile_is_full(g_beg)) load_col8<THREADS, G, true>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
		else load_col8<THREADS, G, false>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
	}
	__syncthreads();
	long long clk_t = 0;
#define SCATTER_MARK(k)                                                              \
	if (CLK && threadIdx.x == 512) {                                                 \
		const long long now_ = clock64();                                            \
		atomicAdd(&g_scatter_clk[k], (unsigned long long)(now_ - clk_t));            \
		clk_t = now_;                                                                \
	}
	if (CLK && threadIdx.x == 512) clk_t = clock64();
	uint32_t t = 0;
	for (uint64_t g0 = g_beg; g0 < g_end; g0 += kGroupsPerTile, ++t) {
		const uint64_t g1 = g0 + kGroupsPerTile;
		const bool last = g1 >= g_end;
		const bool full = tile_is_full(g0);
		const uint32_t par = t & 1;
#pragma unroll
		for (int e = 0; e < IT; ++e) key[e] = nkey[e];
		ok = nok;
		uint32_t vok;
		if (full) load_col8<THREADS, G, true>(val, vok, vals, g0, g_end, r.beg, r.end, n);
		else load_col8<THREADS, G, false>(val, vok, vals, g0, g_end, r.beg, r.end, n);
		if (!last) {
			if (tile_is_full(g1)) load_col8<THREADS, G, true>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
		}
		// (1) rank
		uint32_t dr[IT];
		if (full) {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				dr[e] = (d << 16) | atomicAdd(&cnt[d], 1u);
			}
		} else {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				dr[e] = (ok >> e) & 1u ? (d << 16) | atomicAdd(&cnt[d], 1u) : 0xFFFFFFFFu;
			}
		}
		SCATTER_MARK(0)
		__syncthreads();                                          // A: ranks of tile t taken; tile t-1 fully placed (C of t-1)
		SCATTER_MARK(1)
		// (2) plan of tile t by the first 8 warps (one digit per thread) ...
		if (threadIdx.x < 256) {
			const uint32_t p = threadIdx.x;
			const uint32_t c = p < F ? cnt[p] : 0;
			const uint32_t incl = warp_inclusive_scan_u32(c);
			if (lane_id() == 31) warp_totals[p >> 5] = incl;
			asm volatile("bar.sync 1, 256;" ::: "memory");
			const uint32_t wt = lane_id() < 8 ? warp_totals[lane_id()] : 0;
			const uint32_t tincl = warp_inclusive_scan_u32(wt);
			const uint32_t before = __shfl_sync(kFullMask, tincl - wt, p >> 5);
			if (p < F) {
				const uint32_t run = before + incl - c;
				const uint32_t w = wpos[p], pe = pend[p], endpos = w + pe + c;
				uint32_t lim = last ? endpos : (endpos & ~(kCarry - 1));
				const bool flush = lim > w;
				if (!flush) lim = w;
				base[p] = run;
				golim[par * F + p] = make_uint2(w + pe - run, lim);
				fpos[p] = w;
				oldp[p] = flush ? pe : 0;
				wpos[p] = lim;
				pend[p] = endpos - lim;
				cnt[p] = 0;
			}
			const uint32_t tile_total = __shfl_sync(kFullMask, tincl, 31);
			if (p == 0) {
				s_tile_n[par] = tile_total;
				s_next[par] = 0;
			}
		}
		// ... while everybody (the planners once they are done) streams tile t-1
		if (t > 0) stream(par ^ 1);
		SCATTER_MARK(2)
		__syncthreads();                                          // B: plan of tile t visible, tile t-1 streamed (buf and carry free)
		SCATTER_MARK(3)
		// (3) place tile t, flush the carried tuples of every digit that reached a boundary
		if (full) {
#pragma unroll
			for (int e = 0; e < IT; ++e) buf[base[dr[e] >> 16] + (dr[e] & 0xFFFFu)] = make_uint2(key[e], val[e]);
		} else {
#pragma unroll
			for (int e = 0; e < IT; ++e)
				if (dr[e] != 0xFFFFFFFFu) buf[base[dr[e] >> 16] + (dr[e] & 0xFFFFu)] = make_uint2(key[e], val[e]);
		}
		for (uint32_t i = threadIdx.x; i < F * kCarry; i += THREADS) {
			const uint32_t d = i / kCarry, j = i % kCarry;
			if (j < oldp[d]) {
				const uint2 kv = carry[i];
				const uint32_t dst = fpos[d] + j;
				keys_out[dst] = kv.x;
				vals_out[dst] = kv.y;
			}
		}
		SCATTER_MARK(4)
		__syncthreads();                                          // C: tile t placed, its carried tuples flushed
		SCATTER_MARK(5)
	}
	if (t > 0) stream((t - 1) & 1);                                // the item's last tile (its plan flushed every digit completely)
#undef SCATTER_MARK
}

// ------------------------------------------------------------------ local scatter with fixed digit regions
//
// The phase clocks of k_scatter (HJB_SCATTER_CLOCKS: rank 18 %, plan 12 %, place + carry flush 31 %,
// stream 27 %, barrier waits 11 % of a tile's ~10750 cycles) say the tile is a chain of phases each
// waiting on shared memory.  Here the digit-grouped tile is not packed: digit d owns the slots
// [d REG, (d+1) REG) with REG = 2 TILE / F, twice its expected share, so
//   * the rank atomic gives the tuple's slot at once and the tuple is stored there in the same step
//     (no second look-up of a digit base, no separate placement phase, payloads prefetched like keys);
//   * the plan needs no prefix scan over the digits: every digit's thread decides alone how far its run
//     may be flushed;
//   * a warp streams whole digit regions (the digit is the loop index, not re-hashed from the key).
// A digit with more than REG tuples in a tile (skewed keys) keeps the excess in registers and writes
// it to its final position directly after the plan -- slower, never wrong.
// dynamic shared memory: cnt scnt fpos oldp wpos pend [F] | golim[F] (uint2) | buf[2 TILE] (uint2) | carry[2][F*8] (uint2)
template <bool CLK>
__global__ void __launch_bounds__(1024, 1)
k_scatter_fx(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
             const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
             uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets,
             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
	constexpr int THREADS = 1024, G = 2, IT = 4 * G;
	constexpr uint32_t kCarry = kLocalCarry, TILE = THREADS * IT, kGroupsPerTile = TILE / 4;
	extern __shared__ __align__(16) uint32_t s_mem[];
	const uint32_t F = 1u << bits, mask = F - 1;
	const int reg_shift = 14 - bits;                            // REG = 2 * 8192 / F slots per digit
	const uint32_t REG = 1u << reg_shift;
	uint32_t *cnt = s_mem, *scnt = cnt + F, *fpos = scnt + F, *oldp = fpos + F, *wpos = oldp + F, *pend = wpos + F;
	uint2 *golim = reinterpret_cast<uint2 *>(pend + F);         // x: global position of the region's slot 0, y: flush limit
	uint2 *buf = golim + F;
	uint2 *carry = buf + 2 * TILE;                              // [2][F * kCarry]
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	const uint32_t *row = offsets + (size_t)blockIdx.x * F;
	for (uint32_t p = threadIdx.x; p < F; p += THREADS) {
		wpos[p] = row[p];
		pend[p] = 0;
		cnt[p] = 0;
	}
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	auto tile_is_full = [&](uint64_t g0) { return (g0 << 2) >= r.beg && ((g0 + kGroupsPerTile) << 2) <= r.end; };
	uint32_t key[IT], val[IT], nkey[IT], nval[IT], ok = 0, nok = 0, vok;
	if (g_beg < g_end) {
		if (tile_is_full(g_beg)) {
			load_col8<THREADS, G, true>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
			load_col8<THREADS, G, true>(nval, vok, vals, g_beg, g_end, r.beg, r.end, n);
		} else {
			load_col8<THREADS, G, false>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
			load_col8<THREADS, G, false>(nval, vok, vals, g_beg, g_end, r.beg, r.end, n);
		}
	}
	__syncthreads();
	long long clk_t = 0;
#define SCATTER_MARK(k)                                                              \
	if (CLK && threadIdx.x == 512) {                                                 \
		const long long now_ = clock64();                                            \
		atomicAdd(&g_scatter_clk[k], (unsigned long long)(now_ - clk_t));            \
		clk_t = now_;                                                                \
	}
	if (CLK && threadIdx.x == 512) clk_t = clock64();
	uint32_t t = 0;
	for (uint64_t g0 = g_beg; g0 < g_end; g0 += kGroupsPerTile, ++t) {
		const uint64_t g1 = g0 + kGroupsPerTile;
		const bool last = g1 >= g_end;
		const bool full = tile_is_full(g0);
		uint2 *const oc = carry + (t & 1) * F * kCarry, *const nc = carry + ((t & 1) ^ 1) * F * kCarry;   // carried in / out
#pragma unroll
		for (int e = 0; e < IT; ++e) {
			key[e] = nkey[e];
			val[e] = nval[e];
		}
		ok = nok;
		if (!last) {
			if (tile_is_full(g1)) {
				load_col8<THREADS, G, true>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
				load_col8<THREADS, G, true>(nval, vok, vals, g1, g_end, r.beg, r.end, n);
			} else {
				load_col8<THREADS, G, false>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
				load_col8<THREADS, G, false>(nval, vok, vals, g1, g_end, r.beg, r.end, n);
			}
		}
		// (1) rank and place in one step; a tuple beyond its digit's region stays in registers (excess != 0)
		uint32_t dr[IT], excess = 0;
#pragma unroll
		for (int e = 0; e < IT; ++e) {
			if (full || ((ok >> e) & 1u)) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				const uint32_t rk_ = atomicAdd(&cnt[d], 1u);
				dr[e] = (d << 16) | rk_;
				if (rk_ < REG) buf[(d << reg_shift) + rk_] = make_uint2(key[e], val[e]);
				else excess |= 1u << e;
			}
		}
		SCATTER_MARK(0)
		__syncthreads();                                          // A
		SCATTER_MARK(1)
		// (2) plan: every digit on its own
		for (uint32_t p = threadIdx.x; p < F; p += THREADS) {
			const uint32_t c = cnt[p], w = wpos[p], pe = pend[p], endpos = w + pe + c;
			uint32_t lim = last ? endpos : (endpos & ~(kCarry - 1));
			const bool flush = lim > w;
			if (!flush) lim = w;
			golim[p] = make_uint2(w + pe, lim);
			scnt[p] = min(c, REG);
			fpos[p] = w;
			oldp[p] = flush ? pe : 0x80000000u | pe;              // top bit: the carried tuples stay carried
			wpos[p] = lim;
			pend[p] = endpos - lim;
			cnt[p] = 0;
		}
		SCATTER_MARK(2)
		__syncthreads();                                          // B
		SCATTER_MARK(3)
		// (3a) the excess of over-full digits, straight from registers
		if (excess) {
#pragma unroll
			for (int e = 0; e < IT; ++e)
				if ((excess >> e) & 1u) {
					const uint32_t d = dr[e] >> 16;
					const uint2 gl = golim[d];
					const uint32_t pos = gl.x + (dr[e] & 0xFFFFu);
					if (pos < gl.y) {
						keys_out[pos] = key[e];
						vals_out[pos] = val[e];
					} else {
						nc[d * kCarry + (pos - gl.y)] = make_uint2(key[e], val[e]);
					}
				}
		}
		// (3b) the tuples carried in: to their place if the digit flushes, else on to the next tile
		for (uint32_t i = threadIdx.x; i < F * kCarry; i += THREADS) {
			const uint32_t d = i / kCarry, j = i % kCarry, op = oldp[d];
			if (j < (op & 0x7FFFFFFFu)) {
				const uint2 kv = oc[i];
				if (op & 0x80000000u) {
					nc[i] = kv;
				} else {
					const uint32_t dst = fpos[d] + j;
					keys_out[dst] = kv.x;
					vals_out[dst] = kv.y;
				}
			}
		}
		// (3c) stream the regions: warp w takes digits w, w + 32, ...; a region's slots are consecutive positions
		for (uint32_t d = threadIdx.x >> 5; d < F; d += THREADS / 32) {
			const uint32_t c = scnt[d];
			const uint2 gl = golim[d];
			const uint2 *reg = buf + (d << reg_shift);
			for (uint32_t j = lane_id(); j < c; j += 32) {
				const uint2 kv = reg[j];
				const uint32_t pos = gl.x + j;
				if (pos < gl.y) {
					keys_out[pos] = kv.x;
					vals_out[pos] = kv.y;
				} else {
					nc[d * kCarry + (pos - gl.y)] = kv;
				}
			}
		}
		SCATTER_MARK(4)
		__syncthreads();                                          // C: regions and the old carry buffer are free again
		SCATTER_MARK(5)
	}
#undef SCATTER_MARK
}

// ------------------------------------------------------------------ local scatter, two sub-tiles per plan
//
// Experiment 12: the plan and the three CTA barriers cost ~2500 of a tile's ~10750 cycles and do not depend
// on the tile's size, so rank TWO 8192-tuple sub-tiles before one plan and place / stream 16384 tuples per
// round.  Only the ranks stay in registers between the rank and the placement; the keys are fetched again
// for the placement (an L2 hit: they were read a few microseconds earlier) and the payloads then for the
// first time.  128 KB tile, fan-out <= 256.
// dynamic shared memory: cnt base fpos oldp wpos pend [F] | golim[F] (uint2) | buf[2 * 8192] (uint2) | carry[F*8] (uint2)
__global__ void __launch_bounds__(1024, 1)
k_scatter_2x(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
             const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
             uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets,
             uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
	constexpr int THREADS = 1024, G = 2, IT = 4 * G;
	constexpr uint32_t kCarry = kLocalCarry, SUB = THREADS * IT, TILE = 2 * SUB, kGroupsPerSub = SUB / 4;
	extern __shared__ __align__(16) uint32_t s_mem[];
	__shared__ uint32_t warp_totals[8];
	__shared__ uint32_t s_tile_n;
	const uint32_t F = 1u << bits, mask = F - 1;
	uint32_t *cnt = s_mem, *base = cnt + F, *fpos = base + F, *oldp = fpos + F, *wpos = oldp + F, *pend = wpos + F;
	uint2 *golim = reinterpret_cast<uint2 *>(pend + F);
	uint2 *buf = golim + F;
	uint2 *carry = buf + TILE;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	const uint32_t *row = offsets + (size_t)blockIdx.x * F;
	for (uint32_t p = threadIdx.x; p < F; p += THREADS) {
		wpos[p] = row[p];
		pend[p] = 0;
		cnt[p] = 0;
	}
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	auto sub_is_full = [&](uint64_t g0) { return (g0 << 2) >= r.beg && ((g0 + kGroupsPerSub) << 2) <= r.end; };
	auto load = [&](uint32_t (&x)[IT], uint32_t &ok, const uint32_t *col, uint64_t g0) {
		if (sub_is_full(g0)) load_col8<THREADS, G, true>(x, ok, col, g0, g_end, r.beg, r.end, n);
		else load_col8<THREADS, G, false>(x, ok, col, g0, g_end, r.beg, r.end, n);
	};
	auto rank = [&](const uint32_t (&k)[IT], uint32_t ok, uint32_t (&dr)[IT]) {
#pragma unroll
		for (int e = 0; e < IT; ++e) {
			const uint32_t d = radix_digit(hash_mul(k[e], factor), rshift, mask);
			dr[e] = (ok >> e) & 1u ? (d << 16) | atomicAdd(&cnt[d], 1u) : 0xFFFFFFFFu;
		}
	};
	auto place = [&](const uint32_t (&k)[IT], const uint32_t (&v)[IT], const uint32_t (&dr)[IT]) {
#pragma unroll
		for (int e = 0; e < IT; ++e)
			if (dr[e] != 0xFFFFFFFFu) buf[base[dr[e] >> 16] + (dr[e] & 0xFFFFu)] = make_uint2(k[e], v[e]);
	};
	uint32_t nkey[IT], nok = 0;
	if (g_beg < g_end) load(nkey, nok, keys, g_beg);
	__syncthreads();
	for (uint64_t g0 = g_beg; g0 < g_end; g0 += 2 * kGroupsPerSub) {
		const uint64_t gB = g0 + kGroupsPerSub, g1 = g0 + 2 * kGroupsPerSub;
		const bool last = g1 >= g_end;
		uint32_t drA[IT], drB[IT], valA[IT], tmp[IT], okB, vok;
		load(tmp, okB, keys, gB);                                  // sub-tile B's keys on their way while A is ranked
		rank(nkey, nok, drA);
		load(valA, vok, vals, g0);                                 // A's payloads: in flight across the plan
		rank(tmp, okB, drB);
		if (!last) load(nkey, nok, keys, g1);                      // next round's sub-tile A
		__syncthreads();
		// plan: one digit per thread in the first eight warps (fan-out <= 256)
		if (threadIdx.x < 256) {
			const uint32_t p = threadIdx.x;
			const uint32_t c = p < F ? cnt[p] : 0;
			const uint32_t incl = warp_inclusive_scan_u32(c);
			if (lane_id() == 31) warp_totals[p >> 5] = incl;
			asm volatile("bar.sync 1, 256;" ::: "memory");
			const uint32_t wt = lane_id() < 8 ? warp_totals[lane_id()] : 0;
			const uint32_t tincl = warp_inclusive_scan_u32(wt);
			const uint32_t before = __shfl_sync(kFullMask, tincl - wt, p >> 5);
			const uint32_t tile_total = __shfl_sync(kFullMask, tincl, 31);
			if (p < F) {
				const uint32_t run = before + incl - c;
				const uint32_t w = wpos[p], pe = pend[p], endpos = w + pe + c;
				uint32_t lim = last ? endpos : (endpos & ~(kCarry - 1));
				const bool flush = lim > w;
				if (!flush) lim = w;
				base[p] = run;
				golim[p] = make_uint2(w + pe - run, lim);
				fpos[p] = w;
				oldp[p] = flush ? pe : 0;
				wpos[p] = lim;
				pend[p] = endpos - lim;
				cnt[p] = 0;
			}
			if (p == 0) s_tile_n = tile_total;
		}
		__syncthreads();
		// place A (keys fetched again), then B (keys again, payloads for the first time)
		{
			uint32_t okA;
			load(tmp, okA, keys, g0);
			place(tmp, valA, drA);
		}
		{
			uint32_t okk;
			load(tmp, okk, keys, gB);
			load(valA, vok, vals, gB);
			place(tmp, valA, drB);
		}
		for (uint32_t i = threadIdx.x; i < F * kCarry; i += THREADS) {
			const uint32_t d = i / kCarry, j = i % kCarry;
			if (j < oldp[d]) {
				const uint2 kv = carry[i];
				const uint32_t dst = fpos[d] + j;
				keys_out[dst] = kv.x;
				vals_out[dst] = kv.y;
			}
		}
		__syncthreads();
		const uint32_t tile_n = s_tile_n;
#pragma unroll 8
		for (int it = 0; it < 2 * IT; ++it) {
			const uint32_t i = threadIdx.x + it * THREADS;
			if (i < tile_n) {
				const uint2 kv = buf[i];
				const uint32_t d = radix_digit(hash_mul(kv.x, factor), rshift, mask);
				const uint2 gl = golim[d];
				const uint32_t pos = gl.x + i;
				if (pos < gl.y) {
					keys_out[pos] = kv.x;
					vals_out[pos] = kv.y;
				} else {
					carry[d * kCarry + (pos - gl.y)] = kv;
				}
			}
		}
		// the next round's rank touches only cnt; its plan (after a barrier) rewrites golim and its placement buf
	}
}

// ------------------------------------------------------------------ peer scatter with TMA bulk stores
//
// The GPU-assign pass of CPRA has a small fan-out (one digit per GPU), so a digit's run in a tile
// is thousands of tuples long.  Issuing it as per-lane 4-byte remote stores ties up the SM's
// store path (measured: throughput proportional to the number of CTAs, ~3.5 GB/s per SM).  Here
// the tile is digit-grouped into two SoA shared-memory buffers laid out with the SAME 128-byte
// alignment as the destination rows in the owner's buffer, and ONE thread per digit hands each
// run to the TMA unit as a bulk shared->global copy (cp.async.bulk, UBLKCP in SASS) that then
// crosses NVLink without occupying the SM.  Tuples beyond a digit's last whole 128-byte line are
// carried to the next tile (software write-combining as above); an item's first / last few tuples
// per digit that are not 16-byte aligned take scalar stores.  Fan-out <= 64.
// dynamic shared memory: cnt wpos pend [64] | place[64] (uint4) | strm[64] (uint4) | cin[64] (uint2) |
//                        skeys[PAD] svals[PAD] | carry_k[2][32 F] carry_v[2][32 F],  PAD = TILE + 64 F
constexpr uint32_t kBulkGranule = kPeerCarry;             // 32 tuples = 128 bytes per column

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
	             "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
	             : "memory");
}

template <int THREADS, int GRAN, int MAXF, bool PEER, bool TMA = true>
__global__ void __launch_bounds__(THREADS, 1)
k_scatter_bulk(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
               const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
               uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets, uint32_t *__restrict__ keys_out,
               uint32_t *__restrict__ vals_out, const PeerTable peers)
{
	constexpr int G = 2, IT = 4 * G;
	constexpr uint32_t TILE = THREADS * IT, kGroupsPerTile = TILE / 4, GR = GRAN;
	extern __shared__ __align__(128) uint32_t s_bulk[];
	__shared__ uint32_t warp_totals[MAXF / 32];
	__shared__ uint32_t s_slots;                  // slots of the digit-grouped tile in use (TMA = false)
	__shared__ uint32_t *s_pk[PEER ? 64 : 1], *s_pv[PEER ? 64 : 1];
	const uint32_t F = 1u << bits, mask = F - 1;
	const uint32_t PAD = TILE + 2 * GR * F;
	// !TMA: slots of a region that hold no tuple (before an item's first position, behind its last) get a key
	// that hashes to the region's digit, so that the stream phase can read the digit from any slot
	uint32_t finv = factor;                       // inverse of the odd factor modulo 2^32 (Newton)
#pragma unroll
	for (int t = 0; t < 5; ++t) finv *= 2u - factor * finv;
	uint32_t *cnt = s_bulk, *wpos = cnt + MAXF, *pend = wpos + MAXF;
	uint4 *place = reinterpret_cast<uint4 *>(pend + MAXF);     // x: slot of new rank 0, y: new tuples that fit the region, z: carry index of rank 0
	uint4 *strm = place + MAXF;                                // x: global position of slot 0, y: first valid position, z: end of valid positions
	uint2 *cin = reinterpret_cast<uint2 *>(strm + MAXF);       // carried-in tuples: x: destination slot 0 (0xFFFFFFFF: stay carried), y: how many
	uint32_t *skeys = reinterpret_cast<uint32_t *>(cin + MAXF);        // byte offset 52 * MAXF, a multiple of 128
	uint32_t *svals = skeys + PAD;
	uint32_t *carry_k = svals + PAD, *carry_v = carry_k + 2 * GR * F;
	if (PEER && threadIdx.x < 64) {
		s_pk[threadIdx.x] = peers.k[threadIdx.x];
		s_pv[threadIdx.x] = peers.v[threadIdx.x];
	}
	for (uint32_t item = blockIdx.x;; item += gridDim.x) {
		ItemRange r;
		if (!locate_item(item_prefix, np, parent_off, n, chunk, item, &r)) break;
		__syncthreads();
		const uint32_t *row = offsets + (size_t)item * F;
		if (threadIdx.x < F) {
			wpos[threadIdx.x] = row[threadIdx.x] + (PEER ? peers.bias[threadIdx.x & 63] : 0u);
			pend[threadIdx.x] = 0;
			cnt[threadIdx.x] = 0;
		}
		const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
		auto tile_is_full = [&](uint64_t g0) { return (g0 << 2) >= r.beg && ((g0 + kGroupsPerTile) << 2) <= r.end; };
		uint32_t key[IT], nkey[IT], val[IT], ok = 0, nok = 0;
		if (g_beg < g_end) {
			if (tile_is_full(g_beg)) load_col8<THREADS, G, true>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
		}
		__syncthreads();
		uint32_t tile_no = 0;
		for (uint64_t g0 = g_beg; g0 < g_end; g0 += kGroupsPerTile, ++tile_no) {
			const uint64_t g1 = g0 + kGroupsPerTile;
			const bool last = g1 >= g_end;
			const bool full = tile_is_full(g0);
			uint32_t *const oc_k = carry_k + (tile_no & 1) * GR * F, *const oc_v = carry_v + (tile_no & 1) * GR * F;            // carried in
			uint32_t *const nc_k = carry_k + ((tile_no & 1) ^ 1) * GR * F, *const nc_v = carry_v + ((tile_no & 1) ^ 1) * GR * F;  // carried out
#pragma unroll
			for (int e = 0; e < IT; ++e) key[e] = nkey[e];
			ok = nok;
			uint32_t vok;
			if (full) load_col8<THREADS, G, true>(val, vok, vals, g0, g_end, r.beg, r.end, n);
			else load_col8<THREADS, G, false>(val, vok, vals, g0, g_end, r.beg, r.end, n);
			if (!last) {
				if (tile_is_full(g1)) load_col8<THREADS, G, true>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
				else load_col8<THREADS, G, false>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
			}
			// (1) rank
			uint32_t dr[IT];
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				if (bits <= 4) dr[e] = rank_aggregated(cnt, d, full || ((ok >> e) & 1u));
				else dr[e] = (full || ((ok >> e) & 1u)) ? (d << 16) | atomicAdd(&cnt[d], 1u) : 0xFFFFFFFFu;
			}
			// the bulk copies of the previous tile must have read their shared-memory source before it is reused
			if (TMA && threadIdx.x < F) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
			__syncthreads();
			// (2) plan: one digit per thread in the first MAXF / 32 warps
			if (threadIdx.x < MAXF) {
				const uint32_t p = threadIdx.x;
				uint32_t c = 0, w = 0, pe = 0, wg = 0, lim = 0, slots = 0;
				bool flush = false;
				if (p < F) {
					c = cnt[p];
					w = wpos[p];
					pe = pend[p];
					const uint32_t endpos = w + pe + c;
					lim = last ? endpos : (endpos & ~(GR - 1));
					flush = lim > w;
					if (!flush) lim = w;
					wg = w & ~(GR - 1);
					slots = flush ? ((lim + GR - 1) & ~(GR - 1)) - wg : 0;     // region covers [wg, align_up(lim)), whole lines
				}
				const uint32_t incl = warp_inclusive_scan_u32(slots);
				if (lane_id() == 31) warp_totals[p >> 5] = incl;
				__syncwarp();
				asm volatile("bar.sync 1, %0;" ::"n"(MAXF) : "memory");        // the planning warps only
				uint32_t rb = incl - slots;                                    // region start, multiple of GR
#pragma unroll
				for (int wq = 0; wq < MAXF / 32; ++wq) rb += (wq < (int)(p >> 5)) ? warp_totals[wq] : 0u;
				if (p < F) {
					const uint32_t room = flush ? lim - (w + pe) : 0;          // new tuples that go to the region
					place[p] = make_uint4(rb + (w - wg) + pe, room, p * GR + (flush ? 0u - room : pe), 0);
					strm[p] = make_uint4(wg - rb, w, lim, flush ? 1u : 0u);
					cin[p] = make_uint2(flush ? rb + (w - wg) : 0xFFFFFFFFu, pe);
					wpos[p] = lim;
					pend[p] = w + pe + c - lim;
					cnt[p] = 0;
					if (!TMA) {
						if (p == F - 1) s_slots = rb + slots;
						if (flush && (((w | lim) & (GR - 1)) != 0)) {
							const uint32_t marker = (p << rshift) * finv;
							for (uint32_t q = rb; q < rb + (w - wg); ++q) skeys[q] = marker;
							for (uint32_t q = rb + (lim - wg); q < rb + slots; ++q) skeys[q] = marker;
						}
					}
				}
			}
			__syncthreads();
			// (3) place the new tuples and move the carried-in ones
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				if (dr[e] != 0xFFFFFFFFu) {
					const uint32_t d = dr[e] >> 16, rk_ = dr[e] & 0xFFFFu;
					const uint4 pl = place[d];
					if (rk_ < pl.y) {
						skeys[pl.x + rk_] = key[e];
						svals[pl.x + rk_] = val[e];
					} else {
						nc_k[pl.z + rk_] = key[e];
						nc_v[pl.z + rk_] = val[e];
					}
				}
			}
			for (uint32_t t = threadIdx.x; t < F * GR; t += THREADS) {
				const uint32_t d = t / GR, q = t % GR;
				const uint2 ci = cin[d];
				if (q < ci.y) {
					const uint32_t ck = oc_k[t], cv = oc_v[t];
					if (ci.x != 0xFFFFFFFFu) {
						skeys[ci.x + q] = ck;
						svals[ci.x + q] = cv;
					} else {
						nc_k[t] = ck;
						nc_v[t] = cv;
					}
				}
			}
			if (TMA) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk copy
			__syncthreads();
			if (!TMA) {
				// (4) every thread: four slots at a time, 16-byte stores per column where the four are all valid
				const uint32_t groups = s_slots >> 2;
				for (uint32_t q = threadIdx.x; q < groups; q += THREADS) {
					const uint4 k4 = *reinterpret_cast<const uint4 *>(skeys + 4 * q);
					const uint4 v4 = *reinterpret_cast<const uint4 *>(svals + 4 * q);
					const uint32_t d = radix_digit(hash_mul(k4.x, factor), rshift, mask);
					const uint4 st = strm[d];
					const uint32_t pos = st.x + 4 * q;                     // global position of the group's first slot
					if (pos >= st.y && pos + 4 <= st.z) {
						*reinterpret_cast<uint4 *>(keys_out + pos) = k4;
						*reinterpret_cast<uint4 *>(vals_out + pos) = v4;
					} else {
						const uint32_t kk[4] = {k4.x, k4.y, k4.z, k4.w}, vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
						for (int e = 0; e < 4; ++e)
							if (pos + e >= st.y && pos + e < st.z) {
								keys_out[pos + e] = kk[e];
								vals_out[pos + e] = vv[e];
							}
					}
				}
			}
			// (4) one thread per digit: whole run as bulk copies, unaligned ends as scalar stores
			if (TMA && threadIdx.x < F) {
				const uint32_t d = threadIdx.x;
				const uint4 st = strm[d];
				if (st.w) {
					uint32_t *const ko = PEER ? s_pk[d & 63] : keys_out, *const vo = PEER ? s_pv[d & 63] : vals_out;
					const uint32_t lo = st.y, hi = st.z;                       // valid global positions [lo, hi)
					uint32_t blo = (lo + 3) & ~3u, bhi = hi & ~3u;             // 16-byte aligned body
					if (blo > bhi) blo = bhi = hi;                              // fewer than four tuples: all scalar
					for (uint32_t pos = lo; pos < min(blo, hi); ++pos) {
						ko[pos] = skeys[pos - st.x];
						vo[pos] = svals[pos - st.x];
					}
					if (bhi > blo) {
						bulk_store(ko + blo, skeys + (blo - st.x), (bhi - blo) * 4);
						bulk_store(vo + blo, svals + (blo - st.x), (bhi - blo) * 4);
					}
					for (uint32_t pos = max(bhi, blo); pos < hi; ++pos) {
						ko[pos] = skeys[pos - st.x];
						vo[pos] = svals[pos - st.x];
					}
				}
				asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			}
		}
		// the item's last bulk copies must complete before its shared memory is reused / the CTA exits
		if (TMA && threadIdx.x < F) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}

// ------------------------------------------------------------------ host launchers

size_t radix_scratch_bytes(uint64_t n, uint32_t np, int bits, uint32_t *chunk, uint32_t *max_items, uint32_t *tiles)
{
	// ~1.2K items (eight per SM of a B200): enough to balance the SMs over a pass, few enough that the
	// counts matrix and its scan stay small (measured: 1024-1184 items 4.23 ms per config-2 step, 2048
	// 4.27, 4096 4.35); chunk is a multiple of the scatter tile
	static int target = -1;                       // experiment knob: HJB_ITEMS=<work items per pass>
	if (target < 0) target = getenv("HJB_ITEMS") ? atoi(getenv("HJB_ITEMS")) : 1184;
	if (target < 64) target = 1184;
	uint64_t c = (n + target - 1) / target;
	c = (c + 8191) / 8192 * 8192;
	if (c < 16384) c = 16384;
	if (c > (1u << 24)) c = 1u << 24;
	*chunk = (uint32_t)c;
	const uint64_t mi = n / c + np + 1;
	*max_items = (uint32_t)mi;
	const uint64_t E = mi << bits;
	*tiles = (uint32_t)((E + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems));
	size_t bytes = 0;
	bytes += ((size_t)(np + 1) * 4 + 255) / 256 * 256;           // item_prefix
	bytes += ((size_t)E * 4 + 255) / 256 * 256;                   // counts / offsets
	bytes += ((size_t)*tiles * 8 + 255) / 256 * 256 + 256;        // scan status + counter
	return bytes + 1024;                                            // per-array 256-byte padding of the bump allocator
}

static void scatter_attrs()
{
	// the attribute is per device: a process that drives several GPUs (host/cpra.cpp) sets it once on each
	static std::atomic<unsigned long long> done_mask{0};
	int dev = 0;
	cudaGetDevice(&dev);
	const unsigned long long bit = 1ull << (dev & 63);
	if (done_mask.fetch_or(bit) & bit) return;
	const int big = 2048 * 32 + 1024 * 8 * 8;
	cudaFuncSetAttribute(k_scatter<512, 2, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<512, 3, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<512, 3, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<1024, 1, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<1024, 2, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<1024, 1, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<1024, 2, false, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<1024, 2, true, false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<512, 2, false, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter<1024, 1, true, false, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter_ov<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter_ov<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big);
	cudaFuncSetAttribute(k_scatter_fx<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 + 2 * 8192 * 8 + 2 * 256 * 8 * 8);
	cudaFuncSetAttribute(k_scatter_fx<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 + 2 * 8192 * 8 + 2 * 256 * 8 * 8);
	cudaFuncSetAttribute(k_scatter_2x, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 32 + 2 * 8192 * 8 + 256 * 8 * 8);
	cudaFuncSetAttribute(k_scatter_bulk<1024, kBulkGranule, 64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                     (int)(52 * 64 + (8192 + 2 * kBulkGranule * 64) * 8 + 64 * kBulkGranule * 16));
	cudaFuncSetAttribute(k_scatter_bulk<1024, 8, 256, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                     (int)(52 * 256 + (8192 + 2 * 8 * 256) * 8 + 256 * 8 * 16));
	cudaFuncSetAttribute(k_scatter_bulk<1024, 8, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
	                     (int)(52 * 256 + (8192 + 2 * 8 * 256) * 8 + 256 * 8 * 16));
}

// make_items + histogram + scan: after this a.counts holds every item's start offset per digit
// and a.child_off the partition offsets
int launch_radix_count(const RadixPassArgs &a, cudaStream_t s, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	const uint32_t F = 1u << a.bits;
	const uint32_t tiles = (uint32_t)((((uint64_t)a.max_items << a.bits) + kScanThreads * kScanItems - 1) /
	                                  (kScanThreads * kScanItems));
	cudaMemsetAsync(a.scan_status, 0, (size_t)tiles * 8, s);
	cudaMemsetAsync(a.scan_counter, 0, 4, s);
	t->start(KK_MAKE_ITEMS, s);
	k_make_items<<<1, 1024, 0, s>>>(a.parent_off, a.np, a.n, a.chunk, a.item_prefix, a.child_off, a.np << a.bits);
	t->stop(s);
	t->start(KK_HIST, s);
	if (a.bits <= 3)
		k_hist_small<<<a.max_items, kHistThreads, 0, s>>>(a.keys, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,
		                                                  a.factor, a.rshift, a.bits, a.counts);
	else
		k_hist<<<a.max_items, kHistThreads, F * 4, s>>>(a.keys, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,
		                                                a.factor, a.rshift, a.bits, a.counts);
	t->stop(s);
	t->start(KK_SCAN, s);
	k_scan<<<tiles, kScanThreads, 0, s>>>(a.item_prefix, a.np, a.bits, a.counts, a.child_off, a.scan_status,
	                                      a.scan_counter);
	t->stop(s);
	return 3;
}

int launch_radix_scatter(const RadixPassArgs &a, cudaStream_t s, KernelTimer *t, const PeerTable *peers)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	const uint32_t F = 1u << a.bits;
	scatter_attrs();
	// CTA shape and register budget; HJB_SCATTER_VARIANT picks alternatives for experiments
	static int variant = -1;
	if (variant < 0) {
		const char *e = getenv("HJB_SCATTER_VARIANT");
		variant = e ? atoi(e) : 3;       // measured best on B200: one 1024-thread CTA per SM, 8192-tuple tiles
		if (variant < 0 || variant > 12) variant = 3;
	}
	const int threads = ((variant >= 3 && variant != 9) || peers) ? 1024 : 512;
	const int items = (!peers && (variant == 5 || variant == 6)) ? 4 : (!peers && variant == 9) ? 16 : 8;
	const size_t smem = (size_t)F * 32 + (size_t)threads * items * 8 + (F <= 256 ? (size_t)F * (peers ? kPeerCarry : kLocalCarry) * 8 : 0);
	static const PeerTable no_peers = {};
	t->start(KK_SCATTER, s);
	const uint32_t grid = a.max_items;
#define HJB_LAUNCH_SCATTER(T, M, P, PEER, TABLE, ...)                                                                          \
	k_scatter<T, M, P, PEER, ##__VA_ARGS__><<<grid, T, smem, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,  \
	                                                      a.factor, a.rshift, a.bits, a.counts, a.keys_out, a.vals_out, TABLE)
	static int peer_bulk = -1;
	if (peer_bulk < 0) peer_bulk = getenv("HJB_PEER_BULK") ? atoi(getenv("HJB_PEER_BULK")) : 1;
	if (peers && peer_bulk && F <= 64) {
		const size_t pad = 8192 + 2 * (size_t)kBulkGranule * F;
		const size_t smem_b = 52 * 64 + pad * 8 + (size_t)F * kBulkGranule * 16;
		// the bulk kernel walks the items with a grid stride: HJB_PEER_CTAS (experiments) can leave SMs to other streams
		const uint32_t grid_b = (a.peer_ctas && a.peer_ctas < grid) ? a.peer_ctas : grid;
		k_scatter_bulk<1024, kBulkGranule, 64, true><<<grid_b, 1024, smem_b, s>>>(
		    a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor, a.rshift, a.bits, a.counts, nullptr, nullptr, *peers);
	} else if (peers) {
		HJB_LAUNCH_SCATTER(1024, 1, true, true, *peers);
	} else if (variant == 8 && F <= 256) {
		// aligned SoA tile, 16-byte stores per column
		const size_t smem_b = 52 * 256 + (8192 + 2 * 8 * (size_t)F) * 8 + (size_t)F * 8 * 16;
		k_scatter_bulk<1024, 8, 256, false, false><<<grid, 1024, smem_b, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix,
		                                                                    a.chunk, a.factor, a.rshift, a.bits, a.counts, a.keys_out,
		                                                                    a.vals_out, no_peers);
	} else if (variant == 7 && F <= 256) {
		// experiment: the local scatter with 32-byte granules and one bulk copy per digit, tile and column
		const size_t smem_b = 52 * 256 + (8192 + 2 * 8 * (size_t)F) * 8 + (size_t)F * 8 * 16;
		k_scatter_bulk<1024, 8, 256, false><<<grid, 1024, smem_b, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,
		                                                             a.factor, a.rshift, a.bits, a.counts, a.keys_out, a.vals_out, no_peers);
	} else {
		switch (variant) {
		case 0: HJB_LAUNCH_SCATTER(512, 2, true, false, no_peers); break;
		case 1: HJB_LAUNCH_SCATTER(512, 3, true, false, no_peers); break;
		case 2: HJB_LAUNCH_SCATTER(512, 3, false, false, no_peers); break;
		case 4: HJB_LAUNCH_SCATTER(1024, 2, false, false, no_peers); break;
		case 5: HJB_LAUNCH_SCATTER(1024, 2, false, false, no_peers, 1); break;
		case 6: HJB_LAUNCH_SCATTER(1024, 2, true, false, no_peers, 1); break;
		case 9: HJB_LAUNCH_SCATTER(512, 2, false, false, no_peers, 4); break;     // two co-resident CTAs, 16 tuples per thread
		case 10: {
			// experiment: the stream of tile t-1 overlapped with the plan of tile t (k_scatter_ov); measured 2.81 vs 2.59 ms
			static int clk = -1;
			if (clk < 0) clk = getenv("HJB_SCATTER_CLOCKS") ? 1 : 0;
			if (F > 256) {
				HJB_LAUNCH_SCATTER(1024, 1, true, false, no_peers);
				break;
			}
			const size_t smem_ov = smem + (size_t)F * 8;               // second golim buffer
			if (clk) k_scatter_ov<true><<<grid, 1024, smem_ov, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor,
			                                                      a.rshift, a.bits, a.counts, a.keys_out, a.vals_out);
			else k_scatter_ov<false><<<grid, 1024, smem_ov, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor,
			                                                     a.rshift, a.bits, a.counts, a.keys_out, a.vals_out);
			break;
		}
		case 12: {
			// two sub-tiles per plan (k_scatter_2x)
			if (F > 256) {
				HJB_LAUNCH_SCATTER(1024, 1, true, false, no_peers);
				break;
			}
			const size_t smem_2x = (size_t)F * 32 + 2 * 8192 * 8 + (size_t)F * kLocalCarry * 8;
			k_scatter_2x<<<grid, 1024, smem_2x, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor, a.rshift,
			                                         a.bits, a.counts, a.keys_out, a.vals_out);
			break;
		}
		case 11: {
			// fixed digit regions (k_scatter_fx): fan-outs 32..256
			static int clk = -1;
			if (clk < 0) clk = getenv("HJB_SCATTER_CLOCKS") ? 1 : 0;
			if (F > 256 || F < 32) {
				HJB_LAUNCH_SCATTER(1024, 1, true, false, no_peers);
				break;
			}
			const size_t smem_fx = (size_t)F * 32 + 2 * 8192 * 8 + 2 * (size_t)F * kLocalCarry * 8;
			if (clk) k_scatter_fx<true><<<grid, 1024, smem_fx, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor,
			                                                      a.rshift, a.bits, a.counts, a.keys_out, a.vals_out);
			else k_scatter_fx<false><<<grid, 1024, smem_fx, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor,
			                                                     a.rshift, a.bits, a.counts, a.keys_out, a.vals_out);
			break;
		}
		default: {
			static int clk = -1;
			if (clk < 0) clk = getenv("HJB_SCATTER_CLOCKS") ? 1 : 0;
			if (clk) HJB_LAUNCH_SCATTER(1024, 1, true, false, no_peers, 2, true);
			else HJB_LAUNCH_SCATTER(1024, 1, true, false, no_peers);
			break;
		}
		}
	}
#undef HJB_LAUNCH_SCATTER
	t->stop(s);
	return 1;
}

// debug: read (and clear) the scatter phase clocks
void scatter_phase_clocks(unsigned long long *out8)
{
	cudaMemcpyFromSymbol(out8, g_scatter_clk, 64);
	unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
	cudaMemcpyToSymbol(g_scatter_clk, z, 64);
}

int launch_radix_pass(const RadixPassArgs &a, cudaStream_t s, int /*sms*/, KernelTimer *t)
{
	return launch_radix_count(a, s, t) + launch_radix_scatter(a, s, t, nullptr);
}

int launch_histogram_only(const uint32_t *keys, uint64_t n, uint32_t *counts_dev, uint32_t factor, int rshift,
                          int bits, cudaStream_t s, int sms)
{
	const uint32_t F = 1u << bits;
	cudaMemsetAsync(counts_dev, 0, (size_t)F * 4, s);
	uint64_t groups = (n + 3) / 4;
	uint32_t grid = (uint32_t)((groups + kHistThreads - 1) / kHistThreads);
	if (grid > (uint32_t)sms * 4) grid = (uint32_t)sms * 4;
	if (grid == 0) grid = 1;
	k_hist_global<<<grid, kHistThreads, F * 4, s>>>(keys, n, factor, rshift, bits, counts_dev);
	return 1;
}

}  // namespace hjb
