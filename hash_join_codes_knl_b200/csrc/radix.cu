// radix.cu -- one radix-partitioning pass = four kernels on one stream:
//   k_make_items  : cut every parent partition into work items of <= chunk tuples
//   k_hist_tiles  : per-item digit histogram in shared memory, and the digit   (reference: histogram, cpra2.cpp:801-880)
//                   counts of every 16384-tuple TILE of the item as a by-product
//   k_scan        : single-pass decoupled-look-back prefix sum over the         (reference: interleave, phj.cpp:1263-1291)
//                   (parent, digit, item) ordered counts
//   k_scatter_tc  : tile-wise shared-memory reorder + run-wise coalesced stores (reference: partition / partition_shared +
//                   with sector-granular software write-combining                flush, cpra2.cpp:882-1075, phj.cpp:877-1028)
// The work item plays the role of the reference's thread: it owns a contiguous chunk of the
// input, its counts row is the thread's counts[] and the scan hands it one start offset per
// digit, so items never synchronise while scattering.
//
// Because the histogram kernel has already seen every tile, the scatter knows each tile's digit
// counts BEFORE it loads the tile: a tuple's slot in the digit-grouped shared-memory tile is one
// atomicAdd on a cursor that starts at the digit's tile offset (no separate rank -> plan -> place
// chain), and the plan of tile t+1 is computed by eight warps while the others already stream tile t.
//
// Also here:
//   k_hist / k_scatter : the same pass for fan-outs of 1024..2048 (no tile counts, no write-combining)
// A pass may run over a RANGE of the parents whose members are unions of several row ranges (RadixPassArgs::seg / out_base):
// the local pass of the staged CPRA exchange (stage.cu), where a received sub-partition is one range per sender.
//   k_hist_small   : histogram for <= 8 digits (CPRA's GPU-assign pass), register counters
//   k_scatter_bulk : the scatter whose output columns live in other GPUs' memory; digit runs leave the SM as
//                    TMA bulk copies (cp.async.bulk) -- the fused exchange of CPRA (product path for N > 1)
//   k_hist_global  : whole-column histogram behind the public hjb_histogram
#include "hj_device.cuh"
#include "hj_internal.h"
#include <atomic>
#include <mutex>
#include <stdlib.h>

namespace hjb {

// ------------------------------------------------------------------ work items

__global__ void __launch_bounds__(1024)
k_make_items(const uint32_t *__restrict__ parent_off, uint32_t np, uint64_t n, uint32_t chunk,
             uint32_t *__restrict__ item_prefix, uint32_t *__restrict__ child_off, uint32_t child_total_idx,
             const uint32_t *__restrict__ seg, uint32_t nseg)
{
	__shared__ uint32_t warp_totals[34];
	const uint32_t per = (np + blockDim.x - 1) / blockDim.x;
	const uint32_t q0 = threadIdx.x * per;
	// seg (may be null): parent q is the union of nseg ranges (first row, rows) at seg[(q * nseg + s) * 2] -- what a GPU
	// received of one sub-partition, one piece per sender (staged CPRA exchange); every range is cut into its own items
	auto items_of = [&](uint32_t q) -> uint32_t {
		uint32_t items = 0;
		if (seg) {
			for (uint32_t sgm = 0; sgm < nseg; ++sgm) items += (seg[((size_t)q * nseg + sgm) * 2 + 1] + chunk - 1) / chunk;
		} else {
			const uint64_t size = parent_off ? (uint64_t)(parent_off[q + 1] - parent_off[q]) : n;
			items = (uint32_t)((size + chunk - 1) / chunk);
		}
		return items ? items : 1u;                     // empty parents keep one (empty) item
	};
	uint32_t local = 0;
	for (uint32_t q = q0; q < q0 + per && q < np; ++q) local += items_of(q);
	uint32_t total;
	uint32_t run = block_exclusive_scan(local, warp_totals, &total);
	for (uint32_t q = q0; q < q0 + per && q < np; ++q) {
		item_prefix[q] = run;
		run += items_of(q);
	}
	if (threadIdx.x == 0) {
		item_prefix[np] = total;
		child_off[child_total_idx] = parent_off ? parent_off[np] : (uint32_t)n;       // end sentinel of the child offsets
	}
}

struct ItemRange {
	uint64_t beg, end;
};

__device__ __forceinline__ bool locate_item(const uint32_t *item_prefix, uint32_t np, const uint32_t *parent_off,
                                            uint64_t n, uint32_t chunk, uint32_t item, ItemRange *r,
                                            const uint32_t *seg = nullptr, uint32_t nseg = 0)
{
	if (item >= item_prefix[np]) return false;
	const uint32_t q = upper_parent(item_prefix, np, item);
	uint32_t j = item - item_prefix[q];
	if (seg) {
		const uint32_t *sq = seg + (size_t)q * nseg * 2;
		for (uint32_t sgm = 0; sgm < nseg; ++sgm) {
			const uint32_t first = sq[2 * sgm], len = sq[2 * sgm + 1], cnt = (len + chunk - 1) / chunk;
			if (j < cnt) {
				r->beg = (uint64_t)first + (uint64_t)j * chunk;
				r->end = min(r->beg + chunk, (uint64_t)first + len);
				return true;
			}
			j -= cnt;
		}
		r->beg = r->end = 0;                            // the one empty item of a parent without tuples
		return true;
	}
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
       uint32_t *__restrict__ child_off, uint64_t *__restrict__ status, uint32_t *__restrict__ tile_counter,
       const uint32_t *__restrict__ out_base)
{
	// out_base (may be null): *out_base is the output position of the first parent's first tuple -- a pass over a RANGE of
	// the parents (the staged CPRA exchange handles the received sub-partitions in parts) continues where the range begins
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
	uint32_t run = s_excl + thread_excl + (out_base ? *out_base : 0u);
#pragma unroll
	for (uint32_t k = 0; k < kScanItems; ++k) {
		if (addr[k] != 0xFFFFFFFFu) {
			counts[addr[k]] = run;
			if (child[k] != 0xFFFFFFFFu) child_off[child[k]] = run;
			run += v[k];
		}
	}
}


// ------------------------------------------------------------------ histogram with tile counts

constexpr uint32_t kTile = 8192;                 // tuples per tile of k_scatter / k_scatter_bulk
constexpr uint32_t kTileGroups = kTile / 4;      // absolutely aligned groups of four tuples per tile
constexpr uint32_t kTcTile = 16384;              // tuples per tile of the tile-count path (k_hist_tiles / k_scatter_tc)
constexpr uint32_t kTcMaxFanout = 512;           // widest pass of the tile-count path

// One more count for digit d: a shared-memory atomic per lane, or -- for fan-outs <= 16, where the lanes
// of a warp would serialise on a handful of counters -- one atomic per warp and digit (match.any).
// Warp-collective when `aggregate` is set.
__device__ __forceinline__ void hist_add(uint32_t *h, uint32_t d, bool valid, bool aggregate)
{
	if (!aggregate) {
		if (valid) atomicAdd(&h[d], 1u);
		return;
	}
	const unsigned m = __match_any_sync(kFullMask, valid ? d : 0xFFFFFFFFu);
	if (valid && (int)lane_id() == __ffs(m) - 1) atomicAdd(&h[d], (uint32_t)__popc(m));
}

// k_hist for fan-outs <= 512.  The item is walked in the scatter's tiles (TILE tuples from the item's
// first aligned group on); besides the item's counts row, the digit counts of every tile go to
// tile_counts[item][tile][digit] (uint16: a tile holds at most 16384 tuples).  Two shared-memory histograms
// take turns so that one barrier per tile suffices.
template <uint32_t TILE>
__global__ void __launch_bounds__(kHistThreads)
k_hist_tiles(const uint32_t *__restrict__ keys, uint64_t n, uint32_t np, const uint32_t *__restrict__ parent_off,
             const uint32_t *__restrict__ item_prefix, uint32_t chunk, uint32_t factor, int rshift, int bits,
             uint32_t *__restrict__ counts, uint16_t *__restrict__ tile_counts, uint32_t tiles_per_item,
             const uint32_t *__restrict__ seg, uint32_t nseg)
{
	__shared__ uint32_t s_hist[2][kTcMaxFanout];
	constexpr uint32_t TILE_GROUPS = TILE / 4;
	constexpr int GPT = TILE_GROUPS / kHistThreads;            // groups per thread and tile
	static_assert(GPT * kHistThreads == TILE_GROUPS, "tile must be a whole number of rounds");
	const uint32_t F = 1u << bits, mask = F - 1;
	const bool aggregate = bits <= 4;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r, seg, nseg)) return;
	if (threadIdx.x < kTcMaxFanout) {
		s_hist[0][threadIdx.x] = 0;
		s_hist[1][threadIdx.x] = 0;
	}
	__syncthreads();
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	uint16_t *trow = tile_counts + (size_t)blockIdx.x * tiles_per_item * F;
	uint32_t total = 0, j = 0;
	for (uint64_t g0 = g_beg; g0 < g_end; g0 += TILE_GROUPS, ++j) {
		uint32_t *h = s_hist[j & 1];
		const bool full = (g0 << 2) >= r.beg && ((g0 + TILE_GROUPS) << 2) <= r.end;     // r.end <= n: vector loads stay inside
		if (full) {
			uint4 w[GPT];
#pragma unroll
			for (int t = 0; t < GPT; ++t) w[t] = ldg_stream_u4(reinterpret_cast<const uint4 *>(keys) + g0 + threadIdx.x + t * kHistThreads);
#pragma unroll
			for (int t = 0; t < GPT; ++t) {
				hist_add(h, radix_digit(hash_mul(w[t].x, factor), rshift, mask), true, aggregate);
				hist_add(h, radix_digit(hash_mul(w[t].y, factor), rshift, mask), true, aggregate);
				hist_add(h, radix_digit(hash_mul(w[t].z, factor), rshift, mask), true, aggregate);
				hist_add(h, radix_digit(hash_mul(w[t].w, factor), rshift, mask), true, aggregate);
			}
		} else {
#pragma unroll
			for (int t = 0; t < GPT; ++t) {
				const uint64_t g = g0 + threadIdx.x + t * kHistThreads;
				uint32_t k[4] = {0, 0, 0, 0};
				if (g < g_end) load_group4(keys, g, n, k);
#pragma unroll
				for (int e = 0; e < 4; ++e) {
					const uint64_t idx = (g << 2) + e;
					hist_add(h, radix_digit(hash_mul(k[e], factor), rshift, mask), g < g_end && idx >= r.beg && idx < r.end, aggregate);
				}
			}
		}
		__syncthreads();
		if (threadIdx.x < F) {
			const uint32_t c = h[threadIdx.x];
			h[threadIdx.x] = 0;                        // ready for tile j + 2 (the barrier of tile j + 1 lies in between)
			trow[(size_t)j * F + threadIdx.x] = (uint16_t)c;
			total += c;
		}
	}
	if (threadIdx.x < F) counts[(size_t)blockIdx.x * F + threadIdx.x] = total;
}

// ------------------------------------------------------------------ scatter (fan-out <= 512, tile counts known)
//
// Per tile of 16384 tuples:
//   place   every tuple takes its slot in the digit-grouped shared-memory tile with ONE shared-memory atomicAdd
//           on its digit's cursor (the cursors start at the digits' tile offsets, known from k_hist_tiles'
//           counts), and goes there with one 8-byte store; the tuples a digit carried over from the previous
//           tile are flushed to global memory
//   stream  the tile is written out, neighbouring threads writing neighbouring addresses of one
//           partition's run
// and beside the stream, on the first F threads: the plan of the NEXT tile (one digit per thread: tile offset,
// global position, flush limit), double-buffered, so it never sits between two phases.  Two barriers per tile.
// The loads of tile t+1 are issued before tile t is streamed.
//
// Software write-combining (the reference's per-partition staging buffers, cpra2.cpp:976-1008, flush
// cpra2.cpp:711-729): a digit's run is only written up to the last 32-byte sector boundary of its output
// position; the < 8 tuples beyond it wait in a per-digit carry buffer and lead the digit's run of the next
// tile.  Every store but an item's first and last per digit then ends on a sector boundary, so L2 does not
// have to fetch the rest of a half-written sector from HBM (measured in round 1: 2.70 -> 2.19 GB of DRAM
// traffic per 2^27-tuple launch).
//
// dynamic shared memory: cursor[MAXF] | cf[MAXF] (uint2) | golim[2][MAXF] (uint2) | buf[TILE] (uint2) | carry[MAXF * 8] (uint2)
// MAXF = 256, or 512 for the 9-bit passes of the staged CPRA exchange (174 KB with the 16384-tuple tile)
constexpr uint32_t kLocalCarry = 8;    // tuples per 32-byte sector of a 4-byte column
constexpr size_t scatter_tc_smem(uint32_t tile, uint32_t maxf) { return (size_t)maxf * (4 + 8 + 16 + kLocalCarry * 8) + (size_t)tile * 8; }

// One tile = THREADS * 4 G tuples = THREADS * G absolutely aligned groups; thread t owns groups t,
// t + THREADS, ... of the tile (coalesced 128-bit loads).
template <int THREADS, int G, bool FULL>
__device__ __forceinline__ void load_tile_col(uint32_t (&x)[4 * G], uint32_t &ok, const uint32_t *col, uint64_t g0, uint64_t g_end,
                                              uint64_t beg, uint64_t end, uint64_t n)
{
	ok = 0;
#pragma unroll
	for (int h = 0; h < G; ++h) {
		const uint64_t g = g0 + threadIdx.x + (uint64_t)h * THREADS;
		if (FULL) {
			const uint4 w = ldg_stream_u4(reinterpret_cast<const uint4 *>(col) + g);
			x[4 * h + 0] = w.x; x[4 * h + 1] = w.y; x[4 * h + 2] = w.z; x[4 * h + 3] = w.w;
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
	if (FULL) ok = 0xFFFFFFFFu;
}

// `shift` (may be null): added to every digit's output positions -- the staged CPRA exchange lays the digits'
// runs out with the 16-byte phase of their destination rows in the owners' buffers (k_stage_bases)
template <int THREADS, int G, int MINB, uint32_t MAXF>
__global__ void __launch_bounds__(THREADS, MINB)
k_scatter_tc(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
             const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
             uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets,
             const uint16_t *__restrict__ tile_counts, uint32_t tiles_per_item, const int32_t *__restrict__ shift,
             const uint32_t *__restrict__ seg, uint32_t nseg, uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
	constexpr int IT = 4 * G;                                   // tuples per thread and tile
	constexpr uint32_t TILE = THREADS * IT, TILE_GROUPS = TILE / 4;
	extern __shared__ __align__(16) uint32_t s_mem[];
	__shared__ uint32_t warp_totals[MAXF / 32];
	__shared__ uint32_t s_tile_n[2];
	uint32_t *cursor = s_mem;
	uint2 *cf = reinterpret_cast<uint2 *>(cursor + MAXF);         // carried tuples to flush: x = global position of the first, y = how many
	uint2 *golim = cf + MAXF;                                     // [2][MAXF]; x: global position of tile slot 0, y: flush limit
	uint2 *buf = golim + 2 * MAXF;
	uint2 *carry = buf + TILE;
	const uint32_t F = 1u << bits, mask = F - 1;
	const bool aggregate = bits <= 4;
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r, seg, nseg)) return;
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	const uint32_t ntiles = (uint32_t)((g_end - g_beg + TILE_GROUPS - 1) / TILE_GROUPS);
	if (ntiles == 0) return;
	const uint16_t *trow = tile_counts + (size_t)blockIdx.x * tiles_per_item * F;
	// per-digit state lives in the registers of thread d: next output position, tuples waiting in the carry buffer,
	// and the digit's count in the tile to be planned next
	const uint32_t plan_threads = F <= 32 ? 32u : F;
	uint32_t wpos = 0, pend = 0, c_next = 0;
	if (threadIdx.x < F) {
		wpos = offsets[(size_t)blockIdx.x * F + threadIdx.x] + (shift ? (uint32_t)shift[threadIdx.x] : 0u);
		c_next = trow[threadIdx.x];
	}
	auto tile_is_full = [&](uint64_t g0) {
		return (g0 << 2) >= r.beg && ((g0 + TILE_GROUPS) << 2) <= r.end;    // r.end <= n: vector loads stay inside
	};
	// plan of tile j, by the first plan_threads threads (whole warps): tile offsets by an exclusive scan over the
	// digits, then per digit how far the item's output may be flushed and what stays carried
	auto plan = [&](uint32_t j, bool last) {
		const uint32_t p = threadIdx.x, c = p < F ? c_next : 0u;
		const uint32_t incl = warp_inclusive_scan_u32(c);
		uint32_t before = 0, total;
		if (plan_threads > 32) {
			if (lane_id() == 31) warp_totals[p >> 5] = incl;
			asm volatile("bar.sync 1, %0;" ::"r"(plan_threads) : "memory");          // the planning warps only
			const uint32_t t = lane_id() < (plan_threads >> 5) ? warp_totals[lane_id()] : 0u;
			const uint32_t tincl = warp_inclusive_scan_u32(t);
			before = __shfl_sync(kFullMask, tincl - t, p >> 5);
			total = __shfl_sync(kFullMask, tincl, 31);
		} else {
			total = __shfl_sync(kFullMask, incl, 31);
		}
		if (p == 0) s_tile_n[j & 1] = total;
		if (p < F) {
			const uint32_t lbase = before + incl - c;
			const uint32_t endpos = wpos + pend + c;
			uint32_t lim = last ? endpos : (endpos & ~(kLocalCarry - 1));
			const bool flush = lim > wpos;
			if (!flush) lim = wpos;
			cursor[p] = lbase;
			golim[(j & 1) * MAXF + p] = make_uint2(wpos + pend - lbase, lim);
			cf[p] = make_uint2(wpos, flush ? pend : 0u);
			wpos = lim;
			pend = endpos - lim;
		}
	};
	uint32_t key[IT], val[IT], ok;
	if (tile_is_full(g_beg)) {
		load_tile_col<THREADS, G, true>(key, ok, keys, g_beg, g_end, r.beg, r.end, n);
		load_tile_col<THREADS, G, true>(val, ok, vals, g_beg, g_end, r.beg, r.end, n);
	} else {
		load_tile_col<THREADS, G, false>(key, ok, keys, g_beg, g_end, r.beg, r.end, n);
		load_tile_col<THREADS, G, false>(val, ok, vals, g_beg, g_end, r.beg, r.end, n);
	}
	if (threadIdx.x < plan_threads) {
		plan(0, ntiles == 1);
		if (threadIdx.x < F && ntiles > 1) c_next = trow[F + threadIdx.x];
	}
	__syncthreads();
	for (uint32_t j = 0; j < ntiles; ++j) {
		const uint64_t g1 = g_beg + (uint64_t)(j + 1) * TILE_GROUPS;
		// ---- place
		if (aggregate) {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				const bool valid = (ok >> e) & 1u;
				const unsigned m = __match_any_sync(kFullMask, valid ? d : 0xFFFFFFFFu);
				const int leader = __ffs(m) - 1;
				uint32_t base = 0;
				if (valid && (int)lane_id() == leader) base = atomicAdd(&cursor[d], (uint32_t)__popc(m));
				base = __shfl_sync(kFullMask, base, leader);
				if (valid) buf[base + __popc(m & lanemask_lt())] = make_uint2(key[e], val[e]);
			}
		} else if (ok == 0xFFFFFFFFu) {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				buf[atomicAdd(&cursor[d], 1u)] = make_uint2(key[e], val[e]);
			}
		} else {
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				if ((ok >> e) & 1u) buf[atomicAdd(&cursor[d], 1u)] = make_uint2(key[e], val[e]);
			}
		}
		// the carried tuples of every digit that reaches a sector boundary with this tile
		for (uint32_t i = threadIdx.x; i < F * kLocalCarry; i += THREADS) {
			const uint32_t d = i / kLocalCarry, q = i % kLocalCarry;
			const uint2 c = cf[d];
			if (q < c.y) {
				const uint2 kv = carry[i];
				keys_out[c.x + q] = kv.x;
				vals_out[c.x + q] = kv.y;
			}
		}
		__syncthreads();
		// ---- next tile: loads on their way, plan on the first F threads
		if (j + 1 < ntiles) {
			if (tile_is_full(g1)) {
				load_tile_col<THREADS, G, true>(key, ok, keys, g1, g_end, r.beg, r.end, n);
				load_tile_col<THREADS, G, true>(val, ok, vals, g1, g_end, r.beg, r.end, n);
			} else {
				load_tile_col<THREADS, G, false>(key, ok, keys, g1, g_end, r.beg, r.end, n);
				load_tile_col<THREADS, G, false>(val, ok, vals, g1, g_end, r.beg, r.end, n);
			}
			if (threadIdx.x < plan_threads) {
				plan(j + 1, j + 2 == ntiles);
				if (threadIdx.x < F && j + 2 < ntiles) c_next = trow[(size_t)(j + 2) * F + threadIdx.x];
			}
		}
		// ---- stream tile j: below the digit's limit to global memory, beyond it into the carry buffer
		const uint32_t tile_n = s_tile_n[j & 1];
		const uint2 *gl_tab = golim + (j & 1) * MAXF;
#pragma unroll 4
		for (int it = 0; it < IT; ++it) {
			const uint32_t i = threadIdx.x + it * THREADS;
			if (i < tile_n) {
				const uint2 kv = buf[i];
				const uint32_t d = radix_digit(hash_mul(kv.x, factor), rshift, mask);
				const uint2 gl = gl_tab[d];
				const uint32_t pos = gl.x + i;
				if (pos < gl.y) {
					keys_out[pos] = kv.x;
					vals_out[pos] = kv.y;
				} else {
					carry[d * kLocalCarry + (pos - gl.y)] = kv;
				}
			}
		}
		__syncthreads();
	}
}

// ------------------------------------------------------------------ scatter (fan-out 1024 .. 2048)
//
// Per tile: (1) every tuple takes a rank inside its digit with a shared-memory atomicAdd, (2) the digit
// counts become tile offsets, (3) tuples are placed into shared memory grouped by digit, (4) the tile is
// streamed out run by run.  No write-combining: the carry buffers of 2048 digits do not fit.
// dynamic shared memory: cnt base wpos [F] | buf[kTile] (uint2)
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1)
k_scatter(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
          const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
          uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets,
          uint32_t *__restrict__ keys_out, uint32_t *__restrict__ vals_out)
{
	constexpr int G = kTile / 4 / THREADS, IT = 4 * G;
	extern __shared__ __align__(16) uint32_t s_mem[];
	__shared__ uint32_t warp_totals[34];
	__shared__ uint32_t s_tile_n;
	const uint32_t F = 1u << bits, mask = F - 1;
	uint32_t *cnt = s_mem, *base = cnt + F, *wpos = base + F;
	uint2 *buf = reinterpret_cast<uint2 *>(wpos + F);
	ItemRange r;
	if (!locate_item(item_prefix, np, parent_off, n, chunk, blockIdx.x, &r)) return;
	const uint32_t *row = offsets + (size_t)blockIdx.x * F;
	for (uint32_t p = threadIdx.x; p < F; p += THREADS) {
		wpos[p] = row[p];
		cnt[p] = 0;
	}
	const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
	auto tile_is_full = [&](uint64_t g0) { return (g0 << 2) >= r.beg && ((g0 + kTileGroups) << 2) <= r.end; };
	__syncthreads();
	for (uint64_t g0 = g_beg; g0 < g_end; g0 += kTileGroups) {
		uint32_t key[IT], val[IT], ok;
		if (tile_is_full(g0)) {
			load_tile_col<THREADS, G, true>(key, ok, keys, g0, g_end, r.beg, r.end, n);
			load_tile_col<THREADS, G, true>(val, ok, vals, g0, g_end, r.beg, r.end, n);
		} else {
			load_tile_col<THREADS, G, false>(key, ok, keys, g0, g_end, r.beg, r.end, n);
			load_tile_col<THREADS, G, false>(val, ok, vals, g0, g_end, r.beg, r.end, n);
		}
		uint32_t dr[IT];                              // digit << 16 | rank-in-digit (rank < kTile <= 2^16, digit < 2^11)
#pragma unroll
		for (int e = 0; e < IT; ++e) {
			const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
			dr[e] = (ok >> e) & 1u ? (d << 16) | atomicAdd(&cnt[d], 1u) : 0xFFFFFFFFu;
		}
		__syncthreads();
		const uint32_t ept = (F + THREADS - 1) / THREADS;
		const uint32_t p0 = threadIdx.x * ept;
		uint32_t local = 0;
		for (uint32_t p = p0; p < p0 + ept && p < F; ++p) local += cnt[p];
		uint32_t tile_total;
		uint32_t run = block_exclusive_scan(local, warp_totals, &tile_total);
		for (uint32_t p = p0; p < p0 + ept && p < F; ++p) {
			const uint32_t c = cnt[p];
			base[p] = run;
			run += c;
			cnt[p] = 0;
		}
		if (threadIdx.x == 0) s_tile_n = tile_total;
		__syncthreads();
#pragma unroll
		for (int e = 0; e < IT; ++e)
			if (dr[e] != 0xFFFFFFFFu) buf[base[dr[e] >> 16] + (dr[e] & 0xFFFFu)] = make_uint2(key[e], val[e]);
		__syncthreads();
		const uint32_t tile_n = s_tile_n;
#pragma unroll
		for (int it = 0; it < IT; ++it) {
			const uint32_t i = threadIdx.x + it * THREADS;
			if (i < tile_n) {
				const uint2 kv = buf[i];
				const uint32_t d = radix_digit(hash_mul(kv.x, factor), rshift, mask);
				const uint32_t pos = wpos[d] + (i - base[d]);
				keys_out[pos] = kv.x;
				vals_out[pos] = kv.y;
			}
		}
		__syncthreads();
		for (uint32_t p = p0; p < p0 + ept && p < F; ++p) {
			const uint32_t nxt = p + 1 < F ? base[p + 1] : tile_n;
			wpos[p] += nxt - base[p];
		}
		// base[] is rewritten after the next tile's rank barrier, which orders it behind these reads
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
// The digit-grouped tile is multi-buffered (three tiles for fan-outs <= 8, two up to 32): the bulk copies of
// tile t drain over NVLink while tiles t+1, t+2 are ranked, planned AND placed into the other buffers.
// dynamic shared memory: cnt wpos pend [64] | place[64] (uint4) | strm[64] (uint4) | cin[64] (uint2) |
//                        NB x { skeys[PAD] svals[PAD] } | carry_k[2][32 F] carry_v[2][32 F],  PAD = kTile + 64 F
constexpr uint32_t kBulkGranule = kPeerCarry;             // 32 tuples = 128 bytes per column

// staging tiles of the peer scatter: as many as fit (HJB_BULK_BUFFERS caps it for A/B runs)
static inline uint32_t bulk_buffers(uint32_t F)
{
	static const int cap = getenv("HJB_BULK_BUFFERS") ? atoi(getenv("HJB_BULK_BUFFERS")) : 3;
	const uint32_t fit = F <= 8 ? 3u : F <= 32 ? 2u : 1u;
	return cap >= 1 && (uint32_t)cap < fit ? (uint32_t)cap : fit;
}
static inline size_t bulk_smem_bytes(uint32_t F)
{
	const size_t pad = kTile + 2 * (size_t)kBulkGranule * F;
	return 52 * 64 + bulk_buffers(F) * pad * 8 + (size_t)F * kBulkGranule * 16;
}

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes)
{
	asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
	             "r"((uint32_t)__cvta_generic_to_shared(ssrc)), "r"(bytes)
	             : "memory");
}

template <int THREADS, int MAXF>
__global__ void __launch_bounds__(THREADS, 1)
k_scatter_bulk(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, uint32_t np,
               const uint32_t *__restrict__ parent_off, const uint32_t *__restrict__ item_prefix, uint32_t chunk,
               uint32_t factor, int rshift, int bits, const uint32_t *__restrict__ offsets, const PeerTable peers, uint32_t nbuf)
{
	constexpr int G = kTile / 4 / THREADS, IT = 4 * G;
	constexpr uint32_t GR = kBulkGranule;
	extern __shared__ __align__(128) uint32_t s_bulk[];
	__shared__ uint32_t warp_totals[MAXF / 32];
	__shared__ uint32_t *s_pk[64], *s_pv[64];
	__shared__ uint32_t s_bias[64];
	const uint32_t F = 1u << bits, mask = F - 1;
	const uint32_t PAD = kTile + 2 * GR * F;
	// nbuf staging tiles (bulk_buffers): the bulk copies of up to nbuf - 1 earlier tiles may still be draining
	uint32_t *cnt = s_bulk, *wpos = cnt + MAXF, *pend = wpos + MAXF;
	uint4 *place = reinterpret_cast<uint4 *>(pend + MAXF);     // x: slot of new rank 0, y: new tuples that fit the region, z: carry index of rank 0
	uint4 *strm = place + MAXF;                                // x: global position of slot 0, y: first valid position, z: end of valid positions
	uint2 *cin = reinterpret_cast<uint2 *>(strm + MAXF);       // carried-in tuples: x: destination slot 0 (0xFFFFFFFF: stay carried), y: how many
	uint32_t *tile0 = reinterpret_cast<uint32_t *>(cin + MAXF);        // byte offset 52 * MAXF, a multiple of 128
	uint32_t *carry_k = tile0 + 2 * nbuf * PAD, *carry_v = carry_k + 2 * GR * F;
	if (peers.abort_flag && *peers.abort_flag) return;
	// The scan's positions start at sender_off[g] for owner g; they must land at base[g] of the owner's columns.
	// The kernel counts positions as sender_off + bias, congruent to the physical row modulo the write-combining
	// granule, so that its flush boundaries are line boundaries in the owner's buffer.
	if (threadIdx.x < F) {
		const int64_t shift = (int64_t)peers.base[threadIdx.x] - (int64_t)peers.sender_off[threadIdx.x];
		const int64_t bias = ((shift % (int64_t)GR) + GR) % GR;
		s_bias[threadIdx.x] = (uint32_t)bias;
		s_pk[threadIdx.x] = peers.k[threadIdx.x] + (shift - bias);
		s_pv[threadIdx.x] = peers.v[threadIdx.x] + (shift - bias);
	}
	uint32_t tile_no = 0;                                       // over all items of this CTA: buffer parity
	for (uint32_t item = blockIdx.x;; item += gridDim.x) {
		ItemRange r;
		if (!locate_item(item_prefix, np, parent_off, n, chunk, item, &r)) break;
		__syncthreads();
		const uint32_t *row = offsets + (size_t)item * F;
		if (threadIdx.x < F) {
			wpos[threadIdx.x] = row[threadIdx.x] + s_bias[threadIdx.x];
			pend[threadIdx.x] = 0;
			cnt[threadIdx.x] = 0;
		}
		const uint64_t g_beg = r.beg >> 2, g_end = (r.end + 3) >> 2;
		auto tile_is_full = [&](uint64_t g0) { return (g0 << 2) >= r.beg && ((g0 + kTileGroups) << 2) <= r.end; };
		uint32_t key[IT], nkey[IT], val[IT], ok = 0, nok = 0;
		if (g_beg < g_end) {
			if (tile_is_full(g_beg)) load_tile_col<THREADS, G, true>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
			else load_tile_col<THREADS, G, false>(nkey, nok, keys, g_beg, g_end, r.beg, r.end, n);
		}
		__syncthreads();
		for (uint64_t g0 = g_beg; g0 < g_end; g0 += kTileGroups, ++tile_no) {
			const uint64_t g1 = g0 + kTileGroups;
			const bool last = g1 >= g_end;
			uint32_t *const skeys = tile0 + (tile_no % nbuf) * 2 * PAD, *const svals = skeys + PAD;
			uint32_t *const oc_k = carry_k + (tile_no & 1) * GR * F, *const oc_v = carry_v + (tile_no & 1) * GR * F;            // carried in
			uint32_t *const nc_k = carry_k + ((tile_no & 1) ^ 1) * GR * F, *const nc_v = carry_v + ((tile_no & 1) ^ 1) * GR * F;  // carried out
#pragma unroll
			for (int e = 0; e < IT; ++e) key[e] = nkey[e];
			ok = nok;
			uint32_t vok;
			if (ok == 0xFFFFFFFFu) load_tile_col<THREADS, G, true>(val, vok, vals, g0, g_end, r.beg, r.end, n);
			else load_tile_col<THREADS, G, false>(val, vok, vals, g0, g_end, r.beg, r.end, n);
			if (!last) {
				if (tile_is_full(g1)) load_tile_col<THREADS, G, true>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
				else load_tile_col<THREADS, G, false>(nkey, nok, keys, g1, g_end, r.beg, r.end, n);
			}
			// (1) rank
			uint32_t dr[IT];
#pragma unroll
			for (int e = 0; e < IT; ++e) {
				const uint32_t d = radix_digit(hash_mul(key[e], factor), rshift, mask);
				if (bits <= 4) dr[e] = rank_aggregated(cnt, d, (ok >> e) & 1u);
				else dr[e] = ((ok >> e) & 1u) ? (d << 16) | atomicAdd(&cnt[d], 1u) : 0xFFFFFFFFu;
			}
			// the bulk copies that last read the buffer about to be refilled must have read their shared-memory source
			if (threadIdx.x < F) {
				if (nbuf == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
				else if (nbuf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
				else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
			}
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
			asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the bulk copy
			__syncthreads();
			// (4) one thread per digit: whole run as bulk copies, unaligned ends as scalar stores
			if (threadIdx.x < F) {
				const uint32_t d = threadIdx.x;
				const uint4 st = strm[d];
				if (st.w) {
					uint32_t *const ko = s_pk[d & 63], *const vo = s_pv[d & 63];
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
	}
	// the last bulk copies must complete before the CTA exits
	if (threadIdx.x < MAXF) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// ------------------------------------------------------------------ host launchers

// experiment knob, read once (host/cpra.cpp calls the launchers from one thread per GPU)
static int radix_items()
{
	static const int items = [] {
		const int v = getenv("HJB_ITEMS") ? atoi(getenv("HJB_ITEMS")) : 1184;      // work items per pass
		return v < 64 ? 1184 : v;
	}();
	return items;
}

size_t radix_scratch_bytes(uint64_t n, uint32_t np, int bits, uint32_t *chunk, uint32_t *max_items, uint32_t *tiles,
                           uint32_t *tiles_per_item, uint32_t chunk_div)
{
	// chunk_div > 1: the pass will see about n / chunk_div of the n tuples the array can hold (one part of a relation that
	// is processed in parts); items are sized for that, the bound on their number still covers all n
	// ~1.2K items (eight per SM of a B200): enough to balance the SMs over a pass, few enough that the
	// counts matrix and its scan stay small (measured: 1024-1184 items 4.23 ms per config-2 step, 2048
	// 4.27, 4096 4.35); chunk is a multiple of the scatter tile
	const int target = radix_items();
	const uint32_t tile = kTcTile;
	uint64_t c = (n / (chunk_div ? chunk_div : 1u) + target - 1) / target;
	c = (c + kTcTile - 1) / kTcTile * kTcTile;                 // a multiple of every tile size in use
	if (c < kTcTile) c = kTcTile;
	if (c > (1u << 24)) c = 1u << 24;
	*chunk = (uint32_t)c;
	const uint64_t mi = n / c + np + 1;
	*max_items = (uint32_t)mi;
	const uint64_t E = mi << bits;
	*tiles = (uint32_t)((E + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems));
	const uint32_t tpi = (uint32_t)(c / tile) + 1;                // an item starts inside an aligned group: one tile more than chunk / tile
	if (tiles_per_item) *tiles_per_item = tpi;
	size_t bytes = 0;
	bytes += ((size_t)(np + 1) * 4 + 255) / 256 * 256;           // item_prefix
	bytes += ((size_t)E * 4 + 255) / 256 * 256;                   // counts / offsets
	bytes += ((size_t)*tiles * 8 + 255) / 256 * 256 + 256;        // scan status + counter
	if ((1u << bits) <= kTcMaxFanout) bytes += ((size_t)E * tpi * 2 + 255) / 256 * 256;   // tile counts
	return bytes + 1280;                                            // per-array 256-byte padding of the bump allocator
}

void radix_carve(RadixPassArgs &a, char *scratch, bool tile_counts)
{
	uint32_t tiles;
	radix_scratch_bytes(a.n, a.np * (a.nseg ? a.nseg : 1u), a.bits, &a.chunk, &a.max_items, &tiles, &a.tiles_per_item,
	                    a.chunk_div);                       // every range of a parent may end in a short item
	size_t off = 0;
	auto take = [&](size_t bytes) {
		char *p = scratch + off;
		off += (bytes + 255) & ~(size_t)255;
		return p;
	};
	a.item_prefix = reinterpret_cast<uint32_t *>(take(((size_t)a.np + 1) * 4));
	a.counts = reinterpret_cast<uint32_t *>(take(((size_t)a.max_items << a.bits) * 4));
	a.scan_status = reinterpret_cast<uint64_t *>(take((size_t)tiles * 8));
	a.scan_counter = reinterpret_cast<uint32_t *>(take(4));
	a.tile_counts = nullptr;
	if (tile_counts && (1u << a.bits) <= kTcMaxFanout)
		a.tile_counts = reinterpret_cast<uint16_t *>(take(((size_t)a.max_items << a.bits) * a.tiles_per_item * 2));
}

static void scatter_attrs()
{
	// the attribute is per device: a process that drives several GPUs (host/cpra.cpp) sets it once on each
	static std::atomic<unsigned long long> done_mask{0};
	int dev = 0;
	cudaGetDevice(&dev);
	const unsigned long long bit = 1ull << (dev & 63);
	if (done_mask.fetch_or(bit) & bit) return;
	cudaFuncSetAttribute(k_scatter_tc<1024, 4, 1, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_tc_smem(kTcTile, 256));
	cudaFuncSetAttribute(k_scatter_tc<1024, 4, 1, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scatter_tc_smem(kTcTile, 512));
	cudaFuncSetAttribute(k_scatter<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2048 * 12 + (int)kTile * 8);
	size_t bulk_max = 0;
	for (uint32_t f = 2; f <= 64; f *= 2) bulk_max = bulk_smem_bytes(f) > bulk_max ? bulk_smem_bytes(f) : bulk_max;
	cudaFuncSetAttribute(k_scatter_bulk<1024, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bulk_max);
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
	k_make_items<<<1, 1024, 0, s>>>(a.parent_off, a.np, a.n, a.chunk, a.item_prefix, a.child_off, a.np << a.bits, a.seg, a.nseg);
	t->stop(s);
	t->start(KK_HIST, s);
	if (a.tile_counts)
		k_hist_tiles<kTcTile><<<a.max_items, kHistThreads, 0, s>>>(a.keys, a.n, a.np, a.parent_off, a.item_prefix, a.chunk, a.factor,
		                                                           a.rshift, a.bits, a.counts, a.tile_counts, a.tiles_per_item, a.seg, a.nseg);
	else if (a.bits <= 3)
		k_hist_small<<<a.max_items, kHistThreads, 0, s>>>(a.keys, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,
		                                                  a.factor, a.rshift, a.bits, a.counts);
	else
		k_hist<<<a.max_items, kHistThreads, F * 4, s>>>(a.keys, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,
		                                                a.factor, a.rshift, a.bits, a.counts);
	t->stop(s);
	t->start(KK_SCAN, s);
	k_scan<<<tiles, kScanThreads, 0, s>>>(a.item_prefix, a.np, a.bits, a.counts, a.child_off, a.scan_status,
	                                      a.scan_counter, a.out_base);
	t->stop(s);
	return 3;
}

// the scatter of a counted pass: into a.keys_out / a.vals_out, or -- peers given -- into the owners' columns
int launch_radix_scatter(const RadixPassArgs &a, cudaStream_t s, KernelTimer *t, const PeerTable *peers)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	const uint32_t F = 1u << a.bits;
	scatter_attrs();
	const uint32_t grid = a.max_items;
	if (peers) {
		if (F > 64) return -1;
		// the bulk kernel walks the items with a grid stride: peer_ctas can leave SMs to other streams
		const uint32_t grid_b = (a.peer_ctas && a.peer_ctas < grid) ? a.peer_ctas : grid;
		t->start(KK_SCATTER_PEER, s);
		k_scatter_bulk<1024, 64><<<grid_b, 1024, bulk_smem_bytes(F), s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix, a.chunk,
		                                                                   a.factor, a.rshift, a.bits, a.counts, *peers, bulk_buffers(F));
		t->stop(s);
		return 1;
	}
	t->start(KK_SCATTER, s);
	if (a.tile_counts && F > 256)
		k_scatter_tc<1024, 4, 1, 512><<<grid, 1024, scatter_tc_smem(kTcTile, 512), s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix,
		                                                                                a.chunk, a.factor, a.rshift, a.bits, a.counts, a.tile_counts,
		                                                                                a.tiles_per_item, a.shift, a.seg, a.nseg, a.keys_out, a.vals_out);
	else if (a.tile_counts)
		k_scatter_tc<1024, 4, 1, 256><<<grid, 1024, scatter_tc_smem(kTcTile, 256), s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix,
		                                                                                a.chunk, a.factor, a.rshift, a.bits, a.counts, a.tile_counts,
		                                                                                a.tiles_per_item, a.shift, a.seg, a.nseg, a.keys_out, a.vals_out);
	else
		k_scatter<1024><<<grid, 1024, (size_t)F * 12 + (size_t)kTile * 8, s>>>(a.keys, a.vals, a.n, a.np, a.parent_off, a.item_prefix,
		                                                                       a.chunk, a.factor, a.rshift, a.bits, a.counts, a.keys_out,
		                                                                       a.vals_out);
	t->stop(s);
	return 1;
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
