// skew.cu -- heavy-hitter handling for the multi-GPU join (CPRA).
//
// The reference gives thread t the partitions [t P/T, (t+1) P/T) whatever they hold (par_start / par_end,
// cpra2.cpp:1868-1872): a probe side with a few very frequent keys (write.cpp's `zipf` knob, write.cpp:1685-1689)
// loads one owner with every tuple of such a key.  Here the probe tuples of the hottest keys never travel: they
// are split off the chunk before the GPU-assign pass, the few build tuples with those keys are replicated to
// every GPU, and each GPU joins its own hot probe tuples against them.  The rest of the chunk takes the normal
// path, so the owners receive balanced shares.
//   k_split_hot   S chunk -> (cold S, hot S) by membership in the hot-key set         8 B read + 8 B written per tuple
//   k_select_hot  R chunk -> its tuples with hot keys (a handful)                      4 B read per tuple
//   k_hot_join    hot S x (all GPUs' hot R tuples): table in shared memory, rows appended to the join's result
#include "hj_device.cuh"
#include "hj_internal.h"

namespace hjb {

constexpr uint32_t kHotSetSlots = 1024;          // open-addressing set of <= kMaxHotKeys keys
constexpr uint64_t kHotValid = 1ull << 32;

// every thread of the CTA calls this; set: kHotSetSlots uint64 of shared memory
__device__ __forceinline__ void hot_set_build(uint64_t *set, const uint32_t *__restrict__ hot, uint32_t n_hot)
{
	for (uint32_t i = threadIdx.x; i < kHotSetSlots; i += blockDim.x) set[i] = 0;
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < n_hot; i += blockDim.x) {
		const uint64_t e = kHotValid | hot[i];
		uint32_t h = (hot[i] * 0x9E3779B1u) >> 22;
		while (true) {
			const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long *>(&set[h]), 0ull, (unsigned long long)e);
			if (old == 0 || old == e) break;
			h = (h + 1) & (kHotSetSlots - 1);
		}
	}
	__syncthreads();
}

__device__ __forceinline__ bool hot_set_has(const uint64_t *set, uint32_t key)
{
	const uint64_t e = kHotValid | key;
	uint32_t h = (key * 0x9E3779B1u) >> 22;
	while (true) {
		const uint64_t s = set[h];
		if (s == e) return true;
		if (s == 0) return false;
		h = (h + 1) & (kHotSetSlots - 1);
	}
}

constexpr int kSplitThreads = 256, kSplitItems = 4;

// Two dense outputs from one pass: the CTA reserves the rows of a round (THREADS x ITEMS tuples) in both with one
// atomicAdd each; inside the reservation the warps and lanes take ballot-ranked positions.
__global__ void __launch_bounds__(kSplitThreads)
k_split_hot(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, const uint32_t *__restrict__ hot,
            uint32_t n_hot, uint32_t *__restrict__ cold_k, uint32_t *__restrict__ cold_v, uint32_t *__restrict__ hot_k,
            uint32_t *__restrict__ hot_v, unsigned long long *__restrict__ cursors /* [0] cold rows, [1] hot rows */)
{
	__shared__ uint64_t set[kHotSetSlots];
	__shared__ uint32_t wt[2][2][kSplitThreads / 32];          // [parity][cold / hot][warp]
	__shared__ unsigned long long base[2][2];
	hot_set_build(set, hot, n_hot);
	constexpr uint32_t kRound = kSplitThreads * kSplitItems;
	const uint64_t rounds = (n + kRound - 1) / kRound;
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
	const unsigned lt = lanemask_lt();
	uint32_t par = 0;
	for (uint64_t rd = blockIdx.x; rd < rounds; rd += gridDim.x, par ^= 1) {
		const uint64_t wbase = rd * kRound + (uint64_t)warp * 32 * kSplitItems + lane;
		uint32_t k[kSplitItems], v[kSplitItems];
		bool in[kSplitItems], isHot[kSplitItems];
		uint32_t ncold = 0, nhot = 0;
#pragma unroll
		for (int t = 0; t < kSplitItems; ++t) {
			const uint64_t i = wbase + (uint64_t)t * 32;
			in[t] = i < n;
			k[t] = in[t] ? ldg_stream_u32(&keys[i]) : 0;
			v[t] = in[t] ? ldg_stream_u32(&vals[i]) : 0;
		}
#pragma unroll
		for (int t = 0; t < kSplitItems; ++t) {
			isHot[t] = in[t] && hot_set_has(set, k[t]);
			nhot += __popc(__ballot_sync(kFullMask, isHot[t]));
			ncold += __popc(__ballot_sync(kFullMask, in[t] && !isHot[t]));
		}
		if (lane == 0) {
			wt[par][0][warp] = ncold;
			wt[par][1][warp] = nhot;
		}
		__syncthreads();
		uint32_t before[2], total[2];
#pragma unroll
		for (int c = 0; c < 2; ++c) {
			const uint32_t mine = lane < kSplitThreads / 32 ? wt[par][c][lane] : 0u;
			const uint32_t incl = warp_inclusive_scan_u32(mine);
			total[c] = __shfl_sync(kFullMask, incl, 31);
			before[c] = __shfl_sync(kFullMask, incl - mine, warp);
		}
		if (threadIdx.x == 0 && total[0]) base[par][0] = atomicAdd(&cursors[0], (unsigned long long)total[0]);
		if (threadIdx.x == 32 && total[1]) base[par][1] = atomicAdd(&cursors[1], (unsigned long long)total[1]);
		__syncthreads();
		unsigned long long pc = base[par][0] + before[0], ph = base[par][1] + before[1];
#pragma unroll
		for (int t = 0; t < kSplitItems; ++t) {
			const unsigned mh = __ballot_sync(kFullMask, isHot[t]), mc = __ballot_sync(kFullMask, in[t] && !isHot[t]);
			if (isHot[t]) {
				hot_k[ph + __popc(mh & lt)] = k[t];
				hot_v[ph + __popc(mh & lt)] = v[t];
			} else if (in[t]) {
				cold_k[pc + __popc(mc & lt)] = k[t];
				cold_v[pc + __popc(mc & lt)] = v[t];
			}
			ph += __popc(mh);
			pc += __popc(mc);
		}
	}
}

__global__ void __launch_bounds__(256)
k_select_hot(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n, const uint32_t *__restrict__ hot,
             uint32_t n_hot, uint32_t *__restrict__ out_k, uint32_t *__restrict__ out_v, uint32_t capacity,
             unsigned long long *__restrict__ cursor)
{
	__shared__ uint64_t set[kHotSetSlots];
	hot_set_build(set, hot, n_hot);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint32_t k = ldg_stream_u32(&keys[i]);
		if (hot_set_has(set, k)) {
			const unsigned long long r = atomicAdd(cursor, 1ull);       // a handful of tuples per relation
			if (r < capacity) {
				out_k[r] = k;
				out_v[r] = vals[i];
			}
		}
	}
}

// one row, reservation aggregated over whichever lanes of the warp are here together
__device__ __forceinline__ void hot_emit_row(const OutCols &out, uint32_t key, uint32_t oval, uint32_t ival)
{
	const unsigned m = __activemask();
	const int leader = __ffs(m) - 1;
	unsigned long long base = 0;
	if ((int)lane_id() == leader) base = atomicAdd(out.cursor, (unsigned long long)__popc(m));
	base = __shfl_sync(m, base, leader);
	const uint64_t r = base + __popc(m & lanemask_lt());
	if (r < out.cap) {
		out.k[r] = key;
		out.o[r] = oval;
		out.i[r] = ival;
	}
}

// hot S x hot R.  Table: 2 * kMaxHotBuild slots of (payload << 32 | key), linear probing, every build tuple a slot
// of its own (equal keys allowed), so a probe walks to the first empty slot and emits every equal key on the way:
// the first match of every tuple goes out with the CTA's round (one reservation per 1024 probe tuples -- a hot
// key's tuples ALL match, and same-address atomics are served one per ~ns), further matches one by one.
// The pair (0xFFFFFFFF, 0xFFFFFFFF) looks like an empty slot: the caller never declares key 0xFFFFFFFF hot.
constexpr int kHotThreads = 256, kHotItems = 4;
__global__ void __launch_bounds__(kHotThreads)
k_hot_join(const uint32_t *__restrict__ sk, const uint32_t *__restrict__ sv, uint64_t ns, const uint32_t *__restrict__ rk,
           const uint32_t *__restrict__ rv, uint32_t nr, uint32_t factor, OutCols out, unsigned long long *__restrict__ sums)
{
	extern __shared__ __align__(16) uint64_t table[];          // 2 * kMaxHotBuild
	__shared__ uint64_t scratch[4 * 32];
	__shared__ __align__(8) uint32_t s_emit[2 * (kHotThreads / 32 + 2) + 4];
	constexpr uint32_t kSlots = 2 * kMaxHotBuild, kMask = kSlots - 1;
	for (uint32_t i = threadIdx.x; i < kSlots; i += blockDim.x) table[i] = kEmptySlot;
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < nr; i += blockDim.x) {
		const uint64_t pair = ((uint64_t)rv[i] << 32) | rk[i];
		uint32_t h = (rk[i] * factor) & kMask;
		while (atomicCAS(reinterpret_cast<unsigned long long *>(&table[h]), (unsigned long long)kEmptySlot, (unsigned long long)pair) != kEmptySlot)
			h = (h + 1) & kMask;
	}
	__syncthreads();
	JoinSums acc;
	acc.zero();
	constexpr uint32_t kRound = kHotThreads * kHotItems;
	const uint64_t rounds = (ns + kRound - 1) / kRound;
	uint32_t emit_rounds = 0;
	for (uint64_t rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
		const uint64_t wbase = rd * kRound + (uint64_t)(threadIdx.x >> 5) * 32 * kHotItems + lane_id();
		uint32_t k[kHotItems], v[kHotItems], ival[kHotItems], h[kHotItems];
		bool found[kHotItems];
#pragma unroll
		for (int t = 0; t < kHotItems; ++t) {
			const uint64_t i = wbase + (uint64_t)t * 32;
			found[t] = i < ns;
			k[t] = found[t] ? ldg_stream_u32(&sk[i]) : 0;
			v[t] = found[t] ? ldg_stream_u32(&sv[i]) : 0;
		}
#pragma unroll
		for (int t = 0; t < kHotItems; ++t) {
			bool hit = false;
			h[t] = (k[t] * factor) & kMask;
			ival[t] = 0;
			while (found[t]) {
				const uint64_t slot = table[h[t]];
				if (slot == kEmptySlot) break;
				h[t] = (h[t] + 1) & kMask;
				if ((uint32_t)slot == k[t]) {
					ival[t] = (uint32_t)(slot >> 32);
					hit = true;
					break;
				}
			}
			found[t] = hit;
			if (hit) acc.add(k[t], v[t], ival[t]);
		}
		if (out.cap) emit_round_cta<kHotItems>(out, s_emit, emit_rounds++, found, k, v, ival);
		// equal build keys: the walk goes on behind the first match
#pragma unroll
		for (int t = 0; t < kHotItems; ++t) {
			while (found[t]) {
				const uint64_t slot = table[h[t]];
				if (slot == kEmptySlot) break;
				h[t] = (h[t] + 1) & kMask;
				if ((uint32_t)slot == k[t]) {
					acc.add(k[t], v[t], (uint32_t)(slot >> 32));
					if (out.cap) hot_emit_row(out, k[t], v[t], (uint32_t)(slot >> 32));
				}
			}
		}
	}
	acc.reduce_to_global(sums, scratch);
}

int launch_split_hot(const uint32_t *keys, const uint32_t *vals, uint64_t n, const uint32_t *hot, uint32_t n_hot, uint32_t *cold_k,
                     uint32_t *cold_v, uint32_t *hot_k, uint32_t *hot_v, unsigned long long *cursors, cudaStream_t s, int sms)
{
	cudaMemsetAsync(cursors, 0, 16, s);
	const uint64_t rounds = (n + kSplitThreads * kSplitItems - 1) / (kSplitThreads * kSplitItems);
	uint64_t grid = rounds < (uint64_t)sms * 8 ? rounds : (uint64_t)sms * 8;
	if (grid == 0) grid = 1;
	k_split_hot<<<(uint32_t)grid, kSplitThreads, 0, s>>>(keys, vals, n, hot, n_hot, cold_k, cold_v, hot_k, hot_v, cursors);
	return 1;
}

int launch_select_hot(const uint32_t *keys, const uint32_t *vals, uint64_t n, const uint32_t *hot, uint32_t n_hot, uint32_t *out_k,
                      uint32_t *out_v, uint32_t capacity, unsigned long long *cursor, cudaStream_t s, int sms)
{
	cudaMemsetAsync(cursor, 0, 8, s);
	uint64_t grid = (n + 255) / 256;
	if (grid > (uint64_t)sms * 8) grid = (uint64_t)sms * 8;
	if (grid == 0) grid = 1;
	k_select_hot<<<(uint32_t)grid, 256, 0, s>>>(keys, vals, n, hot, n_hot, out_k, out_v, capacity, cursor);
	return 1;
}

int launch_hot_join(const uint32_t *sk, const uint32_t *sv, uint64_t ns, const uint32_t *rk, const uint32_t *rv, uint32_t nr,
                    uint32_t factor, uint32_t *out_k, uint32_t *out_o, uint32_t *out_i, uint64_t out_cap, unsigned long long *scalars,
                    cudaStream_t s, int sms)
{
	if (ns == 0 || nr == 0) return 0;
	cudaFuncSetAttribute(k_hot_join, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * kMaxHotBuild * 8);      // per device
	OutCols out;
	out.k = out_k;
	out.o = out_o;
	out.i = out_i;
	out.cursor = scalars;
	out.cap = out_cap;
	uint64_t grid = (ns + kHotThreads * kHotItems - 1) / (kHotThreads * kHotItems);
	if (grid > (uint64_t)sms * 3) grid = (uint64_t)sms * 3;
	k_hot_join<<<(uint32_t)grid, kHotThreads, 2 * kMaxHotBuild * 8, s>>>(sk, sv, ns, rk, rv, nr, factor, out, scalars + 1);
	return 1;
}

}  // namespace hjb
