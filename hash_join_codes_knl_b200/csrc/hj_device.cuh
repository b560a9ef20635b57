// hj_device.cuh -- device-side helpers shared by every kernel of libhjb200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hjb {

constexpr uint64_t kEmptySlot = 0xFFFFFFFFFFFFFFFFull;   // all-ones: cudaMemset(0xFF) initialises a table
constexpr uint32_t kSentinelKey = 0xFFFFFFFFu;           // the one (key, payload) pair equal to kEmptySlot
constexpr unsigned kFullMask = 0xFFFFFFFFu;

// multiplicative hash of the reference: x = (uint32)(key * factor); h(key, f, N) = (x * N) >> 32
// (npj.cpp:200-201, simd_hash npj.cpp:90-106).  Radix digits are bit fields of the same x, i.e.
// h(key, f, 2^B) for B consumed bits.
__device__ __forceinline__ uint32_t hash_mul(uint32_t key, uint32_t factor) { return key * factor; }
__device__ __forceinline__ uint32_t hash_range(uint32_t key, uint32_t factor, uint32_t n)
{
	return __umulhi(key * factor, n);
}
// digit = bits [rshift, rshift + bits) of x; rshift = 32 - consumed - bits
__device__ __forceinline__ uint32_t radix_digit(uint32_t x, int rshift, uint32_t mask)
{
	return (x >> rshift) & mask;
}

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ uint32_t lanemask_lt()
{
	uint32_t m;
	asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
	return m;
}

// streaming 128-bit load of read-once input columns (keeps L1 for tables / staging)
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4 *p)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
	return r;
}
__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t *p)
{
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
	return r;
}
// L2-coherent loads (bypass L1) for data other CTAs write during the same kernel
__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t *p)
{
	uint64_t r;
	asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p)
{
	uint64_t r;
	asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(r) : "l"(p));
	return r;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v)
{
	asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ uint64_t warp_sum_u64(uint64_t v)
{
#pragma unroll
	for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFullMask, v, o);
	return v;
}
__device__ __forceinline__ uint32_t warp_inclusive_scan_u32(uint32_t v)
{
	const uint32_t lane = lane_id();
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		uint32_t t = __shfl_up_sync(kFullMask, v, o);
		if (lane >= (uint32_t)o) v += t;
	}
	return v;
}

// CTA-wide exclusive scan of one value per thread; `warp_totals` holds >= blockDim.x/32 + 1
// uint32 of shared memory.  Returns the exclusive prefix; *total receives the CTA sum.
// Ends with a __syncthreads(), so warp_totals may be reused right after.
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *warp_totals, uint32_t *total)
{
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
	uint32_t incl = warp_inclusive_scan_u32(v);
	if (lane == 31) warp_totals[warp] = incl;
	__syncthreads();
	if (warp == 0) {
		uint32_t w = lane < nwarps ? warp_totals[lane] : 0;
		uint32_t wi = warp_inclusive_scan_u32(w);
		if (lane < nwarps) warp_totals[lane] = wi - w;
		if (lane == 31) warp_totals[nwarps] = wi;
	}
	__syncthreads();
	uint32_t excl = warp_totals[warp] + incl - v;
	*total = warp_totals[nwarps];
	__syncthreads();
	return excl;
}

// Per-thread running result statistics, folded into four global uint64 at kernel end.
struct JoinSums {
	uint64_t count, key, outer, inner;
	__device__ __forceinline__ void zero() { count = key = outer = inner = 0; }
	__device__ __forceinline__ void add(uint32_t k, uint32_t o, uint32_t i)
	{
		count += 1;
		key += k;
		outer += o;
		inner += i;
	}
	// branch-free form: hit is 0 or 1; one 32x32+64 multiply-add per sum (IMAD.WIDE.U32)
	__device__ __forceinline__ void add_if(uint32_t hit, uint32_t k, uint32_t o, uint32_t i)
	{
		asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(count) : "r"(hit), "r"(1u));
		asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(key) : "r"(k), "r"(hit));
		asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(outer) : "r"(o), "r"(hit));
		asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(inner) : "r"(i), "r"(hit));
	}
	// every thread of the CTA calls this; scratch: 4 * 32 uint64 of shared memory
	__device__ __forceinline__ void reduce_to_global(unsigned long long *g /* [4] */, uint64_t *scratch)
	{
		const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
		uint64_t a = warp_sum_u64(count), b = warp_sum_u64(key), c = warp_sum_u64(outer), d = warp_sum_u64(inner);
		__syncthreads();
		if (lane == 0) {
			scratch[warp * 4 + 0] = a;
			scratch[warp * 4 + 1] = b;
			scratch[warp * 4 + 2] = c;
			scratch[warp * 4 + 3] = d;
		}
		__syncthreads();
		if (threadIdx.x < 4) {
			uint64_t s = 0;
			for (uint32_t w = 0; w < nwarps; ++w) s += scratch[w * 4 + threadIdx.x];
			if (s) atomicAdd(&g[threadIdx.x], (unsigned long long)s);
		}
	}
};

struct OutCols {
	uint32_t *k, *o, *i;          // global result columns
	unsigned long long *cursor;   // global row cursor
	uint64_t cap;                 // rows the columns can hold
};


// largest q in [0, n) with a[q] <= x, for a non-decreasing a[0..n] with a[0] <= x < a[n]
__device__ __forceinline__ uint32_t upper_parent(const uint32_t *a, uint32_t n, uint32_t x)
{
	uint32_t lo = 0, hi = n;
	while (hi - lo > 1) {
		const uint32_t mid = (lo + hi) >> 1;
		if (a[mid] <= x) lo = mid;
		else hi = mid;
	}
	return lo;
}

// warp-collective: at most one match per item and lane, rows of item t of all lanes contiguous.
// Rows that would not fit the result columns are not written at all: the cursor still counts them, the
// host sees count > capacity, grows the columns and runs the join phase again.
template <int ITEMS>
__device__ __forceinline__ void emit_round(const OutCols &out, const bool (&found)[ITEMS], const uint32_t (&key)[ITEMS],
                                           const uint32_t (&val)[ITEMS], const uint32_t (&ival)[ITEMS])
{
	uint32_t total = 0;                            // the ballots are recomputed below instead of kept in registers
#pragma unroll
	for (int t = 0; t < ITEMS; ++t) total += __popc(__ballot_sync(kFullMask, found[t]));
	if (total == 0) return;
	unsigned long long base = 0;
	if (lane_id() == 0) base = atomicAdd(out.cursor, (unsigned long long)total);
	base = __shfl_sync(kFullMask, base, 0);
	if (base + total > out.cap) return;
	const unsigned lt = lanemask_lt();
	uint32_t *const ck = out.k + base, *const co = out.o + base, *const ci = out.i + base;
	uint32_t off = 0;
#pragma unroll
	for (int t = 0; t < ITEMS; ++t) {
		const unsigned mt = __ballot_sync(kFullMask, found[t]);
		const uint32_t r = off + __popc(mt & lt);
		if (found[t]) {
			ck[r] = key[t];
			co[r] = val[t];
			ci[r] = ival[t];
		}
		off += __popc(mt);
	}
}

// The same, CTA-collective: EVERY thread of the CTA calls it once per round (warps without a match too), and the
// CTA reserves the rows of all its warps with ONE global atomicAdd.  Same-address atomics are served one after the
// other by the L2 (measured on B200: ~1 ns each, i.e. a reservation per warp round of 128 tuples caps a join at
// ~120 G probe tuples/s whatever else it does); per CTA round of THREADS x ITEMS tuples the cap is 8x higher.
// stage: 2 x (nwarps + 2) uint32 of shared memory, `round` alternates its halves so that two barriers per call suffice.
template <int ITEMS>
__device__ __forceinline__ void emit_round_cta(const OutCols &out, uint32_t *stage, uint32_t round, const bool (&found)[ITEMS],
                                               const uint32_t (&key)[ITEMS], const uint32_t (&val)[ITEMS],
                                               const uint32_t (&ival)[ITEMS])
{
	const uint32_t lane = lane_id(), warp = threadIdx.x >> 5, nwarps = (blockDim.x + 31) >> 5;
	uint32_t *wt = stage + (round & 1u) * (nwarps + 2);
	unsigned long long *basep = reinterpret_cast<unsigned long long *>(stage + 2 * (nwarps + 2)) + (round & 1u);
	uint32_t total = 0;
#pragma unroll
	for (int t = 0; t < ITEMS; ++t) total += __popc(__ballot_sync(kFullMask, found[t]));
	if (lane == 0) wt[warp] = total;
	__syncthreads();
	const uint32_t mine = lane < nwarps ? wt[lane] : 0u;
	const uint32_t incl = warp_inclusive_scan_u32(mine);
	const uint32_t cta_total = __shfl_sync(kFullMask, incl, 31);
	const uint32_t before = __shfl_sync(kFullMask, incl - mine, warp);
	if (cta_total == 0) return;                    // uniform over the CTA
	if (threadIdx.x == 0) *basep = atomicAdd(out.cursor, (unsigned long long)cta_total);
	__syncthreads();
	const unsigned long long cta_base = *basep;
	if (cta_base + cta_total > out.cap || total == 0) return;
	const unsigned lt = lanemask_lt();
	const unsigned long long base = cta_base + before;
	uint32_t *const ck = out.k + base, *const co = out.o + base, *const ci = out.i + base;
	uint32_t off = 0;
#pragma unroll
	for (int t = 0; t < ITEMS; ++t) {
		const unsigned mt = __ballot_sync(kFullMask, found[t]);
		const uint32_t r = off + __popc(mt & lt);
		if (found[t]) {
			ck[r] = key[t];
			co[r] = val[t];
			ci[r] = ival[t];
		}
		off += __popc(mt);
	}
}

}  // namespace hjb
