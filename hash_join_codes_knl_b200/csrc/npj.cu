// npj.cu -- non-partitioned hash join: one global open-addressing table in HBM
// (reference: set / build / probe of npj.cpp:366-380, 190-212, 216-364).
//
// Table layout: `buckets` x 4 slots, slot = payload<<32 | key, all-ones = empty.  A bucket is
// one 32-byte DRAM sector, fetched with ONE 256-bit load (LDG.E.256 on sm_100a), so a probe costs one
// sector and one L1 wavefront however many of its four slots are in use; buckets chain linearly
// (bucket b full -> b+1).  Build claims slots with a 64-bit atomicCAS as the reference does
// (npj.cpp:206); probe walks buckets until it sees an empty slot and emits every equal key on the way
// (all duplicates, npj.cpp:288-290).
//
// A table larger than L2 is built and probed in PHASES: phase p handles only the tuples whose home
// bucket lies in the p-th slice of the table, a slice small enough to stay L2-resident (its loads
// carry an evict-last policy, the streamed columns and the result rows evict-first).  The relations
// are read once per phase -- sequential HBM reads at full bandwidth -- instead of every probe
// fetching its sector from DRAM (measured in round 1: 84 bytes of DRAM traffic per probe, 3.4x the
// algorithmic bytes).
#include "hj_device.cuh"
#include "hj_internal.h"
#include <stdlib.h>

namespace hjb {

struct Bucket {
	uint64_t s[4];
};

// L2 eviction policies: the table slice of the current phase should stay, the streamed columns should not
// (hints == 0: no preference either way, for A/B runs)
__device__ __forceinline__ uint64_t policy_evict_last(int hints)
{
	uint64_t p;
	if (hints) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ uint64_t policy_evict_first(int hints)
{
	uint64_t p;
	if (hints) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
	else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
	return p;
}
// one whole bucket, read-only path (probe)
__device__ __forceinline__ Bucket ld_bucket_nc(const uint64_t *table, uint32_t b, uint64_t policy)
{
	Bucket r;
	asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;"
	             : "=l"(r.s[0]), "=l"(r.s[1]), "=l"(r.s[2]), "=l"(r.s[3])
	             : "l"(table + (uint64_t)b * 4), "l"(policy));
	return r;
}
// one whole bucket, L2-coherent (build: other CTAs insert into the same table)
__device__ __forceinline__ Bucket ld_bucket_cg(const uint64_t *table, uint32_t b, uint64_t policy)
{
	Bucket r;
	asm volatile("ld.global.cg.L2::cache_hint.v4.u64 {%0, %1, %2, %3}, [%4], %5;"
	             : "=l"(r.s[0]), "=l"(r.s[1]), "=l"(r.s[2]), "=l"(r.s[3])
	             : "l"(table + (uint64_t)b * 4), "l"(policy)
	             : "memory");
	return r;
}
__device__ __forceinline__ uint32_t ldg_first_u32(const uint32_t *p, uint64_t policy)
{
	uint32_t r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(policy));
	return r;
}
__device__ __forceinline__ uint4 ldg_first_u4(const uint4 *p, uint64_t policy)
{
	uint4 r;
	asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0, %1, %2, %3}, [%4], %5;"
	             : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(policy));
	return r;
}

// Build.  Every thread takes four tuples per round and fetches the home buckets of all of them before it
// inserts any (the inserts are bound by the latency of random sector reads).  Only tuples whose home bucket
// lies in [b_lo, b_hi) are inserted by this launch (one phase; [0, buckets) = everything).
__global__ void __launch_bounds__(kNpjThreads)
k_npj_build(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
            uint64_t *__restrict__ table, uint32_t buckets, uint32_t factor, uint32_t b_lo, uint32_t b_hi, int hints,
            unsigned long long *__restrict__ flags /* [0] sentinel pairs, [1] duplicate build keys seen */)
{
	const uint64_t groups = (n + 3) >> 2;
	const uint64_t keep = policy_evict_last(hints), pass = policy_evict_first(hints);
	const bool phased = b_lo != 0 || b_hi != buckets;
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t k[4], v[4], b[4];
		bool mine[4];
		const uint64_t idx = g << 2;
		if (idx + 3 < n) {
			const uint4 kk = ldg_first_u4(reinterpret_cast<const uint4 *>(keys) + g, pass);
			k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e) k[e] = idx + e < n ? keys[idx + e] : 0;
		}
		bool any = false;
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			b[e] = hash_range(k[e], factor, buckets);
			mine[e] = idx + e < n && b[e] >= b_lo && b[e] < b_hi;
			any |= mine[e];
		}
		if (!any) continue;
		if (idx + 3 < n && !phased) {
			const uint4 vv = ldg_first_u4(reinterpret_cast<const uint4 *>(vals) + g, pass);
			v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e) v[e] = mine[e] ? vals[idx + e] : 0;
		}
		Bucket home[4];
#pragma unroll
		for (int e = 0; e < 4; ++e)
			if (mine[e]) home[e] = ld_bucket_cg(table, b[e], keep);
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			if (!mine[e]) continue;
			const uint64_t pair = ((uint64_t)v[e] << 32) | k[e];
			if (pair == kEmptySlot) {
				atomicAdd(&flags[0], 1ull);
				continue;
			}
			uint32_t bb = b[e];
			Bucket cur = home[e];
			bool done = false;
			while (!done) {
				int free_slot = -1;
#pragma unroll
				for (int z = 3; z >= 0; --z) {
					if (cur.s[z] == kEmptySlot) free_slot = z;                      // lowest free slot
					else if ((uint32_t)cur.s[z] == k[e]) flags[1] = 1;              // equal build keys: probes walk whole chains
				}
				if (free_slot < 0) {
					bb = bb + 1 == buckets ? 0 : bb + 1;
					cur = ld_bucket_cg(table, bb, keep);
					continue;
				}
				done = atomicCAS(reinterpret_cast<unsigned long long *>(&table[(uint64_t)bb * 4 + free_slot]),
				                 (unsigned long long)kEmptySlot, (unsigned long long)pair) == kEmptySlot;
				if (!done) cur = ld_bucket_cg(table, bb, keep);                     // lost the race for that slot: look at the same bucket again
			}
		}
	}
}

// one row, reservation aggregated over whichever lanes of the warp are here together
__device__ __forceinline__ void npj_emit_row(const OutCols &out, uint32_t key, uint32_t oval, uint32_t ival)
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

// Probe.  Each thread takes kNpjItems probe tuples per round (a warp owns 32*kNpjItems consecutive
// tuples, item t of lane l = tuple 32*t + l, so every load and every result store is a coalesced
// 128-byte run) and fetches the home buckets of all of them before looking at any: the probe is
// bound by the latency of random 32-byte sector reads, so loads in flight are what counts.
// Only tuples whose home bucket lies in [b_lo, b_hi) are probed by this launch (one phase).
// Unique build keys (flags[1] == 0, detected by the build): a lane stops at its first match, the
// warp reserves rows with one atomicAdd per round and stores them ballot-ranked from registers.
// Otherwise every match is emitted as it is met.
template <bool MATERIALIZE, bool CTAEMIT>
__global__ void __launch_bounds__(kNpjThreads, 3)
k_npj_probe(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
            const uint64_t *__restrict__ table, uint32_t buckets, uint32_t factor, uint32_t b_lo, uint32_t b_hi, int hints, OutCols out,
            unsigned long long *__restrict__ sums, const unsigned long long *__restrict__ flags)
{
	__shared__ uint64_t scratch[4 * 32];
	__shared__ __align__(8) uint32_t s_emit[2 * (kNpjThreads / 32 + 2) + 4];
	uint32_t emit_rounds = 0;
	JoinSums acc;
	acc.zero();
	const uint32_t sentinels = (uint32_t)flags[0];
	const bool slow = flags[1] != 0 || sentinels != 0;
	const uint64_t keep = policy_evict_last(hints), pass = policy_evict_first(hints);
	constexpr uint32_t kRound = kNpjThreads * kNpjItems;
	const uint64_t rounds = (n + kRound - 1) / kRound;
	for (uint64_t rd = blockIdx.x; rd < rounds; rd += gridDim.x) {
		const uint64_t wbase = rd * kRound + (threadIdx.x & ~31u) * kNpjItems + lane_id();
		uint32_t k[kNpjItems], v[kNpjItems], ival[kNpjItems], b[kNpjItems];
		bool found[kNpjItems];
		Bucket home[kNpjItems];
#pragma unroll
		for (int t = 0; t < kNpjItems; ++t) {
			const uint64_t i = wbase + (uint64_t)t * 32;
			found[t] = i < n;                               // "valid" until probed
			k[t] = found[t] ? ldg_first_u32(&keys[i], pass) : 0;
		}
#pragma unroll
		for (int t = 0; t < kNpjItems; ++t) {
			b[t] = hash_range(k[t], factor, buckets);
			found[t] = found[t] && b[t] >= b_lo && b[t] < b_hi;
			if (found[t]) home[t] = ld_bucket_nc(table, b[t], keep);
		}
#pragma unroll
		for (int t = 0; t < kNpjItems; ++t) {
			const uint64_t i = wbase + (uint64_t)t * 32;
			v[t] = found[t] ? ldg_first_u32(&vals[i], pass) : 0;
			ival[t] = 0;
		}
		if (!slow) {
			// home bucket without branches: four key compares, the payload picked by selects.  Only a full
			// bucket without the key (the chain goes on) and the all-ones probe key (its low word equals an
			// empty slot's) take the loop below.
#pragma unroll
			for (int t = 0; t < kNpjItems; ++t) {
				bool hit = false;
				if (found[t]) {
					const uint64_t s0 = home[t].s[0], s1 = home[t].s[1], s2 = home[t].s[2], s3 = home[t].s[3];
					const bool m0 = (uint32_t)s0 == k[t], m1 = (uint32_t)s1 == k[t], m2 = (uint32_t)s2 == k[t], m3 = (uint32_t)s3 == k[t];
					ival[t] = m0 ? (uint32_t)(s0 >> 32) : m1 ? (uint32_t)(s1 >> 32) : m2 ? (uint32_t)(s2 >> 32) : (uint32_t)(s3 >> 32);
					hit = m0 || m1 || m2 || m3;
					const bool special = k[t] == 0xFFFFFFFFu;
					if (special || (!hit && s3 != kEmptySlot)) {
						hit = false;
						uint32_t bb = b[t];
						if (!special) bb = bb + 1 == buckets ? 0 : bb + 1;
						while (true) {
							const Bucket q = ld_bucket_nc(table, bb, keep);
#pragma unroll
							for (int z = 0; z < 4; ++z)
								if (!hit && (uint32_t)q.s[z] == k[t] && q.s[z] != kEmptySlot) {
									ival[t] = (uint32_t)(q.s[z] >> 32);
									hit = true;
								}
							if (hit || q.s[3] == kEmptySlot) break;           // slots fill lowest-first: a free last slot ends the chain
							bb = bb + 1 == buckets ? 0 : bb + 1;
						}
					}
				}
				found[t] = hit;
				acc.add_if(hit ? 1u : 0u, k[t], v[t], ival[t]);
			}
			if (MATERIALIZE) {
				if (CTAEMIT) emit_round_cta<kNpjItems>(out, s_emit, emit_rounds++, found, k, v, ival);
				else emit_round<kNpjItems>(out, found, k, v, ival);
			}
		} else {
#pragma unroll
			for (int t = 0; t < kNpjItems; ++t) {
				if (!found[t]) continue;
				uint32_t bb = b[t];
				Bucket cur = home[t];
				while (true) {
					bool full = true;
#pragma unroll
					for (int z = 0; z < 4; ++z) {
						if (cur.s[z] == kEmptySlot) {
							full = false;
						} else if ((uint32_t)cur.s[z] == k[t]) {
							acc.add(k[t], v[t], (uint32_t)(cur.s[z] >> 32));
							if (MATERIALIZE) npj_emit_row(out, k[t], v[t], (uint32_t)(cur.s[z] >> 32));
						}
					}
					if (!full) break;                              // an empty slot ends the chain
					bb = bb + 1 == buckets ? 0 : bb + 1;
					cur = ld_bucket_nc(table, bb, keep);
				}
				if (k[t] == kSentinelKey)
					for (uint32_t c = 0; c < sentinels; ++c) {
						acc.add(k[t], v[t], kSentinelKey);
						if (MATERIALIZE) npj_emit_row(out, k[t], v[t], kSentinelKey);
					}
			}
		}
	}
	acc.reduce_to_global(sums, scratch);
}

// phases of a table of `buckets` buckets: slices of at most HJB_NPJ_PHASE_MB (default 48) megabytes
uint32_t npj_phases(uint64_t buckets)
{
	static const long long mb = getenv("HJB_NPJ_PHASE_MB") ? atoll(getenv("HJB_NPJ_PHASE_MB")) : 0;
	if (mb <= 0) return 1;
	const uint64_t slice = (uint64_t)mb << 20;
	const uint64_t bytes = buckets * 32;
	uint64_t p = (bytes + slice - 1) / slice;
	if (bytes <= (96ull << 20)) p = 1;                    // fits L2 as a whole
	if (p > 64) p = 64;
	return (uint32_t)(p ? p : 1);
}

static int npj_hints()
{
	static const int h = getenv("HJB_NPJ_HINTS") ? atoi(getenv("HJB_NPJ_HINTS")) : 1;
	return h;
}

static void phase_range(uint64_t buckets, uint32_t phases, uint32_t p, uint32_t *lo, uint32_t *hi)
{
	*lo = (uint32_t)(buckets * p / phases);
	*hi = (uint32_t)(buckets * (p + 1) / phases);
}

int launch_npj_build(const NpjArgs &a, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	t->start(KK_NPJ_BUILD, s);
	cudaMemsetAsync(a.table, 0xFF, (size_t)a.buckets * 32, s);               // set(), npj.cpp:366-380
	const uint64_t groups = (a.nr + 3) / 4;
	uint64_t grid = (groups + kNpjThreads - 1) / kNpjThreads;
	if (grid > (uint64_t)sms * 16) grid = (uint64_t)sms * 16;
	if (grid == 0) grid = 1;
	const uint32_t phases = a.phases ? a.phases : 1;
	for (uint32_t p = 0; p < phases; ++p) {
		uint32_t lo, hi;
		phase_range(a.buckets, phases, p, &lo, &hi);
		k_npj_build<<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.rk, a.rv, a.nr, a.table, (uint32_t)a.buckets, a.factor, lo, hi,
		                                                   npj_hints(), a.scalars + 5);
	}
	t->stop(s);
	return (int)phases;
}

int launch_npj_probe(const NpjArgs &a, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	int per_sm = 0;
	static const int cta_emit = getenv("HJB_NPJ_CTA_EMIT") ? atoi(getenv("HJB_NPJ_CTA_EMIT")) : 0;
	if (a.materialize) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_npj_probe<true, true>, kNpjThreads, 0);
	else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_npj_probe<false, false>, kNpjThreads, 0);
	if (per_sm < 1) per_sm = 1;
	const uint64_t rounds = (a.ns + kNpjThreads * kNpjItems - 1) / (kNpjThreads * kNpjItems);
	uint64_t grid = rounds;
	if (grid > (uint64_t)sms * per_sm) grid = (uint64_t)sms * per_sm;
	if (grid == 0) grid = 1;
	OutCols out;
	out.k = a.out_k;
	out.o = a.out_o;
	out.i = a.out_i;
	out.cursor = a.scalars;
	out.cap = a.materialize ? a.out_cap : 0;
	const uint32_t phases = a.phases ? a.phases : 1;
	t->start(KK_NPJ_PROBE, s);
	for (uint32_t p = 0; p < phases; ++p) {
		uint32_t lo, hi;
		phase_range(a.buckets, phases, p, &lo, &hi);
		if (a.materialize && cta_emit)
			k_npj_probe<true, true><<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.sk, a.sv, a.ns, a.table, (uint32_t)a.buckets, a.factor, lo, hi,
			                                                               npj_hints(), out, a.scalars + 1, a.scalars + 5);
		else if (a.materialize)
			k_npj_probe<true, false><<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.sk, a.sv, a.ns, a.table, (uint32_t)a.buckets, a.factor, lo, hi,
			                                                                npj_hints(), out, a.scalars + 1, a.scalars + 5);
		else
			k_npj_probe<false, false><<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.sk, a.sv, a.ns, a.table, (uint32_t)a.buckets, a.factor, lo, hi,
			                                                                 npj_hints(), out, a.scalars + 1, a.scalars + 5);
	}
	t->stop(s);
	return (int)phases;
}

}  // namespace hjb
