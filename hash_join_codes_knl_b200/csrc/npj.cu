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
// Measured and dropped (round 2, profiles/README.md): building and probing a table larger than L2 in phases, one
// L2-sized slice of the table at a time with evict-last / evict-first cache hints -- every phase re-reads the whole
// relation and the probes of a slice were no faster than before (the table's sectors did not stay resident beside
// the streamed columns): 11.3-17.3 ms against 5.9 ms for config 1's probe.
#include "hj_device.cuh"
#include "hj_internal.h"
#include <stdlib.h>

namespace hjb {

struct Bucket {
	uint64_t s[4];
};

// one whole bucket, read-only path (probe)
__device__ __forceinline__ Bucket ld_bucket_nc(const uint64_t *table, uint32_t b)
{
	Bucket r;
	asm volatile("ld.global.nc.v4.u64 {%0, %1, %2, %3}, [%4];"
	             : "=l"(r.s[0]), "=l"(r.s[1]), "=l"(r.s[2]), "=l"(r.s[3])
	             : "l"(table + (uint64_t)b * 4));
	return r;
}
// the same as two 128-bit loads through L1: faster when the table is cache-resident (config 3: 9.1 against 10.7 ms --
// the second half of the sector is an L1 hit, and the 256-bit form is served past L1)
__device__ __forceinline__ Bucket ld_bucket_halves(const uint64_t *table, uint32_t b)
{
	const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(table + (uint64_t)b * 4);
	const ulonglong2 lo = __ldg(bp), hi = __ldg(bp + 1);
	Bucket r;
	r.s[0] = lo.x; r.s[1] = lo.y; r.s[2] = hi.x; r.s[3] = hi.y;
	return r;
}
// one whole bucket, L2-coherent (build: other CTAs insert into the same table)
__device__ __forceinline__ Bucket ld_bucket_cg(const uint64_t *table, uint32_t b)
{
	Bucket r;
	asm volatile("ld.global.cg.v4.u64 {%0, %1, %2, %3}, [%4];"
	             : "=l"(r.s[0]), "=l"(r.s[1]), "=l"(r.s[2]), "=l"(r.s[3])
	             : "l"(table + (uint64_t)b * 4)
	             : "memory");
	return r;
}
// Build.  Every thread takes four tuples per round and fetches the home buckets of all of them before it
// inserts any (the inserts are bound by the latency of random sector reads).
__global__ void __launch_bounds__(kNpjThreads)
k_npj_build(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
            uint64_t *__restrict__ table, uint32_t buckets, uint32_t factor,
            unsigned long long *__restrict__ flags /* [0] sentinel pairs, [1] duplicate build keys seen */)
{
	const uint64_t groups = (n + 3) >> 2;
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t k[4], v[4], b[4];
		bool mine[4];
		const uint64_t idx = g << 2;
		if (idx + 3 < n) {
			const uint4 kk = ldg_stream_u4(reinterpret_cast<const uint4 *>(keys) + g);
			k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e) k[e] = idx + e < n ? keys[idx + e] : 0;
		}
		bool any = false;
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			b[e] = hash_range(k[e], factor, buckets);
			mine[e] = idx + e < n;
			any |= mine[e];
		}
		if (!any) continue;
		if (idx + 3 < n) {
			const uint4 vv = ldg_stream_u4(reinterpret_cast<const uint4 *>(vals) + g);
			v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e) v[e] = mine[e] ? vals[idx + e] : 0;
		}
		Bucket home[4];
#pragma unroll
		for (int e = 0; e < 4; ++e)
			if (mine[e]) home[e] = ld_bucket_cg(table, b[e]);
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
					cur = ld_bucket_cg(table, bb);
					continue;
				}
				done = atomicCAS(reinterpret_cast<unsigned long long *>(&table[(uint64_t)bb * 4 + free_slot]),
				                 (unsigned long long)kEmptySlot, (unsigned long long)pair) == kEmptySlot;
				if (!done) cur = ld_bucket_cg(table, bb);                     // lost the race for that slot: look at the same bucket again
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
// Unique build keys (flags[1] == 0, detected by the build): a lane stops at its first match, the
// warp reserves rows with one atomicAdd per round and stores them ballot-ranked from registers.
// Otherwise every match is emitted as it is met.
template <bool MATERIALIZE, bool WIDE>
__global__ void __launch_bounds__(kNpjThreads, 4)
k_npj_probe(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
            const uint64_t *__restrict__ table, uint32_t buckets, uint32_t factor, OutCols out,
            unsigned long long *__restrict__ sums, const unsigned long long *__restrict__ flags)
{
	__shared__ uint64_t scratch[4 * 32];
	JoinSums acc;
	acc.zero();
	const uint32_t sentinels = (uint32_t)flags[0];
	const bool slow = flags[1] != 0 || sentinels != 0;
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
			k[t] = found[t] ? ldg_stream_u32(&keys[i]) : 0;
			v[t] = found[t] ? ldg_stream_u32(&vals[i]) : 0;
			ival[t] = 0;
		}
#pragma unroll
		for (int t = 0; t < kNpjItems; ++t) {
			b[t] = hash_range(k[t], factor, buckets);
			home[t] = WIDE ? ld_bucket_nc(table, b[t]) : ld_bucket_halves(table, b[t]);   // also for lanes past the end: bucket of key 0, never used
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
							const Bucket q = WIDE ? ld_bucket_nc(table, bb) : ld_bucket_halves(table, bb);
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
			if (MATERIALIZE) emit_round<kNpjItems>(out, found, k, v, ival);
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
					cur = WIDE ? ld_bucket_nc(table, bb) : ld_bucket_halves(table, bb);
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
	k_npj_build<<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.rk, a.rv, a.nr, a.table, (uint32_t)a.buckets, a.factor, a.scalars + 5);
	t->stop(s);
	return 1;
}

int launch_npj_probe(const NpjArgs &a, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	int per_sm = 0;
	// 256-bit bucket loads for tables that live in DRAM, two 128-bit loads through L1 for cache-resident ones
	const bool wide = a.buckets * 32 > (96ull << 20);
	if (a.materialize) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_npj_probe<true, true>, kNpjThreads, 0);
	else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_npj_probe<false, true>, kNpjThreads, 0);
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
	t->start(KK_NPJ_PROBE, s);
	auto launch = [&](auto kernel) {
		kernel<<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.sk, a.sv, a.ns, a.table, (uint32_t)a.buckets, a.factor, out, a.scalars + 1, a.scalars + 5);
	};
	if (a.materialize && wide) launch(k_npj_probe<true, true>);
	else if (a.materialize) launch(k_npj_probe<true, false>);
	else if (wide) launch(k_npj_probe<false, true>);
	else launch(k_npj_probe<false, false>);
	t->stop(s);
	return 1;
}

}  // namespace hjb
