// npj.cu -- non-partitioned hash join: one global open-addressing table in HBM
// (reference: set / build / probe of npj.cpp:366-380, 190-212, 216-364).
//
// Table layout: `buckets` x 4 slots, slot = payload<<32 | key, all-ones = empty.  A bucket is
// one 32-byte DRAM sector, so a probe costs one sector fetch however many of its four slots
// are in use; buckets chain linearly (bucket b full -> b+1).  Build claims slots with a 64-bit
// atomicCAS as the reference does (npj.cpp:206); probe walks buckets until it sees an empty
// slot and emits every equal key on the way (all duplicates, npj.cpp:288-290).
#include "hj_device.cuh"
#include "hj_internal.h"

namespace hjb {

__global__ void __launch_bounds__(kNpjThreads)
k_npj_build(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
            uint64_t *__restrict__ table, uint32_t buckets, uint32_t factor,
            unsigned long long *__restrict__ sentinel_count)
{
	const uint64_t groups = (n + 3) >> 2;
	for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < groups; g += (uint64_t)gridDim.x * blockDim.x) {
		uint32_t k[4], v[4];
		const uint64_t idx = g << 2;
		if (idx + 3 < n) {
			const uint4 kk = ldg_stream_u4(reinterpret_cast<const uint4 *>(keys) + g);
			const uint4 vv = ldg_stream_u4(reinterpret_cast<const uint4 *>(vals) + g);
			k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
			v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
		} else {
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				k[e] = idx + e < n ? keys[idx + e] : 0;
				v[e] = idx + e < n ? vals[idx + e] : 0;
			}
		}
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			if (idx + e >= n) continue;
			const uint64_t pair = ((uint64_t)v[e] << 32) | k[e];
			if (pair == kEmptySlot) {
				atomicAdd(sentinel_count, 1ull);
				continue;
			}
			uint32_t b = hash_range(k[e], factor, buckets);
			bool done = false;
			while (!done) {
				uint64_t *slots = table + (uint64_t)b * 4;
				int free_slot = -1;
#pragma unroll
				for (int z = 3; z >= 0; --z)
					if (ld_cg_u64(&slots[z]) == kEmptySlot) free_slot = z;      // lowest free slot
				if (free_slot < 0) {
					b = b + 1 == buckets ? 0 : b + 1;
					continue;
				}
				done = atomicCAS(reinterpret_cast<unsigned long long *>(&slots[free_slot]),
				                 (unsigned long long)kEmptySlot, (unsigned long long)pair) == kEmptySlot;
				// lost the race for that slot: look at the same bucket again
			}
		}
	}
}

// dynamic shared memory: stage k,o,i [kStageCap] | scratch 4*32 uint64
template <bool MATERIALIZE>
__global__ void __launch_bounds__(kNpjThreads)
k_npj_probe(const uint32_t *__restrict__ keys, const uint32_t *__restrict__ vals, uint64_t n,
            const uint64_t *__restrict__ table, uint32_t buckets, uint32_t factor, OutCols out,
            unsigned long long *__restrict__ sums, const unsigned long long *__restrict__ sentinel_count)
{
	extern __shared__ __align__(16) unsigned char s_raw[];
	uint32_t *stage_mem = reinterpret_cast<uint32_t *>(s_raw);
	uint64_t *scratch = reinterpret_cast<uint64_t *>(stage_mem + 3 * kStageCap);
	__shared__ uint32_t s_cnt;
	__shared__ unsigned long long s_base;
	MatchStage st;
	st.k = stage_mem;
	st.o = stage_mem + kStageCap;
	st.i = stage_mem + 2 * kStageCap;
	st.cnt = &s_cnt;
	st.cap = kStageCap;
	if (threadIdx.x == 0) s_cnt = 0;
	__syncthreads();
	JoinSums acc;
	acc.zero();
	const uint32_t sentinels = (uint32_t)*sentinel_count;
	const uint64_t groups = (n + 3) >> 2;
	// one round = one absolutely aligned group of four probe tuples per thread
	for (uint64_t g0 = (uint64_t)blockIdx.x * kNpjThreads; g0 < groups; g0 += (uint64_t)gridDim.x * kNpjThreads) {
		const uint64_t g = g0 + threadIdx.x, idx = g << 2;
		uint32_t k[4] = {0, 0, 0, 0}, v[4] = {0, 0, 0, 0};
		if (g < groups) {
			if (idx + 3 < n) {
				const uint4 kk = ldg_stream_u4(reinterpret_cast<const uint4 *>(keys) + g);
				const uint4 vv = ldg_stream_u4(reinterpret_cast<const uint4 *>(vals) + g);
				k[0] = kk.x; k[1] = kk.y; k[2] = kk.z; k[3] = kk.w;
				v[0] = vv.x; v[1] = vv.y; v[2] = vv.z; v[3] = vv.w;
			} else {
#pragma unroll
				for (int e = 0; e < 4; ++e) {
					k[e] = idx + e < n ? keys[idx + e] : 0;
					v[e] = idx + e < n ? vals[idx + e] : 0;
				}
			}
		}
		// first bucket of all four tuples in flight together: the probe is latency bound
		uint32_t b[4];
		ulonglong2 lo[4], hi[4];
#pragma unroll
		for (int e = 0; e < 4; ++e) {
			b[e] = hash_range(k[e], factor, buckets);
			const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(table + (uint64_t)b[e] * 4);
			lo[e] = __ldg(bp);
			hi[e] = __ldg(bp + 1);
		}
		for (int mode = 0; mode < 2; ++mode) {         // 0: staged; 1: direct, only after a stage overflow
#pragma unroll
			for (int e = 0; e < 4; ++e) {
				bool active = g < groups && idx + e < n;
				uint32_t bb = b[e];
				uint64_t s0 = lo[e].x, s1 = lo[e].y, s2 = hi[e].x, s3 = hi[e].y;
				while (__any_sync(kFullMask, active)) {
					const uint64_t slot[4] = {s0, s1, s2, s3};
					bool full = true;
#pragma unroll
					for (int z = 0; z < 4; ++z) {
						const bool used = slot[z] != kEmptySlot;
						full = full && used;
						const bool hit = active && used && (uint32_t)slot[z] == k[e];
						const uint32_t ival = (uint32_t)(slot[z] >> 32);
						if (mode == 0) {
							if (hit) acc.add(k[e], v[e], ival);
							if (MATERIALIZE) st.emit(hit, k[e], v[e], ival);
						} else {
							emit_direct(out, hit, k[e], v[e], ival);
						}
					}
					active = active && full;               // an empty slot ends the chain
					if (active) {
						bb = bb + 1 == buckets ? 0 : bb + 1;
						const ulonglong2 *bp = reinterpret_cast<const ulonglong2 *>(table + (uint64_t)bb * 4);
						const ulonglong2 a = __ldg(bp), c = __ldg(bp + 1);
						s0 = a.x; s1 = a.y; s2 = c.x; s3 = c.y;
					}
				}
				if (sentinels && g < groups && idx + e < n && k[e] == kSentinelKey) {
					for (uint32_t c = 0; c < sentinels; ++c) {
						if (mode == 0) {
							acc.add(k[e], v[e], kSentinelKey);
							if (MATERIALIZE) st.emit_one(k[e], v[e], kSentinelKey);
						} else {
							const unsigned long long r = atomicAdd(out.cursor, 1ull);
							if (r < out.cap) {
								out.k[r] = k[e];
								out.o[r] = v[e];
								out.i[r] = kSentinelKey;
							}
						}
					}
				}
			}
			if (!MATERIALIZE || mode == 1) break;
			if (stage_flush(st, out, &s_base)) break;
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
	k_npj_build<<<(uint32_t)grid, kNpjThreads, 0, s>>>(a.rk, a.rv, a.nr, a.table, (uint32_t)a.buckets, a.factor,
	                                                   a.scalars + 5);
	t->stop(s);
	return 1;
}

int launch_npj_probe(const NpjArgs &a, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	const size_t smem = (size_t)3 * kStageCap * 4 + 4 * 32 * 8;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(k_npj_probe<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		cudaFuncSetAttribute(k_npj_probe<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		attr_set = true;
	}
	int per_sm = 0;
	if (a.materialize) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_npj_probe<true>, kNpjThreads, smem);
	else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_npj_probe<false>, kNpjThreads, smem);
	if (per_sm < 1) per_sm = 1;
	const uint64_t groups = (a.ns + 3) / 4;
	uint64_t grid = (groups + kNpjThreads - 1) / kNpjThreads;
	if (grid > (uint64_t)sms * per_sm) grid = (uint64_t)sms * per_sm;
	if (grid == 0) grid = 1;
	OutCols out;
	out.k = a.out_k;
	out.o = a.out_o;
	out.i = a.out_i;
	out.cursor = a.scalars;
	out.cap = a.materialize ? a.out_cap : 0;
	t->start(KK_NPJ_PROBE, s);
	if (a.materialize)
		k_npj_probe<true><<<(uint32_t)grid, kNpjThreads, smem, s>>>(a.sk, a.sv, a.ns, a.table, (uint32_t)a.buckets,
		                                                            a.factor, out, a.scalars + 1, a.scalars + 5);
	else
		k_npj_probe<false><<<(uint32_t)grid, kNpjThreads, smem, s>>>(a.sk, a.sv, a.ns, a.table, (uint32_t)a.buckets,
		                                                             a.factor, out, a.scalars + 1, a.scalars + 5);
	t->stop(s);
	return 1;
}

}  // namespace hjb
