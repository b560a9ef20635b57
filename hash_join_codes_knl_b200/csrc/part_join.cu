// part_join.cu -- per-partition build + probe with the partition's hash table resident in
// shared memory (reference: the cache-resident join loop, phj.cpp:1880-1923 / cpra2.cpp:1907-1969,
// build phj.cpp:307-397, probe phj.cpp:399-571).
//
// Persistent CTAs pull tasks from an atomic counter.  A task = (partition p, one slice of at
// most s_task probe tuples of p): uniform inputs give one task per partition, a skewed probe
// side (Zipf) is cut into many tasks that each rebuild p's small table, so no CTA is stuck
// with a hot partition.  A build partition larger than the table is joined in several fills
// (block nested loop), so any input is handled -- duplicates included.
#include "hj_device.cuh"
#include "hj_internal.h"

namespace hjb {

// tasks per partition -> exclusive prefix; P <= 2^22, one CTA
__global__ void __launch_bounds__(1024)
k_join_tasks(const uint32_t *__restrict__ r_off, const uint32_t *__restrict__ s_off, uint32_t P, uint32_t s_task,
             uint32_t *__restrict__ task_prefix)
{
	__shared__ uint32_t warp_totals[34];
	const uint32_t per = (P + blockDim.x - 1) / blockDim.x;
	const uint32_t p0 = threadIdx.x * per;
	auto tasks_of = [&](uint32_t p) -> uint32_t {
		const uint32_t rc = r_off[p + 1] - r_off[p], sc = s_off[p + 1] - s_off[p];
		return (rc && sc) ? (sc + s_task - 1) / s_task : 0u;     // an empty side cannot match
	};
	uint32_t local = 0;
	for (uint32_t p = p0; p < p0 + per && p < P; ++p) local += tasks_of(p);
	uint32_t total;
	uint32_t run = block_exclusive_scan(local, warp_totals, &total);
	for (uint32_t p = p0; p < p0 + per && p < P; ++p) {
		task_prefix[p] = run;
		run += tasks_of(p);
	}
	if (threadIdx.x == 0) task_prefix[P] = total;
}

// dynamic shared memory: table[kJoinSlots] (uint64) | stage k,o,i [kStageCap] | scratch
template <bool MATERIALIZE>
__global__ void __launch_bounds__(kJoinThreads)
k_partition_join(const uint32_t *__restrict__ rk, const uint32_t *__restrict__ rv,
                 const uint32_t *__restrict__ sk, const uint32_t *__restrict__ sv,
                 const uint32_t *__restrict__ r_off, const uint32_t *__restrict__ s_off, uint32_t P,
                 const uint32_t *__restrict__ task_prefix, uint32_t *__restrict__ task_counter, uint32_t s_task,
                 uint32_t table_factor, OutCols out, unsigned long long *__restrict__ sums)
{
	extern __shared__ __align__(16) unsigned char s_raw[];
	uint64_t *table = reinterpret_cast<uint64_t *>(s_raw);
	uint32_t *stage_mem = reinterpret_cast<uint32_t *>(table + kJoinSlots);
	uint64_t *scratch = reinterpret_cast<uint64_t *>(stage_mem + 3 * kStageCap);   // 4 * 32 uint64
	__shared__ uint32_t s_task_id, s_cnt, s_sentinels;
	__shared__ unsigned long long s_base;
	MatchStage st;
	st.k = stage_mem;
	st.o = stage_mem + kStageCap;
	st.i = stage_mem + 2 * kStageCap;
	st.cnt = &s_cnt;
	st.cap = kStageCap;
	constexpr uint32_t kMask = kJoinSlots - 1, kFill = kJoinSlots / 2;   // build tuples per table fill
	constexpr int kTableShift = 32 - 13;
	static_assert(kJoinSlots == 1u << 13, "table hash takes the top 13 bits");
	JoinSums acc;
	acc.zero();
	if (threadIdx.x == 0) s_cnt = 0;
	const uint32_t total_tasks = task_prefix[P];

	while (true) {
		__syncthreads();
		if (threadIdx.x == 0) s_task_id = atomicAdd(task_counter, 1u);
		__syncthreads();
		const uint32_t task = s_task_id;
		if (task >= total_tasks) break;
		const uint32_t p = upper_parent(task_prefix, P, task);
		const uint32_t slice = task - task_prefix[p];
		const uint32_t r_beg = r_off[p], r_end = r_off[p + 1];
		uint32_t s_beg = s_off[p] + slice * s_task, s_end = s_off[p + 1];
		if (s_end - s_beg > s_task) s_end = s_beg + s_task;

		for (uint32_t fb = r_beg; fb < r_end; fb += kFill) {
			const uint32_t fe = min(fb + kFill, r_end);
			// ---- build (reference build(): double hashing into a prime table; here linear probing
			// into a power-of-two table indexed by the top bits of key * table_factor)
			for (uint32_t h = threadIdx.x; h < kJoinSlots; h += kJoinThreads) table[h] = kEmptySlot;
			if (threadIdx.x == 0) s_sentinels = 0;
			__syncthreads();
			for (uint32_t i = fb + threadIdx.x; i < fe; i += kJoinThreads) {
				const uint32_t key = rk[i];
				const uint64_t pair = ((uint64_t)rv[i] << 32) | key;
				if (pair == kEmptySlot) {                  // the one pair that looks like an empty slot
					atomicAdd(&s_sentinels, 1u);
					continue;
				}
				uint32_t h = hash_mul(key, table_factor) >> kTableShift;
				while (atomicCAS(reinterpret_cast<unsigned long long *>(&table[h]), (unsigned long long)kEmptySlot,
				                 (unsigned long long)pair) != kEmptySlot)
					h = (h + 1) & kMask;
			}
			__syncthreads();
			const uint32_t sentinels = s_sentinels;
			// ---- probe, kJoinThreads * kJoinItems tuples per round
			for (uint32_t sb = s_beg; sb < s_end; sb += kJoinThreads * kJoinItems) {
				uint32_t key[kJoinItems], val[kJoinItems];
				bool valid[kJoinItems];
#pragma unroll
				for (uint32_t t = 0; t < kJoinItems; ++t) {
					const uint32_t i = sb + t * kJoinThreads + threadIdx.x;
					valid[t] = i < s_end;
					key[t] = valid[t] ? ldg_stream_u32(&sk[i]) : 0;
					val[t] = valid[t] ? ldg_stream_u32(&sv[i]) : 0;
				}
				for (int mode = 0; mode < 2; ++mode) {     // 0: staged; 1: direct, only after a stage overflow
#pragma unroll
					for (uint32_t t = 0; t < kJoinItems; ++t) {
						bool active = valid[t];
						uint32_t h = hash_mul(key[t], table_factor) >> kTableShift;
						while (__any_sync(kFullMask, active)) {
							const uint64_t slot = active ? table[h] : kEmptySlot;
							active = active && slot != kEmptySlot;
							const bool hit = active && (uint32_t)slot == key[t];
							const uint32_t ival = (uint32_t)(slot >> 32);
							if (mode == 0) {
								if (hit) acc.add(key[t], val[t], ival);
								if (MATERIALIZE) st.emit(hit, key[t], val[t], ival);
							} else {
								emit_direct(out, hit, key[t], val[t], ival);
							}
							h = (h + 1) & kMask;
						}
						if (sentinels && valid[t] && key[t] == kSentinelKey) {
							for (uint32_t c = 0; c < sentinels; ++c) {
								if (mode == 0) {
									acc.add(key[t], val[t], kSentinelKey);
									if (MATERIALIZE) st.emit_one(key[t], val[t], kSentinelKey);
								} else {
									const unsigned long long r = atomicAdd(out.cursor, 1ull);
									if (r < out.cap) {
										out.k[r] = key[t];
										out.o[r] = val[t];
										out.i[r] = kSentinelKey;
									}
								}
							}
						}
					}
					if (!MATERIALIZE || mode == 1) break;
					if (stage_flush(st, out, &s_base)) break;      // common case: rows copied out, done
				}
			}
			__syncthreads();
		}
	}
	acc.reduce_to_global(sums, scratch);
}

int launch_partition_join(const JoinArgs &a, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	const size_t smem = (size_t)kJoinSlots * 8 + (size_t)3 * kStageCap * 4 + 4 * 32 * 8;
	static bool attr_set = false;
	if (!attr_set) {
		cudaFuncSetAttribute(k_partition_join<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		cudaFuncSetAttribute(k_partition_join<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		attr_set = true;
	}
	cudaMemsetAsync(a.task_counter, 0, 4, s);
	t->start(KK_JOIN_TASKS, s);
	k_join_tasks<<<1, 1024, 0, s>>>(a.r_off, a.s_off, a.P, a.s_task, a.task_prefix);
	t->stop(s);
	int per_sm = 0;
	if (a.materialize)
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_partition_join<true>, kJoinThreads, smem);
	else
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_partition_join<false>, kJoinThreads, smem);
	if (per_sm < 1) per_sm = 1;
	const uint32_t grid = (uint32_t)(sms * per_sm);
	OutCols out;
	out.k = a.out_k;
	out.o = a.out_o;
	out.i = a.out_i;
	out.cursor = a.scalars;
	out.cap = a.materialize ? a.out_cap : 0;
	t->start(KK_PART_JOIN, s);
	if (a.materialize)
		k_partition_join<true><<<grid, kJoinThreads, smem, s>>>(a.rk, a.rv, a.sk, a.sv, a.r_off, a.s_off, a.P,
		                                                        a.task_prefix, a.task_counter, a.s_task,
		                                                        a.table_factor, out, a.scalars + 1);
	else
		k_partition_join<false><<<grid, kJoinThreads, smem, s>>>(a.rk, a.rv, a.sk, a.sv, a.r_off, a.s_off, a.P,
		                                                         a.task_prefix, a.task_counter, a.s_task,
		                                                         a.table_factor, out, a.scalars + 1);
	t->stop(s);
	return 2;
}

}  // namespace hjb
