// part_join.cu -- per-partition build + probe with the partition's table resident in shared
// memory (reference: the cache-resident join loop, phj.cpp:1880-1923 / cpra2.cpp:1907-1969,
// build phj.cpp:307-397, probe phj.cpp:399-571).
//
// Persistent CTAs pull tasks from an atomic counter.  A task = (partition p, one slice of at
// most s_task probe tuples of p): uniform inputs give one task per partition, a skewed probe
// side (Zipf) is cut into many tasks that each rebuild p's small table, so no CTA is stuck
// with a hot partition.  A build partition larger than one table fill is joined in several
// fills (block nested loop), so any input is handled -- duplicates included.
//
// Two table forms, chosen per fill:
//  * DIRECT (needs <= 16 hash bits left below the radix bits).  x = key * factor is a
//    bijection on 32-bit keys (factor is odd), and all tuples of a partition share x's top
//    bits, so the remaining low bits of x identify the key: a 2^16-bit bitmap says whether
//    a key is present, and the rank of its bit (popcount prefix) addresses its payload in a
//    dense array.  No key comparison, no collision chain, no divergent loop: build is an
//    OR reduction + one store, probe is two loads + a popcount.  A fill in which two build
//    tuples have the SAME key (the rank scan counts fewer set bits than tuples) falls back to:
//  * HASH: open addressing with linear probing over 64-bit slots (payload<<32 | key),
//    atomicCAS insert as in the reference's build (npj.cpp:206); fills without equal build
//    keys stop a probe at its first match, the others walk each chain to its end and emit
//    every match (no _UNIQUE, npj.cpp:288-290).
//
// Result rows: every lane keeps its matches in registers; the warp reserves rows with ONE
// global atomicAdd per round of 32 x kJoinItems probe tuples and stores the three result
// columns straight from registers, one coalesced 128-byte run per item (ballot-ranked).
#include "hj_device.cuh"
#include "hj_internal.h"
#include <stdlib.h>

namespace hjb {

// tasks per partition -> exclusive prefix; P <= 2^22.  A CTA takes its block of 1024 partitions from an atomic
// ticket (so the blocks START in prefix order whatever order the hardware dispatches CTAs in), publishes its
// block total as a status word (bit 63 = ready), then sums the status words of ALL lower blocks -- their
// owners hold earlier tickets, i.e. are running, so waiting on them cannot deadlock, and the waits are
// independent loads rather than a chain.
constexpr uint32_t kTaskBlock = 1024;
__global__ void __launch_bounds__(1024)
k_join_tasks(const uint32_t *__restrict__ r_off, const uint32_t *__restrict__ s_off, uint32_t P, uint32_t s_task,
             uint32_t *__restrict__ task_prefix, unsigned long long *__restrict__ status, uint32_t *__restrict__ ticket)
{
	__shared__ uint32_t warp_totals[34];
	__shared__ uint64_t s_before[32];
	__shared__ uint32_t s_block;
	if (threadIdx.x == 0) s_block = atomicAdd(ticket, 1u);
	__syncthreads();
	const uint32_t block = s_block;
	const uint32_t p = block * kTaskBlock + threadIdx.x;
	uint32_t v = 0;
	if (p < P) {
		const uint32_t rc = r_off[p + 1] - r_off[p], sc = s_off[p + 1] - s_off[p];
		v = (rc && sc) ? (sc + s_task - 1) / s_task : 0u;        // an empty side cannot match
	}
	uint32_t tot;
	const uint32_t excl = block_exclusive_scan(v, warp_totals, &tot);
	if (threadIdx.x == 0) {
		atomicExch(&status[block], (1ull << 63) | tot);
		__threadfence();
	}
	uint64_t before = 0;
	for (uint32_t b = threadIdx.x; b < block; b += 1024) {
		unsigned long long w;
		do {
			w = *reinterpret_cast<volatile unsigned long long *>(&status[b]);
		} while (!(w >> 63));
		before += w & 0xFFFFFFFFull;
	}
	before = warp_sum_u64(before);
	if (lane_id() == 0) s_before[threadIdx.x >> 5] = before;
	__syncthreads();
	uint32_t base = 0;
#pragma unroll 8
	for (int w = 0; w < 32; ++w) base += (uint32_t)s_before[w];
	if (p < P) task_prefix[p] = base + excl;
	if (block == gridDim.x - 1 && threadIdx.x == 0) task_prefix[P] = base + tot;
}

// one row, reservation aggregated over whichever lanes of the warp are here together
__device__ __forceinline__ void emit_row_opportunistic(const OutCols &out, uint32_t key, uint32_t oval, uint32_t ival)
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

constexpr uint32_t kDirectWords = 4096;                  // 2^16 presence bits, 16 per word; the word's high half holds the rank prefix
constexpr uint32_t kDirectFill = 6144;                   // build tuples per DIRECT fill (payload array, 24 KB)
constexpr uint32_t kHashSlots = 1u << kJoinLog2Slots;    // HASH table slots (filled to <= 0.75)
constexpr size_t kDirectBytes = (size_t)kDirectWords * 4 + (size_t)kDirectFill * 4;
constexpr size_t kJoinSmemBytes = (size_t)kHashSlots * 8 > kDirectBytes ? (size_t)kHashSlots * 8 : kDirectBytes;

// FILL / WORDS: build tuples per DIRECT fill and words of the presence bitmap.  The default serves partitions that
// average 4096 build tuples below <= 16 hash bits; 12288 / 1024 (52 KB, four CTAs per SM) serves the staged CPRA
// exchange at 2^31 tuples, where two 9-bit passes leave partitions of 8192 build tuples and 14 hash bits.
template <int THREADS, int ITEMS, int MINB, bool MATERIALIZE, bool OWNER, bool CTAEMIT, uint32_t FILL = kDirectFill,
          int LOG2HASH = kJoinLog2Slots, uint32_t WORDS = kDirectWords>
__global__ void __launch_bounds__(THREADS, MINB)
k_partition_join(const uint32_t *__restrict__ rk, const uint32_t *__restrict__ rv,
                 const uint32_t *__restrict__ sk, const uint32_t *__restrict__ sv,
                 const uint32_t *__restrict__ r_off, const uint32_t *__restrict__ s_off, uint32_t P,
                 const uint32_t *__restrict__ task_prefix, uint32_t *__restrict__ task_counter, uint32_t s_task,
                 uint32_t radix_factor, uint32_t table_factor, int rem_bits, uint32_t owner, int owner_bits,
                 OutCols out, unsigned long long *__restrict__ sums)
{
	extern __shared__ __align__(16) unsigned char s_raw[];
	// DIRECT view
	// one word per 16 keys: bits 0..15 presence, bits 16..31 the number of present keys before the word -- a probe
	// (or a payload placement) is ONE shared-memory load + a popcount
	uint32_t *dtab = reinterpret_cast<uint32_t *>(s_raw);              // WORDS
	uint32_t *dvals = dtab + WORDS;                             // FILL
	constexpr uint32_t kHashSlots = 1u << LOG2HASH, kHashFill = kHashSlots / 4 * 3;
	// HASH view
	uint64_t *table = reinterpret_cast<uint64_t *>(s_raw);             // kHashSlots
	__shared__ uint64_t scratch[4 * 32];
	__shared__ uint32_t warp_totals[34];
	__shared__ uint32_t s_task_id[2], s_hdups, s_sentinels;
	__shared__ __align__(8) uint32_t s_emit[2 * (THREADS / 32 + 2) + 4];
	uint32_t emit_rounds = 0;                      // CTA-uniform: which half of s_emit the next round uses
	constexpr uint32_t kMask = kHashSlots - 1;
	constexpr int kShift = 32 - LOG2HASH;
	const bool direct_ok = (1ull << rem_bits) <= (unsigned long long)WORDS * 16;
	const uint32_t rem_mask = rem_bits >= 32 ? 0xFFFFFFFFu : (1u << rem_bits) - 1;
	JoinSums acc;
	acc.zero();
	// CPRA local join: every tuple must hash into this GPU's owner range (top owner_bits bits of
	// key * radix_factor), else the DIRECT tables would confuse keys; violations are reported, not joined
	const int owner_shift = OWNER ? 32 - owner_bits : 0;
	uint32_t foreign = 0;
	const uint32_t total_tasks = task_prefix[P];
	// Every thread resolves a task's ranges itself from uniform (broadcast) loads -- no barrier, no
	// serial search by one thread.  With one task per partition (no skew) task == partition.
	auto resolve = [&](uint32_t task, uint32_t &rb, uint32_t &re, uint32_t &sb, uint32_t &se) {
		uint32_t p = task < P ? task : P - 1;
		if (!(task_prefix[p] <= task && task < task_prefix[p + 1])) p = upper_parent(task_prefix, P, task);
		rb = r_off[p];
		re = r_off[p + 1];
		sb = s_off[p] + (task - task_prefix[p]) * s_task;
		se = min(s_off[p + 1], sb + s_task);
	};
	// pull the 128-byte lines of [beg, end) of a column towards L2 (one line per thread and round)
	auto prefetch_col = [&](const uint32_t *col, uint32_t beg, uint32_t end) {
		const char *lo = reinterpret_cast<const char *>(col + beg), *hi = reinterpret_cast<const char *>(col + end);
		for (const char *q = lo + (size_t)threadIdx.x * 128; q < hi; q += (size_t)THREADS * 128)
			asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
	};
	if (threadIdx.x == 0) s_task_id[0] = atomicAdd(task_counter, 1u);
	__syncthreads();
	uint32_t r_beg = 0, r_end = 0, s_beg = 0, s_end = 0;
	if (s_task_id[0] < total_tasks) resolve(s_task_id[0], r_beg, r_end, s_beg, s_end);
	for (uint32_t it = 0;; ++it) {
		const uint32_t task = s_task_id[it & 1];
		if (task >= total_tasks) break;
		if (threadIdx.x == 0) s_task_id[(it + 1) & 1] = atomicAdd(task_counter, 1u);    // visible after this task's first barrier
		uint32_t nr_beg = 0, nr_end = 0, ns_beg = 0, ns_end = 0;
		bool next_pending = true;
		// after the task's first barrier: resolve the NEXT task and start its tuples on their way from HBM
		auto look_ahead = [&]() {
			if (!next_pending) return;
			next_pending = false;
			const uint32_t ntask = s_task_id[(it + 1) & 1];
			if (ntask >= total_tasks) return;
			resolve(ntask, nr_beg, nr_end, ns_beg, ns_end);
			prefetch_col(rk, nr_beg, min(nr_end, nr_beg + FILL));
			prefetch_col(rv, nr_beg, min(nr_end, nr_beg + FILL));
			// at most the first rounds of a long probe slice: 740 CTAs x 128 KB of prefetched lines do not survive in L2
			// beside the streamed rows (config 3: 31.0 GB of DRAM traffic for 21.5 GB algorithmic with whole slices)
			prefetch_col(sk, ns_beg, min(ns_end, ns_beg + 4096u));
			prefetch_col(sv, ns_beg, min(ns_end, ns_beg + 4096u));
		};

		uint32_t fb = r_beg;
		while (fb < r_end) {
			bool use_hash = !direct_ok;
			uint32_t fe = min(fb + (use_hash ? kHashFill : FILL), r_end);
			if (!use_hash) {
				// ---- DIRECT build, step 1: presence bits (equal keys set the same bit: step 2 counts fewer bits than tuples)
				for (uint32_t w = threadIdx.x; w < WORDS / 4; w += THREADS)
					reinterpret_cast<uint4 *>(dtab)[w] = make_uint4(0, 0, 0, 0);
				__syncthreads();
				look_ahead();
				// build tuples are fetched kBatch at a time so that their loads are in flight together; the
				// first batch (all of a typical partition) stays in registers for step 3
				constexpr int kBatch = THREADS >= 512 ? 4 : THREADS >= 256 ? 8 : 16;
				uint32_t k0[kBatch], v0[kBatch];
#pragma unroll
				for (int t = 0; t < kBatch; ++t) {
					const uint32_t i = fb + threadIdx.x + t * THREADS;
					k0[t] = i < fe ? rk[i] : 0;
					v0[t] = i < fe ? rv[i] : 0;
				}
#pragma unroll
				for (int t = 0; t < kBatch; ++t)
					if (fb + threadIdx.x + t * THREADS < fe) {
						const uint32_t x = hash_mul(k0[t], radix_factor), lo = x & rem_mask;
						const uint32_t bit = 1u << (lo & 15);
						if (OWNER) foreign |= (x >> owner_shift) ^ owner;
						atomicOr(&dtab[lo >> 4], bit);              // result unused: a reduction, no return path
					}
				for (uint32_t i0 = fb + threadIdx.x + THREADS * kBatch; i0 < fe; i0 += THREADS * kBatch) {
					uint32_t bk[kBatch];
#pragma unroll
					for (int t = 0; t < kBatch; ++t) bk[t] = i0 + t * THREADS < fe ? rk[i0 + t * THREADS] : 0;
#pragma unroll
					for (int t = 0; t < kBatch; ++t)
						if (i0 + t * THREADS < fe) {
							const uint32_t x = hash_mul(bk[t], radix_factor), lo = x & rem_mask;
							const uint32_t bit = 1u << (lo & 15);
							if (OWNER) foreign |= (x >> owner_shift) ^ owner;
							atomicOr(&dtab[lo >> 4], bit);              // result unused: a reduction, no return path
						}
				}
				__syncthreads();
				{
					// step 2: rank structure -- the high half of word w = set bits before it; fewer bits than build
					// tuples means equal keys: that fill is redone with the hash table
					// Thread t owns words t, t + THREADS, ... (conflict-free accesses); ranks are counted in (thread, word)
					// order -- any fixed order serves, build and probe read the same prefixes.
					constexpr uint32_t kPer = WORDS / THREADS;
					uint32_t local = 0;
#pragma unroll
					for (uint32_t j = 0; j < kPer; ++j) local += __popc(dtab[threadIdx.x + j * THREADS]);
					uint32_t tot;
					uint32_t run = block_exclusive_scan(local, warp_totals, &tot);
					use_hash = tot != fe - fb;
#pragma unroll
					for (uint32_t j = 0; j < kPer; ++j) {                  // read again rather than kept: 16 registers less across the scan
						const uint32_t w = dtab[threadIdx.x + j * THREADS];
						dtab[threadIdx.x + j * THREADS] = w | (run << 16);
						run += __popc(w);
					}
					__syncthreads();
				}
				if (!use_hash) {
					// step 3: payloads in rank order
					// all look-ups before the first store: as far as the compiler knows dvals[] may alias the
					// bitmap, so look-up and store in one loop would run as a chain
					uint32_t dpos[kBatch];
#pragma unroll
					for (int t = 0; t < kBatch; ++t) {
						const uint32_t lo = hash_mul(k0[t], radix_factor) & rem_mask;
						const uint32_t e = dtab[lo >> 4];
						dpos[t] = (e >> 16) + __popc(e & ((1u << (lo & 15)) - 1));
					}
#pragma unroll
					for (int t = 0; t < kBatch; ++t)
						if (fb + threadIdx.x + t * THREADS < fe) dvals[dpos[t]] = v0[t];
					for (uint32_t i0 = fb + threadIdx.x + THREADS * kBatch; i0 < fe; i0 += THREADS * kBatch) {
						uint32_t bk[kBatch], bv[kBatch];
#pragma unroll
						for (int t = 0; t < kBatch; ++t) {
							const bool in = i0 + t * THREADS < fe;
							bk[t] = in ? rk[i0 + t * THREADS] : 0;
							bv[t] = in ? rv[i0 + t * THREADS] : 0;
						}
						uint32_t bpos[kBatch];
#pragma unroll
						for (int t = 0; t < kBatch; ++t) {
							const uint32_t lo = hash_mul(bk[t], radix_factor) & rem_mask;
							const uint32_t e = dtab[lo >> 4];
							bpos[t] = (e >> 16) + __popc(e & ((1u << (lo & 15)) - 1));
						}
#pragma unroll
						for (int t = 0; t < kBatch; ++t)
							if (i0 + t * THREADS < fe) dvals[bpos[t]] = bv[t];
					}
					__syncthreads();
					// ---- DIRECT probe
					for (uint32_t sb = s_beg; sb < s_end; sb += THREADS * ITEMS) {
						uint32_t key[ITEMS], val[ITEMS], ival[ITEMS];
						bool found[ITEMS];
						const uint32_t wbase = sb + (threadIdx.x & ~31u) * ITEMS + lane_id();   // a warp owns 32*ITEMS consecutive tuples
#pragma unroll
						for (int t = 0; t < ITEMS; ++t) {
							const uint32_t i = wbase + t * 32;
							found[t] = i < s_end;
							key[t] = found[t] ? ldg_stream_u32(&sk[i]) : 0;
							val[t] = found[t] ? ldg_stream_u32(&sv[i]) : 0;
						}
#pragma unroll
						for (int t = 0; t < ITEMS; ++t) {
							const uint32_t x = hash_mul(key[t], radix_factor), lo = x & rem_mask;
							if (OWNER && found[t]) foreign |= (x >> owner_shift) ^ owner;
							const uint32_t e = dtab[lo >> 4];
							const uint32_t hit = found[t] ? (e >> (lo & 15)) & 1u : 0u;
							// rank < fill size whenever the bit is set; a miss may compute fill size itself: clamp
							const uint32_t pos = min((e >> 16) + (uint32_t)__popc(e & ((1u << (lo & 15)) - 1)), FILL - 1);
							ival[t] = dvals[pos];
							found[t] = hit != 0;
							acc.add_if(hit, key[t], val[t], ival[t]);
						}
						if (MATERIALIZE) {
							if (CTAEMIT) emit_round_cta<ITEMS>(out, s_emit, emit_rounds++, found, key, val, ival);
							else emit_round<ITEMS>(out, found, key, val, ival);
						}
					}
					__syncthreads();            // the bitmap is cleared next; also publishes the prefetched task id
					fb = fe;
					continue;
				}
				fe = min(fb + kHashFill, r_end);       // equal build keys in this fill: redo it with the hash table
			}
			// ---- HASH build (reference build(): double hashing into a prime table; here linear
			// probing into a power-of-two table indexed by the top bits of key * table_factor)
			{
				ulonglong2 *t2 = reinterpret_cast<ulonglong2 *>(table);
				for (uint32_t h = threadIdx.x; h < kHashSlots / 2; h += THREADS) t2[h] = make_ulonglong2(kEmptySlot, kEmptySlot);
			}
			if (threadIdx.x == 0) {
				s_hdups = 0;
				s_sentinels = 0;
			}
			__syncthreads();
			look_ahead();
			for (uint32_t i = fb + threadIdx.x; i < fe; i += THREADS) {
				const uint32_t key = rk[i];
				const uint64_t pair = ((uint64_t)rv[i] << 32) | key;
				if (pair == kEmptySlot) {                  // the one pair that looks like an empty slot
					atomicAdd(&s_sentinels, 1u);
					continue;
				}
				uint32_t h = hash_mul(key, table_factor) >> kShift;
				while (true) {
					const uint64_t old = atomicCAS(reinterpret_cast<unsigned long long *>(&table[h]),
					                               (unsigned long long)kEmptySlot, (unsigned long long)pair);
					if (old == kEmptySlot) break;
					if ((uint32_t)old == key) s_hdups = 1;  // equal build keys: probes must walk to the chain's end
					h = (h + 1) & kMask;
				}
			}
			__syncthreads();
			const bool slow = s_hdups != 0 || s_sentinels != 0;
			const uint32_t sentinels = s_sentinels;
			for (uint32_t sb = s_beg; sb < s_end; sb += THREADS * ITEMS) {
				uint32_t key[ITEMS], val[ITEMS], ival[ITEMS];
				bool found[ITEMS];
				const uint32_t wbase = sb + (threadIdx.x & ~31u) * ITEMS + lane_id();
#pragma unroll
				for (int t = 0; t < ITEMS; ++t) {
					const uint32_t i = wbase + t * 32;
					found[t] = i < s_end;                   // "valid" until probed
					key[t] = found[t] ? ldg_stream_u32(&sk[i]) : 0;
					val[t] = found[t] ? ldg_stream_u32(&sv[i]) : 0;
					ival[t] = 0;
				}
				if (!slow) {
#pragma unroll
					for (int t = 0; t < ITEMS; ++t) {
						bool hit = false;
						if (found[t]) {
							uint32_t h = hash_mul(key[t], table_factor) >> kShift;
							while (true) {
								const uint64_t slot = table[h];
								if (slot == kEmptySlot) break;
								if ((uint32_t)slot == key[t]) {
									ival[t] = (uint32_t)(slot >> 32);
									hit = true;
									break;
								}
								h = (h + 1) & kMask;
							}
						}
						found[t] = hit;
						if (hit) acc.add(key[t], val[t], ival[t]);
					}
					if (MATERIALIZE) emit_round<ITEMS>(out, found, key, val, ival);
				} else {
#pragma unroll
					for (int t = 0; t < ITEMS; ++t) {
						if (!found[t]) continue;
						uint32_t h = hash_mul(key[t], table_factor) >> kShift;
						while (true) {
							const uint64_t slot = table[h];
							if (slot == kEmptySlot) break;
							if ((uint32_t)slot == key[t]) {
								acc.add(key[t], val[t], (uint32_t)(slot >> 32));
								if (MATERIALIZE) emit_row_opportunistic(out, key[t], val[t], (uint32_t)(slot >> 32));
							}
							h = (h + 1) & kMask;
						}
						if (key[t] == kSentinelKey)
							for (uint32_t c = 0; c < sentinels; ++c) {
								acc.add(key[t], val[t], kSentinelKey);
								if (MATERIALIZE) emit_row_opportunistic(out, key[t], val[t], kSentinelKey);
							}
					}
				}
			}
			__syncthreads();          // the table is cleared next; also publishes the prefetched task id
			fb = fe;
		}
		r_beg = nr_beg;
		r_end = nr_end;
		s_beg = ns_beg;
		s_end = ns_end;
	}
	acc.reduce_to_global(sums, scratch);
	if (foreign) sums[6] = 1;                    // scalars[7]: foreign tuple seen
}

int launch_partition_join(const JoinArgs &a, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	const uint32_t task_blocks = (a.P + kTaskBlock - 1) / kTaskBlock;
	cudaMemsetAsync(a.task_counter, 0, 256 + (size_t)task_blocks * 8, s);
	t->start(KK_JOIN_TASKS, s);
	k_join_tasks<<<task_blocks, 1024, 0, s>>>(a.r_off, a.s_off, a.P, a.s_task, a.task_prefix,
	                                          reinterpret_cast<unsigned long long *>(a.task_counter + 64), a.task_counter + 1);
	t->stop(s);
	OutCols out;
	out.k = a.out_k;
	out.o = a.out_o;
	out.i = a.out_i;
	out.cursor = a.scalars;
	out.cap = a.materialize ? a.out_cap : 0;
	// five 256-thread CTAs per SM, 48 registers: measured best in round 1 (512-thread CTAs 1.30-1.59 vs 1.11 ms, six CTAs spill)
	t->start(KK_PART_JOIN, s);
	auto launch_shape = [&](auto kernel, int threads, size_t smem) {
		cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
		int per_sm = 0;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
		if (per_sm < 1) per_sm = 1;
		kernel<<<(uint32_t)(sms * per_sm), threads, smem, s>>>(a.rk, a.rv, a.sk, a.sv, a.r_off, a.s_off, a.P, a.task_prefix, a.task_counter,
		                                                       a.s_task, a.radix_factor, a.table_factor, a.rem_bits, a.owner, a.owner_bits,
		                                                       out, a.scalars + 1);
	};
	auto launch = [&](auto kernel) { launch_shape(kernel, kJoinThreads, kJoinSmemBytes); };
	static const int cta_emit = getenv("HJB_CTA_EMIT") ? atoi(getenv("HJB_CTA_EMIT")) : 1;
	constexpr size_t kBigBytes = 1024 * 4 + 12288 * 4;
	if (a.big_fill && a.rem_bits <= 14 && a.materialize)
		launch_shape(k_partition_join<kJoinThreads, kJoinItems, 4, true, true, true, 12288, kJoinLog2Slots, 1024>, kJoinThreads, kBigBytes);
	else if (a.big_fill && a.rem_bits <= 14)
		launch_shape(k_partition_join<kJoinThreads, kJoinItems, 4, false, true, false, 12288, kJoinLog2Slots, 1024>, kJoinThreads, kBigBytes);
	else if (a.materialize && a.owner_bits && cta_emit) launch(k_partition_join<kJoinThreads, kJoinItems, 5, true, true, true>);
	else if (a.materialize && a.owner_bits) launch(k_partition_join<kJoinThreads, kJoinItems, 5, true, true, false>);
	else if (a.materialize && cta_emit) launch(k_partition_join<kJoinThreads, kJoinItems, 5, true, false, true>);
	else if (a.materialize) launch(k_partition_join<kJoinThreads, kJoinItems, 5, true, false, false>);
	else if (a.owner_bits) launch(k_partition_join<kJoinThreads, kJoinItems, 5, false, true, false>);
	else launch(k_partition_join<kJoinThreads, kJoinItems, 5, false, false, false>);
	t->stop(s);
	return 2;
}

}  // namespace hjb
