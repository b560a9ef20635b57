// hj_internal.h -- host-side declarations shared by the translation units of libhjb200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/hjb200.h"

namespace hjb {

constexpr int kMaxPasses = 4;
constexpr int kMaxRadixBits = 11;            // fan-out per pass <= 2048 (shared-memory counters)
constexpr uint32_t kDefaultPartTuples = 2048; // planner: build tuples per final partition
constexpr int kJoinLog2Slots = 12;            // shared-memory table slots per CTA (4096 x 8 B = 32 KB)
constexpr int kJoinThreads = 256;
constexpr int kJoinItems = 4;                 // probe tuples per thread per round
constexpr uint32_t kHistThreads = 512;
constexpr uint32_t kScanThreads = 256;
constexpr uint32_t kScanItems = 8;
constexpr int kNpjThreads = 256;
constexpr int kNpjItems = 4;                  // probe tuples per thread per round (their home buckets are in flight together)

// optional per-launch CUDA-event timing on the launching stream (bench.py's roofline line)
enum KernelKind { KK_MAKE_ITEMS = 0, KK_HIST, KK_SCAN, KK_SCATTER, KK_JOIN_TASKS, KK_PART_JOIN, KK_NPJ_BUILD,
                  KK_NPJ_PROBE, KK_SCATTER_PEER, KK_PEER_COPY, KK_COUNT };
struct KernelTimer {
	static constexpr int kMaxLaunches = 384;     // a sliced host join: 16 probe slices of ~20 kernels + the build side
	cudaEvent_t beg[kMaxLaunches], end[kMaxLaunches];
	int kind[kMaxLaunches];
	int n;
	int dropped;                                  // launches that found the event pool full (reported by hjb_kernel_times)
	bool enabled;
	float ms[KK_COUNT];
	uint32_t launches[KK_COUNT];
	void start(int k, cudaStream_t s)
	{
		if (!enabled) return;
		if (n >= kMaxLaunches) {
			++dropped;
			return;
		}
		kind[n] = k;
		cudaEventRecord(beg[n], s);
	}
	void stop(cudaStream_t s)
	{
		if (!enabled || n >= kMaxLaunches) return;
		cudaEventRecord(end[n], s);
		++n;
	}
};

struct RadixPassArgs {
	const uint32_t *keys, *vals;      // input columns (whole array, 16-byte aligned base)
	uint32_t *keys_out, *vals_out;    // output columns
	uint64_t n;                       // tuples in the array
	uint32_t np;                      // parents
	const uint32_t *parent_off;       // np + 1 entries, device; nullptr: one parent [0, n)
	uint32_t *child_off;              // np * 2^bits + 1 entries, device
	uint32_t factor;
	int rshift, bits;                 // digit = (key*factor >> rshift) & (2^bits - 1)
	uint32_t chunk;                   // tuples per work item
	uint32_t max_items;               // upper bound n/chunk + np
	uint32_t peer_ctas;               // peer scatter only: CTAs to launch (0 = one per item)
	uint32_t tiles_per_item;          // rows of tile_counts per item
	// scratch (device)
	uint32_t *item_prefix;            // np + 1
	uint32_t *counts;                 // max_items * 2^bits  (histogram, then offsets in place)
	uint64_t *scan_status;            // tiles
	uint32_t *scan_counter;           // 1
	uint16_t *tile_counts;            // max_items * tiles_per_item * 2^bits digit counts per scatter tile; nullptr: none
	                                  // (fan-out > 512, or the pass feeds the peer scatter, which ranks its tiles itself)
	const uint32_t *seg;              // device or nullptr: parent q is the union of nseg ranges, (first row, rows) at seg[(q * nseg + s) * 2]
	uint32_t nseg;                    // (tile-count path only; parent_off then holds the parents' cumulative sizes)
	const uint32_t *out_base;         // device or nullptr: output position of the first parent's first tuple (a pass over a range of parents)
	uint32_t chunk_div;               // 0 / 1, or: the pass sees about n / chunk_div tuples (sizes the work items)
	const int32_t *shift;             // device, 2^bits entries or nullptr: added to every digit's output positions (staged CPRA
	                                  // exchange, np == 1: each digit's run starts with the 16-byte phase of its destination row)
};
// per-owner output columns of the fused GPU-assign pass (CPRA): the owners' receive buffers as mapped into this
// process, and -- in device memory, so that no host round trip sits between the count exchange and the scatter --
// the first row of every owner's buffer reserved for this sender
struct PeerTable {
	uint32_t *k[64];
	uint32_t *v[64];
	const uint32_t *base;        // [fan-out] device: row of owner g's columns where this sender's tuples start
	const uint32_t *sender_off;  // [fan-out + 1] device: the scan's owner offsets of this sender's chunk (child_off)
	const uint32_t *abort_flag;  // device, may be null: non-zero = some owner's buffer is too small, scatter nothing
};
constexpr uint32_t kPeerCarry = 32;       // peer stores are combined to whole 128-byte lines (fan-out <= 64)
size_t radix_scratch_bytes(uint64_t n, uint32_t np, int bits, uint32_t *chunk, uint32_t *max_items, uint32_t *tiles,
                           uint32_t *tiles_per_item = nullptr, uint32_t chunk_div = 1);
// fills a.chunk / max_items / tiles_per_item and carves a's scratch arrays out of `scratch` (radix_scratch_bytes of
// a.n, a.np, a.bits); tile_counts = false leaves a.tile_counts null
void radix_carve(RadixPassArgs &a, char *scratch, bool tile_counts);
// launches make_items + histogram + scan + scatter; returns kernels launched
int launch_radix_pass(const RadixPassArgs &a, cudaStream_t s, int sms, KernelTimer *t = nullptr);
int launch_radix_count(const RadixPassArgs &a, cudaStream_t s, KernelTimer *t = nullptr);
int launch_radix_scatter(const RadixPassArgs &a, cudaStream_t s, KernelTimer *t = nullptr, const PeerTable *peers = nullptr);
int launch_histogram_only(const uint32_t *keys, uint64_t n, uint32_t *counts_dev, uint32_t factor,
                          int rshift, int bits, cudaStream_t s, int sms);

// ---- staged CPRA exchange (stage.cu)
// device words written by k_stage_bases (uint32), per relation: what stage A's scatter adds to the scan's positions per digit;
// per owner this sender's run: first row in stage A's output columns, rows, first row in the owner's columns; the ranges
// (first row, rows) every received sub-partition consists of, [sub-partition][sender]; the sub-partitions' cumulative sizes
// The runs leave in up to kStageMaxParts parts (ranges of sub-partitions): [part][owner].
constexpr int kStageMaxParts = 8;
enum { SD_SHIFT = 0, SD_OWN_SRC = 512, SD_OWN_LEN = 1024, SD_OWN_DST = 1536, SD_SEG = 2048, SD_POFF = 3072,
       SD_REL_R = 0, SD_REL_S = 4096, SD_WORDS = 8192 };
// first sub-partition of part k of K: equal ranges (cutting four parts 3 : 3 : 1 : 1, so that the last pieces -- whose pass and
// join the copies cannot hide -- are the small ones, measured no better at 8 GPUs: 12.98 vs 12.84 ms)
__host__ __device__ inline uint32_t stage_part_lo(uint32_t k, uint32_t K, uint32_t nsub) { return k * nsub / K; }
struct PeerCols {
	uint32_t *k[64];
	uint32_t *v[64];
};
int launch_stage_counts(const uint32_t *r_off, const uint32_t *s_off, int abits, unsigned long long *counts, cudaStream_t s);
int launch_stage_bases(const unsigned long long *M, int G, int me, int abits, int gbits, uint64_t cap_r, uint64_t cap_s,
                       const uint32_t *child_r, const uint32_t *child_s, uint32_t stage_base_r, uint32_t stage_base_s, int inplace,
                       int nparts, uint32_t *out, uint32_t *status, cudaStream_t s);
int launch_peer_copy(const uint32_t *sk, const uint32_t *sv, const PeerCols &peers, const uint32_t *desc, const uint32_t *abort_flag,
                     int gbits, int me, int skip_me, cudaStream_t s, KernelTimer *t = nullptr);

struct JoinArgs {
	const uint32_t *rk, *rv, *sk, *sv;       // partitioned columns
	const uint32_t *r_off, *s_off;           // P + 1 entries each
	uint32_t P;
	uint32_t radix_factor;                   // the partitions are radix digits of key * radix_factor
	int rem_bits;                            // hash bits of key * radix_factor below the partition id
	uint32_t table_factor;
	uint32_t owner;                          // CPRA local join: this GPU's owner id and the bits that encode it
	int owner_bits;
	uint32_t *task_prefix;                   // P + 1 scratch
	uint32_t *task_counter;                  // 1 counter at [0], 64-bit block status words from byte 256 on; zeroed by the launcher
	uint32_t s_task;                         // probe tuples per task
	uint32_t *out_k, *out_o, *out_i;
	uint64_t out_cap;
	unsigned long long *scalars;             // [0] cursor, [1..4] count, sum_key, sum_outer, sum_inner
	int materialize;
	int big_fill;                            // partitions average 8192 build tuples (needs rem_bits <= 14): 12288-tuple DIRECT fills
};
int launch_partition_join(const JoinArgs &a, cudaStream_t s, int sms, KernelTimer *t = nullptr);

struct NpjArgs {
	const uint32_t *rk, *rv, *sk, *sv;
	uint64_t nr, ns;
	uint64_t *table;
	uint64_t buckets;                         // x 4 slots
	uint32_t factor;
	uint32_t *out_k, *out_o, *out_i;
	uint64_t out_cap;
	unsigned long long *scalars;              // [0] cursor, [1..4] sums, [5] sentinel build tuples, [6] duplicate build keys seen
	int materialize;
};
int launch_npj_build(const NpjArgs &a, cudaStream_t s, int sms, KernelTimer *t = nullptr);
int launch_npj_probe(const NpjArgs &a, cudaStream_t s, int sms, KernelTimer *t = nullptr);


// heavy-hitter handling of the multi-GPU join (skew.cu)
constexpr uint32_t kMaxHotKeys = 256;        // keys that may be declared hot
constexpr uint32_t kMaxHotBuild = 4096;      // build tuples with hot keys, all GPUs together
int launch_split_hot(const uint32_t *keys, const uint32_t *vals, uint64_t n, const uint32_t *hot, uint32_t n_hot, uint32_t *cold_k,
                     uint32_t *cold_v, uint32_t *hot_k, uint32_t *hot_v, unsigned long long *cursors, cudaStream_t s, int sms);
int launch_select_hot(const uint32_t *keys, const uint32_t *vals, uint64_t n, const uint32_t *hot, uint32_t n_hot, uint32_t *out_k,
                      uint32_t *out_v, uint32_t capacity, unsigned long long *cursor, cudaStream_t s, int sms);
int launch_hot_join(const uint32_t *sk, const uint32_t *sv, uint64_t ns, const uint32_t *rk, const uint32_t *rv, uint32_t nr,
                    uint32_t factor, uint32_t *out_k, uint32_t *out_o, uint32_t *out_i, uint64_t out_cap, unsigned long long *scalars,
                    cudaStream_t s, int sms);

int launch_generate(const hjb_gen &g, uint32_t *keys, uint32_t *vals, cudaStream_t s);
int launch_column_sum(const uint32_t *col, uint64_t n, unsigned long long *out_dev, cudaStream_t s, int sms);
int launch_rows_fingerprint(const uint32_t *k, const uint32_t *o, const uint32_t *iv, uint64_t n, unsigned long long *out_dev,
                            cudaStream_t s, int sms);

}  // namespace hjb
