/* hjb200.h -- C ABI of the B200-native hash-join engine (libhjb200.so).
 *
 * Drop-in boundary for the hot path of xtcyclist/hash_join_codes_KNL: the three join
 * programs NPJ (npj.cpp), PHJ (phj.cpp), CPRA (cpra2.cpp) and the kernels they are made of.
 * The reference has no FFI; its interface is (1) the CLI `./npj|./phj|./cpra [#threads]
 * [outer] [inner]` over four raw uint32 files and (2) C-linkage-style free functions over SoA
 * uint32 columns (hj.h:73-83 and the definitions cited below).  Every entry point here names
 * the reference interface it replaces.  Plain pointers and sizes only; no C++/torch types.
 *
 * Conventions (as in the reference): R = inner = build side, S = outer = probe side
 * (npj.cpp:933-934); a tuple is a 32-bit key + 32-bit payload held in two separate columns;
 * a result row is (key, outer/S payload, inner/R payload) in three columns (npj.cpp:998-1000);
 * every (r, s) pair with equal keys is emitted (no _UNIQUE, npj.cpp:288-290).
 * Unlike the reference no key value is reserved: key 0 is legal (the reference uses it as
 * the empty sentinel, npj.cpp:205,583).
 *
 * All functions return 0 (HJB_OK) or a negative error code; hjb_last_error(ctx) gives text.
 * Nothing here ever runs the join on the CPU: without a CUDA device hjb_create fails.
 */
#ifndef HJB200_H
#define HJB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HJB_VERSION 100

enum {
	HJB_OK = 0,
	HJB_E_INVALID = -1,   /* bad argument (null pointer, misaligned device column, size > 2^32-1) */
	HJB_E_CUDA = -2,      /* CUDA runtime error, see hjb_last_error */
	HJB_E_NOMEM = -3,     /* device or host allocation failed */
	HJB_E_IO = -4,        /* relation file missing or of the wrong size */
	HJB_E_NODEVICE = -5,  /* no CUDA device: there is no CPU fallback */
	HJB_E_CAPACITY = -6   /* hjb_cpra_finish: a receive buffer was too small; nothing was exchanged, re-bind larger ones */
};

typedef struct hjb_ctx hjb_ctx;        /* one per GPU: device, stream, workspace, output buffers */

/* One relation: two columns of `tuples` uint32 each.  Host or device pointers depending on
 * the entry point.  Device columns must be 16-byte aligned.  Replaces the inner_keys /
 * inner_vals / outer_keys / outer_vals members of info_t_hj (hj.h:15-18). */
typedef struct hjb_rel {
	const uint32_t *keys;
	const uint32_t *vals;
	uint64_t tuples;                   /* <= 2^32 - 2^16 per GPU (the reference: uint32 counts, phj.cpp:1722-1727) */
} hjb_rel;

/* Tunables the reference hard-codes in main() (npj.cpp:944-945, phj.cpp:1976-1979,
 * cpra2.cpp:2023,2031-2034).  Zero / 0.0 selects the default. */
typedef struct hjb_opts {
	int materialize;       /* 1 (reference behaviour): write the dense 3-column result; 0: count + checksums only */
	uint32_t seed;         /* hash factors are drawn from it (npj.cpp:975-977); 0 -> fixed default */
	double npj_load;       /* NPJ table load factor (reference 0.90); default 0.75 */
	int radix_bits[4];     /* PHJ/CPRA fan-out bits per pass (reference: planner phj.cpp:1791-1808); all 0 -> planner */
	uint32_t part_tuples;  /* planner target for build tuples per final partition (reference hash_table_limit 6400) */
	uint64_t out_capacity; /* rows the result buffers may hold; 0 -> max(|S|, |R|); grown and retried on overflow */
	int reserved[8];
} hjb_opts;

/* Result of one join.  Replaces join_keys / join_outer_vals / join_inner_vals and
 * join_tuples of info_t_hj (hj.h:11,19-21) and the value close_gaps returns (npj.cpp:905).
 * The rows are dense ([0, count)), in no particular order.  The buffers belong to the
 * context and stay valid until the next join on it or hjb_destroy. */
typedef struct hjb_result {
	uint64_t count;        /* number of matching pairs */
	uint64_t sum_key;      /* sum of key over the rows, uint64 wrap */
	uint64_t sum_outer;    /* sum of S payloads */
	uint64_t sum_inner;    /* sum of R payloads */
	const uint32_t *keys;        /* NULL when !materialize */
	const uint32_t *outer_vals;
	const uint32_t *inner_vals;
	int rows_on_device;    /* 1: the three pointers are device memory; 0: (pinned) host memory */
	double seconds;        /* device-timed region: the reference's definition (npj.cpp:861-918): table
	                          init + every partition pass + join + materialisation, inputs resident */
	double seconds_e2e;    /* host entry points: wall time including H2D of the inputs and D2H of the rows */
	float phase_ms[8];     /* [0] table init / pass-1 R, [1] build / pass-1 S, [2] probe / pass-2 R, [3] pass-2 S,
	                          [4] per-partition join, [5] H2D, [6] D2H, [7] spare */
	uint32_t kernel_launches;  /* kernels of this library launched by the call */
	uint32_t partitions;   /* final partition count (PHJ / CPRA), NPJ: buckets */
} hjb_result;

/* ---- context ---------------------------------------------------------------------- */
int hjb_version(void);
int hjb_create(int device, hjb_ctx **ctx);
int hjb_destroy(hjb_ctx *ctx);
const char *hjb_last_error(const hjb_ctx *ctx);    /* ctx may be NULL: error of the last failed hjb_create */
/* run on a caller-owned CUDA stream (cudaStream_t as void*), e.g. torch's current stream */
int hjb_set_stream(hjb_ctx *ctx, void *cuda_stream);
int hjb_synchronize(hjb_ctx *ctx);
/* Per-kernel device times of the LAST whole join, from CUDA events recorded around every launch on
 * the launching stream (replaces the reference's per-phase times[] / -DTIMELOG, npj.cpp:878-915,
 * hj.h:69-70).  hjb_kernel_times fills ms[k] / launches[k] for kernel kind k and returns the number
 * of kinds; hjb_kernel_name(k) names kind k. */
int hjb_set_profiling(hjb_ctx *ctx, int on);
int hjb_kernel_times(hjb_ctx *ctx, float *ms, uint32_t *launches, int max_kinds);
const char *hjb_kernel_name(int kind);

/* ---- whole joins: replace run() / run_hj() + main()'s allocation ------------------- */
/* NPJ, npj.cpp:769-927.  Columns in device memory. */
int hjb_npj_device(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out);
/* PHJ, phj.cpp:1646-1949 (with the join phase of phj.cpp:1869-1924). */
int hjb_phj_device(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out);
/* Same, columns in host memory (what the reference's main() holds after fread,
 * npj.cpp:1036-1039): copies them in, joins, copies the rows out to pinned host memory. */
int hjb_npj_host(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out);
int hjb_phj_host(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out);

/* ---- CPRA, cpra2.cpp:1697-1986, one context (= one GPU = one of the reference's "threads")
 * per process.  The reference partitions each thread's chunk locally, then thread t gathers
 * the pieces of the partitions it owns (cpra2.cpp:1868-1906).  Here: GPU g owns the hash
 * range [g/G, (g+1)/G); hjb_cpra_split radix-partitions this GPU's chunk by owner (the
 * GPU-assign pass) into ctx-owned send buffers, the caller exchanges them (NCCL all-to-all or
 * peer copies -- plumbing, not part of this library), and hjb_cpra_join_local joins what
 * arrived exactly like PHJ. */
typedef struct hjb_split {
	const uint32_t *r_keys, *r_vals;   /* device, grouped by owner GPU */
	const uint32_t *s_keys, *s_vals;
	uint64_t r_offsets[65];            /* owner g holds [r_offsets[g], r_offsets[g+1]) ; ngpus <= 64 */
	uint64_t s_offsets[65];
	float ms;
} hjb_split;
int hjb_cpra_split(hjb_ctx *ctx, const hjb_rel *R_chunk, const hjb_rel *S_chunk, int ngpus,
                   const hjb_opts *opts, hjb_split *out);
/* hash range restriction for the local join: the tuples all hash into owner `gpu` of `ngpus` */
int hjb_cpra_join_local(hjb_ctx *ctx, const hjb_rel *R_recv, const hjb_rel *S_recv, int gpu, int ngpus,
                        const hjb_opts *opts, hjb_result *out);

/* ---- CPRA with the exchange FUSED into the GPU-assign pass: instead of split -> all-to-all,
 * the scatter kernel stores every tuple straight into its owner's receive buffer over NVLink
 * (the owners' buffers are mapped into this process through CUDA IPC).  Per join:
 *   hjb_cpra_count         histogram + scan of this GPU's chunk by owner; counts to the host
 *   (caller)               all-gather the counts; sender s gets rows [base, base + count) of owner g's buffers
 *   hjb_cpra_scatter_peer  the scatter, one coalesced sector-aligned NVLink store stream per owner
 *   (caller)               barrier, then hjb_cpra_join_local on the own receive buffers
 * This replaces the reference's per-partition memcpy gather (cpra2.cpp:1896-1904, :1951-1958). */
typedef struct hjb_recv {
	uint32_t *r_keys, *r_vals, *s_keys, *s_vals;   /* this GPU's receive buffers (device, owned by ctx) */
	uint64_t r_capacity, s_capacity;               /* rows */
	unsigned char ipc[4][64];                      /* cudaIpcMemHandle_t of the four buffers, in that order */
} hjb_recv;
int hjb_cpra_recv_alloc(hjb_ctx *ctx, uint64_t r_capacity, uint64_t s_capacity, hjb_recv *out);
int hjb_ipc_open(hjb_ctx *ctx, const unsigned char *handle64, void **dev_ptr);
int hjb_ipc_close(hjb_ctx *ctx, void *dev_ptr);
/* r_counts / s_counts: ngpus entries each.  Keeps the chunk pointers and the scan state in ctx
 * until hjb_cpra_scatter_peer; the chunks must stay untouched in between. */
int hjb_cpra_count(hjb_ctx *ctx, const hjb_rel *R_chunk, const hjb_rel *S_chunk, int ngpus, const hjb_opts *opts,
                   uint64_t *r_counts, uint64_t *s_counts);
/* peer_*[g]: owner g's receive column as seen from this process (own pointer for g == this GPU,
 * hjb_ipc_open'ed otherwise); r_base[g] / s_base[g]: first row of owner g reserved for this sender */
int hjb_cpra_scatter_peer(hjb_ctx *ctx, int ngpus, void *const *peer_r_keys, void *const *peer_r_vals,
                          void *const *peer_s_keys, void *const *peer_s_vals, const uint64_t *r_base,
                          const uint64_t *s_base, float *ms);

/* ---- the same exchange, STREAM-ORDERED: count, scatter and local join are enqueued on the context's stream and
 * read their sizes from device memory, so the host is not involved between them.  The caller's collectives (NCCL
 * on the same stream) order the GPUs.  One CPRA step (replaces run_hj, cpra2.cpp:1697-1986, its barriers at
 * cpra2.cpp:1811,1828,1860 becoming the two collectives):
 *   hjb_cpra_bind           once per set of receive buffers: every owner's four columns as mapped into this process
 *   hjb_cpra_count_async    histogram + scan by owner; this sender's 2*ngpus counts (R per owner, then S) -> counts_dev
 *   (caller)                all-gather of the counts into matrix_dev[ngpus][2*ngpus] (uint64, device)
 *   hjb_cpra_scatter_async  bases from the matrix (device), then the fused scatter into the owners' buffers
 *   (caller)                a collective every rank enqueues after its scatter (e.g. a 1-element all-reduce): when it
 *                           completes here, every sender's stores have landed
 *   hjb_cpra_join_async     local join of what arrived; the received counts are taken from the matrix on the device
 *   (caller, optional)      all-reduce of hjb_cpra_sums_dev (count + 3 checksums, uint64[4]) over the ranks
 *   hjb_cpra_finish         the only synchronisation: waits, returns this GPU's result, rows received and the
 *                           largest row counts any owner received (for sizing the buffers)
 * If some owner's buffer is too small (every sender sees that in the same matrix) nothing is scattered or joined
 * and hjb_cpra_finish returns HJB_E_CAPACITY with `largest` set. */
int hjb_cpra_bind(hjb_ctx *ctx, int gpu, int ngpus, void *const *peer_r_keys, void *const *peer_r_vals,
                  void *const *peer_s_keys, void *const *peer_s_vals, uint64_t r_capacity, uint64_t s_capacity);
int hjb_cpra_count_async(hjb_ctx *ctx, const hjb_rel *R_chunk, const hjb_rel *S_chunk, const hjb_opts *opts,
                         uint64_t *counts_dev);
int hjb_cpra_scatter_async(hjb_ctx *ctx, const uint64_t *matrix_dev);
/* r_expect / s_expect: the row counts the plan is made for (0: the capacities) */
int hjb_cpra_join_async(hjb_ctx *ctx, const hjb_opts *opts, uint64_t r_expect, uint64_t s_expect);
void *hjb_cpra_sums_dev(hjb_ctx *ctx);
int hjb_cpra_finish(hjb_ctx *ctx, hjb_result *out, uint64_t received[2], uint64_t largest[2]);

/* ---- the STAGED exchange (the default for N > 1 in cpra.py / bench.py): two radix passes over the data instead of the
 * fused path's three, and an exchange that leaves the SMs alone.  It is the reference's own order -- chunk-local passes
 * first (cpra2.cpp:1783-1827), then the per-owner gather of whole partition pieces with memcpy
 * (cpra2.cpp:1861-1905,1940-1959):
 *   hjb_cpra_stage_plan          how one step's radix bits are split (same arguments on every rank): stage A takes
 *                                abits = owner bits + sub-partition bits (<= 9), the local pass bbits (1..9); big_fill: the
 *                                partitions average 8192 build tuples and the join takes 12288-tuple fills (2^31 tuples
 *                                on 8 GPUs).  HJB_E_INVALID: two passes do not suffice, use the fused path
 *   hjb_cpra_stage_count_async   stage A's histogram + scan of both chunks; this sender's 2 * 2^abits counts (R per digit,
 *                                then S) -> counts_dev.  nparts (a power of two <= 8): the runs will leave, and be
 *                                processed by their owners, in that many parts (ranges of sub-partitions)
 *   (caller)                     all-gather into matrix_dev[ngpus][2][2^abits] (uint64, device)
 *   hjb_cpra_stage_scatter_async rel 0: from the matrix, this sender's run in every owner's columns (an owner receives
 *                                sender-major: one run per sender with its sub-partitions in order -- the pieces of the
 *                                reference's gather, cpra2.cpp:1896-1904) and the per-sender ranges every received
 *                                sub-partition consists of; then stage A's scatter of R -- the run this GPU owns itself
 *                                straight to its final rows (receive buffers from hjb_cpra_recv_alloc), the others into
 *                                the staging region behind them, each starting with the 128-byte phase of its destination.
 *                                rel 1: S.  One host synchronisation (rel 0, the stream holds the counting kernels only)
 *                                unless HJB_STAGE_COPY=tma
 *   hjb_cpra_stage_copy_async    part `part` of one relation's runs -> the owners' columns, on `cuda_stream` (null: the
 *                                context's stream): cudaMemcpyAsync per owner and column (copy engines, no SM), or with
 *                                HJB_STAGE_COPY=tma the kernel k_peer_copy (TMA bulk copies global -> shared -> peer).  On a
 *                                side stream that waits for the scatter the copies cross NVLink beside the passes
 *   (caller)                     per piece, a collective after the copy: when it completes every sender's piece is in
 *   hjb_cpra_stage_local_async   rel 0: the local pass over part `part` of what arrived of R (every sub-partition is the
 *                                union of one range per sender); rel 1: the same for S, then the join of the part's
 *                                partitions.  R's part before S's; any order of parts
 *   hjb_cpra_finish              after the last part: as above
 * hjb_cpra_bind precedes as for the fused path; capacity failures are reported the same way. */
int hjb_cpra_stage_plan(hjb_ctx *ctx, int ngpus, uint64_t r_expect, uint64_t s_expect, const hjb_opts *opts, int *abits,
                        int *bbits, int *big_fill);
int hjb_cpra_stage_count_async(hjb_ctx *ctx, const hjb_rel *R_chunk, const hjb_rel *S_chunk, const hjb_opts *opts, int abits,
                               int nparts, uint64_t *counts_dev);
int hjb_cpra_stage_scatter_async(hjb_ctx *ctx, const uint64_t *matrix_dev, int rel);
int hjb_cpra_stage_copy_async(hjb_ctx *ctx, int rel, int part, void *cuda_stream);
int hjb_cpra_stage_local_async(hjb_ctx *ctx, const hjb_opts *opts, int bbits, int big_fill, int rel, int part);

/* Skew (write.cpp's `zipf` knob, write.cpp:1685-1689; the reference's static ownership par_start / par_end,
 * cpra2.cpp:1868-1872, sends every tuple of a frequent key to one thread).  The probe tuples of a small set of
 * hot keys (<= 256, chosen by the caller, e.g. from a sample of the probe chunks) stay with their sender:
 *   hjb_cpra_split_hot    S chunk -> cold part (takes the normal step) + hot part; both in context-owned memory,
 *                         valid until the next split on this context
 *   hjb_cpra_select_hot   this sender's R tuples with hot keys (the caller all-gathers them: <= 4096 in total)
 *   hjb_cpra_hot_join     after hjb_cpra_join_async: this GPU's hot S tuples x all hot R tuples, rows and checksums
 *                         appended to the step's result (enqueued; hjb_cpra_finish returns the union)
 * Key 0xFFFFFFFF must not be declared hot. */
int hjb_cpra_split_hot(hjb_ctx *ctx, const hjb_rel *S_chunk, const uint32_t *hot_keys_dev, uint32_t n_hot, hjb_rel *cold,
                       hjb_rel *hot);
int hjb_cpra_select_hot(hjb_ctx *ctx, const hjb_rel *R_chunk, const uint32_t *hot_keys_dev, uint32_t n_hot,
                        uint32_t *keys_out_dev, uint32_t *vals_out_dev, uint32_t capacity, uint64_t *found);
int hjb_cpra_hot_join(hjb_ctx *ctx, const hjb_rel *S_hot, const hjb_rel *R_hot);

/* The same step with the chunk in HOST memory (what the reference's main() holds after fread, cpra2.cpp:2128-2136):
 * hjb_cpra_count_async_host copies the chunk in on the stream before counting, hjb_cpra_finish_host copies this GPU's
 * rows out to pinned host memory (hjb_result: rows_on_device = 0, phase_ms[5] H2D, phase_ms[6] D2H).  The columns
 * should be pinned: cudaHostAlloc'ed, or registered once with hjb_host_register (page-locks the caller's own
 * allocation, e.g. the reference's mamalloc'ed columns, npj.cpp:118-126); pageable memory works but copies slower.
 * The same holds for hjb_npj_host / hjb_phj_host. */
int hjb_host_register(void *ptr, size_t bytes);
int hjb_host_unregister(void *ptr);
int hjb_cpra_count_async_host(hjb_ctx *ctx, const hjb_rel *R_chunk_host, const hjb_rel *S_chunk_host, const hjb_opts *opts,
                              uint64_t *counts_dev);
int hjb_cpra_finish_host(hjb_ctx *ctx, hjb_result *out, uint64_t received[2], uint64_t largest[2]);

/* ---- the kernels, one call each, device pointers: mirror the reference's free functions
 * so intermediate products can be compared with the oracle -------------------------- */
/* hash h(key,f,N) = ((uint32)(key*f) * N) >> 32, npj.cpp:200-201 / simd_hash npj.cpp:90-106;
 * the factor the library derives from `seed` for stage `which` (0: radix, 1: table) */
uint32_t hjb_hash_factor(uint32_t seed, int which);
/* histogram(), cpra2.cpp:801-802: counts[p] = #keys with radix digit p.  The digit is bits
 * [32-shift-bits, 32-shift) of key*factor, i.e. h(key,f,2^(shift+bits)) mod 2^bits. */
int hjb_histogram(hjb_ctx *ctx, const uint32_t *keys, uint64_t size, uint32_t *counts,
                  uint32_t factor, int shift, int bits);
/* histogram + interleave + partition, cpra2.cpp:801-1075, phj.cpp:1263-1291: one radix pass
 * of `bits` bits below `shift` already-partitioned bits.  parent_offsets has 2^shift + 1
 * entries (NULL when shift == 0); child_offsets receives 2^(shift+bits) + 1. */
int hjb_partition_pass(hjb_ctx *ctx, const uint32_t *keys, const uint32_t *vals, uint64_t size,
                       const uint32_t *parent_offsets, uint32_t *keys_out, uint32_t *vals_out,
                       uint32_t *child_offsets, uint32_t factor, int shift, int bits);
/* set + build, npj.cpp:366-380,190-212: table of `buckets` x 4 slots of (payload<<32 | key),
 * all-ones = empty.  table must hold buckets*4 uint64. */
int hjb_npj_build(hjb_ctx *ctx, const uint32_t *keys, const uint32_t *vals, uint64_t size,
                  uint64_t *table, uint64_t buckets, uint32_t factor);

/* ---- relation files, write.cpp:1824-1865 / npj.cpp:1013-1039: headerless little-endian
 * uint32[n]; <dir>/ik_<n>.txt iv_<n>.txt (inner) and ok_<n>.txt ov_<n>.txt (outer) ------ */
int hjb_relation_write(const char *dir, int outer, uint64_t tuples, const uint32_t *keys, const uint32_t *vals);
int hjb_relation_read(const char *dir, int outer, uint64_t tuples, uint32_t *keys, uint32_t *vals);

/* ---- deterministic generator on the device (replaces write.cpp / generate_data_for_join,
 * cpra2.cpp:1578-1696; seeded, real Zipf -- SURVEY.md §2 note on W).  kind: 0 = unique keys
 * (R, or S as a permutation of R's key set), 1 = foreign keys into a build side of
 * `domain` keys (every key at least once when tuples >= domain, the rest uniform),
 * 2 = Zipf(theta) ranks over `domain` keys of which a `selectivity` fraction of TUPLES hit. */
typedef struct hjb_gen {
	int kind;
	uint64_t tuples;       /* size of this column pair */
	uint64_t domain;       /* |R| the keys refer to */
	uint64_t first;        /* global index of element 0 (for sharded generation) */
	uint64_t total;        /* global size of the relation (permutation domain) */
	uint32_t seed;         /* selects the key SET: rank r -> key is a function of (seed, r) only */
	uint32_t order_seed;   /* selects the order / the uniform picks, so R and S share keys but not order */
	uint32_t payload_factor;   /* payload = key * factor (cpra2.cpp:1663-1674) */
	uint32_t pad_;
	double theta, selectivity;
} hjb_gen;
int hjb_generate(hjb_ctx *ctx, const hjb_gen *g, uint32_t *keys_dev, uint32_t *vals_dev);
/* sum of a device column as uint64 (input checksums, cpra2.cpp:1628,1650) */
int hjb_column_sum(hjb_ctx *ctx, const uint32_t *col_dev, uint64_t size, uint64_t *sum);
/* Verifier (the reference has none): order-independent fingerprint of `rows`
 * result rows held in three DEVICE columns -- fp[0] = sum, fp[1] = xor over the rows of
 * mix64(key | outer_val << 32, inner_val) (splitmix64 finaliser).  Two joins produced the same
 * multiset of rows iff (with overwhelming probability) count and both words agree, whatever the
 * row order; lets results of 2^27..2^31 rows be compared across NPJ / PHJ / CPRA and against a
 * CPU restatement without copying or sorting them. */
int hjb_rows_fingerprint(hjb_ctx *ctx, const uint32_t *keys_dev, const uint32_t *outer_vals_dev,
                         const uint32_t *inner_vals_dev, uint64_t rows, uint64_t fp[2]);

#ifdef __cplusplus
}
#endif
#endif
