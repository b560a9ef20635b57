/* hj_oracle.h -- CPU ORACLE for the hash-join hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain-C restatement of the reference's algorithms (xtcyclist/hash_join_codes_KNL:
 * npj.cpp, phj.cpp, cpra2.cpp, write.cpp).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (hash_join_codes_knl_b200/, include/hjb200.h) never links or calls it.
 *
 * Pinning: the reference ships no tests, golden vectors or printed results (SURVEY.md
 * §8c), so this oracle is pinned against (1) the reference's OWN functions compiled from
 * /root/reference into oracle/_ref/ (tests/test_oracle_ref.py; fixtures committed under
 * tests/golden/ so the pin also holds where /root/reference is absent), and (2) an
 * independent numpy sort-merge join.
 */
#ifndef HJ_ORACLE_H
#define HJ_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- MT19937 (npj.cpp:133-175, rand.h:40-47) ---- */
typedef struct hjo_rand32 {
	uint32_t num[625];
	size_t index;
} hjo_rand32;
void hjo_rand32_seed(hjo_rand32 *st, uint32_t seed);
uint32_t hjo_rand32_next(hjo_rand32 *st);

/* ---- hash: ((uint32)(key*factor) * n) >> 32  (npj.cpp:200-201, simd_hash npj.cpp:90-106) ---- */
uint32_t hjo_hash(uint32_t key, uint32_t factor, uint64_t n);
int hjo_odd_prime(uint64_t x);                                   /* cpra2.cpp:281-289 */
size_t hjo_thread_beg(size_t size, size_t alignment, size_t thread, size_t threads); /* npj.cpp:516-521 */
size_t hjo_thread_end(size_t size, size_t alignment, size_t thread, size_t threads); /* npj.cpp:523-529 */

/* ---- generator (cpra2.cpp:1530-1696; payloads :1663-1674; file format write.cpp:1824-1865) ---- */
void hjo_shuffle(uint32_t *data, size_t size, hjo_rand32 *gen);   /* cpra2.cpp:1530-1542 */
/* fills keys[0..size) with distinct keys != empty; table must hold `buckets` zeroed (== empty) words */
void hjo_unique(uint32_t *keys, size_t size, uint32_t *table, size_t buckets,
                uint32_t factor, uint32_t empty, hjo_rand32 *gen); /* cpra2.cpp:1544-1570 */
/* Whole generator, `threads` emulated sequentially, explicit seed (the reference seeds from
 * time(NULL), cpra2.cpp:2069).  selectivity scales join_distinct (write.cpp:1685-1689).
 * Returns 0 on success. */
int hjo_generate(size_t inner_tuples, size_t outer_tuples, double selectivity,
                 int threads, uint32_t seed,
                 uint32_t *inner_keys, uint32_t *inner_vals,
                 uint32_t *outer_keys, uint32_t *outer_vals,
                 uint32_t *inner_factor_out, uint32_t *outer_factor_out);
int hjo_relation_write(const char *dir, const char *prefix /* "i" or "o" */, size_t tuples,
                       const uint32_t *keys, const uint32_t *vals);
int hjo_relation_read(const char *dir, const char *prefix, size_t tuples,
                      uint32_t *keys, uint32_t *vals);

/* ---- NPJ kernels (npj.cpp:190-212 build, :412-445 scalar probe) ---- */
void hjo_npj_build(const uint32_t *keys, const uint32_t *vals, size_t size,
                   volatile uint64_t *table, size_t buckets, uint32_t factor, uint32_t empty);

/* ---- radix partitioning (cpra2.cpp:730-796 histogram_s/partition_s; interleave :1426-1440) ---- */
void hjo_histogram(const uint32_t *keys, size_t size, uint32_t *counts,
                   uint32_t factor, size_t partitions);
void hjo_partition(const uint32_t *keys, const uint32_t *vals, size_t size,
                   const uint32_t *counts, uint32_t *keys_out, uint32_t *vals_out,
                   uint32_t factor, size_t partitions);
size_t hjo_interleave(uint32_t **counts, uint32_t *offsets, uint32_t *aggr_counts,
                      size_t partitions, size_t thread, size_t threads);
/* planner (cpra2.cpp:1757-1772 / phj.cpp:1791-1808): fills fanout[], returns #passes */
size_t hjo_plan_fanout(size_t partitions, size_t fanout[6]);

/* ---- per-partition double-hash table (cpra2.cpp:640-710 build_s/probe_s) ---- */
void hjo_dh_build(const uint32_t *keys, const uint32_t *vals, size_t size,
                  uint64_t *table, size_t buckets, const uint32_t factor[2], uint32_t empty);

/* ---- whole joins ---- */
typedef struct hjo_result {
	uint64_t count;       /* number of (r,s) pairs with r.key == s.key */
	uint64_t sum_key;     /* sum of key over result rows, uint64 wrap */
	uint64_t sum_outer;   /* sum of probe-side (S) payloads */
	uint64_t sum_inner;   /* sum of build-side (R) payloads */
	uint32_t *keys;       /* materialised rows (malloc'd) when requested, else NULL */
	uint32_t *outer_vals;
	uint32_t *inner_vals;
	double seconds;       /* wall time of the timed region (reference definition) */
} hjo_result;

/* R = inner = build side; S = outer = probe side (npj.cpp:933-934).  threads >= 1.
 * Return 0 on success, non-zero on invalid input (key == 0 is the reference's empty
 * sentinel, npj.cpp:205,583). */
int hjo_npj(const uint32_t *rk, const uint32_t *rv, size_t nr,
            const uint32_t *sk, const uint32_t *sv, size_t ns,
            int threads, uint32_t seed, int materialize, hjo_result *out);
int hjo_phj(const uint32_t *rk, const uint32_t *rv, size_t nr,
            const uint32_t *sk, const uint32_t *sv, size_t ns,
            int threads, uint32_t seed, int materialize, hjo_result *out);
int hjo_cpra(const uint32_t *rk, const uint32_t *rv, size_t nr,
             const uint32_t *sk, const uint32_t *sv, size_t ns,
             int threads, uint32_t seed, int materialize, hjo_result *out);
void hjo_result_free(hjo_result *r);

#ifdef __cplusplus
}
#endif
#endif
