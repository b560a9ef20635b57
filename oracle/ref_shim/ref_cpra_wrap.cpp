/* ref_cpra_wrap.cpp -- exposes the REFERENCE's own CPRA functions (compiled from
 * /root/reference/cpra2.cpp where it lies, nothing copied) under C names.
 * TEST INFRASTRUCTURE ONLY; built by oracle/Makefile into oracle/_ref/ (git-ignored). */
#define main hjref_cpra_main
#include "cpra2.cpp"
#undef main

extern "C" {

int hjref_odd_prime(uint64_t x) { return odd_prime(x); }

/* AVX-512 histogram / partition (cpra2.cpp:801-880, 882-1075) */
void hjref_histogram(const uint32_t *keys, size_t size, uint32_t *counts,
                     uint32_t factor, size_t partitions)
{
	histogram(keys, size, counts, factor, partitions);
}

void hjref_partition(const uint32_t *keys, const uint32_t *vals, size_t size,
                     const uint32_t *counts, uint32_t *keys_out, uint32_t *vals_out,
                     uint32_t factor, size_t partitions)
{
	partition(keys, vals, size, counts, keys_out, vals_out, factor, partitions);
}

/* scalar twins the oracle restates (cpra2.cpp:640-796) */
void hjref_histogram_s(const uint32_t *keys, size_t size, uint32_t *counts,
                       uint32_t factor, size_t partitions)
{
	histogram_s(keys, size, counts, factor, partitions);
}

void hjref_partition_s(const uint32_t *keys, const uint32_t *vals, size_t size,
                       const uint32_t *counts, uint32_t *keys_out, uint32_t *vals_out,
                       uint32_t factor, size_t partitions)
{
	partition_s(keys, vals, size, counts, keys_out, vals_out, factor, partitions);
}

void hjref_dh_build(const uint32_t *keys, const uint32_t *vals, size_t size,
                    uint64_t *table, size_t buckets, const uint32_t *factor)
{
	build(keys, vals, size, table, buckets, factor, 0);
}

void hjref_dh_build_s(const uint32_t *keys, const uint32_t *vals, size_t size,
                      uint64_t *table, size_t buckets, const uint32_t *factor)
{
	build_s(keys, vals, size, table, buckets, factor, 0);
}

struct hjref_run_arg {
	info_t_hj *info;
};

/* The whole of run_hj (cpra2.cpp:1697-1986) on ONE thread: with threads == 1 the thread owns
 * every partition and passes flush = 1 on the last one, so the materialised rows are complete
 *  and final_offsets[0] is the match count.  Output columns must
 * hold block_limit*65536 entries, 64-byte aligned.  The reference's stdout line ("copy:") is
 * the caller's to swallow. */
size_t hjref_cpra_join(uint32_t *rk, uint32_t *rv, size_t nr,
                       uint32_t *sk, uint32_t *sv, size_t ns, int seed,
                       uint32_t *keys_out, uint32_t *svals_out, uint32_t *rvals_out,
                       size_t block_limit)
{
	NUM_PARTITIONS = 4096;                                           /* cpra2.cpp:2023 */
	static info_t_hj info;
	memset(&info, 0, sizeof info);
	pthread_barrier_t barrier[64];
	for (int b = 0; b != 64; ++b) pthread_barrier_init(&barrier[b], NULL, 1);
	size_t final_offsets[1] = { 0 };
	volatile size_t block_counter = 0, close_gaps_counter = 0;
	uint32_t *rk2 = (uint32_t *)mamalloc((nr + 64) * sizeof(uint32_t));
	uint32_t *rv2 = (uint32_t *)mamalloc((nr + 64) * sizeof(uint32_t));
	uint32_t *sk2 = (uint32_t *)mamalloc((ns + 64) * sizeof(uint32_t));
	uint32_t *sv2 = (uint32_t *)mamalloc((ns + 64) * sizeof(uint32_t));
	info.thread = 0;
	info.threads = 1;
	info.seed = seed;
	info.block_limit = block_limit;
	info.outer_tuples = ns;
	info.inner_tuples = nr;
	info.inner_keys[0] = rk; info.inner_keys[1] = rk2;
	info.inner_vals[0] = rv; info.inner_vals[1] = rv2;
	info.outer_keys[0] = sk; info.outer_keys[1] = sk2;
	info.outer_vals[0] = sv; info.outer_vals[1] = sv2;
	info.final_offsets = final_offsets;
	info.join_keys = keys_out;
	info.join_inner_vals = rvals_out;
	info.join_outer_vals = svals_out;
	info.thread_factor = 0x9e3779b1u;
	info.hash_table_load = 0.4;                                      /* cpra2.cpp:2031-2034 */
	info.hash_table_limit = 300;
	info.buffer_size = 256;
	info.block_size = 256 * 256;
	info.block_counter = &block_counter;
	info.close_gaps_counter = &close_gaps_counter;
	info.barrier = barrier;
	pthread_create(&info.id, NULL, run_hj, (void *)&info);
	pthread_join(info.id, NULL);
	for (int b = 0; b != 64; ++b) pthread_barrier_destroy(&barrier[b]);
	free(rk2); free(rv2); free(sk2); free(sv2);
	return final_offsets[0];
}

} /* extern "C" */
