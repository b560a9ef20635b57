/* ref_phj_wrap.cpp -- exposes the REFERENCE's own PHJ kernels (compiled from
 * /root/reference/phj.cpp where it lies, nothing copied) under C names.  The shipped
 * run_hj comments its join phase out (phj.cpp:1869-1924), so what can be pinned here are the
 * kernels that phase would call: local histogram / partition and the double-hash
 * build / probe, strung together for ONE thread the way the commented block does.
 * TEST INFRASTRUCTURE ONLY; built by oracle/Makefile into oracle/_ref/ (git-ignored). */
#define main hjref_phj_main
#include "phj.cpp"
#undef main

extern "C" {

void hjref_phj_histogram(const uint32_t *keys, size_t size, uint32_t *counts,
                         uint32_t factor, size_t partitions)
{
	histogram(keys, size, counts, factor, partitions);
}

void hjref_phj_partition(const uint32_t *keys, const uint32_t *vals, size_t size,
                         const uint32_t *counts, uint32_t *keys_out, uint32_t *vals_out,
                         uint32_t factor, size_t partitions)
{
	partition(keys, vals, size, counts, keys_out, vals_out, factor, partitions);
}

/* One radix pass (histogram + partition, phj.cpp:1832-1836) followed by the per-partition
 * build + probe loop of phj.cpp:1880-1923 with the reference's own table policy constants
 * (load 0.4, odd-prime buckets).  Single thread, so blocks are consecutive and the returned
 * offset is the match count; flush = 1 on the last partition.  Inputs and outputs 64-byte
 * aligned; outputs hold block_limit*65536 entries. */
size_t hjref_phj_local_join(const uint32_t *rk, const uint32_t *rv, size_t nr,
                            const uint32_t *sk, const uint32_t *sv, size_t ns,
                            size_t fanout, uint32_t part_factor, const uint32_t *join_factors,
                            uint32_t *keys_out, uint32_t *svals_out, uint32_t *rvals_out,
                            size_t block_limit)
{
	BUFFER_SIZE = 64;                                                /* phj.cpp:1967 */
	const size_t buffer_size = 256, block_size = 256 * 256;          /* phj.cpp:1978-1979 */
	uint32_t *rk2 = (uint32_t *)mamalloc((nr + 64) * sizeof(uint32_t));
	uint32_t *rv2 = (uint32_t *)mamalloc((nr + 64) * sizeof(uint32_t));
	uint32_t *sk2 = (uint32_t *)mamalloc((ns + 64) * sizeof(uint32_t));
	uint32_t *sv2 = (uint32_t *)mamalloc((ns + 64) * sizeof(uint32_t));
	uint32_t *rc = (uint32_t *)malloc(fanout * sizeof(uint32_t));
	uint32_t *sc = (uint32_t *)malloc(fanout * sizeof(uint32_t));
	histogram(rk, nr, rc, part_factor, fanout);
	partition(rk, rv, nr, rc, rk2, rv2, part_factor, fanout);
	histogram(sk, ns, sc, part_factor, fanout);
	partition(sk, sv, ns, sc, sk2, sv2, part_factor, fanout);
	uint32_t *keys_buf = (uint32_t *)mamalloc((buffer_size + 16) * sizeof(uint32_t));
	uint32_t *vals_buf = (uint32_t *)mamalloc((buffer_size + 16) * sizeof(uint32_t));
	uint32_t *tabs_buf = (uint32_t *)mamalloc((buffer_size + 16) * sizeof(uint32_t));
	volatile size_t counter = 0;
	size_t offset = __sync_fetch_and_add(&counter, 1) * block_size;
	uint64_t *table = NULL;
	size_t i = 0, o = 0;
	for (size_t p = 0; p != fanout; ++p) {
		size_t buckets = rc[p] / 0.4;
		for (buckets |= 1; !odd_prime(buckets); buckets += 2);
		table = (uint64_t *)realloc(table, buckets * sizeof(uint64_t));
		build(&rk2[i], &rv2[i], rc[p], table, buckets, join_factors, 0);
		i += rc[p];
		offset = probe(&sk2[o], &sv2[o], sc[p], table, buckets, join_factors, 0,
		               keys_buf, vals_buf, tabs_buf, keys_out, svals_out, rvals_out,
		               offset, buffer_size, block_size, block_limit, &counter, p + 1 == fanout);
		o += sc[p];
	}
	free(table); free(keys_buf); free(vals_buf); free(tabs_buf);
	free(rk2); free(rv2); free(sk2); free(sv2); free(rc); free(sc);
	return offset;
}

} /* extern "C" */
