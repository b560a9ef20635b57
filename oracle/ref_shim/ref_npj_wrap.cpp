/* ref_npj_wrap.cpp -- exposes the REFERENCE's own NPJ functions (compiled from
 * /root/reference/npj.cpp where it lies, nothing copied) under C names so the tests can run
 * them on chosen inputs and pin oracle/hj_oracle.c against them.
 * TEST INFRASTRUCTURE ONLY; built by oracle/Makefile into oracle/_ref/ (git-ignored). */
#define main hjref_npj_main
#define inner_keys_1 inner_keys
#define inner_vals_1 inner_vals
#define outer_keys_1 outer_keys
#define outer_vals_1 outer_vals
#include "npj.cpp"
#undef main

extern "C" {

uint32_t hjref_rand32_stream(uint32_t seed, uint32_t *out, size_t n)
{
	rand32_t *g = rand32_init(seed);
	for (size_t i = 0; i != n; ++i) out[i] = rand32_next(g);
	free(g);
	return n ? out[n - 1] : 0;
}

void hjref_shuffle(uint32_t *data, size_t size, uint32_t seed)
{
	rand32_t *g = rand32_init(seed);
	shuffle(data, size, g);
	free(g);
}

/* buckets must be prime (npj.cpp:576); table must hold `buckets` zeroed words */
void hjref_unique(uint32_t *keys, size_t size, uint32_t *table, size_t buckets,
                  uint32_t factor, uint32_t seed)
{
	rand32_t *g = rand32_init(seed);
	unique(keys, size, table, buckets, factor, 0, g);
	free(g);
}

void hjref_npj_build(const uint32_t *keys, const uint32_t *vals, size_t size,
                     uint64_t *table, size_t buckets, uint32_t factor)
{
	/* ratio = 0: every slot index lives in `table` (npj.cpp:195,202) */
	build(keys, vals, size, table, table, buckets, factor, 0, 0.0);
}

/* One-thread NPJ exactly as run() strings it together (npj.cpp:861-911): set, build,
 * the AVX-512 probe, close_gaps.  Output columns must hold block_limit*65536 entries and be
 * 64-byte aligned.  Returns the match count (close_gaps' return value). */
size_t hjref_npj_join(const uint32_t *rk, const uint32_t *rv, size_t nr,
                      const uint32_t *sk, const uint32_t *sv, size_t ns,
                      size_t buckets, uint32_t factor,
                      uint32_t *keys_out, uint32_t *svals_out, uint32_t *rvals_out,
                      size_t block_limit)
{
	const size_t block_size = 256 * 256;                             /* npj.cpp:945 */
	uint64_t *table = (uint64_t *)mamalloc(buckets * sizeof(uint64_t));
	volatile size_t counters[2] = { 0, 0 };
	set(table, buckets, 0);
	build(rk, rv, nr, table, table, buckets, factor, 0, 0.0);
	size_t final_offset = probe(sk, sv, ns, table, table, buckets, factor, 0,
	                            keys_out, svals_out, rvals_out,
	                            block_size, block_limit, &counters[0], 0.0);
	size_t count = close_gaps(keys_out, svals_out, rvals_out, &final_offset, 1,
	                          block_size, &counters[1]);
	free(table);
	return count;
}

} /* extern "C" */
