// write -- the reference's ./write (write.cpp:1677-1888): ./write [#threads] [outer] [inner] [selc] [zipf]
// generates the two relations and writes ik_<inner>.txt iv_<inner>.txt ok_<outer>.txt
// ov_<outer>.txt (raw uint32, write.cpp:1824-1865) into the working directory.
// Semantics follow the intact generator (cpra2.cpp:1578-1696): distinct non-zero 32-bit build
// keys; probe keys cover every build key once, the rest uniform picks; payload = key * odd
// factor.  Deviations, all documented in DESIGN.md: seeded (HJB_SEED, default 42) instead of
// time(NULL) (write.cpp:1737); `selc` is the fraction of probe TUPLES that find a partner and
// `zipf` really skews the probe keys (the reference's Zipf knob is inert, write.cpp:1553-1571);
// generated on the GPU (csrc/gen.cu) because 2^31 tuples take minutes with a serial shuffle.
#include "hj_host.h"

int main(int argc, char **argv)
{
	info_t_gpu d;
	parse_join_args(argc, argv, &d);
	const double selc = argc > 4 ? atof(argv[4]) : 1.0, zipf = argc > 5 ? atof(argv[5]) : 0.0;
	const uint32_t seed = getenv("HJB_SEED") ? (uint32_t)strtoul(getenv("HJB_SEED"), NULL, 0) : 42;
	int rc = hjb_create(0, &d.ctx);
	if (rc) die("hjb_create", rc, NULL);
	const size_t n[2] = {d.inner_tuples, d.outer_tuples};
	const size_t domain = d.inner_tuples < d.outer_tuples ? d.inner_tuples : d.outer_tuples;   // inner_distinct, cpra2.cpp:2024-2026
	for (int outer = 0; outer < 2; ++outer) {
		hjb_gen g;
		memset(&g, 0, sizeof g);
		g.tuples = g.total = n[outer];
		g.domain = domain ? domain : 1;
		g.seed = seed;
		g.order_seed = seed * 2 + 1 + outer;
		g.payload_factor = outer ? 0xDF56B8FBu : 0x6587F97Du;
		g.theta = zipf;
		g.selectivity = selc;
		if (!outer) g.kind = n[0] <= domain ? 0 : 1;                   // R: unique, or duplicates when |R| > |S|
		else if (selc >= 1.0 && zipf == 0.0) g.kind = n[1] == domain ? 0 : 1;
		else g.kind = 2;
		uint32_t *dk, *dv, *hk, *hv;
		const size_t bytes = (n[outer] ? n[outer] : 1) * 4;
		if (cudaMalloc(&dk, bytes) || cudaMalloc(&dv, bytes) || cudaHostAlloc(&hk, bytes, 0) || cudaHostAlloc(&hv, bytes, 0)) {
			fprintf(stderr, "out of memory for %zu tuples\n", n[outer]);
			return 1;
		}
		if (n[outer] && (rc = hjb_generate(d.ctx, &g, dk, dv))) die("hjb_generate", rc, d.ctx);
		hjb_synchronize(d.ctx);
		cudaMemcpy(hk, dk, n[outer] * 4, cudaMemcpyDeviceToHost);
		cudaMemcpy(hv, dv, n[outer] * 4, cudaMemcpyDeviceToHost);
		if ((rc = hjb_relation_write(".", outer, n[outer], hk, hv))) die("hjb_relation_write", rc, d.ctx);
		cudaFree(dk); cudaFree(dv); cudaFreeHost(hk); cudaFreeHost(hv);
	}
	hjb_destroy(d.ctx);
	return EXIT_SUCCESS;
}
