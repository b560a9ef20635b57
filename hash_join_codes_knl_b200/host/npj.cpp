// npj -- the reference's ./npj (npj.cpp:929-1125) on a B200: ./npj [#threads] [outer] [inner] [ratio]
// reads ./ik_<inner>.txt ./iv_<inner>.txt ./ok_<outer>.txt ./ov_<outer>.txt, joins, prints the
// reference's "%.4f\n" seconds line (npj.cpp:1114) and its stderr phase lines (npj.cpp:1111),
// then one JSON line with the count and checksums the reference never printed.
#include "hj_host.h"

int main(int argc, char **argv)
{
	info_t_gpu d;
	parse_join_args(argc, argv, &d);
	load_relations(&d);
	int rc = hjb_create(0, &d.ctx);
	if (rc) die("hjb_create", rc, NULL);
	hjb_rel R = {d.inner_keys, d.inner_vals, d.inner_tuples}, S = {d.outer_keys, d.outer_vals, d.outer_tuples};
	hjb_opts o;
	memset(&o, 0, sizeof o);
	o.materialize = 1;
	o.seed = d.seed;
	if ((rc = hjb_npj_host(d.ctx, &R, &S, &o, &d.result))) die("hjb_npj_host", rc, d.ctx);
	const hjb_result *r = &d.result;
	const double total = r->phase_ms[1] + r->phase_ms[2];
	fprintf(stderr, "Phase 1: %5.2f%% (%.4f)\n", total > 0 ? 100.0 * r->phase_ms[1] / total : 0.0, r->phase_ms[1] * 1e-3);
	fprintf(stderr, "Phase 2: %5.2f%% (%.4f)\n", total > 0 ? 100.0 * r->phase_ms[2] / total : 0.0, r->phase_ms[2] * 1e-3);
	printf("%.4f\n", r->seconds);
	print_json("npj", &d, r, 1);
	hjb_destroy(d.ctx);
	return EXIT_SUCCESS;
}
