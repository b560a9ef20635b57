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
	run_join_single(&d, "hjb_npj", hjb_npj_device, hjb_npj_host);
	const hjb_result *r = &d.result;
	const double total = r->phase_ms[1] + r->phase_ms[2];
	fprintf(stderr, "Phase 1: %5.2f%% (%.4f)\n", total > 0 ? 100.0 * r->phase_ms[1] / total : 0.0, r->phase_ms[1] * 1e-3);
	fprintf(stderr, "Phase 2: %5.2f%% (%.4f)\n", total > 0 ? 100.0 * r->phase_ms[2] / total : 0.0, r->phase_ms[2] * 1e-3);
	printf("%.4f\n", r->seconds);
	print_json("npj", &d, r, 1);
	hjb_destroy(d.ctx);
	return EXIT_SUCCESS;
}
