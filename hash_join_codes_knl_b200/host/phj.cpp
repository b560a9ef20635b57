// phj -- the reference's ./phj (phj.cpp:1959-2231) on a B200: ./phj [#threads] [outer] [inner] [ratio].
// The reference prints "max\tinfo[0]\tinfo[128]" seconds (phj.cpp:2197, the last an out-of-bounds
// read); here the three fields are total, partitioning, join.  Unlike the shipped phj.cpp this
// one performs the join (phj.cpp:1869-1924 is commented out there).
#include "hj_host.h"

int main(int argc, char **argv)
{
	info_t_gpu d;
	parse_join_args(argc, argv, &d);
	load_relations(&d);
	run_join_single(&d, "hjb_phj", hjb_phj_device, hjb_phj_host);
	const hjb_result *r = &d.result;
	printf("%lf\t%lf\t%lf\t\n", r->seconds, (r->phase_ms[0] + r->phase_ms[1]) * 1e-3, r->phase_ms[4] * 1e-3);
	print_json("phj", &d, r, 1);
	hjb_destroy(d.ctx);
	return EXIT_SUCCESS;
}
