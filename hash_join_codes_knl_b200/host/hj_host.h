// hj_host.h -- what the four host programs share, in the style of the reference's hj.h: one
// plain struct that carries a run's arguments and buffers (info_t_hj, hj.h:1-72) and a few
// free functions.  The programs keep the reference's command lines
//     ./npj|./phj|./cpra [#threads] [outer tuples] [inner tuples] [ratio]      (npj.cpp:932-935)
//     ./write            [#threads] [outer tuples] [inner tuples] [selc] [zipf] (write.cpp:1680-1686)
// and the four raw uint32 relation files in the working directory (write.cpp:1824-1865), and
// drive the CUDA engine through the C ABI of include/hjb200.h only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "hjb200.h"

typedef struct info_t_gpu {
	int threads;              // argv[1]: kept for CLI compatibility; host loader threads, unused by the kernels
	size_t outer_tuples;      // argv[2] = |S| (probe side)
	size_t inner_tuples;      // argv[3] = |R| (build side)
	double ratio;             // argv[4]: the reference's DDR/MCDRAM split -- accepted and ignored (one memory tier)
	int gpus;                 // HJB_GPUS (cpra only)
	uint32_t seed;            // HJB_SEED
	uint32_t *inner_keys, *inner_vals, *outer_keys, *outer_vals;    // pinned host columns
	hjb_ctx *ctx;
	hjb_result result;
} info_t_gpu;

static inline void die(const char *what, int rc, const hjb_ctx *ctx)
{
	fprintf(stderr, "%s failed (%d): %s\n", what, rc, hjb_last_error(ctx));
	exit(1);
}

static inline void parse_join_args(int argc, char **argv, info_t_gpu *d)
{
	memset(d, 0, sizeof *d);
	d->threads = argc > 1 ? atoi(argv[1]) : 1;
	d->outer_tuples = argc > 2 ? (size_t)atoll(argv[2]) : 200 * 1000 * 1000;   // defaults of npj.cpp:933-934
	d->inner_tuples = argc > 3 ? (size_t)atoll(argv[3]) : 200 * 1000 * 1000;
	d->ratio = argc > 4 ? atof(argv[4]) : 1;
	d->gpus = getenv("HJB_GPUS") ? atoi(getenv("HJB_GPUS")) : 1;
	d->seed = getenv("HJB_SEED") ? (uint32_t)strtoul(getenv("HJB_SEED"), NULL, 0) : 0;
}

// main()'s fopen/fread block (npj.cpp:1013-1039) with the checks it lacks; pinned so the copy in runs at link speed
static inline void load_relations(info_t_gpu *d)
{
	uint32_t **cols[4] = {&d->inner_keys, &d->inner_vals, &d->outer_keys, &d->outer_vals};
	const size_t n[4] = {d->inner_tuples, d->inner_tuples, d->outer_tuples, d->outer_tuples};
	for (int c = 0; c < 4; ++c)
		if (cudaHostAlloc((void **)cols[c], (n[c] ? n[c] : 1) * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) {
			fprintf(stderr, "cannot allocate pinned host memory for %zu tuples\n", n[c]);
			exit(1);
		}
	int rc = hjb_relation_read(".", 0, d->inner_tuples, d->inner_keys, d->inner_vals);
	if (rc) {
		fprintf(stderr, "cannot read ./ik_%zu.txt / ./iv_%zu.txt (need exactly %zu bytes each)\n", d->inner_tuples,
		        d->inner_tuples, d->inner_tuples * 4);
		exit(1);
	}
	rc = hjb_relation_read(".", 1, d->outer_tuples, d->outer_keys, d->outer_vals);
	if (rc) {
		fprintf(stderr, "cannot read ./ok_%zu.txt / ./ov_%zu.txt (need exactly %zu bytes each)\n", d->outer_tuples,
		        d->outer_tuples, d->outer_tuples * 4);
		exit(1);
	}
}

// npj / phj on one GPU.  The reference times the join with both relations already in the memory its
// threads read (npj.cpp:861-918); so: copy in, one untimed pass that sizes the workspaces (the
// reference allocates before its timed region), then the timed pass on the resident columns
// (result.seconds and the phase times).  result.seconds_e2e is a further pass through the host entry
// point -- copy in, join, rows back to host memory -- which must reproduce the same result.
typedef int (*hjb_join_fn)(hjb_ctx *, const hjb_rel *, const hjb_rel *, const hjb_opts *, hjb_result *);
static inline void run_join_single(info_t_gpu *d, const char *name, hjb_join_fn on_device, hjb_join_fn on_host)
{
	setenv("HJB_GRAPHS", "0", 0);          // a one-shot program gains nothing from graph replay, and eager launches keep the phase times
	int rc = hjb_create(0, &d->ctx);
	if (rc) die("hjb_create", rc, NULL);
	uint32_t *dev[4];
	const uint32_t *host[4] = {d->inner_keys, d->inner_vals, d->outer_keys, d->outer_vals};
	const size_t n[4] = {d->inner_tuples, d->inner_tuples, d->outer_tuples, d->outer_tuples};
	for (int c = 0; c < 4; ++c)
		if (cudaMalloc((void **)&dev[c], (n[c] ? n[c] : 1) * 4) != cudaSuccess ||
		    cudaMemcpy(dev[c], host[c], n[c] * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
			fprintf(stderr, "cannot place %zu tuples in device memory\n", n[c]);
			exit(1);
		}
	hjb_rel dR = {dev[0], dev[1], d->inner_tuples}, dS = {dev[2], dev[3], d->outer_tuples};
	hjb_rel hR = {d->inner_keys, d->inner_vals, d->inner_tuples}, hS = {d->outer_keys, d->outer_vals, d->outer_tuples};
	hjb_opts o;
	memset(&o, 0, sizeof o);
	o.materialize = 1;
	o.seed = d->seed;
	hjb_result warm, e2e;
	if ((rc = on_device(d->ctx, &dR, &dS, &o, &warm))) die(name, rc, d->ctx);
	if ((rc = on_device(d->ctx, &dR, &dS, &o, &d->result))) die(name, rc, d->ctx);
	if ((rc = on_host(d->ctx, &hR, &hS, &o, &e2e))) die(name, rc, d->ctx);
	if (e2e.count != d->result.count || e2e.sum_key != d->result.sum_key || e2e.sum_outer != d->result.sum_outer ||
	    e2e.sum_inner != d->result.sum_inner) {
		fprintf(stderr, "%s: the host entry point and the device entry point disagree\n", name);
		exit(1);
	}
	d->result.seconds_e2e = e2e.seconds_e2e;
	for (int c = 0; c < 4; ++c) cudaFree(dev[c]);
}

// the machine-readable line that follows the reference's own stdout line
static inline void print_json(const char *algo, const info_t_gpu *d, const hjb_result *r, int gpus)
{
	const double tuples = (double)d->inner_tuples + (double)d->outer_tuples;
	printf("{\"algorithm\": \"%s\", \"gpus\": %d, \"inner_tuples\": %zu, \"outer_tuples\": %zu, \"join_tuples\": %llu, "
	       "\"sum_key\": %llu, \"sum_outer\": %llu, \"sum_inner\": %llu, \"seconds\": %.6f, \"seconds_e2e\": %.6f, "
	       "\"tuples_per_second\": %.4e, \"kernel_launches\": %u, \"partitions\": %u}\n",
	       algo, gpus, d->inner_tuples, d->outer_tuples, (unsigned long long)r->count, (unsigned long long)r->sum_key,
	       (unsigned long long)r->sum_outer, (unsigned long long)r->sum_inner, r->seconds, r->seconds_e2e,
	       r->seconds > 0 ? tuples / r->seconds : 0.0, r->kernel_launches, r->partitions);
}
