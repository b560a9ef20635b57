// cpra -- the reference's ./cpra (cpra2.cpp:2017-2231) on 1..8 B200s of one box:
//     HJB_GPUS=G ./cpra [#threads] [outer] [inner]
// The reference gives every thread a contiguous chunk, partitions it locally, and lets thread t
// gather the pieces of the partitions it owns with memcpy (cpra2.cpp:1868-1906, printed as
// "copy:").  Here the threads are GPUs: chunk g lives on GPU g, hjb_cpra_split partitions it by
// owner, the gather is a peer copy over NVLink (this single-process program uses
// cudaMemcpyPeerAsync; the one-process-per-GPU path uses NCCL, hash_join_codes_knl_b200/cpra.py),
// and every GPU joins what it received.  Prints "copy:\t<seconds>" and the seconds line like the
// reference (cpra2.cpp:1984,2208), then the JSON line.
#include "hj_host.h"
#include <chrono>
#include <vector>

static double now_s()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

#define CUDA_OK(x)                                                                        \
	do {                                                                                  \
		cudaError_t e_ = (x);                                                             \
		if (e_ != cudaSuccess) {                                                          \
			fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                      \
			exit(1);                                                                      \
		}                                                                                 \
	} while (0)

int main(int argc, char **argv)
{
	info_t_gpu d;
	parse_join_args(argc, argv, &d);
	const int G = d.gpus;
	int ndev = 0;
	cudaGetDeviceCount(&ndev);
	if (G < 1 || G > ndev || (G & (G - 1))) {
		fprintf(stderr, "HJB_GPUS=%d: need a power of two <= %d visible GPUs\n", G, ndev);
		return 1;
	}
	load_relations(&d);
	std::vector<hjb_ctx *> ctx(G);
	for (int g = 0; g < G; ++g) {
		int rc = hjb_create(g, &ctx[g]);
		if (rc) die("hjb_create", rc, NULL);
		cudaSetDevice(g);
		for (int p = 0; p < G; ++p)
			if (p != g) cudaDeviceEnablePeerAccess(p, 0);       // already-enabled is fine
		cudaGetLastError();
	}
	hjb_opts o;
	memset(&o, 0, sizeof o);
	o.materialize = 1;
	o.seed = d.seed;
	// chunk g of both relations -> GPU g (the reference's thread_beg/thread_end chunks, cpra2.cpp:1724-1731)
	std::vector<uint32_t *> dk[2], dv[2];
	std::vector<size_t> beg[2], cnt[2];
	const uint32_t *hk[2] = {d.inner_keys, d.outer_keys}, *hv[2] = {d.inner_vals, d.outer_vals};
	const size_t tot[2] = {d.inner_tuples, d.outer_tuples};
	const double t_e2e = now_s();
	for (int r = 0; r < 2; ++r) {
		dk[r].resize(G); dv[r].resize(G); beg[r].resize(G); cnt[r].resize(G);
		for (int g = 0; g < G; ++g) {
			beg[r][g] = tot[r] / G * g;
			cnt[r][g] = g + 1 == G ? tot[r] - beg[r][g] : tot[r] / G;
			CUDA_OK(cudaSetDevice(g));
			CUDA_OK(cudaMalloc(&dk[r][g], (cnt[r][g] ? cnt[r][g] : 1) * 4));
			CUDA_OK(cudaMalloc(&dv[r][g], (cnt[r][g] ? cnt[r][g] : 1) * 4));
			CUDA_OK(cudaMemcpyAsync(dk[r][g], hk[r] + beg[r][g], cnt[r][g] * 4, cudaMemcpyHostToDevice, 0));
			CUDA_OK(cudaMemcpyAsync(dv[r][g], hv[r] + beg[r][g], cnt[r][g] * 4, cudaMemcpyHostToDevice, 0));
		}
	}
	for (int g = 0; g < G; ++g) { CUDA_OK(cudaSetDevice(g)); CUDA_OK(cudaDeviceSynchronize()); }
	// ---- timed region (inputs resident), as in the reference (cpra2.cpp:1747-1982)
	const double t0 = now_s();
	std::vector<hjb_split> sp(G);
	for (int g = 0; g < G; ++g) {
		hjb_rel R = {dk[0][g], dv[0][g], cnt[0][g]}, S = {dk[1][g], dv[1][g], cnt[1][g]};
		int rc = hjb_cpra_split(ctx[g], &R, &S, G, &o, &sp[g]);
		if (rc) die("hjb_cpra_split", rc, ctx[g]);
	}
	// gather: owner g pulls its piece from every source GPU
	const double t_copy0 = now_s();
	std::vector<uint32_t *> rk(G), rv(G), sk(G), sv(G);
	std::vector<size_t> rn(G, 0), sn(G, 0);
	for (int g = 0; g < G; ++g) {
		for (int s = 0; s < G; ++s) {
			rn[g] += sp[s].r_offsets[g + 1] - sp[s].r_offsets[g];
			sn[g] += sp[s].s_offsets[g + 1] - sp[s].s_offsets[g];
		}
		CUDA_OK(cudaSetDevice(g));
		CUDA_OK(cudaMalloc(&rk[g], (rn[g] ? rn[g] : 1) * 4)); CUDA_OK(cudaMalloc(&rv[g], (rn[g] ? rn[g] : 1) * 4));
		CUDA_OK(cudaMalloc(&sk[g], (sn[g] ? sn[g] : 1) * 4)); CUDA_OK(cudaMalloc(&sv[g], (sn[g] ? sn[g] : 1) * 4));
		size_t ro = 0, so = 0;
		for (int s = 0; s < G; ++s) {
			const size_t rc_ = sp[s].r_offsets[g + 1] - sp[s].r_offsets[g], sc_ = sp[s].s_offsets[g + 1] - sp[s].s_offsets[g];
			CUDA_OK(cudaMemcpyPeerAsync(rk[g] + ro, g, sp[s].r_keys + sp[s].r_offsets[g], s, rc_ * 4, 0));
			CUDA_OK(cudaMemcpyPeerAsync(rv[g] + ro, g, sp[s].r_vals + sp[s].r_offsets[g], s, rc_ * 4, 0));
			CUDA_OK(cudaMemcpyPeerAsync(sk[g] + so, g, sp[s].s_keys + sp[s].s_offsets[g], s, sc_ * 4, 0));
			CUDA_OK(cudaMemcpyPeerAsync(sv[g] + so, g, sp[s].s_vals + sp[s].s_offsets[g], s, sc_ * 4, 0));
			ro += rc_;
			so += sc_;
		}
	}
	for (int g = 0; g < G; ++g) { CUDA_OK(cudaSetDevice(g)); CUDA_OK(cudaDeviceSynchronize()); }
	const double copy_s = now_s() - t_copy0;
	hjb_result total;
	memset(&total, 0, sizeof total);
	for (int g = 0; g < G; ++g) {
		hjb_rel R = {rk[g], rv[g], rn[g]}, S = {sk[g], sv[g], sn[g]};
		hjb_result r;
		int rc = hjb_cpra_join_local(ctx[g], &R, &S, g, G, &o, &r);   // returns after its stream drained; GPUs run one after
		if (rc) die("hjb_cpra_join_local", rc, ctx[g]);               // the other here -- the NCCL path runs them concurrently
		total.count += r.count;
		total.sum_key += r.sum_key;
		total.sum_outer += r.sum_outer;
		total.sum_inner += r.sum_inner;
		total.kernel_launches += r.kernel_launches;
		total.partitions += r.partitions;
	}
	total.seconds = now_s() - t0;
	total.seconds_e2e = now_s() - t_e2e;
	printf("copy:\t%lf\n", copy_s);
	printf("%lf\n", total.seconds);
	print_json("cpra", &d, &total, G);
	for (int g = 0; g < G; ++g) hjb_destroy(ctx[g]);
	return EXIT_SUCCESS;
}
