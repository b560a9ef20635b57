// cpra -- the reference's ./cpra (cpra2.cpp:2017-2231) on 1..8 B200s of one box:
//     HJB_GPUS=G ./cpra [#threads] [outer] [inner]
// The reference starts one pthread per core, gives every thread a contiguous chunk, partitions
// it locally, and lets thread t gather the pieces of the partitions it owns with memcpy
// (cpra2.cpp:1868-1906, printed as "copy:"), the threads meeting at barriers in between
// (cpra2.cpp:1811,1828,1860).  Here the threads are GPUs, one host thread each, same barriers:
//     hjb_cpra_count         chunk g's histogram by owner                    (cpra2.cpp:1783-1811)
//     barrier                the G x G count matrix is complete              (interleave, cpra2.cpp:1813-1827)
//     hjb_cpra_scatter_peer  the GPU-assign pass; its stores ARE the gather: every tuple goes straight
//                            into its owner's buffer over NVLink (TMA bulk copies)
//     barrier                every tuple has landed
//     hjb_cpra_join_local    each GPU joins its share                         (cpra2.cpp:1907-1969)
// The owners' buffers are plain device pointers here (one process, peer access enabled); the
// one-process-per-GPU driver (hash_join_codes_knl_b200/cpra.py) maps them through CUDA IPC instead.
// Prints "copy:\t<seconds>" (the GPU-assign pass incl. exchange) and the seconds line like the
// reference (cpra2.cpp:1984,2208), then the JSON line.
#include "hj_host.h"
#include <chrono>
#include <pthread.h>
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

// what the G worker threads share (the reference's global info[] array, cpra2.cpp:2140-2190)
struct shared_t {
	int G;
	info_t_gpu *d;
	hjb_opts o;
	pthread_barrier_t barrier;
	std::vector<hjb_ctx *> ctx;
	std::vector<uint64_t> counts;          // [src][2][dst]
	std::vector<hjb_recv> recv;            // owner g's receive buffers
	std::vector<hjb_result> result;
	std::vector<double> t_start, t_copy, t_end;
};

struct worker_t {
	shared_t *sh;
	int g;
};

static void *run_gpu(void *arg)
{
	worker_t *w = (worker_t *)arg;
	shared_t *sh = w->sh;
	const int g = w->g, G = sh->G;
	info_t_gpu *d = sh->d;
	hjb_ctx *ctx = sh->ctx[g];
	CUDA_OK(cudaSetDevice(g));
	// chunk g of both relations -> GPU g (the reference's thread_beg / thread_end, cpra2.cpp:1724-1731)
	const uint32_t *hk[2] = {d->inner_keys, d->outer_keys}, *hv[2] = {d->inner_vals, d->outer_vals};
	const size_t tot[2] = {d->inner_tuples, d->outer_tuples};
	uint32_t *dk[2], *dv[2];
	size_t cnt[2];
	for (int r = 0; r < 2; ++r) {
		const size_t beg = tot[r] / G * g;
		cnt[r] = g + 1 == G ? tot[r] - beg : tot[r] / G;
		CUDA_OK(cudaMalloc(&dk[r], (cnt[r] ? cnt[r] : 1) * 4));
		CUDA_OK(cudaMalloc(&dv[r], (cnt[r] ? cnt[r] : 1) * 4));
		CUDA_OK(cudaMemcpy(dk[r], hk[r] + beg, cnt[r] * 4, cudaMemcpyHostToDevice));
		CUDA_OK(cudaMemcpy(dv[r], hv[r] + beg, cnt[r] * 4, cudaMemcpyHostToDevice));
	}
	hjb_rel R = {dk[0], dv[0], cnt[0]}, S = {dk[1], dv[1], cnt[1]};
	// Two passes: the first sizes every workspace (the reference allocates before its timed region,
	// cpra2.cpp:2040-2127), the second is the timed one.
	for (int pass = 0; pass < 2; ++pass) {
		pthread_barrier_wait(&sh->barrier);
		// ---- timed region (inputs resident), as in the reference (cpra2.cpp:1747-1982)
		sh->t_start[g] = now_s();
		int rc;
		if (G == 1) {
			if ((rc = hjb_cpra_join_local(ctx, &R, &S, 0, 1, &sh->o, &sh->result[0]))) die("hjb_cpra_join_local", rc, ctx);
			sh->t_copy[g] = sh->t_start[g];
			sh->t_end[g] = now_s();
			continue;
		}
		uint64_t *mine = &sh->counts[(size_t)g * 2 * G];
		if ((rc = hjb_cpra_count(ctx, &R, &S, G, &sh->o, mine, mine + G))) die("hjb_cpra_count", rc, ctx);
		pthread_barrier_wait(&sh->barrier);                      // the count matrix is complete
		uint64_t recv_n[2] = {0, 0};
		std::vector<uint64_t> base[2];
		for (int r = 0; r < 2; ++r) {
			base[r].assign(G, 0);
			for (int s = 0; s < G; ++s) recv_n[r] += sh->counts[((size_t)s * 2 + r) * G + g];
			for (int dst = 0; dst < G; ++dst)
				for (int s = 0; s < g; ++s) base[r][dst] += sh->counts[((size_t)s * 2 + r) * G + dst];
		}
		if (pass == 0) {
			if ((rc = hjb_cpra_recv_alloc(ctx, recv_n[0] + 1024, recv_n[1] + 1024, &sh->recv[g]))) die("hjb_cpra_recv_alloc", rc, ctx);
			pthread_barrier_wait(&sh->barrier);                  // every owner's buffers exist
		}
		std::vector<void *> pk[2], pv[2];
		for (int dst = 0; dst < G; ++dst) {
			pk[0].push_back(sh->recv[dst].r_keys); pv[0].push_back(sh->recv[dst].r_vals);
			pk[1].push_back(sh->recv[dst].s_keys); pv[1].push_back(sh->recv[dst].s_vals);
		}
		float ms = 0;
		if ((rc = hjb_cpra_scatter_peer(ctx, G, pk[0].data(), pv[0].data(), pk[1].data(), pv[1].data(), base[0].data(),
		                                base[1].data(), &ms)))
			die("hjb_cpra_scatter_peer", rc, ctx);
		pthread_barrier_wait(&sh->barrier);                      // every tuple has landed
		sh->t_copy[g] = now_s();
		hjb_rel Rr = {sh->recv[g].r_keys, sh->recv[g].r_vals, recv_n[0]}, Sr = {sh->recv[g].s_keys, sh->recv[g].s_vals, recv_n[1]};
		if ((rc = hjb_cpra_join_local(ctx, &Rr, &Sr, g, G, &sh->o, &sh->result[g]))) die("hjb_cpra_join_local", rc, ctx);
		sh->t_end[g] = now_s();
	}
	return NULL;
}

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
	shared_t sh;
	sh.G = G;
	sh.d = &d;
	memset(&sh.o, 0, sizeof sh.o);
	sh.o.materialize = 1;
	sh.o.seed = d.seed;
	sh.ctx.resize(G);
	sh.counts.assign((size_t)G * 2 * G, 0);
	sh.recv.resize(G);
	sh.result.resize(G);
	sh.t_start.assign(G, 0); sh.t_copy.assign(G, 0); sh.t_end.assign(G, 0);
	for (int g = 0; g < G; ++g) {
		int rc = hjb_create(g, &sh.ctx[g]);
		if (rc) die("hjb_create", rc, NULL);
		cudaSetDevice(g);
		for (int p = 0; p < G; ++p)
			if (p != g) cudaDeviceEnablePeerAccess(p, 0);       // already-enabled is fine
		cudaGetLastError();
		memset(&sh.result[g], 0, sizeof(hjb_result));
	}
	pthread_barrier_init(&sh.barrier, NULL, (unsigned)G);
	const double t_e2e = now_s();
	std::vector<pthread_t> th(G);
	std::vector<worker_t> wk(G);
	for (int g = 0; g < G; ++g) {
		wk[g].sh = &sh;
		wk[g].g = g;
		pthread_create(&th[g], NULL, run_gpu, &wk[g]);          // one host thread per GPU, like the reference's pthread per core
	}
	for (int g = 0; g < G; ++g) pthread_join(th[g], NULL);
	hjb_result total;
	memset(&total, 0, sizeof total);
	double t0 = sh.t_start[0], t1 = sh.t_end[0], copy_s = 0;
	for (int g = 0; g < G; ++g) {
		total.count += sh.result[g].count;
		total.sum_key += sh.result[g].sum_key;
		total.sum_outer += sh.result[g].sum_outer;
		total.sum_inner += sh.result[g].sum_inner;
		total.kernel_launches += sh.result[g].kernel_launches;
		total.partitions += sh.result[g].partitions;
		if (sh.t_start[g] < t0) t0 = sh.t_start[g];
		if (sh.t_end[g] > t1) t1 = sh.t_end[g];
		if (sh.t_copy[g] - sh.t_start[g] > copy_s) copy_s = sh.t_copy[g] - sh.t_start[g];
	}
	total.seconds = t1 - t0;
	total.seconds_e2e = now_s() - t_e2e;
	printf("copy:\t%lf\n", copy_s);
	printf("%lf\n", total.seconds);
	print_json("cpra", &d, &total, G);
	pthread_barrier_destroy(&sh.barrier);
	for (int g = 0; g < G; ++g) hjb_destroy(sh.ctx[g]);
	return EXIT_SUCCESS;
}
