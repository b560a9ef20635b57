// peer_copy_harness.cu -- the product's k_peer_copy (csrc/stage.cu, linked in) on hand-made run descriptors, devices 0 -> 1
// of ONE process (peer access, no IPC): per-CTA throughput of the copy kernel without the rest of the step.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I hash_join_codes_knl_b200/csrc peer_copy_harness.cu \
//        ../../hash_join_codes_knl_b200/csrc/stage.cu -o _build/peer_copy_harness
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <string>
#include <cstring>
#include <unistd.h>
#include <sys/wait.h>
#include "hj_internal.h"
using namespace hjb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__global__ void k_fill(uint32_t *p, size_t n, uint32_t salt)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
		uint32_t x = (uint32_t)i * 2654435761u + salt;
		x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
		p[i] = x;
	}
}

static std::string hex(const void *p, size_t n)
{
	std::string o;
	char b[3];
	for (size_t i = 0; i < n; ++i) { snprintf(b, 3, "%02x", ((const unsigned char *)p)[i]); o += b; }
	return o;
}
static void unhex(const char *s, void *p, size_t n)
{
	for (size_t i = 0; i < n; ++i) { unsigned v; sscanf(s + 2 * i, "%2x", &v); ((unsigned char *)p)[i] = (unsigned char)v; }
}

// usage: harness            same-process peer access
//        harness ipc        the parent allocates on device 1 and exports legacy IPC handles, a child process copies into them
int main(int argc, char **argv)
{
	int nd = 0;
	CK(cudaGetDeviceCount(&nd));
	const int peer = nd > 1 ? 1 : 0;
	if (argc > 1 && !strcmp(argv[1], "ipc")) {
		const size_t rows0 = (size_t)((1u << 28) / 512 + 37) * 512 + 64 * 512;
		void *a, *b;
		CK(cudaSetDevice(peer));
		CK(cudaMalloc(&a, rows0 * 4)); CK(cudaMalloc(&b, rows0 * 4));
		cudaIpcMemHandle_t ha, hb;
		CK(cudaIpcGetMemHandle(&ha, a)); CK(cudaIpcGetMemHandle(&hb, b));
		const std::string sa = hex(&ha, sizeof ha), sb = hex(&hb, sizeof hb);
		pid_t pid = fork();                       // the child execs at once: no CUDA state is used across the fork
		if (pid == 0) { execl(argv[0], argv[0], "child", sa.c_str(), sb.c_str(), (char *)nullptr); _exit(127); }
		int st = 0;
		waitpid(pid, &st, 0);
		return st;
	}
	const bool child = argc > 3 && !strcmp(argv[1], "child");
	const int gbits = 1, G = 2, F = 512, nsub = F / G;
	const uint32_t per_digit = (1u << 28) / F + 37;            // rows per run (not a multiple of 32)
	const size_t rows = (size_t)per_digit * F + 64 * F;
	uint32_t *sk, *sv, *dk[2], *dv[2];
	CK(cudaSetDevice(0));
	if (peer && !(argc > 3 && !strcmp(argv[1], "child"))) CK(cudaDeviceEnablePeerAccess(peer, 0));
	CK(cudaMalloc(&sk, rows * 4)); CK(cudaMalloc(&sv, rows * 4));
	CK(cudaMemset(sk, 1, rows * 4)); CK(cudaMemset(sv, 2, rows * 4));
	if (getenv("HARNESS_RANDOM")) {
		k_fill<<<1024, 256>>>(sk, rows, 1);
		k_fill<<<1024, 256>>>(sv, rows, 2);
		CK(cudaDeviceSynchronize());
		printf("source columns: pseudo-random words\n");
	}
	CK(cudaMalloc(&dk[0], rows * 4)); CK(cudaMalloc(&dv[0], rows * 4));
	if (child) {
		cudaIpcMemHandle_t ha, hb;
		unhex(argv[2], &ha, sizeof ha); unhex(argv[3], &hb, sizeof hb);
		CK(cudaIpcOpenMemHandle((void **)&dk[1], ha, cudaIpcMemLazyEnablePeerAccess));
		CK(cudaIpcOpenMemHandle((void **)&dv[1], hb, cudaIpcMemLazyEnablePeerAccess));
		printf("destination: legacy IPC handles opened in a second process\n");
	} else {
		CK(cudaSetDevice(peer));
		CK(cudaMalloc(&dk[1], rows * 4)); CK(cudaMalloc(&dv[1], rows * 4));
		CK(cudaSetDevice(0));
	}
	std::vector<uint32_t> desc(SD_WORDS, 0);
	uint32_t s = 0;
	for (int o = 0; o < G; ++o) {                                    // one run per owner; destination rows start unaligned
		const uint32_t t = 5 + o, len = per_digit * nsub;
		desc[SD_OWN_LEN + o] = len;
		desc[SD_OWN_DST + o] = t;
		s = ((s + 31) & ~31u) + (t & 31);
		desc[SD_OWN_SRC + o] = s;
		s += len;
	}
	uint32_t *ddesc, *dabort;
	CK(cudaMalloc(&ddesc, SD_WORDS * 4)); CK(cudaMalloc(&dabort, 4));
	CK(cudaMemcpy(ddesc, desc.data(), SD_WORDS * 4, cudaMemcpyHostToDevice));
	CK(cudaMemset(dabort, 0, 4));
	PeerCols pc = {};
	for (int g = 0; g < G; ++g) { pc.k[g] = dk[g]; pc.v[g] = dv[g]; }
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	if (getenv("HARNESS_LOOP")) {
		// many launches, each timed: run two of these processes at once (CUDA_VISIBLE_DEVICES=0,1 and 1,0) for both directions
		const int iters = atoi(getenv("HARNESS_LOOP"));
		std::vector<float> t;
		for (int i = 0; i < iters; ++i) {
			CK(cudaEventRecord(e0));
			launch_peer_copy(sk, sv, pc, ddesc, dabort, gbits, 0, 1, 0, nullptr);
			CK(cudaEventRecord(e1));
			CK(cudaDeviceSynchronize());
			float ms;
			CK(cudaEventElapsedTime(&ms, e0, e1));
			t.push_back(ms);
		}
		std::vector<float> u = t;
		std::sort(u.begin(), u.end());
		const double bytes = (double)per_digit * nsub * 8;
		printf("loop of %d (HJB_COPY_CTAS=%s): min %.3f ms (%.0f GB/s) median %.3f ms (%.0f GB/s) max %.3f ms\n", iters, getenv("HJB_COPY_CTAS"), u[0],
		       bytes / u[0] / 1e6, u[iters / 2], bytes / u[iters / 2] / 1e6, u.back());
		return 0;
	}
	for (int skip = 1; skip >= 0; --skip) {
		launch_peer_copy(sk, sv, pc, ddesc, dabort, gbits, 0, skip, 0, nullptr);
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(e0));
		for (int i = 0; i < 3; ++i) launch_peer_copy(sk, sv, pc, ddesc, dabort, gbits, 0, skip, 0, nullptr);
		CK(cudaEventRecord(e1));
		CK(cudaDeviceSynchronize());
		CK(cudaGetLastError());
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		const double bytes = (double)per_digit * (skip ? nsub : F) * 8;
		printf("k_peer_copy %s: %.3f ms  %.1f GB/s (HJB_COPY_CTAS=%s)\n", skip ? "remote runs only" : "remote + own runs", ms / 3,
		       bytes / (ms / 3 * 1e-3) / 1e9, getenv("HJB_COPY_CTAS") ? getenv("HJB_COPY_CTAS") : "default");
	}
	return 0;
}
