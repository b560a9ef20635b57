// peer_copy_harness.cu -- the product's k_peer_copy (csrc/stage.cu, linked in) on hand-made run descriptors, devices 0 -> 1
// of ONE process (peer access, no IPC): per-CTA throughput of the copy kernel without the rest of the step.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I hash_join_codes_knl_b200/csrc peer_copy_harness.cu \
//        ../../hash_join_codes_knl_b200/csrc/stage.cu -o _build/peer_copy_harness
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "hj_internal.h"
using namespace hjb;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

int main(int argc, char **argv)
{
	int nd = 0;
	CK(cudaGetDeviceCount(&nd));
	const int peer = nd > 1 ? 1 : 0;
	const int abits = 9, gbits = 1, G = 2, F = 512, nsub = F / G;
	const uint32_t per_digit = (1u << 28) / F + 37;            // rows per run (not a multiple of 32)
	const size_t rows = (size_t)per_digit * F + 64 * F;
	uint32_t *sk, *sv, *dk[2], *dv[2];
	CK(cudaSetDevice(0));
	if (peer) CK(cudaDeviceEnablePeerAccess(peer, 0));
	CK(cudaMalloc(&sk, rows * 4)); CK(cudaMalloc(&sv, rows * 4));
	CK(cudaMemset(sk, 1, rows * 4)); CK(cudaMemset(sv, 2, rows * 4));
	CK(cudaMalloc(&dk[0], rows * 4)); CK(cudaMalloc(&dv[0], rows * 4));
	CK(cudaSetDevice(peer));
	CK(cudaMalloc(&dk[1], rows * 4)); CK(cudaMalloc(&dv[1], rows * 4));
	CK(cudaSetDevice(0));
	std::vector<uint32_t> desc(SD_WORDS, 0);
	uint32_t s = 0, t[2] = {5, 5};                                   // destination rows start unaligned
	for (int d = 0; d < F; ++d) {
		const int o = d / nsub;
		desc[SD_N + d] = per_digit;
		desc[SD_T0 + d] = t[o];
		s = ((s + 31) & ~31u) + (t[o] & 31);
		desc[SD_S0 + d] = s;
		s += per_digit;
		t[o] += per_digit + 11;                                      // other senders' rows in between
	}
	uint32_t *ddesc, *dabort;
	CK(cudaMalloc(&ddesc, SD_WORDS * 4)); CK(cudaMalloc(&dabort, 4));
	CK(cudaMemcpy(ddesc, desc.data(), SD_WORDS * 4, cudaMemcpyHostToDevice));
	CK(cudaMemset(dabort, 0, 4));
	PeerCols pc = {};
	for (int g = 0; g < G; ++g) { pc.k[g] = dk[g]; pc.v[g] = dv[g]; }
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
	for (int skip = 1; skip >= 0; --skip) {
		launch_peer_copy(sk, sv, pc, ddesc, dabort, abits, gbits, 0, skip, 0, nullptr);
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(e0));
		for (int i = 0; i < 3; ++i) launch_peer_copy(sk, sv, pc, ddesc, dabort, abits, gbits, 0, skip, 0, nullptr);
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
