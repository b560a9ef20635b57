// peer_copy_bench.cu -- how many SMs does a pure copy into a peer's memory need to fill NVLink?
//   (a) TMA pipeline: cp.async.bulk global->shared (mbarrier), cp.async.bulk shared->peer global, one thread per CTA
//   (b) SIMT 16-byte loads / stores
//   (c) cudaMemcpyPeerAsync (copy engine)
// one process, devices 0 and 1.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a peer_copy_bench.cu -o peer_copy_bench
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void mbar_init(uint64_t *b, int n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(n)); }
__device__ __forceinline__ void mbar_expect(uint64_t *b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t *b, uint32_t phase)
{
	asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}" ::"r"((uint32_t)__cvta_generic_to_shared(b)), "r"(phase) : "memory");
}

template <int NS>
__global__ void __launch_bounds__(32) k_copy_tma(const char *__restrict__ src, char *__restrict__ dst, size_t bytes, uint32_t chunk)
{
	extern __shared__ __align__(128) char smem[];
	__shared__ uint64_t full[NS];
	if (threadIdx.x != 0) return;
	for (int i = 0; i < NS; ++i) mbar_init(&full[i], 1);
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	const size_t nchunks = (bytes + chunk - 1) / chunk;
	// chunk c of this CTA: global chunk blockIdx.x + c * gridDim.x
	size_t issued = 0, done = 0;
	const size_t mine = nchunks > blockIdx.x ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
	auto issue_load = [&](size_t c) {
		const int st = (int)(c % NS);
		const size_t off = (blockIdx.x + c * gridDim.x) * (size_t)chunk;
		const uint32_t sz = (uint32_t)(bytes - off < chunk ? bytes - off : chunk);
		mbar_expect(&full[st], sz);
		asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"((uint32_t)__cvta_generic_to_shared(smem + (size_t)st * chunk)),
		             "l"(src + off), "r"(sz), "r"((uint32_t)__cvta_generic_to_shared(&full[st])) : "memory");
	};
	for (; issued < mine && issued < NS - 1; ++issued) issue_load(issued);
	for (; done < mine; ++done) {
		const int st = (int)(done % NS);
		mbar_wait(&full[st], (uint32_t)((done / NS) & 1));
		const size_t off = (blockIdx.x + done * gridDim.x) * (size_t)chunk;
		const uint32_t sz = (uint32_t)(bytes - off < chunk ? bytes - off : chunk);
		asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + off), "r"((uint32_t)__cvta_generic_to_shared(smem + (size_t)st * chunk)), "r"(sz) : "memory");
		asm volatile("cp.async.bulk.commit_group;" ::: "memory");
		if (issued < mine) {
			// the stage that load `issued` refills was last read by the store of chunk issued - NS: at most NS - 2... stores may still be reading
			asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(1) : "memory");    // NS - (NS - 1) stores may still read
			issue_load(issued);
			++issued;
		}
	}
	asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

__global__ void __launch_bounds__(512) k_copy_simt(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t n16)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char **argv)
{
	int nd = 0;
	CK(cudaGetDeviceCount(&nd));
	const int peer = nd > 1 ? 1 : 0;
	printf("devices %d, copying 0 -> %d\n", nd, peer);
	const size_t bytes = (size_t)1 << 30;
	char *src, *dst, *dst_local;
	CK(cudaSetDevice(0));
	if (peer) CK(cudaDeviceEnablePeerAccess(peer, 0));
	CK(cudaMalloc(&src, bytes));
	CK(cudaMalloc(&dst_local, bytes));
	CK(cudaMemset(src, 1, bytes));
	CK(cudaSetDevice(peer));
	CK(cudaMalloc(&dst, bytes));
	CK(cudaSetDevice(0));
	cudaEvent_t e0, e1;
	CK(cudaEventCreate(&e0));
	CK(cudaEventCreate(&e1));
	auto timeit = [&](const char *name, auto fn) {
		fn();
		CK(cudaDeviceSynchronize());
		CK(cudaEventRecord(e0));
		for (int i = 0; i < 3; ++i) fn();
		CK(cudaEventRecord(e1));
		CK(cudaDeviceSynchronize());
		CK(cudaGetLastError());
		float ms;
		CK(cudaEventElapsedTime(&ms, e0, e1));
		printf("%-44s %8.3f ms  %7.1f GB/s\n", name, ms / 3, bytes / (ms / 3 * 1e-3) / 1e9);
		fflush(stdout);
	};
	char name[128];
	for (int target = 0; target < 2; ++target) {
		char *d = target ? dst : dst_local;
		const char *tn = target ? "peer" : "local";
		snprintf(name, sizeof name, "memcpyAsync %s", tn);
		timeit(name, [&] { CK(cudaMemcpyAsync(d, src, bytes, cudaMemcpyDeviceToDevice, 0)); });
		for (int ctas : {8, 16, 24, 32, 48, 64, 148}) {
			for (uint32_t chunk : {8192u, 16384u, 32768u}) {
				CK(cudaFuncSetAttribute(k_copy_tma<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 32768));
				CK(cudaFuncSetAttribute(k_copy_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768));
				snprintf(name, sizeof name, "tma %s ctas %3d chunk %5u stages 4", tn, ctas, chunk);
				timeit(name, [&] { k_copy_tma<4><<<ctas, 32, 4 * chunk>>>(src, d, bytes, chunk); });
				snprintf(name, sizeof name, "tma %s ctas %3d chunk %5u stages 6", tn, ctas, chunk);
				timeit(name, [&] { k_copy_tma<6><<<ctas, 32, 6 * chunk>>>(src, d, bytes, chunk); });
			}
			snprintf(name, sizeof name, "simt %s ctas %3d x 512", tn, ctas);
			timeit(name, [&] { k_copy_simt<<<ctas, 512>>>((const uint4 *)src, (uint4 *)d, bytes / 16); });
		}
		// several TMA CTAs per SM (each 32 threads, 4 x 8 KB): does the per-SM limit move?
		for (int ctas : {32, 64, 128, 296, 592}) {
			snprintf(name, sizeof name, "tma %s ctas %3d chunk 8192 stages 4 (co-res)", tn, ctas);
			timeit(name, [&] { k_copy_tma<4><<<ctas, 32, 4 * 8192>>>(src, d, bytes, 8192); });
		}
	}
	// both directions at once (what the exchange does): device 1 copies into device 0 while device 0 copies into device 1
	if (peer) {
		char *src1, *dst0;
		CK(cudaSetDevice(1));
		CK(cudaDeviceEnablePeerAccess(0, 0));
		CK(cudaMalloc(&src1, bytes));
		CK(cudaMemset(src1, 2, bytes));
		cudaStream_t s1;
		CK(cudaStreamCreate(&s1));
		CK(cudaSetDevice(0));
		CK(cudaMalloc(&dst0, bytes));
		cudaStream_t s0;
		CK(cudaStreamCreate(&s0));
		for (int ctas : {12, 16, 20, 24, 32, 48, 148}) {
			for (uint32_t chunk : {8192u, 16384u, 32768u}) {
				snprintf(name, sizeof name, "tma BIDIR ctas %3d chunk %5u stages 6", ctas, chunk);
				timeit(name, [&] {
					CK(cudaSetDevice(1));
					CK(cudaFuncSetAttribute(k_copy_tma<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * 32768));
					k_copy_tma<6><<<ctas, 32, 6 * chunk, s1>>>(src1, dst0, bytes, chunk);
					CK(cudaSetDevice(0));
					k_copy_tma<6><<<ctas, 32, 6 * chunk, 0>>>(src, dst, bytes, chunk);
				});
				CK(cudaSetDevice(1));
				CK(cudaDeviceSynchronize());
				CK(cudaSetDevice(0));
			}
		}
		snprintf(name, sizeof name, "memcpyAsync BIDIR");
		timeit(name, [&] {
			CK(cudaSetDevice(1));
			CK(cudaMemcpyAsync(dst0, src1, bytes, cudaMemcpyDeviceToDevice, s1));
			CK(cudaSetDevice(0));
			CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, 0));
		});
		CK(cudaSetDevice(1));
		CK(cudaDeviceSynchronize());
		CK(cudaSetDevice(0));
	}
	// check
	CK(cudaMemset(dst, 0, bytes));
	k_copy_tma<4><<<32, 32, 4 * 16384>>>(src, dst, bytes - 48, 16384);
	CK(cudaDeviceSynchronize());
	unsigned char h[64];
	CK(cudaMemcpy(h, dst + bytes - 64, 64, cudaMemcpyDeviceToHost));
	printf("tail check: %d %d %d (expect 1 1 0)\n", h[0], h[15], h[16]);
	return 0;
}
