// EXPERIMENT, not part of libhjb200 (measured and dropped in round 2): config 3 (2^16 x 2^30) took
// 173 ms with this kernel against 10.1 ms for k_npj_probe on the global, L2-resident table -- correct (rows equal the
// oracle's for cluster sizes 1, 2, 4 and 8, misses, the all-ones key), but one 512-thread CTA per SM cannot hide the
// chains of dependent shared-memory loads: 0.5 us per 2048-tuple tile for the ownership scan, ~10 us for the probes.
// To build it again: add it to KERNEL_SOURCES, give k_npj_probe a device-side gate, launch it from launch_npj_probe.
// npj_cluster.cu -- NPJ's probe for a SMALL build side (reference: probe, npj.cpp:216-364; BASELINE config 3:
// |R| = 64K, |S| = 1G, "cache-resident build side").
//
// The global table of npj.cu costs one L2 round trip per probe tuple (config 3: 1.02 G requests, 10.1 ms for 2^30 probes,
// issue slots idle behind the latency).  A build side of up to ~110 K tuples fits into the shared memory of a thread-block
// CLUSTER instead: the C = 1, 2, 4 or 8 CTAs of a cluster each keep the keys whose top log2(C) hash bits equal their rank in
// an open-addressing table of 20480 64-bit slots (160 KB), and all of them look at EVERY probe tuple of the cluster's
// tiles, each picking out its own share:
//   load    a tile of 2048 probe tuples is fetched ONCE per cluster: every CTA issues a TMA bulk copy of 1/C of the tile
//           that is MULTICAST into the shared memory of all C CTAs (cp.async.bulk ... .multicast::cluster), completion
//           on each CTA's own mbarrier; a stage is refilled when all C CTAs have released it (remote mbarrier arrives)
//   scan    every thread hashes its keys of the tile; the tile indices of the keys this CTA owns are queued
//   probe   the queue is worked off with full warps: table look-up in shared memory, rows reserved once per CTA round
// HBM sees the probe relation once and the result rows once; L2 serves one multicast read per tile.
// Unique build keys without the all-ones pair only (the global build of npj.cu has checked both by then, flags[]); a
// table share that does not fit makes every cluster give up before the first row is written.  In both cases -- and for
// the last, partial tile -- k_npj_probe does the work against the global table.
#include "hj_device.cuh"
#include "hj_internal.h"
#include <stdio.h>
#include <stdlib.h>

namespace hjb {

constexpr int kNcThreads = 512;
constexpr uint32_t kNcTile = 2048;             // probe tuples per tile
constexpr int kNcStages = 3;
constexpr uint32_t kNcSlots = 20480;           // 64-bit slots per CTA: 160 KB
constexpr uint32_t kNcMaxFill = 17408;         // 0.85
constexpr int kNcItems = 2;                    // queue entries per thread and probe round
constexpr size_t kNcSmem = (size_t)kNcSlots * 8 + (size_t)kNcStages * 2 * kNcTile * 4 + (size_t)kNcTile * 2;

__device__ __forceinline__ uint32_t nc_smem(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
	return r;
}
__device__ __forceinline__ uint32_t cluster_id_x()
{
	uint32_t r;
	asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
	return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void nc_mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n.reg .pred p;\nNCW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra NCD_%=;\nbra NCW_%=;\nNCD_%=:\n}" ::"r"(
	                 nc_smem(bar)),
	             "r"(parity)
	             : "memory");
}
// one arrival on the barrier at the same shared-memory offset in CTA `rank` of this cluster
__device__ __forceinline__ void nc_remote_arrive(uint64_t *bar, uint32_t rank)
{
	uint32_t remote;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(nc_smem(bar)), "r"(rank));
	asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}

template <bool MATERIALIZE>
__global__ void __launch_bounds__(kNcThreads, 1)
k_npj_cluster(const uint32_t *__restrict__ rk, const uint32_t *__restrict__ rv, uint32_t nr, const uint32_t *__restrict__ sk,
              const uint32_t *__restrict__ sv, uint64_t tiles, uint32_t factor, int log2c, OutCols out,
              unsigned long long *__restrict__ sums, const unsigned long long *__restrict__ flags, unsigned long long *__restrict__ giveup,
              int dbg)
{
	extern __shared__ __align__(128) unsigned char nc_raw[];
	uint64_t *table = reinterpret_cast<uint64_t *>(nc_raw);
	uint32_t *tile = reinterpret_cast<uint32_t *>(nc_raw + (size_t)kNcSlots * 8);         // [stage][keys | vals][kNcTile]
	uint16_t *queue = reinterpret_cast<uint16_t *>(tile + kNcStages * 2 * kNcTile);
	__shared__ uint64_t full[kNcStages], empty[kNcStages];
	__shared__ uint64_t scratch[4 * 32];
	__shared__ uint32_t s_qlen[2], s_fill;
	__shared__ __align__(8) uint32_t s_emit[2 * (kNcThreads / 32 + 2) + 4];
	// equal build keys or the all-ones pair: the probe of npj.cu handles those (uniform over the grid, before any cluster operation)
	if (flags[0] | flags[1]) return;
	const uint32_t C = 1u << log2c, rank = cluster_ctarank(), nclusters = gridDim.x >> log2c, cid = cluster_id_x();
	const int qshift = 32 - log2c;
	// ---- this CTA's share of the build side
	for (uint32_t h = threadIdx.x; h < kNcSlots / 2; h += kNcThreads)
		reinterpret_cast<ulonglong2 *>(table)[h] = make_ulonglong2(kEmptySlot, kEmptySlot);
	if (threadIdx.x == 0) {
		for (int i = 0; i < kNcStages; ++i) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(nc_smem(&full[i])));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(nc_smem(&empty[i])), "r"(C));
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		s_fill = 0;
		s_qlen[0] = s_qlen[1] = 0;
	}
	__syncthreads();
	for (uint32_t i = threadIdx.x; i < nr; i += kNcThreads) {
		const uint32_t key = rk[i], x = hash_mul(key, factor);
		if (log2c && (x >> qshift) != rank) continue;
		if (atomicAdd(&s_fill, 1u) >= kNcMaxFill) continue;              // does not fit: every cluster will give up
		const uint64_t pair = ((uint64_t)rv[i] << 32) | key;
		uint32_t h = __umulhi(x << log2c, kNcSlots);
		while (atomicCAS(reinterpret_cast<unsigned long long *>(&table[h]), (unsigned long long)kEmptySlot, (unsigned long long)pair) != kEmptySlot)
			h = h + 1 == kNcSlots ? 0 : h + 1;
	}
	__syncthreads();
	if (threadIdx.x == 0 && s_fill > kNcMaxFill) {
		*reinterpret_cast<volatile unsigned long long *>(giveup) = 1;
		__threadfence();
	}
	cluster_sync_all();                       // every CTA's barriers are initialised, its table is built, its verdict written
	if (*reinterpret_cast<volatile unsigned long long *>(giveup)) return;        // the same in every cluster: the tables are the same
	// ---- probe: tiles cid, cid + nclusters, ...
	const uint64_t mine = tiles > cid ? (tiles - cid + nclusters - 1) / nclusters : 0;
	const uint32_t slice = kNcTile >> log2c;                                      // tuples of a tile this CTA fetches for the cluster
	auto issue = [&](uint64_t k) {
		const int st = (int)(k % kNcStages);
		const uint64_t t = cid + k * nclusters;
		uint32_t *dk = tile + (size_t)st * 2 * kNcTile + rank * slice, *dv = dk + kNcTile;
		const uint32_t *gk = sk + t * kNcTile + rank * slice, *gv = sv + t * kNcTile + rank * slice;
		asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(nc_smem(&full[st])), "r"(kNcTile * 8u) : "memory");
		if (log2c) {
			const uint16_t mask = (uint16_t)((1u << C) - 1);
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(nc_smem(dk)),
			             "l"(gk), "r"(slice * 4u), "r"(nc_smem(&full[st])), "h"(mask) : "memory");
			asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(nc_smem(dv)),
			             "l"(gv), "r"(slice * 4u), "r"(nc_smem(&full[st])), "h"(mask) : "memory");
		} else {
			asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(nc_smem(dk)), "l"(gk),
			             "r"(slice * 4u), "r"(nc_smem(&full[st])) : "memory");
			asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(nc_smem(dv)), "l"(gv),
			             "r"(slice * 4u), "r"(nc_smem(&full[st])) : "memory");
		}
	};
	if (threadIdx.x == 0)
		for (uint64_t k = 0; k < mine && k < (uint64_t)(kNcStages - 1); ++k) issue(k);
	JoinSums acc;
	acc.zero();
	uint32_t emit_rounds = 0;
	const unsigned lt = lanemask_lt();
	for (uint64_t k = 0; k < mine; ++k) {
		const int st = (int)(k % kNcStages);
		if (!(dbg & 2) && threadIdx.x == 0 && k + kNcStages - 1 < mine) {
			const uint64_t kk = k + kNcStages - 1;
			// its stage was last used by tile kk - kNcStages: every CTA of the cluster must have released it
			if (kk >= (uint64_t)kNcStages) nc_mbar_wait(&empty[kk % kNcStages], (uint32_t)((kk / kNcStages - 1) & 1));
			issue(kk);
		}
		if (!(dbg & 2) || k < (uint64_t)(kNcStages - 1)) nc_mbar_wait(&full[st], (uint32_t)((k / kNcStages) & 1));
		const uint32_t *tk = tile + (size_t)st * 2 * kNcTile, *tv = tk + kNcTile;
		uint32_t *qlen = &s_qlen[k & 1];
		// ---- scan: queue the tuples whose keys this CTA owns
#pragma unroll
		for (uint32_t e = 0; e < ((dbg & 8) ? 0u : kNcTile / kNcThreads); ++e) {
			const uint32_t idx = threadIdx.x + e * kNcThreads;
			const bool own = !log2c || (hash_mul(tk[idx], factor) >> qshift) == rank;
			const unsigned m = __ballot_sync(kFullMask, own);
			if (m) {
				uint32_t base = 0;
				if (lane_id() == (uint32_t)(__ffs(m) - 1)) base = atomicAdd(qlen, (uint32_t)__popc(m));
				base = __shfl_sync(kFullMask, base, __ffs(m) - 1);
				if (own) queue[base + __popc(m & lt)] = (uint16_t)idx;
			}
		}
		__syncthreads();
		if (threadIdx.x == 0) s_qlen[(k + 1) & 1] = 0;
		const uint32_t n = (dbg & 4) ? 0u : *qlen;
		// ---- probe the queue, kNcItems entries per thread and round
		for (uint32_t base = 0; base < n; base += kNcThreads * kNcItems) {
			uint32_t key[kNcItems], val[kNcItems], ival[kNcItems];
			bool found[kNcItems];
#pragma unroll
			for (int t = 0; t < kNcItems; ++t) {
				const uint32_t j = base + (threadIdx.x & ~31u) * kNcItems + t * 32 + lane_id();      // a warp owns 32 * kNcItems consecutive entries
				const bool valid = j < n;
				const uint32_t idx = valid ? queue[j] : 0;
				key[t] = tk[idx];
				val[t] = tv[idx];
				ival[t] = 0;
				bool hit = false;
				if (valid) {
					uint32_t h = __umulhi(hash_mul(key[t], factor) << log2c, kNcSlots);
					while (true) {
						const uint64_t slot = table[h];
						if (slot == kEmptySlot) break;
						if ((uint32_t)slot == key[t]) {
							ival[t] = (uint32_t)(slot >> 32);
							hit = true;
							break;
						}
						h = h + 1 == kNcSlots ? 0 : h + 1;
					}
				}
				found[t] = hit;
				acc.add_if(hit ? 1u : 0u, key[t], val[t], ival[t]);
			}
			if (MATERIALIZE) emit_round_cta<kNcItems>(out, s_emit, emit_rounds++, found, key, val, ival);
		}
		__syncthreads();                      // every thread is done with the stage and the queue
		if (!(dbg & 2) && threadIdx.x < C) nc_remote_arrive(&empty[st], threadIdx.x);
	}
	acc.reduce_to_global(sums, scratch);
	cluster_sync_all();                       // no CTA leaves while a peer may still signal its barriers
}

// how many CTAs share the build side: the smallest cluster whose tables stay below ~0.8 full; 0 = the build side is too large
int npj_cluster_size(uint64_t nr)
{
	static const int enabled = getenv("HJB_NPJ_CLUSTER") ? atoi(getenv("HJB_NPJ_CLUSTER")) : 1;
	if (!enabled) return 0;
	for (int c = 1; c <= 8; c *= 2)
		if (nr <= (uint64_t)c * 16384) return c;
	return 0;
}

// probe of the first `tiles` whole tiles of S; returns kernels launched (0: not applicable here)
int launch_npj_cluster(const NpjArgs &a, uint64_t tiles, int csize, cudaStream_t s, int sms, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	int log2c = 0;
	while ((1 << log2c) < csize) ++log2c;
	auto kernel = a.materialize ? k_npj_cluster<true> : k_npj_cluster<false>;
	if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kNcSmem) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	cudaLaunchConfig_t cfg = {};
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeClusterDimension;
	attr[0].val.clusterDim.x = (unsigned)csize;
	attr[0].val.clusterDim.y = 1;
	attr[0].val.clusterDim.z = 1;
	cfg.blockDim = dim3(kNcThreads);
	cfg.dynamicSmemBytes = kNcSmem;
	cfg.stream = s;
	cfg.attrs = attr;
	cfg.numAttrs = 1;
	cfg.gridDim = dim3((unsigned)(sms / csize * csize));
	int clusters = 0;
	if (cudaOccupancyMaxActiveClusters(&clusters, kernel, &cfg) != cudaSuccess || clusters < 1) {
		cudaGetLastError();
		return 0;
	}
	if (getenv("HJB_DEBUG")) fprintf(stderr, "k_npj_cluster: cluster size %d, %d clusters can be resident, %d SMs\n", csize, clusters, sms);
	// clusters work on disjoint tiles and never wait for one another: a grid larger than what is resident at once is correct
	if ((uint64_t)clusters > tiles) clusters = (int)tiles;
	if (clusters > sms / csize) clusters = sms / csize;
	cfg.gridDim = dim3((unsigned)(clusters * csize));
	OutCols out;
	out.k = a.out_k;
	out.o = a.out_o;
	out.i = a.out_i;
	out.cursor = a.scalars;
	out.cap = a.materialize ? a.out_cap : 0;
	t->start(KK_NPJ_CLUSTER, s);
	const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, a.rk, a.rv, (uint32_t)a.nr, a.sk, a.sv, tiles, a.factor, log2c, out, a.scalars + 1,
	                                         (const unsigned long long *)(a.scalars + 5), a.scalars + 13,
	                                         getenv("HJB_NC_DBG") ? atoi(getenv("HJB_NC_DBG")) : 0);
	t->stop(s);
	if (e != cudaSuccess) {
		cudaGetLastError();
		return -1;
	}
	return 1;
}

uint32_t npj_cluster_tile() { return kNcTile; }

}  // namespace hjb
