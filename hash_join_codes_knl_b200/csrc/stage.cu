// stage.cu -- the STAGED exchange of CPRA (reference: the chunk-local passes cpra2.cpp:1783-1827 followed by the
// per-owner memcpy gather cpra2.cpp:1861-1905,1940-1959, timed there as "copy:").
//
// Stage A partitions this GPU's chunk LOCALLY by the top `abits` bits of key * factor with the ordinary pass of
// radix.cu (digit = owner << sub_bits | sub-partition), so that an owner's share is already cut into 2^sub_bits
// sub-partitions when it leaves.  The exchange is then nothing but copies of whole digit runs:
//   k_stage_counts : this sender's tuples per digit, R then S -- the all-gather's input
//   k_stage_bases  : from the all-gathered G x 2 x 2^abits count matrix: the first row of this sender's run in every
//                    owner's columns (an owner receives sender-major: one run per sender, its sub-partitions in order
//                    inside -- the pieces the reference's gather copies, cpra2.cpp:1896-1904), the ranges every
//                    sub-partition this GPU receives consists of (one per sender: the parents of the local pass),
//                    whether every owner's buffer is large enough, and where stage A must START each run in the
//                    staging columns so that source and destination rows have the same 128-byte phase
//   k_peer_copy    : a handful of two-warp CTAs drive the TMA unit: cp.async.bulk global -> shared (mbarrier) and
//                    shared -> the owner's global memory over NVLink, 16 KB per column and piece, two stages, three CTAs per SM.
//                    Measured (scripts/r2/peer_copy_bench.cu, 2 GPUs): 16 such CTAs move 705 GB/s, the copy
//                    engine 776 GB/s -- the SMs beside them stay free for stage A of the other relation and the
//                    local pass of the relation that has arrived.
// The receiver continues with ONE local pass over the 2^sub_bits sub-partitions (parents with device-resident
// offsets) and the shared-memory join.
#include "hj_device.cuh"
#include "hj_internal.h"
#include <atomic>
#include <stdlib.h>

namespace hjb {

__global__ void __launch_bounds__(512)
k_stage_counts(const uint32_t *__restrict__ r_off, const uint32_t *__restrict__ s_off, uint32_t F,
               unsigned long long *__restrict__ counts)
{
	const uint32_t d = threadIdx.x;
	if (d < F) {
		counts[d] = r_off[d + 1] - r_off[d];
		counts[F + d] = s_off[d + 1] - s_off[d];
	}
}

// exclusive scan of one uint64 per thread over a 512-thread CTA; *total = the sum
__device__ __forceinline__ unsigned long long block_exclusive_scan_u64(unsigned long long v, unsigned long long *warp_tot,
                                                                      unsigned long long *total)
{
	unsigned long long incl = v;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1) {
		const unsigned long long t = __shfl_up_sync(kFullMask, incl, o);
		if ((int)lane_id() >= o) incl += t;
	}
	__syncthreads();                                  // warp_tot may still be read from an earlier call
	if (lane_id() == 31) warp_tot[threadIdx.x >> 5] = incl;
	__syncthreads();
	unsigned long long before = 0, all = 0;
	for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) {
		const unsigned long long t = warp_tot[w];
		if (w < (threadIdx.x >> 5)) before += t;
		all += t;
	}
	*total = all;
	return before + incl - v;
}

// Source and destination rows of a run agree modulo kCopyPhase rows = 128 bytes: the bulk copies then move whole,
// identically aligned 128-byte lines on both sides.
constexpr uint32_t kCopyPhase = 32;

// M[src][rel][digit] (uint64): the all-gathered counts.  One CTA of 512 threads, one digit per thread, R then S.
// stage_base_*: first row of the staging region in stage A's output columns.  inplace: those columns ARE this GPU's
// receive columns (staging region behind the receive region): the run this GPU owns itself is then scattered
// straight to its final rows and never copied.
__global__ void __launch_bounds__(512)
k_stage_bases(const unsigned long long *__restrict__ M, int G, int me, int abits, int gbits, unsigned long long cap_r,
              unsigned long long cap_s, const uint32_t *__restrict__ child_r, const uint32_t *__restrict__ child_s,
              uint32_t stage_base_r, uint32_t stage_base_s, int inplace, int nparts, uint32_t *__restrict__ out,
              uint32_t *__restrict__ status)
{
	// nparts: the runs leave in nparts pieces, piece k = the sub-partitions [stage_part_lo(k), stage_part_lo(k + 1))
	__shared__ unsigned long long warp_tot[16];
	__shared__ unsigned long long s_tot[2][64], s_bef[2][64], s_len[2][64];     // per owner: rows it receives, rows of the senders before me, my rows
	__shared__ unsigned long long s_max[2];
	__shared__ uint32_t s_src[64];
	__shared__ int s_abort;
	const uint32_t F = 1u << abits, nsub = F >> gbits, d = threadIdx.x;
	if (d < 64) {
		for (int rel = 0; rel < 2; ++rel) s_tot[rel][d] = s_bef[rel][d] = s_len[rel][d] = 0;
	}
	if (d == 0) {
		s_abort = 0;
		s_max[0] = s_max[1] = 0;
	}
	__syncthreads();
	for (int rel = 0; rel < 2; ++rel) {
		unsigned long long col = 0, before = 0, mine = 0;
		if (d < F) {
			for (int src = 0; src < G; ++src) {
				const unsigned long long c = M[((size_t)src * 2 + rel) * F + d];
				col += c;
				if (src < me) before += c;
				if (src == me) mine = c;
			}
			const uint32_t o = d / nsub;
			atomicAdd(&s_tot[rel][o], col);
			atomicAdd(&s_bef[rel][o], before);
			atomicAdd(&s_len[rel][o], mine);
		}
		// the cumulative sizes of the sub-partitions this GPU receives (the parents of the local pass)
		const bool my_sub = d < F && d / nsub == (uint32_t)me;
		unsigned long long total;
		const unsigned long long e = block_exclusive_scan_u64(my_sub ? col : 0ull, warp_tot, &total);
		uint32_t *poff = out + (rel ? SD_REL_S : SD_REL_R) + SD_POFF;
		if (my_sub) poff[d - me * nsub] = (uint32_t)e;
		if (d == 0) poff[nsub] = (uint32_t)total;
	}
	__syncthreads();
	if (d < (uint32_t)G)
		for (int rel = 0; rel < 2; ++rel) {
			atomicMax(&s_max[rel], s_tot[rel][d]);
			if (s_tot[rel][d] > (rel ? cap_s : cap_r)) s_abort = 1;
		}
	__syncthreads();
	const int abort = s_abort;
	for (int rel = 0; rel < 2; ++rel) {
		uint32_t *o = out + (rel ? SD_REL_S : SD_REL_R);
		// what this GPU receives: sender-major, and inside a sender's block its sub-partitions in order.  Range (q, src) of
		// parent q: one exclusive scan over the sub-partitions per sender.
		unsigned long long sender_base = 0;
		const bool my_sub = d < F && d / nsub == (uint32_t)me;
		for (int src = 0; src < G; ++src) {
			const unsigned long long v = my_sub ? M[((size_t)src * 2 + rel) * F + d] : 0ull;
			unsigned long long total;
			const unsigned long long pre = block_exclusive_scan_u64(v, warp_tot, &total);
			if (my_sub) {
				const uint32_t q = d - me * nsub;
				o[SD_SEG + ((size_t)q * G + src) * 2] = (uint32_t)(sender_base + pre);
				o[SD_SEG + ((size_t)q * G + src) * 2 + 1] = abort ? 0u : (uint32_t)v;
			}
			sender_base += total;
		}
		if (abort && d <= nsub) o[SD_POFF + d] = 0;
		// my run for every owner: where stage A puts it (own run: its final rows; the others: the staging region, each
		// starting with the 128-byte phase of its destination), its length, its first row in the owner's columns
		if (d == 0) {
			uint32_t next = rel ? stage_base_s : stage_base_r;
			for (int g = 0; g < G; ++g) {
				const uint32_t dst = (uint32_t)s_bef[rel][g], len = (uint32_t)s_len[rel][g];
				const bool home = inplace && !abort && g == me;
				uint32_t src;
				if (home) src = dst;
				else {
					src = next + (dst & (kCopyPhase - 1));
					next = (src + len + kCopyPhase - 1) & ~(kCopyPhase - 1);
				}
				s_src[g] = src;
			}
		}
		__syncthreads();
		const uint32_t *child = rel ? child_s : child_r;
		if (d < (uint32_t)(nparts * G)) {
			// piece (k, g): the scan's offsets of this sender's chunk say where the sub-partitions begin inside the run
			const uint32_t k = d / G, g = d % G;
			const uint32_t first = child[g * nsub], lo = child[g * nsub + stage_part_lo(k, nparts, nsub)] - first,
			               hi = child[g * nsub + stage_part_lo(k + 1, nparts, nsub)] - first;
			o[SD_OWN_SRC + k * 64 + g] = s_src[g] + lo;
			o[SD_OWN_LEN + k * 64 + g] = hi - lo;
			o[SD_OWN_DST + k * 64 + g] = (uint32_t)s_bef[rel][g] + lo;
		}
		if (d < F) o[SD_SHIFT + d] = s_src[d / nsub] - child[(d / nsub) * nsub];     // a run keeps the scan's order of its sub-partitions
		__syncthreads();
	}
	// what hjb_cpra_finish reports: rows received here, the verdict, the fullest owner's rows
	if (d == 0) {
		status[0] = 0;
		status[1] = abort ? 0u : (uint32_t)s_tot[0][me];
		status[2] = 0;
		status[3] = abort ? 0u : (uint32_t)s_tot[1][me];
		status[4] = abort ? 1u : 0u;
		status[5] = (uint32_t)(s_max[0] > 0xFFFFFFFFull ? 0xFFFFFFFFull : s_max[0]);
		status[6] = (uint32_t)(s_max[1] > 0xFFFFFFFFull ? 0xFFFFFFFFull : s_max[1]);
	}
}

// ------------------------------------------------------------------ the copy

#ifndef HJB_COPY_CHUNK
#define HJB_COPY_CHUNK 4096
#define HJB_COPY_STAGES 2
#define HJB_COPY_DRAIN 1
#endif
constexpr uint32_t kCopyChunk = HJB_COPY_CHUNK;     // tuples per piece and column: 16 KB
constexpr int kCopyStages = HJB_COPY_STAGES;        // stages of 2 x 16 KB: 64 KB + 10 KB of tables per CTA, so three CTAs share an SM.  Measured at 2
                                                    // GPUs (bench.py --exchange staged-serial): one CTA with six stages per SM moves 23.5 GB/s,
                                                    // two with three stages 38, three with two stages 48 -- a pipeline is a chain of waits
constexpr int kCopyDrain = HJB_COPY_DRAIN;          // stores that may still be reading shared memory when a stage is handed back
constexpr size_t kCopySmem = (size_t)kCopyStages * 2 * kCopyChunk * 4;
constexpr uint32_t kCopyEnd = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	asm volatile("{\n.reg .pred p;\nWAIT_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
	                 smem_u32(bar)),
	             "r"(parity)
	             : "memory");
}

struct CopyPiece {
	uint32_t src, dst, rows, owner;            // rows of the 128-byte aligned body piece (may be 0), its first row on both sides
};

// Two warps per CTA.  Warp 0 walks this CTA's pieces and issues the loads (global -> shared, completion on full[stage]);
// warp 1 waits for a piece, issues its two stores (shared -> the owner's columns) and hands a stage back (empty[stage])
// once its store has read shared memory.  Neither waits for the other's bookkeeping.
__global__ void __launch_bounds__(64)
k_peer_copy(const uint32_t *__restrict__ sk, const uint32_t *__restrict__ sv, const PeerCols peers,
            const uint32_t *__restrict__ desc, const uint32_t *__restrict__ abort_flag, int gbits, int me, int skip_me)
{
	const int abits = gbits;                         // one run per owner (an owner's sub-partitions leave as one piece of the staging columns)
	extern __shared__ __align__(128) unsigned char s_stage[];
	__shared__ uint64_t full[kCopyStages], empty[kCopyStages];
	__shared__ uint32_t s_n[64], s_s0[64], s_t0[64], s_end[64], s_total[64], s_cur[64];
	__shared__ uint32_t *s_pk[64], *s_pv[64];
	__shared__ CopyPiece ring[kCopyStages];
	__shared__ uint32_t s_maxp;
	if (*abort_flag) return;
	const uint32_t F = 1u << abits, nsub = F >> gbits, G = 1u << gbits, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const uint32_t *dn = desc + SD_OWN_LEN, *ds0 = desc + SD_OWN_SRC, *dt0 = desc + SD_OWN_DST;
	for (uint32_t d = threadIdx.x; d < F; d += 64) {
		s_n[d] = dn[d];
		s_s0[d] = ds0[d];
		s_t0[d] = dt0[d];
	}
	if (threadIdx.x < 64) {
		s_pk[threadIdx.x] = peers.k[threadIdx.x];
		s_pv[threadIdx.x] = peers.v[threadIdx.x];
		s_cur[threadIdx.x] = 0;
	}
	if (threadIdx.x == 0) {
		for (int i = 0; i < kCopyStages; ++i) {
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[i])));
			asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[i])));
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	// head: rows before the first 128-byte aligned destination row; body: whole 128-byte lines; tail: the rest
	auto head_of = [&](uint32_t d) { return min(s_n[d], (kCopyPhase - (s_t0[d] & (kCopyPhase - 1))) & (kCopyPhase - 1)); };
	auto body_of = [&](uint32_t d) { return (s_n[d] - head_of(d)) & ~(kCopyPhase - 1); };
	if (warp == 0) {
		uint32_t maxp = 0;
		for (uint32_t o = lane; o < G; o += 32) {
			uint32_t run = 0;
			for (uint32_t sub = 0; sub < nsub; ++sub) {
				const uint32_t d = o * nsub + sub;
				if (s_n[d] && !(skip_me && o == (uint32_t)me)) run += max(1u, (body_of(d) + kCopyChunk - 1) / kCopyChunk);
				s_end[d] = run;                       // pieces of the owner's runs up to and including this one
			}
			s_total[o] = run;
			maxp = max(maxp, run);
		}
#pragma unroll
		for (int o = 16; o; o >>= 1) maxp = max(maxp, __shfl_xor_sync(kFullMask, maxp, o));
		if (lane == 0) s_maxp = maxp;
	}
	__syncthreads();
	if (warp == 0) {
		// Virtual piece v = (index within the owner) * na + slot over the na owners that get copies (me + 1, me + 2, ... and
		// this GPU last unless its own runs are already in place): consecutive pieces go to different owners, so every
		// sender spreads its traffic over all receivers at all times.  CTA b takes v = b, b + grid, ...
		const uint32_t na = G - (skip_me ? 1u : 0u);
		const uint32_t vmax = s_maxp * na;
		uint32_t k = 0;
		for (uint32_t v = blockIdx.x; v < vmax; v += gridDim.x) {
			const uint32_t slot = v % na, idx = v / na;
			const uint32_t owner = ((uint32_t)me + 1u + slot) & (G - 1);
			if (idx >= s_total[owner]) continue;
			const uint32_t *end = s_end + owner * nsub;
			uint32_t cur = s_cur[owner];              // this CTA's pieces of an owner come in increasing order
			while (end[cur] <= idx) ++cur;
			__syncwarp();
			if (lane == 0) s_cur[owner] = cur;
			const uint32_t d = owner * nsub + cur, c = idx - (cur ? end[cur - 1] : 0u);
			const uint32_t h = head_of(d), body = body_of(d), off = c * kCopyChunk;
			CopyPiece p;
			p.rows = body > off ? min(kCopyChunk, body - off) : 0u;
			p.src = s_s0[d] + h + off;
			p.dst = s_t0[d] + h + off;
			p.owner = owner;
			if (c == 0) {
				// the unaligned ends of the run: at most 31 rows each, plain stores
				const uint32_t tail = s_n[d] - h - body;
				if (lane < h) {
					s_pk[owner][s_t0[d] + lane] = sk[s_s0[d] + lane];
					s_pv[owner][s_t0[d] + lane] = sv[s_s0[d] + lane];
				}
				if (lane < tail) {
					const uint32_t row = h + body + lane;
					s_pk[owner][s_t0[d] + row] = sk[s_s0[d] + row];
					s_pv[owner][s_t0[d] + row] = sv[s_s0[d] + row];
				}
			}
			if (p.rows == 0) continue;                // a run shorter than one line: its ends were all of it
			const int st = (int)(k % kCopyStages);
			if (lane == 0) {
				if (k >= (uint32_t)kCopyStages) mbar_wait(&empty[st], ((k / kCopyStages) - 1) & 1u);
				ring[st] = p;
				const uint32_t bytes = p.rows * 4;
				unsigned char *buf = s_stage + (size_t)st * 2 * kCopyChunk * 4;
				asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&full[st])), "r"(2 * bytes) : "memory");
				asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf)),
				             "l"(sk + p.src), "r"(bytes), "r"(smem_u32(&full[st])) : "memory");
				asm volatile("cp.async.bulk.shared::cta.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(buf + kCopyChunk * 4)),
				             "l"(sv + p.src), "r"(bytes), "r"(smem_u32(&full[st])) : "memory");
			}
			++k;
		}
		// the end mark travels through the same ring
		if (lane == 0) {
			const int st = (int)(k % kCopyStages);
			if (k >= (uint32_t)kCopyStages) mbar_wait(&empty[st], ((k / kCopyStages) - 1) & 1u);
			ring[st].owner = kCopyEnd;
			asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&full[st])) : "memory");
		}
	} else if (lane == 0) {
		for (uint32_t k = 0;; ++k) {
			const int st = (int)(k % kCopyStages);
			mbar_wait(&full[st], (k / kCopyStages) & 1u);
			const CopyPiece q = ring[st];
			if (q.owner == kCopyEnd) break;
			unsigned char *buf = s_stage + (size_t)st * 2 * kCopyChunk * 4;
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(s_pk[q.owner] + q.dst), "r"(smem_u32(buf)),
			             "r"(q.rows * 4) : "memory");
			asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(s_pv[q.owner] + q.dst),
			             "r"(smem_u32(buf + kCopyChunk * 4)), "r"(q.rows * 4) : "memory");
			asm volatile("cp.async.bulk.commit_group;" ::: "memory");
			if (k >= (uint32_t)kCopyDrain) {
				// stores are committed in order: all but the last kCopyDrain have read their stage
				asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kCopyDrain) : "memory");
				asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[(k - kCopyDrain) % kCopyStages])) : "memory");
			}
		}
		asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
	}
}

int launch_stage_counts(const uint32_t *r_off, const uint32_t *s_off, int abits, unsigned long long *counts, cudaStream_t s)
{
	k_stage_counts<<<1, 512, 0, s>>>(r_off, s_off, 1u << abits, counts);
	return 1;
}

int launch_stage_bases(const unsigned long long *M, int G, int me, int abits, int gbits, uint64_t cap_r, uint64_t cap_s,
                       const uint32_t *child_r, const uint32_t *child_s, uint32_t stage_base_r, uint32_t stage_base_s, int inplace,
                       int nparts, uint32_t *out, uint32_t *status, cudaStream_t s)
{
	k_stage_bases<<<1, 512, 0, s>>>(M, G, me, abits, gbits, cap_r, cap_s, child_r, child_s, stage_base_r, stage_base_s, inplace, nparts, out,
	                                status);
	return 1;
}

int launch_peer_copy(const uint32_t *sk, const uint32_t *sv, const PeerCols &peers, const uint32_t *desc, const uint32_t *abort_flag,
                     int gbits, int me, int skip_me, cudaStream_t s, KernelTimer *t)
{
	KernelTimer off;
	off.enabled = false;
	off.n = 0;
	if (!t) t = &off;
	static const int ctas = [] {
		const int v = getenv("HJB_COPY_CTAS") ? atoi(getenv("HJB_COPY_CTAS")) : 48;      // three per SM: 16 SMs
		return v < 1 ? 48 : v;
	}();
	static std::atomic<unsigned long long> done_mask{0};         // the attribute is per device
	int dev = 0;
	cudaGetDevice(&dev);
	const unsigned long long bit = 1ull << (dev & 63);
	if (!(done_mask.fetch_or(bit) & bit)) cudaFuncSetAttribute(k_peer_copy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCopySmem);
	t->start(KK_PEER_COPY, s);
	k_peer_copy<<<ctas, 64, kCopySmem, s>>>(sk, sv, peers, desc, abort_flag, gbits, me, skip_me);
	t->stop(s);
	return 1;
}

}  // namespace hjb
