// capi.cu -- the extern "C" boundary (include/hjb200.h): context, workspace, and the host-side
// orchestration that replaces main() + run()/run_hj() of the reference (npj.cpp:769-1125,
// phj.cpp:1646-2231, cpra2.cpp:1697-2231): plan, launch the kernels on one stream, read back
// the count and checksums.  No CPU join path exists here.
#include "hj_internal.h"
#include <chrono>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

using namespace hjb;

constexpr int kMaxHostSlices = 16;               // probe-side slices of the pipelined host entry points
constexpr uint64_t kMinHostSlice = 1u << 22;     // tuples; smaller probe sides are copied in one piece

// HJB_HOST_SLICE=<tuples> overrides the minimum slice (tests exercise the slicing on small inputs)
static uint64_t host_slice_min()
{
	const char *e = getenv("HJB_HOST_SLICE");
	const long long v = e ? atoll(e) : 0;
	return v >= 4 ? (uint64_t)v : kMinHostSlice;
}

// what a captured PHJ launch sequence depends on: every pointer and size baked into its kernel arguments
struct PhjGraphKey {
	const void *rk, *rv, *sk, *sv, *ws, *out;
	uint64_t nr, ns, out_cap, out_capacity;
	cudaStream_t stream;
	uint32_t seed, part_tuples, owner;
	int consumed, materialize, radix_bits[4];
};

struct PhjState;

struct hjb_ctx {
	int device, sms;
	cudaStream_t stream;
	bool own_stream;
	char err[512];
	char *ws;                 // workspace arena (partition buffers, tables, scratch)
	size_t ws_bytes;
	uint32_t *out_cols;       // 3 result columns, out_cap rows each
	uint64_t out_cap;
	char *in_buf;             // device copies of host inputs (hjb_*_host)
	size_t in_bytes;
	uint32_t *h_rows;         // pinned host result rows (hjb_*_host)
	uint64_t h_rows_cap;
	char *split_buf;          // CPRA send buffers
	size_t split_bytes;
	unsigned long long *d_scalars, *h_scalars;   // 16 each; h_ pinned
	uint32_t *h_small;        // pinned, 256 uint32
	cudaEvent_t ev[12];
	// hjb_*_host pipeline: copy-in / copy-out streams, per-slice events, row cursor after each slice (pinned)
	cudaStream_t pipe_in, pipe_out;
	cudaEvent_t pipe_ev[2][kMaxHostSlices + 1];
	unsigned long long *h_cursor;
	bool pipe_ready;
	// PHJ launch sequence as a CUDA graph (replayed when the same buffers are joined again)
	PhjGraphKey gkey;
	int gkey_seen;
	bool gvalid;
	cudaGraphExec_t gexec;
	uint32_t glaunches;
	uint32_t launches;
	KernelTimer timer;        // per-kernel times of the last join (hjb_set_profiling)
	char *recv_buf[4];        // CPRA fused exchange: receive columns r_keys r_vals s_keys s_vals
	uint64_t recv_cap[2];
	RadixPassArgs pending[2]; // hjb_cpra_count -> hjb_cpra_scatter_peer
	int pending_gpus;
	uint32_t *cpra_dev;       // device words of the fused exchange: bases, receive ranges, abort flag (enum CD_*)
	// stream-ordered CPRA (hjb_cpra_bind ... hjb_cpra_finish)
	void *bind_peer[4][64];
	int bind_gpu, bind_gpus;
	uint64_t bind_cap[2];
	int step_state;           // 0 idle, 1 counted, 2 scattered, 3 join enqueued
	uint32_t step_launches;
	struct PhjState *step_phj;
	hjb_opts step_opts;
	// staged exchange (hjb_cpra_stage_*): stage A's output columns live in split_buf
	uint32_t *stage_dev;      // device words of k_stage_bases (enum SD_*)
	uint32_t *stage_k[2], *stage_v[2];
	int stage_abits, stage_state[2];   // per relation: 0 idle, 1 counted, 2 scattered
	int stage_parts;          // the runs leave, and are processed by their owners, in this many parts (ranges of sub-partitions)
	uint32_t stage_copied[2], stage_done[2];   // bit k: part k of the relation has been copied / passed (and, for S, joined)
	uint64_t recv_stage_off[2], recv_stage_cap[2];   // the staging region behind the receive region of recv_buf's columns (rows)
	uint32_t *h_stage;        // pinned: the per-owner runs of both relations + the verdict, for the copy-engine form of the exchange
	int stage_copy_engine;    // this step's copies are cudaMemcpyAsync calls (HJB_STAGE_COPY=ce) instead of k_peer_copy
	int stage_inplace;        // stage A writes into recv_buf's columns: own runs at their final rows, the others into the staging region
	uint32_t stage_base[2];
	// heavy-hitter handling (hjb_cpra_split_hot ... hjb_cpra_hot_join)
	char *skew_buf;           // cold / hot copies of the probe chunk
	size_t skew_bytes;
	hjb_rel hot_s, hot_r;     // arguments of the enqueued hot join (re-run when the result columns had to grow)
	bool hot_pending;
};

static char g_create_err[512];


#define CK(call)                                                                                   \
	do {                                                                                           \
		cudaError_t e_ = (call);                                                                   \
		if (e_ != cudaSuccess) {                                                                   \
			snprintf(ctx->err, sizeof ctx->err, "%s:%d %s: %s", __FILE__, __LINE__, #call,         \
			         cudaGetErrorString(e_));                                                      \
			return e_ == cudaErrorMemoryAllocation ? HJB_E_NOMEM : HJB_E_CUDA;                     \
		}                                                                                          \
	} while (0)

static int fail(hjb_ctx *ctx, int code, const char *msg)
{
	if (ctx) snprintf(ctx->err, sizeof ctx->err, "%s", msg);
	return code;
}

static double wall_now()
{
	return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

extern "C" int hjb_version(void) { return HJB_VERSION; }

extern "C" const char *hjb_last_error(const hjb_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int hjb_create(int device, hjb_ctx **out)
{
	if (!out) return HJB_E_INVALID;
	*out = nullptr;
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		snprintf(g_create_err, sizeof g_create_err, "no CUDA device (%s): libhjb200 has no CPU fallback",
		         e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
		return HJB_E_NODEVICE;
	}
	if (device < 0 || device >= ndev) {
		snprintf(g_create_err, sizeof g_create_err, "device %d out of range [0,%d)", device, ndev);
		return HJB_E_INVALID;
	}
	hjb_ctx *ctx = (hjb_ctx *)calloc(1, sizeof(hjb_ctx));
	if (!ctx) return HJB_E_NOMEM;
	ctx->device = device;
	cudaDeviceProp prop;
	if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
		snprintf(g_create_err, sizeof g_create_err, "cannot open device %d", device);
		free(ctx);
		return HJB_E_CUDA;
	}
	if (prop.major < 10) {
		snprintf(g_create_err, sizeof g_create_err, "device %d is sm_%d%d; libhjb200 is built for sm_100a only", device,
		         prop.major, prop.minor);
		free(ctx);
		return HJB_E_NODEVICE;
	}
	ctx->sms = prop.multiProcessorCount;
	bool ok = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) == cudaSuccess;
	ctx->own_stream = true;
	ok = ok && cudaMalloc(&ctx->d_scalars, 16 * 8) == cudaSuccess;
	ok = ok && cudaHostAlloc(&ctx->h_scalars, 16 * 8, cudaHostAllocDefault) == cudaSuccess;
	ok = ok && cudaHostAlloc(&ctx->h_small, 256 * 4, cudaHostAllocDefault) == cudaSuccess;
	for (int i = 0; ok && i < 12; ++i) ok = cudaEventCreate(&ctx->ev[i]) == cudaSuccess;
	for (int i = 0; ok && i < KernelTimer::kMaxLaunches; ++i)
		ok = cudaEventCreate(&ctx->timer.beg[i]) == cudaSuccess && cudaEventCreate(&ctx->timer.end[i]) == cudaSuccess;
	if (!ok) {
		snprintf(g_create_err, sizeof g_create_err, "context setup failed: %s", cudaGetErrorString(cudaGetLastError()));
		free(ctx);
		return HJB_E_CUDA;
	}
	*out = ctx;
	return HJB_OK;
}

extern "C" int hjb_destroy(hjb_ctx *ctx)
{
	if (!ctx) return HJB_E_INVALID;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	cudaFree(ctx->ws);
	cudaFree(ctx->out_cols);
	cudaFree(ctx->in_buf);
	cudaFree(ctx->split_buf);
	for (int i = 0; i < 4; ++i) cudaFree(ctx->recv_buf[i]);
	cudaFree(ctx->cpra_dev);
	cudaFree(ctx->stage_dev);
	cudaFreeHost(ctx->h_stage);
	cudaFree(ctx->skew_buf);
	free(ctx->step_phj);
	cudaFree(ctx->d_scalars);
	cudaFreeHost(ctx->h_scalars);
	cudaFreeHost(ctx->h_small);
	cudaFreeHost(ctx->h_rows);
	if (ctx->gvalid) cudaGraphExecDestroy(ctx->gexec);
	if (ctx->pipe_ready) {
		cudaStreamDestroy(ctx->pipe_in);
		cudaStreamDestroy(ctx->pipe_out);
		for (int i = 0; i < 2; ++i)
			for (int k = 0; k <= kMaxHostSlices; ++k) cudaEventDestroy(ctx->pipe_ev[i][k]);
		cudaFreeHost(ctx->h_cursor);
	}
	for (int i = 0; i < 12; ++i) cudaEventDestroy(ctx->ev[i]);
	for (int i = 0; i < KernelTimer::kMaxLaunches; ++i) {
		cudaEventDestroy(ctx->timer.beg[i]);
		cudaEventDestroy(ctx->timer.end[i]);
	}
	if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
	free(ctx);
	return HJB_OK;
}

extern "C" int hjb_set_stream(hjb_ctx *ctx, void *cuda_stream)
{
	if (!ctx) return HJB_E_INVALID;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
	ctx->stream = (cudaStream_t)cuda_stream;
	ctx->own_stream = false;
	return HJB_OK;
}

extern "C" int hjb_synchronize(hjb_ctx *ctx)
{
	if (!ctx) return HJB_E_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	return HJB_OK;
}

// ------------------------------------------------------------------ buffers

static int grow_device(hjb_ctx *ctx, char **buf, size_t *have, size_t need)
{
	if (need <= *have) return HJB_OK;
	CK(cudaStreamSynchronize(ctx->stream));
	if (*buf) CK(cudaFree(*buf));
	*buf = nullptr;
	*have = 0;
	need = (need + ((size_t)1 << 21) - 1) & ~(((size_t)1 << 21) - 1);
	CK(cudaMalloc(buf, need));
	*have = need;
	return HJB_OK;
}

static int grow_out(hjb_ctx *ctx, uint64_t rows)
{
	if (rows <= ctx->out_cap) return HJB_OK;
	size_t have = (size_t)ctx->out_cap * 12;
	char *p = (char *)ctx->out_cols;
	rows = (rows + 1023) & ~1023ull;
	int rc = grow_device(ctx, &p, &have, (size_t)rows * 12);
	ctx->out_cols = (uint32_t *)p;
	ctx->out_cap = rc == HJB_OK ? have / 12 : 0;
	return rc;
}

struct Bump {
	char *base;
	size_t off;
	template <typename T>
	T *take(size_t count)
	{
		T *p = reinterpret_cast<T *>(base + off);
		off += (count * sizeof(T) + 255) & ~(size_t)255;
		return p;
	}
};
static size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

static int check_rel(hjb_ctx *ctx, const hjb_rel *r, bool device_cols)
{
	if (!r) return fail(ctx, HJB_E_INVALID, "null relation");
	// positions are 32-bit; the last 2^16 are kept free so that "position + one table fill" never wraps
	if (r->tuples > 0xFFFF0000ull) return fail(ctx, HJB_E_INVALID, "more than 2^32 - 2^16 tuples per relation per GPU");
	if (r->tuples && (!r->keys || !r->vals)) return fail(ctx, HJB_E_INVALID, "null column");
	if (device_cols && r->tuples && ((((uintptr_t)r->keys) | ((uintptr_t)r->vals)) & 15))
		return fail(ctx, HJB_E_INVALID, "device columns must be 16-byte aligned");
	return HJB_OK;
}

// NPJ table load factor when the caller does not set one: a table that overflows L2 probes fastest at 0.75 (fewer
// DRAM sectors), a cache-resident one at a low load (few full buckets, so few probes walk into a second bucket --
// a warp pays that walk's latency as soon as one of its lanes takes it)
static double npj_default_load(uint64_t build_tuples)
{
	static const double small = getenv("HJB_NPJ_SMALL_LOAD") ? atof(getenv("HJB_NPJ_SMALL_LOAD")) : 0.5;
	return build_tuples * 16 <= (32u << 20) ? small : 0.75;
}

static const hjb_opts kDefaultOpts = {1, 0, 0.0, {0, 0, 0, 0}, 0, 0, {0}};

extern "C" uint32_t hjb_hash_factor(uint32_t seed, int which)
{
	// odd multipliers; distinct per stage so that the table hash is independent of the radix digits
	static const uint32_t base[4] = {0x9E3779B1u, 0x85EBCA6Bu, 0xC2B2AE35u, 0x27D4EB2Fu};
	uint32_t x = base[which & 3];
	if (seed) {
		x ^= seed * 0x01000193u;
		x ^= x >> 15;
		x *= 0x2C1B3C6Du;
		x ^= x >> 12;
	}
	return x | 1u;
}

static void zero_result(hjb_result *out)
{
	memset(out, 0, sizeof *out);
}

static void timer_reset(hjb_ctx *ctx)
{
	ctx->timer.n = 0;
	ctx->timer.dropped = 0;
	memset(ctx->timer.ms, 0, sizeof ctx->timer.ms);
	memset(ctx->timer.launches, 0, sizeof ctx->timer.launches);
}

// after a stream synchronize: fold the recorded event pairs into per-kernel totals
static void timer_collect(hjb_ctx *ctx)
{
	KernelTimer &t = ctx->timer;
	for (int i = 0; i < t.n; ++i) {
		float ms = 0;
		if (cudaEventElapsedTime(&ms, t.beg[i], t.end[i]) == cudaSuccess) {
			t.ms[t.kind[i]] += ms;
			t.launches[t.kind[i]] += 1;
		}
	}
	t.n = 0;
}

extern "C" int hjb_set_profiling(hjb_ctx *ctx, int on)
{
	if (!ctx) return HJB_E_INVALID;
	ctx->timer.enabled = on != 0;
	timer_reset(ctx);
	return HJB_OK;
}

extern "C" const char *hjb_kernel_name(int kind)
{
	static const char *names[KK_COUNT] = {"k_make_items", "k_hist", "k_scan", "k_scatter", "k_join_tasks",
	                                      "k_partition_join", "k_npj_build", "k_npj_probe", "k_scatter_bulk", "k_peer_copy"};
	return kind >= 0 && kind < KK_COUNT ? names[kind] : nullptr;
}

extern "C" int hjb_kernel_times(hjb_ctx *ctx, float *ms, uint32_t *launches, int max_kinds)
{
	if (!ctx || !ms || !launches) return HJB_E_INVALID;
	if (ctx->timer.dropped) return fail(ctx, HJB_E_NOMEM, "hjb_kernel_times: more launches than event pairs, times incomplete");
	for (int k = 0; k < max_kinds && k < KK_COUNT; ++k) {
		ms[k] = ctx->timer.ms[k];
		launches[k] = ctx->timer.launches[k];
	}
	return KK_COUNT;
}

static int read_scalars(hjb_ctx *ctx, hjb_result *out)
{
	CK(cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	timer_collect(ctx);
	out->count = ctx->h_scalars[1];
	out->sum_key = ctx->h_scalars[2];
	out->sum_outer = ctx->h_scalars[3];
	out->sum_inner = ctx->h_scalars[4];
	return HJB_OK;
}

static void set_rows(hjb_ctx *ctx, hjb_result *out, int materialize)
{
	if (materialize) {
		out->keys = ctx->out_cols;
		out->outer_vals = ctx->out_cols + ctx->out_cap;
		out->inner_vals = ctx->out_cols + 2 * ctx->out_cap;
		out->rows_on_device = 1;
	}
}

// ------------------------------------------------------------------ NPJ

static int npj_device(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *o, hjb_result *out)
{
	int rc;
	if ((rc = check_rel(ctx, R, true)) || (rc = check_rel(ctx, S, true))) return rc;
	zero_result(out);
	CK(cudaSetDevice(ctx->device));
	timer_reset(ctx);
	if (R->tuples == 0 || S->tuples == 0) return HJB_OK;
	// measured on B200: a table that overflows L2 probes fastest at 0.75 (fewer DRAM sectors), a
	// cache-resident one at 0.5 (shorter bucket chains)
	const double load = o->npj_load > 0.0 ? o->npj_load : npj_default_load(R->tuples);
	if (load > 0.95) return fail(ctx, HJB_E_INVALID, "npj_load must be <= 0.95");
	uint64_t buckets = (uint64_t)ceil((double)R->tuples / load / 4.0) + 1;       // +1: at least one empty slot
	if (buckets > 0xFFFFFFFFull) return fail(ctx, HJB_E_INVALID, "table too large");
	if ((rc = grow_device(ctx, &ctx->ws, &ctx->ws_bytes, buckets * 32))) return rc;
	uint64_t cap = 0;
	if (o->materialize) {
		cap = o->out_capacity ? o->out_capacity : (S->tuples > R->tuples ? S->tuples : R->tuples);
		if ((rc = grow_out(ctx, cap))) return rc;
	}
	NpjArgs a;
	a.rk = R->keys; a.rv = R->vals; a.sk = S->keys; a.sv = S->vals;
	a.nr = R->tuples; a.ns = S->tuples;
	a.table = (uint64_t *)ctx->ws;
	a.buckets = buckets;
	a.factor = hjb_hash_factor(o->seed, 1);
	a.scalars = ctx->d_scalars;
	a.materialize = o->materialize;
	cudaStream_t s = ctx->stream;
	uint32_t launches = 0;
	for (int attempt = 0; attempt < 2; ++attempt) {
		a.out_k = ctx->out_cols;
		a.out_o = ctx->out_cols + ctx->out_cap;
		a.out_i = ctx->out_cols + 2 * ctx->out_cap;
		a.out_cap = o->materialize ? ctx->out_cap : 0;
		CK(cudaEventRecord(ctx->ev[0], s));
		CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));
		launches += launch_npj_build(a, s, ctx->sms, &ctx->timer);
		CK(cudaEventRecord(ctx->ev[1], s));
		launches += launch_npj_probe(a, s, ctx->sms, &ctx->timer);
		CK(cudaEventRecord(ctx->ev[2], s));
		CK(cudaGetLastError());
		if ((rc = read_scalars(ctx, out))) return rc;
		if (!o->materialize || out->count <= ctx->out_cap) break;
		if (attempt == 1) return fail(ctx, HJB_E_CUDA, "result overflow after regrow");
		if ((rc = grow_out(ctx, out->count))) return rc;       // duplicates: more rows than max(|R|,|S|)
	}
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2]));
	out->seconds = ms * 1e-3;
	CK(cudaEventElapsedTime(&out->phase_ms[1], ctx->ev[0], ctx->ev[1]));   // table init + build
	CK(cudaEventElapsedTime(&out->phase_ms[2], ctx->ev[1], ctx->ev[2]));   // probe
	out->kernel_launches = launches;
	out->partitions = (uint32_t)buckets;
	set_rows(ctx, out, o->materialize);
	ctx->launches += launches;
	return HJB_OK;
}

extern "C" int hjb_npj_device(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	return npj_device(ctx, R, S, opts ? opts : &kDefaultOpts, out);
}

// ------------------------------------------------------------------ PHJ / CPRA local join

struct Plan {
	int npass;
	int bits[kMaxPasses];
	int total_bits;
};

// Reference planner (phj.cpp:1791-1808): partitions = tuples / hash_table_limit, 1-4 passes of
// equal fan-out.  Here: enough bits that a build partition averages part_tuples (a quarter of
// the shared-memory table), passes of at most 8 bits (256-way scatter keeps per-digit runs long).
static int make_plan(hjb_ctx *ctx, uint64_t nr, uint64_t ns, const hjb_opts *o, int consumed, Plan *p)
{
	memset(p, 0, sizeof *p);
	int user = 0;
	for (int i = 0; i < kMaxPasses; ++i) user += o->radix_bits[i] != 0;
	if (user) {
		for (int i = 0; i < kMaxPasses && o->radix_bits[i]; ++i) {
			if (o->radix_bits[i] < 1 || o->radix_bits[i] > kMaxRadixBits)
				return fail(ctx, HJB_E_INVALID, "radix_bits must be in [1,11]");
			p->bits[p->npass++] = o->radix_bits[i];
			p->total_bits += o->radix_bits[i];
		}
	} else {
		// with <= 16 hash bits left below the partition id the join kernel addresses payloads directly
		// (csrc/part_join.cu: one fill holds 6144 build tuples, so partitions may average 4096); worth
		// two full passes as soon as the input is not tiny.  Otherwise hash tables, 2048 per partition.
		const bool big = !o->part_tuples && nr + ns >= (1u << 22) && nr >= (1u << 20);   // a small build side: one pass, hash tables
		const uint32_t target = o->part_tuples ? o->part_tuples : (big ? 4096u : kDefaultPartTuples);
		int tb = 0;
		while (tb < 28 && (nr >> tb) > target) ++tb;
		if (big && consumed + tb < 16) tb = 16 - consumed;
		p->total_bits = tb;
		p->npass = (tb + 7) / 8;
		for (int i = 0; i < p->npass; ++i) p->bits[i] = tb / p->npass + (i < tb % p->npass ? 1 : 0);
	}
	if (consumed + p->total_bits > 32) return fail(ctx, HJB_E_INVALID, "more than 32 radix bits");
	if (p->total_bits > 22) return fail(ctx, HJB_E_INVALID, "more than 2^22 partitions");
	return HJB_OK;
}

static size_t phj_workspace(uint64_t nr, uint64_t ns, const Plan &p, size_t *radix_scratch, int pre_bits = 0, uint32_t pre_segs = 1,
                            uint32_t chunk_div = 1)
{
	size_t total = 0;
	const int nb = p.npass >= 2 ? 2 : p.npass;
	total += (size_t)nb * 2 * (pad256(nr * 4) + pad256(ns * 4));
	const uint32_t P = 1u << (pre_bits + p.total_bits);
	total += 4 * pad256(((size_t)P + 1) * 4);          // r_off / s_off, two generations each
	total += pad256(((size_t)P + 1) * 4) + pad256(256 + 8 * (((size_t)P >> 10) + 1));   // task prefix, counter + block status words
	size_t rs = 0;
	uint32_t np = 1u << pre_bits;
	for (int i = 0; i < p.npass; ++i) {
		uint32_t chunk, mi, tiles;
		const uint32_t np_eff = i == 0 ? np * pre_segs : np;          // every range of a pre-partitioned parent may end in a short item
		size_t a = radix_scratch_bytes(nr, np_eff, p.bits[i], &chunk, &mi, &tiles, nullptr, i == 0 ? chunk_div : 1u);
		size_t b = radix_scratch_bytes(ns, np_eff, p.bits[i], &chunk, &mi, &tiles, nullptr, i == 0 ? chunk_div : 1u);
		if (a > rs) rs = a;
		if (b > rs) rs = b;
		np <<= p.bits[i];
	}
	*radix_scratch = rs;
	return total + rs + 4096;
}

struct Partitioned {
	const uint32_t *k, *v;
	const uint32_t *off;
};

// all passes over one relation; ping-pongs between two workspace buffers
static int partition_relation(hjb_ctx *ctx, const hjb_rel *rel, const Plan &p, int consumed, uint32_t factor,
                              uint32_t *bufk[2], uint32_t *bufv[2], uint32_t *off[2], char *scratch,
                              Partitioned *res, uint32_t *launches, const uint32_t *dev_range = nullptr, int pre_bits = 0,
                              const uint32_t *seg = nullptr, uint32_t nseg = 0)
{
	// dev_range: {0, tuples} in DEVICE memory -- the relation's size is then only known there (CPRA's receive
	// buffers in the stream-ordered path) and rel->tuples is an upper bound that sizes grids and scratch.
	// pre_bits > 0: the relation arrives cut into 2^pre_bits partitions by the bits below `consumed` (the staged
	// exchange); dev_range then holds their 2^pre_bits + 1 cumulative sizes and seg the nseg ranges each consists of.
	const uint32_t *ink = rel->keys, *inv = rel->vals;
	const uint32_t *parent = dev_range;
	uint32_t np = 1u << pre_bits;
	int used = consumed + pre_bits;
	for (int i = 0; i < p.npass; ++i) {
		RadixPassArgs a = {};
		a.keys = ink; a.vals = inv;
		a.keys_out = bufk[i & 1]; a.vals_out = bufv[i & 1];
		a.n = rel->tuples;
		a.np = np;
		a.parent_off = parent;
		a.child_off = off[i & 1];
		a.factor = factor;
		a.bits = p.bits[i];
		a.rshift = 32 - used - p.bits[i];
		if (i == 0 && seg) {
			a.seg = seg;
			a.nseg = nseg;
		}
		radix_carve(a, scratch, true);
		*launches += launch_radix_pass(a, ctx->stream, ctx->sms, &ctx->timer);
		ink = a.keys_out; inv = a.vals_out;
		parent = a.child_off;
		np <<= a.bits;
		used += a.bits;
	}
	res->k = ink; res->v = inv; res->off = parent;
	return HJB_OK;
}

// One PHJ join in pieces, so that the probe side can arrive in slices (hjb_phj_host overlaps the
// slices' copies with the kernels): phj_setup plans and carves the workspace, phj_build partitions
// R, phj_probe partitions one S slice and joins it against R's partitions.  Rows of successive
// slices are appended (the row cursor d_scalars[0] and the checksums are not reset in between).
struct PhjState {
	Plan plan;
	uint32_t P, owner, radix_factor;
	int consumed, pre_bits, big_fill;
	const uint32_t *seg[2];           // staged exchange: the ranges of the received sub-partitions, R and S
	uint32_t nseg;
	uint32_t *rbk[2], *rbv[2], *sbk[2], *sbv[2], *roff[2], *soff[2], *task_prefix, *task_counter;
	char *scratch;
	Partitioned pr, ps;
	JoinArgs j;
};

static int phj_setup(hjb_ctx *ctx, uint64_t nr, uint64_t ns_slice, uint64_t ns_total, const hjb_opts *o, int consumed,
                     uint32_t owner, PhjState *st, uint64_t nr_plan = 0, uint64_t ns_plan = 0, const Plan *given = nullptr,
                     int pre_bits = 0, uint32_t pre_segs = 1, uint32_t chunk_div = 1)
{
	// nr / ns_slice / ns_total size the buffers; the plan is made for nr_plan / ns_plan tuples when given (sizes
	// that are only upper bounds here, the expected sizes there), or is `given` (the staged exchange: pre_bits of
	// the partition id are already in place when the relation arrives)
	int rc;
	memset(st, 0, sizeof *st);
	if (given) st->plan = *given;
	else if ((rc = make_plan(ctx, nr_plan ? nr_plan : nr, ns_plan ? ns_plan : ns_total, o, consumed, &st->plan))) return rc;
	st->pre_bits = pre_bits;
	const Plan &plan = st->plan;
	size_t rscratch;
	const size_t need = phj_workspace(nr, ns_slice, plan, &rscratch, pre_bits, pre_segs, chunk_div);
	if ((rc = grow_device(ctx, &ctx->ws, &ctx->ws_bytes, need))) return rc;
	if (o->materialize) {
		const uint64_t cap = o->out_capacity ? o->out_capacity : (ns_total > nr ? ns_total : nr);
		if ((rc = grow_out(ctx, cap))) return rc;
	}
	st->P = 1u << (pre_bits + plan.total_bits);
	st->consumed = consumed;
	st->owner = owner;
	st->radix_factor = hjb_hash_factor(o->seed, 0);
	Bump b = {ctx->ws, 0};
	const int nb = plan.npass >= 2 ? 2 : plan.npass;
	for (int i = 0; i < nb; ++i) {
		st->rbk[i] = b.take<uint32_t>(nr); st->rbv[i] = b.take<uint32_t>(nr);
		st->sbk[i] = b.take<uint32_t>(ns_slice); st->sbv[i] = b.take<uint32_t>(ns_slice);
	}
	for (int i = 0; i < 2; ++i) st->roff[i] = b.take<uint32_t>(st->P + 1);
	for (int i = 0; i < 2; ++i) st->soff[i] = b.take<uint32_t>(st->P + 1);
	st->task_prefix = b.take<uint32_t>(st->P + 1);
	st->task_counter = b.take<uint32_t>(64 + 2 * ((size_t)(st->P >> 10) + 1));   // counter, then one 64-bit status word per 1024 partitions
	st->scratch = b.take<char>(rscratch);
	return HJB_OK;
}

__global__ void k_set_pair(uint32_t *dst, uint32_t a, uint32_t b, const uint32_t *src)
{
	dst[0] = src ? src[0] : a;
	dst[1] = src ? src[1] : b;
}

// one side through all radix passes (no pass at all: a single partition [0, n))
static int phj_partition_side(hjb_ctx *ctx, PhjState *st, const hjb_rel *rel, bool build_side, uint32_t *launches,
                              const uint32_t *dev_range = nullptr)
{
	cudaStream_t s = ctx->stream;
	Partitioned *res = build_side ? &st->pr : &st->ps;
	uint32_t **off = build_side ? st->roff : st->soff;
	if (st->plan.npass == 0) {
		// by value in the kernel arguments: nothing on the host that a later call (or a graph replay) could find changed
		k_set_pair<<<1, 1, 0, s>>>(off[0], 0u, (uint32_t)rel->tuples, dev_range);
		res->k = rel->keys; res->v = rel->vals; res->off = off[0];
		return HJB_OK;
	}
	return partition_relation(ctx, rel, st->plan, st->consumed, st->radix_factor, build_side ? st->rbk : st->sbk,
	                          build_side ? st->rbv : st->sbv, off, st->scratch, res, launches, dev_range, st->pre_bits,
	                          st->seg[build_side ? 0 : 1], st->nseg);
}

static int phj_launch_join(hjb_ctx *ctx, PhjState *st, const hjb_opts *o, uint32_t *launches, uint32_t p0 = 0, uint32_t p1 = 0)
{
	// [p0, p1): the partitions to join (default: all)
	if (p1 == 0) p1 = st->P;
	JoinArgs &j = st->j;
	j.rk = st->pr.k; j.rv = st->pr.v; j.sk = st->ps.k; j.sv = st->ps.v;
	j.r_off = st->pr.off + p0; j.s_off = st->ps.off + p0;
	j.P = p1 - p0;
	j.radix_factor = st->radix_factor;
	j.rem_bits = 32 - st->consumed - st->pre_bits - st->plan.total_bits;
	j.big_fill = st->big_fill;
	j.owner = st->owner;
	j.owner_bits = st->consumed;
	j.table_factor = hjb_hash_factor(o->seed, 1);
	j.task_prefix = st->task_prefix;
	j.task_counter = st->task_counter;
	j.s_task = 16384;
	j.scalars = ctx->d_scalars;
	j.materialize = o->materialize;
	j.out_k = ctx->out_cols;
	j.out_o = ctx->out_cols + ctx->out_cap;
	j.out_i = ctx->out_cols + 2 * ctx->out_cap;
	j.out_cap = o->materialize ? ctx->out_cap : 0;
	*launches += launch_partition_join(j, ctx->stream, ctx->sms, &ctx->timer);
	return HJB_OK;
}

static int phj_device(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *o, int consumed,
                      hjb_result *out, uint32_t owner = 0)
{
	int rc;
	if ((rc = check_rel(ctx, R, true)) || (rc = check_rel(ctx, S, true))) return rc;
	zero_result(out);
	CK(cudaSetDevice(ctx->device));
	if (consumed == 0) timer_reset(ctx);       // a CPRA join keeps the times of its count / scatter steps
	if (R->tuples == 0 || S->tuples == 0) return HJB_OK;
	PhjState st;
	if ((rc = phj_setup(ctx, R->tuples, S->tuples, S->tuples, o, consumed, owner, &st))) return rc;
	cudaStream_t s = ctx->stream;
	uint32_t launches = 0;
	// The whole sequence (3 memsets, ~14 kernels) depends on nothing the host reads in between, so a
	// repeated join of the same buffers replays it as ONE graph launch: first call eager, second call
	// captured, from then on replayed.  Not while per-kernel events are wanted (hjb_set_profiling).
	static const int graphs = getenv("HJB_GRAPHS") ? atoi(getenv("HJB_GRAPHS")) : 1;
	PhjGraphKey key;
	memset(&key, 0, sizeof key);
	key.rk = R->keys; key.rv = R->vals; key.sk = S->keys; key.sv = S->vals; key.ws = ctx->ws; key.out = ctx->out_cols;
	key.nr = R->tuples; key.ns = S->tuples; key.out_cap = ctx->out_cap; key.out_capacity = o->out_capacity;
	key.stream = s; key.seed = o->seed; key.part_tuples = o->part_tuples; key.owner = owner;
	key.consumed = consumed; key.materialize = o->materialize;
	for (int i = 0; i < 4; ++i) key.radix_bits[i] = o->radix_bits[i];
	const bool same = memcmp(&key, &ctx->gkey, sizeof key) == 0;
	bool replayed = false;
	if (graphs && !ctx->timer.enabled && same && (ctx->gvalid || ctx->gkey_seen >= 1)) {
		if (!ctx->gvalid) {
			cudaGraph_t g = nullptr;
			uint32_t gl = 0;
			bool ok = cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
			if (ok) {
				ok = cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s) == cudaSuccess;
				ok = ok && phj_partition_side(ctx, &st, R, true, &gl) == HJB_OK;
				ok = ok && phj_partition_side(ctx, &st, S, false, &gl) == HJB_OK;
				ok = ok && phj_launch_join(ctx, &st, o, &gl) == HJB_OK;
				ok = (cudaStreamEndCapture(s, &g) == cudaSuccess) && ok && g;
			}
			if (ok) ok = cudaGraphInstantiate(&ctx->gexec, g, 0) == cudaSuccess;
			if (g) cudaGraphDestroy(g);
			if (ok) {
				ctx->gvalid = true;
				ctx->glaunches = gl;
			} else {
				cudaGetLastError();           // capture refused (e.g. the stream is being captured by the caller): stay eager
				ctx->gkey_seen = -1000000;
			}
		}
		if (ctx->gvalid) {
			CK(cudaEventRecord(ctx->ev[0], s));
			CK(cudaGraphLaunch(ctx->gexec, s));
			CK(cudaEventRecord(ctx->ev[3], s));
			CK(cudaGetLastError());
			if ((rc = read_scalars(ctx, out))) return rc;
			if (ctx->h_scalars[7]) return fail(ctx, HJB_E_INVALID, "hjb_cpra_join_local: a tuple does not hash into this owner's range");
			launches = ctx->glaunches;
			replayed = !(o->materialize && out->count > ctx->out_cap);      // overflow: the eager path below regrows
		}
	}
	if (!replayed) {
		if (!same || ctx->gvalid) {
			if (ctx->gvalid) cudaGraphExecDestroy(ctx->gexec);
			ctx->gvalid = false;
			ctx->gkey = key;
			ctx->gkey_seen = 0;
		}
		ctx->gkey_seen += 1;
		launches = 0;
		CK(cudaEventRecord(ctx->ev[0], s));
		CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));
		if ((rc = phj_partition_side(ctx, &st, R, true, &launches))) return rc;
		CK(cudaEventRecord(ctx->ev[1], s));
		if ((rc = phj_partition_side(ctx, &st, S, false, &launches))) return rc;
		CK(cudaEventRecord(ctx->ev[2], s));
		for (int attempt = 0; attempt < 2; ++attempt) {
			if ((rc = phj_launch_join(ctx, &st, o, &launches))) return rc;
			CK(cudaEventRecord(ctx->ev[3], s));
			CK(cudaGetLastError());
			if ((rc = read_scalars(ctx, out))) return rc;
			if (ctx->h_scalars[7]) return fail(ctx, HJB_E_INVALID, "hjb_cpra_join_local: a tuple does not hash into this owner's range");
			if (!o->materialize || out->count <= ctx->out_cap) break;
			if (attempt == 1) return fail(ctx, HJB_E_CUDA, "result overflow after regrow");
			if ((rc = grow_out(ctx, out->count))) return rc;
			CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));         // rerun the join phase only
		}
	}
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
	out->seconds = ms * 1e-3;
	if (!replayed) {                           // a replayed graph has no phase events inside
		if (st.plan.npass) {
			CK(cudaEventElapsedTime(&out->phase_ms[0], ctx->ev[0], ctx->ev[1]));   // all passes over R
			CK(cudaEventElapsedTime(&out->phase_ms[1], ctx->ev[1], ctx->ev[2]));   // all passes over S
		}
		CK(cudaEventElapsedTime(&out->phase_ms[4], ctx->ev[2], ctx->ev[3]));       // join
	}
	out->kernel_launches = launches;
	out->partitions = st.P;
	set_rows(ctx, out, o->materialize);
	ctx->launches += launches;
	return HJB_OK;
}

extern "C" int hjb_phj_device(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	return phj_device(ctx, R, S, opts ? opts : &kDefaultOpts, 0, out);
}

// ------------------------------------------------------------------ host entry points

typedef int (*device_join_fn)(hjb_ctx *, const hjb_rel *, const hjb_rel *, const hjb_opts *, hjb_result *);

static int phj_device0(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *o, hjb_result *out)
{
	return phj_device(ctx, R, S, o, 0, out);
}

static int host_join_pipelined(hjb_ctx *ctx, bool npj, const hjb_rel *R, const hjb_rel *S, const hjb_opts *o,
                               uint32_t *drk, uint32_t *drv, uint32_t *dsk, uint32_t *dsv, hjb_result *out);
static int grow_host_rows(hjb_ctx *ctx, uint64_t rows);

static int host_join(hjb_ctx *ctx, device_join_fn fn, bool npj, const hjb_rel *R, const hjb_rel *S, const hjb_opts *o,
                     hjb_result *out)
{
	int rc;
	if ((rc = check_rel(ctx, R, false)) || (rc = check_rel(ctx, S, false))) return rc;
	CK(cudaSetDevice(ctx->device));
	const double t0 = wall_now();
	const size_t rb = pad256(R->tuples * 4), sb = pad256(S->tuples * 4);
	if ((rc = grow_device(ctx, &ctx->in_buf, &ctx->in_bytes, 2 * rb + 2 * sb + 256))) return rc;
	cudaStream_t s = ctx->stream;
	uint32_t *drk = (uint32_t *)ctx->in_buf, *drv = (uint32_t *)(ctx->in_buf + rb);
	uint32_t *dsk = (uint32_t *)(ctx->in_buf + 2 * rb), *dsv = (uint32_t *)(ctx->in_buf + 2 * rb + sb);
	static const int pipeline = getenv("HJB_HOST_PIPELINE") ? atoi(getenv("HJB_HOST_PIPELINE")) : 1;   // 0: one copy in, the join, one copy out
	bool on_device = false;
	if (pipeline && R->tuples && S->tuples >= 2 * host_slice_min()) {
		CK(cudaStreamSynchronize(s));
		rc = host_join_pipelined(ctx, npj, R, S, o, drk, drv, dsk, dsv, out);
		if (rc <= 0) {
			if (rc == HJB_OK) out->seconds_e2e = wall_now() - t0;
			return rc;
		}
		on_device = true;                   // more rows than the capacity guess: finish on the plain path
	}
	CK(cudaEventRecord(ctx->ev[8], s));
	if (R->tuples && !on_device) {
		CK(cudaMemcpyAsync(drk, R->keys, R->tuples * 4, cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(drv, R->vals, R->tuples * 4, cudaMemcpyHostToDevice, s));
	}
	if (S->tuples && !on_device) {
		CK(cudaMemcpyAsync(dsk, S->keys, S->tuples * 4, cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(dsv, S->vals, S->tuples * 4, cudaMemcpyHostToDevice, s));
	}
	CK(cudaEventRecord(ctx->ev[9], s));
	hjb_rel dR = {drk, drv, R->tuples}, dS = {dsk, dsv, S->tuples};
	if ((rc = fn(ctx, &dR, &dS, o, out))) return rc;
	float h2d = 0, d2h = 0;
	if (o->materialize && out->count) {
		if ((rc = grow_host_rows(ctx, out->count))) return rc;
		CK(cudaEventRecord(ctx->ev[10], s));
		const uint32_t *cols[3] = {out->keys, out->outer_vals, out->inner_vals};
		for (int c = 0; c < 3; ++c)
			CK(cudaMemcpyAsync(ctx->h_rows + (size_t)c * ctx->h_rows_cap, cols[c], out->count * 4,
			                   cudaMemcpyDeviceToHost, s));
		CK(cudaEventRecord(ctx->ev[11], s));
		CK(cudaStreamSynchronize(s));
		CK(cudaEventElapsedTime(&d2h, ctx->ev[10], ctx->ev[11]));
		out->keys = ctx->h_rows;
		out->outer_vals = ctx->h_rows + ctx->h_rows_cap;
		out->inner_vals = ctx->h_rows + 2 * ctx->h_rows_cap;
	} else {
		out->keys = out->outer_vals = out->inner_vals = nullptr;
		CK(cudaStreamSynchronize(s));          // empty inputs return before any synchronisation
	}
	out->rows_on_device = 0;
	CK(cudaEventElapsedTime(&h2d, ctx->ev[8], ctx->ev[9]));
	out->phase_ms[5] = h2d;
	out->phase_ms[6] = d2h;
	out->seconds_e2e = wall_now() - t0;
	return HJB_OK;
}

// Host columns, probe side in slices: while slice k+1 crosses PCIe towards the GPU, slice k is
// partitioned and joined against the (already resident) build side, and the rows of slice k-1 cross
// PCIe the other way -- three streams, the link busy in both directions.  The reference has no
// counterpart (its relations are already in the memory its threads read, npj.cpp:1013-1039); this is
// the loader a host application needs in front of a GPU join.  Row order is slice order, inside a
// slice unspecified, as for the device entry points.  More rows than the output capacity (a build
// side with many equal keys): the inputs are on the device by then, the plain path finishes the job.
static int pipe_setup(hjb_ctx *ctx)
{
	if (ctx->pipe_ready) return HJB_OK;
	CK(cudaStreamCreateWithFlags(&ctx->pipe_in, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&ctx->pipe_out, cudaStreamNonBlocking));
	for (int i = 0; i < 2; ++i)
		for (int k = 0; k <= kMaxHostSlices; ++k) CK(cudaEventCreateWithFlags(&ctx->pipe_ev[i][k], cudaEventDisableTiming));
	CK(cudaHostAlloc(&ctx->h_cursor, (kMaxHostSlices + 1) * 8, cudaHostAllocDefault));
	ctx->pipe_ready = true;
	return HJB_OK;
}

static int grow_host_rows(hjb_ctx *ctx, uint64_t rows)
{
	if (rows <= ctx->h_rows_cap) return HJB_OK;
	if (ctx->h_rows) CK(cudaFreeHost(ctx->h_rows));
	ctx->h_rows = nullptr;
	ctx->h_rows_cap = 0;
	rows = (rows + 4095) & ~4095ull;
	CK(cudaHostAlloc(&ctx->h_rows, (size_t)rows * 12, cudaHostAllocDefault));
	ctx->h_rows_cap = rows;
	return HJB_OK;
}

// returns 1 when the result did not fit the output capacity (caller falls back), 0 on success, < 0 on error
static int host_join_pipelined(hjb_ctx *ctx, bool npj, const hjb_rel *R, const hjb_rel *S, const hjb_opts *o,
                               uint32_t *drk, uint32_t *drv, uint32_t *dsk, uint32_t *dsv, hjb_result *out)
{
	int rc;
	if ((rc = pipe_setup(ctx))) return rc;
	int K = (int)(S->tuples / host_slice_min());
	if (K > kMaxHostSlices) K = kMaxHostSlices;
	if (K < 1) K = 1;
	const uint64_t slice = ((S->tuples + K - 1) / K + 1023) & ~1023ull;   // whole 4 KB column pieces, 16-byte aligned
	zero_result(out);
	timer_reset(ctx);
	PhjState st;
	NpjArgs a;
	if (npj) {
		const double load = o->npj_load > 0.0 ? o->npj_load : npj_default_load(R->tuples);
		if (load > 0.95) return fail(ctx, HJB_E_INVALID, "npj_load must be <= 0.95");
		const uint64_t buckets = (uint64_t)ceil((double)R->tuples / load / 4.0) + 1;
		if (buckets > 0xFFFFFFFFull) return fail(ctx, HJB_E_INVALID, "table too large");
		if ((rc = grow_device(ctx, &ctx->ws, &ctx->ws_bytes, buckets * 32))) return rc;
		if (o->materialize) {
			const uint64_t cap = o->out_capacity ? o->out_capacity : (S->tuples > R->tuples ? S->tuples : R->tuples);
			if ((rc = grow_out(ctx, cap))) return rc;
		}
		memset(&a, 0, sizeof a);
		a.rk = drk; a.rv = drv; a.nr = R->tuples;
		a.table = (uint64_t *)ctx->ws;
		a.buckets = buckets;
			a.factor = hjb_hash_factor(o->seed, 1);
		a.scalars = ctx->d_scalars;
		a.materialize = o->materialize;
		a.out_k = ctx->out_cols;
		a.out_o = ctx->out_cols + ctx->out_cap;
		a.out_i = ctx->out_cols + 2 * ctx->out_cap;
		a.out_cap = o->materialize ? ctx->out_cap : 0;
	} else {
		if ((rc = phj_setup(ctx, R->tuples, slice < S->tuples ? slice : S->tuples, S->tuples, o, 0, 0, &st))) return rc;
	}
	if (o->materialize && (rc = grow_host_rows(ctx, ctx->out_cap))) return rc;
	cudaStream_t s = ctx->stream, sin = ctx->pipe_in, sout = ctx->pipe_out;
	// copy-in stream: R, then the slices of S
	CK(cudaEventRecord(ctx->ev[8], sin));
	CK(cudaMemcpyAsync(drk, R->keys, R->tuples * 4, cudaMemcpyHostToDevice, sin));
	CK(cudaMemcpyAsync(drv, R->vals, R->tuples * 4, cudaMemcpyHostToDevice, sin));
	CK(cudaEventRecord(ctx->pipe_ev[0][kMaxHostSlices], sin));
	for (int k = 0; k < K; ++k) {
		const uint64_t beg = (uint64_t)k * slice, end = beg + slice < S->tuples ? beg + slice : S->tuples;
		if (beg >= end) { K = k; break; }
		CK(cudaMemcpyAsync(dsk + beg, S->keys + beg, (end - beg) * 4, cudaMemcpyHostToDevice, sin));
		CK(cudaMemcpyAsync(dsv + beg, S->vals + beg, (end - beg) * 4, cudaMemcpyHostToDevice, sin));
		CK(cudaEventRecord(ctx->pipe_ev[0][k], sin));
	}
	CK(cudaEventRecord(ctx->ev[9], sin));
	// compute stream: build side once, then slice after slice
	uint32_t launches = 0;
	CK(cudaStreamWaitEvent(s, ctx->pipe_ev[0][kMaxHostSlices], 0));
	CK(cudaEventRecord(ctx->ev[0], s));
	CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));
	hjb_rel dR = {drk, drv, R->tuples};
	if (npj) launches += launch_npj_build(a, s, ctx->sms, &ctx->timer);
	else if ((rc = phj_partition_side(ctx, &st, &dR, true, &launches))) return rc;
	CK(cudaEventRecord(ctx->ev[1], s));
	for (int k = 0; k < K; ++k) {
		const uint64_t beg = (uint64_t)k * slice, end = beg + slice < S->tuples ? beg + slice : S->tuples;
		CK(cudaStreamWaitEvent(s, ctx->pipe_ev[0][k], 0));
		if (npj) {
			a.sk = dsk + beg; a.sv = dsv + beg; a.ns = end - beg;
			launches += launch_npj_probe(a, s, ctx->sms, &ctx->timer);
		} else {
			hjb_rel dS = {dsk + beg, dsv + beg, end - beg};
			if ((rc = phj_partition_side(ctx, &st, &dS, false, &launches))) return rc;
			if ((rc = phj_launch_join(ctx, &st, o, &launches))) return rc;
		}
		CK(cudaMemcpyAsync(&ctx->h_cursor[k], ctx->d_scalars, 8, cudaMemcpyDeviceToHost, s));
		CK(cudaEventRecord(ctx->pipe_ev[1][k], s));
	}
	CK(cudaEventRecord(ctx->ev[3], s));
	// host: as each slice finishes, send its rows home on the copy-out stream
	uint64_t prev = 0;
	bool overflow = false, first_out = true;
	for (int k = 0; k < K; ++k) {
		CK(cudaEventSynchronize(ctx->pipe_ev[1][k]));
		const uint64_t cur = ctx->h_cursor[k];
		if (!o->materialize) continue;
		if (cur > ctx->out_cap) { overflow = true; break; }
		if (cur > prev) {
			if (first_out) { CK(cudaEventRecord(ctx->ev[10], sout)); first_out = false; }
			for (int c = 0; c < 3; ++c)
				CK(cudaMemcpyAsync(ctx->h_rows + (size_t)c * ctx->h_rows_cap + prev, ctx->out_cols + (size_t)c * ctx->out_cap + prev,
				                   (cur - prev) * 4, cudaMemcpyDeviceToHost, sout));
			prev = cur;
		}
	}
	if (!first_out) CK(cudaEventRecord(ctx->ev[11], sout));
	CK(cudaStreamSynchronize(sout));
	CK(cudaStreamSynchronize(sin));
	CK(cudaGetLastError());
	if ((rc = read_scalars(ctx, out))) return rc;          // synchronizes the compute stream
	ctx->launches += launches;
	if (overflow) return 1;
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
	out->seconds = ms * 1e-3;                              // compute stream, waits for the copies included
	CK(cudaEventElapsedTime(&out->phase_ms[npj ? 1 : 0], ctx->ev[0], ctx->ev[1]));
	CK(cudaEventElapsedTime(&out->phase_ms[npj ? 2 : 4], ctx->ev[1], ctx->ev[3]));
	CK(cudaEventElapsedTime(&out->phase_ms[5], ctx->ev[8], ctx->ev[9]));
	if (!first_out) CK(cudaEventElapsedTime(&out->phase_ms[6], ctx->ev[10], ctx->ev[11]));
	out->kernel_launches = launches;
	out->partitions = npj ? (uint32_t)a.buckets : st.P;
	if (o->materialize && out->count) {
		out->keys = ctx->h_rows;
		out->outer_vals = ctx->h_rows + ctx->h_rows_cap;
		out->inner_vals = ctx->h_rows + 2 * ctx->h_rows_cap;
	}
	out->rows_on_device = 0;
	return HJB_OK;
}

extern "C" int hjb_npj_host(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	return host_join(ctx, npj_device, true, R, S, opts ? opts : &kDefaultOpts, out);
}

extern "C" int hjb_phj_host(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, hjb_result *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	return host_join(ctx, phj_device0, false, R, S, opts ? opts : &kDefaultOpts, out);
}

// ------------------------------------------------------------------ CPRA

static int log2_exact(int x)
{
	int b = 0;
	while ((1 << b) < x) ++b;
	return (1 << b) == x ? b : -1;
}

extern "C" int hjb_cpra_split(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, int ngpus, const hjb_opts *opts,
                              hjb_split *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	int rc;
	if ((rc = check_rel(ctx, R, true)) || (rc = check_rel(ctx, S, true))) return rc;
	const int gbits = log2_exact(ngpus);
	if (gbits < 0 || ngpus > 64) return fail(ctx, HJB_E_INVALID, "ngpus must be a power of two <= 64");
	memset(out, 0, sizeof *out);
	CK(cudaSetDevice(ctx->device));
	timer_reset(ctx);
	if (ngpus == 1) {          // one owner: nothing to split (CPRA on one thread partitions only locally)
		out->r_keys = R->keys; out->r_vals = R->vals; out->s_keys = S->keys; out->s_vals = S->vals;
		out->r_offsets[1] = R->tuples; out->s_offsets[1] = S->tuples;
		return HJB_OK;
	}
	const size_t rb = pad256(R->tuples * 4), sb = pad256(S->tuples * 4), ob = pad256((size_t)(ngpus + 1) * 4);
	uint32_t chunk, mi, tiles;
	size_t rs = radix_scratch_bytes(R->tuples, 1, gbits, &chunk, &mi, &tiles);
	const size_t rs2 = radix_scratch_bytes(S->tuples, 1, gbits, &chunk, &mi, &tiles);
	if (rs2 > rs) rs = rs2;
	if ((rc = grow_device(ctx, &ctx->split_buf, &ctx->split_bytes, 2 * rb + 2 * sb + 2 * ob + 256))) return rc;
	if ((rc = grow_device(ctx, &ctx->ws, &ctx->ws_bytes, rs + 4096))) return rc;
	Bump b = {ctx->split_buf, 0};
	uint32_t *ok[2] = {b.take<uint32_t>(R->tuples), b.take<uint32_t>(S->tuples)};
	uint32_t *ov[2] = {b.take<uint32_t>(R->tuples), b.take<uint32_t>(S->tuples)};
	uint32_t *off[2] = {b.take<uint32_t>(ngpus + 1), b.take<uint32_t>(ngpus + 1)};
	cudaStream_t s = ctx->stream;
	const hjb_rel *rel[2] = {R, S};
	uint32_t launches = 0;
	CK(cudaEventRecord(ctx->ev[4], s));
	for (int r = 0; r < 2; ++r) {
		if (rel[r]->tuples == 0) {
			CK(cudaMemsetAsync(off[r], 0, (size_t)(ngpus + 1) * 4, s));
			continue;
		}
		RadixPassArgs a = {};
		a.keys = rel[r]->keys; a.vals = rel[r]->vals;
		a.keys_out = ok[r]; a.vals_out = ov[r];
		a.n = rel[r]->tuples;
		a.np = 1;
		a.parent_off = nullptr;
		a.child_off = off[r];
		a.factor = hjb_hash_factor(o->seed, 0);
		a.bits = gbits;
		a.rshift = 32 - gbits;
		radix_carve(a, ctx->ws, true);
		launches += launch_radix_pass(a, s, ctx->sms, &ctx->timer);
	}
	CK(cudaEventRecord(ctx->ev[5], s));
	CK(cudaMemcpyAsync(&ctx->h_small[0], off[0], (size_t)(ngpus + 1) * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(&ctx->h_small[128], off[1], (size_t)(ngpus + 1) * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	timer_collect(ctx);
	for (int g = 0; g <= ngpus; ++g) {
		out->r_offsets[g] = ctx->h_small[g];
		out->s_offsets[g] = ctx->h_small[128 + g];
	}
	out->r_keys = ok[0]; out->r_vals = ov[0]; out->s_keys = ok[1]; out->s_vals = ov[1];
	CK(cudaEventElapsedTime(&out->ms, ctx->ev[4], ctx->ev[5]));
	ctx->launches += launches;
	return HJB_OK;
}

// ---- fused exchange ----------------------------------------------------------------

extern "C" int hjb_cpra_recv_alloc(hjb_ctx *ctx, uint64_t r_capacity, uint64_t s_capacity, hjb_recv *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaStreamSynchronize(ctx->stream));
	const uint64_t cap[2] = {r_capacity ? r_capacity : 1, s_capacity ? s_capacity : 1};
	// behind every receive column, in the same allocation: room for the staged exchange's outgoing runs of a chunk of
	// about the same size (stage A then writes this GPU's own runs straight to their final rows with 32-bit positions)
	for (int r = 0; r < 2; ++r) {
		ctx->recv_stage_off[r] = (cap[r] + 63) & ~63ull;
		ctx->recv_stage_cap[r] = cap[r] + 64 * 512 + 64;
	}
	for (int i = 0; i < 4; ++i) {
		if (ctx->recv_buf[i]) CK(cudaFree(ctx->recv_buf[i]));
		ctx->recv_buf[i] = nullptr;
		CK(cudaMalloc(&ctx->recv_buf[i], (ctx->recv_stage_off[i / 2] + ctx->recv_stage_cap[i / 2]) * 4 + 256));
		cudaIpcMemHandle_t h;
		CK(cudaIpcGetMemHandle(&h, ctx->recv_buf[i]));
		static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
		memcpy(out->ipc[i], &h, 64);
	}
	ctx->recv_cap[0] = cap[0];
	ctx->recv_cap[1] = cap[1];
	out->r_keys = (uint32_t *)ctx->recv_buf[0];
	out->r_vals = (uint32_t *)ctx->recv_buf[1];
	out->s_keys = (uint32_t *)ctx->recv_buf[2];
	out->s_vals = (uint32_t *)ctx->recv_buf[3];
	out->r_capacity = cap[0];
	out->s_capacity = cap[1];
	return HJB_OK;
}

extern "C" int hjb_ipc_open(hjb_ctx *ctx, const unsigned char *handle64, void **dev_ptr)
{
	if (!ctx || !handle64 || !dev_ptr) return HJB_E_INVALID;
	CK(cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle64, 64);
	CK(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
	return HJB_OK;
}

extern "C" int hjb_ipc_close(hjb_ctx *ctx, void *dev_ptr)
{
	if (!ctx || !dev_ptr) return HJB_E_INVALID;
	CK(cudaSetDevice(ctx->device));
	CK(cudaIpcCloseMemHandle(dev_ptr));
	return HJB_OK;
}

// hjb_cpra_count's enqueue half: histogram + scan by owner of both chunks; the counted passes stay in ctx->pending
static int cpra_count_enqueue(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, int ngpus, const hjb_opts *o,
                              uint32_t *off_dev[2], uint32_t *launches)
{
	int rc;
	const int gbits = log2_exact(ngpus);
	const hjb_rel *rel[2] = {R, S};
	size_t scratch[2], total = 0;
	for (int r = 0; r < 2; ++r) {
		uint32_t chunk, mi, tiles;
		scratch[r] = radix_scratch_bytes(rel[r]->tuples, 1, gbits, &chunk, &mi, &tiles);
		total += scratch[r] + pad256((size_t)(ngpus + 1) * 4);
	}
	if ((rc = grow_device(ctx, &ctx->split_buf, &ctx->split_bytes, total + 4096))) return rc;
	cudaStream_t s = ctx->stream;
	Bump w = {ctx->split_buf, 0};
	for (int r = 0; r < 2; ++r) {
		RadixPassArgs &a = ctx->pending[r];
		memset(&a, 0, sizeof a);
		a.keys = rel[r]->keys; a.vals = rel[r]->vals;
		a.n = rel[r]->tuples;
		a.np = 1;
		a.factor = hjb_hash_factor(o->seed, 0);
		a.bits = gbits;
		a.rshift = 32 - gbits;
		off_dev[r] = a.child_off = w.take<uint32_t>(ngpus + 1);
		radix_carve(a, w.take<char>(scratch[r]), false);         // the peer scatter ranks its tiles itself: no tile counts
		if (a.n == 0) {
			CK(cudaMemsetAsync(a.child_off, 0, (size_t)(ngpus + 1) * 4, s));
			continue;
		}
		*launches += launch_radix_count(a, s, &ctx->timer);
	}
	return HJB_OK;
}

extern "C" int hjb_cpra_count(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, int ngpus, const hjb_opts *opts,
                              uint64_t *r_counts, uint64_t *s_counts)
{
	if (!ctx || !r_counts || !s_counts) return HJB_E_INVALID;
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	int rc;
	if ((rc = check_rel(ctx, R, true)) || (rc = check_rel(ctx, S, true))) return rc;
	const int gbits = log2_exact(ngpus);
	if (gbits < 1 || ngpus > 64) return fail(ctx, HJB_E_INVALID, "ngpus must be a power of two in [2, 64]");
	CK(cudaSetDevice(ctx->device));
	timer_reset(ctx);
	cudaStream_t s = ctx->stream;
	uint32_t launches = 0;
	uint32_t *off_dev[2];
	if ((rc = cpra_count_enqueue(ctx, R, S, ngpus, o, off_dev, &launches))) return rc;
	CK(cudaMemcpyAsync(&ctx->h_small[0], off_dev[0], (size_t)(ngpus + 1) * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaMemcpyAsync(&ctx->h_small[128], off_dev[1], (size_t)(ngpus + 1) * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	timer_collect(ctx);
	for (int g = 0; g < ngpus; ++g) {
		r_counts[g] = ctx->h_small[g + 1] - ctx->h_small[g];
		s_counts[g] = ctx->h_small[128 + g + 1] - ctx->h_small[128 + g];
	}
	ctx->pending_gpus = ngpus;
	ctx->launches += launches;
	return HJB_OK;
}

// device words of the fused exchange (ctx->cpra_dev, uint32): the scatter kernel and the local join read them there
enum { CD_BASE_R = 0, CD_BASE_S = 64, CD_RANGE_R = 128, CD_RANGE_S = 130, CD_ABORT = 132, CD_MAX_R = 133, CD_MAX_S = 134,
       CD_WORDS = 136 };

static int cpra_dev_alloc(hjb_ctx *ctx)
{
	if (!ctx->cpra_dev) CK(cudaMalloc(&ctx->cpra_dev, CD_WORDS * 4));
	return HJB_OK;
}

// the fused scatter of one counted relation into the owners' columns (asynchronous)
static int cpra_scatter_one(hjb_ctx *ctx, int r, int ngpus, void *const *pk, void *const *pv, bool check_abort,
                            uint32_t *launches)
{
	RadixPassArgs &a = ctx->pending[r];
	if (a.n == 0) return HJB_OK;
	PeerTable t;
	memset(&t, 0, sizeof t);
	for (int g = 0; g < ngpus; ++g) {
		t.k[g] = (uint32_t *)pk[g];
		t.v[g] = (uint32_t *)pv[g];
	}
	t.base = ctx->cpra_dev + (r ? CD_BASE_S : CD_BASE_R);
	t.sender_off = a.child_off;
	t.abort_flag = check_abort ? ctx->cpra_dev + CD_ABORT : nullptr;
	static const int env_ctas = getenv("HJB_PEER_CTAS") ? atoi(getenv("HJB_PEER_CTAS")) : 0;   // experiment knob: limits the peer scatter's grid
	a.peer_ctas = (uint32_t)env_ctas;
	const int n = launch_radix_scatter(a, ctx->stream, &ctx->timer, &t);
	if (n < 0) return fail(ctx, HJB_E_INVALID, "peer scatter: more than 64 owners");
	*launches += n;
	return HJB_OK;
}

extern "C" int hjb_cpra_scatter_peer(hjb_ctx *ctx, int ngpus, void *const *peer_r_keys, void *const *peer_r_vals,
                                     void *const *peer_s_keys, void *const *peer_s_vals, const uint64_t *r_base,
                                     const uint64_t *s_base, float *ms)
{
	if (!ctx || !peer_r_keys || !peer_r_vals || !peer_s_keys || !peer_s_vals || !r_base || !s_base) return HJB_E_INVALID;
	if (ngpus != ctx->pending_gpus || ngpus < 2) return fail(ctx, HJB_E_INVALID, "hjb_cpra_count must precede with the same ngpus");
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	uint32_t launches = 0;
	int rc;
	if ((rc = cpra_dev_alloc(ctx))) return rc;
	for (int g = 0; g < ngpus; ++g) {
		if (r_base[g] > 0xFFFFFFFFull || s_base[g] > 0xFFFFFFFFull) return fail(ctx, HJB_E_INVALID, "base row beyond 2^32-1");
		ctx->h_small[CD_BASE_R + g] = (uint32_t)r_base[g];
		ctx->h_small[CD_BASE_S + g] = (uint32_t)s_base[g];
	}
	CK(cudaMemcpyAsync(ctx->cpra_dev, ctx->h_small, 128 * 4, cudaMemcpyHostToDevice, s));
	CK(cudaEventRecord(ctx->ev[6], s));
	if ((rc = cpra_scatter_one(ctx, 0, ngpus, peer_r_keys, peer_r_vals, false, &launches))) return rc;
	if ((rc = cpra_scatter_one(ctx, 1, ngpus, peer_s_keys, peer_s_vals, false, &launches))) return rc;
	CK(cudaEventRecord(ctx->ev[7], s));
	CK(cudaStreamSynchronize(s));            // the owners may read once every sender has passed this point
	CK(cudaGetLastError());
	timer_collect(ctx);
	if (ms) CK(cudaEventElapsedTime(ms, ctx->ev[6], ctx->ev[7]));
	ctx->pending_gpus = 0;
	ctx->launches += launches;
	return HJB_OK;
}

extern "C" int hjb_cpra_join_local(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, int gpu, int ngpus,
                                   const hjb_opts *opts, hjb_result *out)
{
	if (!ctx || !out) return HJB_E_INVALID;
	const int gbits = log2_exact(ngpus);
	if (gbits < 0 || ngpus > 64 || gpu < 0 || gpu >= ngpus) return fail(ctx, HJB_E_INVALID, "bad gpu / ngpus");
	return phj_device(ctx, R, S, opts ? opts : &kDefaultOpts, gbits, out, (uint32_t)gpu);
}

// ---- the same exchange, stream-ordered: nothing between the count and the result touches the host ----------

// this sender's tuples per owner, R then S, as the all-gather's input
__global__ void k_cpra_counts(const uint32_t *__restrict__ r_off, const uint32_t *__restrict__ s_off, int G,
                              unsigned long long *__restrict__ counts)
{
	const int g = threadIdx.x;
	if (g < G) {
		counts[g] = r_off[g + 1] - r_off[g];
		counts[G + g] = s_off[g + 1] - s_off[g];
	}
}

// From the all-gathered count matrix M[src][2 G] (R counts, then S counts): the first row of every owner's
// columns reserved for this sender (the senders before it come first, as thread t's pieces precede thread
// t+1's in the reference's gather, cpra2.cpp:1896-1904), what this GPU receives in total, and whether
// every owner's buffer is large enough -- every sender reaches the same verdict from the same matrix.
__global__ void k_cpra_bases(const unsigned long long *__restrict__ M, int G, int me, unsigned long long cap_r,
                             unsigned long long cap_s, uint32_t *__restrict__ out)
{
	__shared__ unsigned long long s_max[2];
	const int g = threadIdx.x;
	if (g < 2) s_max[g] = 0;
	__syncthreads();
	unsigned long long base_r = 0, base_s = 0, tot_r = 0, tot_s = 0;
	if (g < G) {
		for (int src = 0; src < G; ++src) {
			const unsigned long long cr = M[(size_t)src * 2 * G + g], cs = M[(size_t)src * 2 * G + G + g];
			if (src < me) {
				base_r += cr;
				base_s += cs;
			}
			tot_r += cr;
			tot_s += cs;
		}
		atomicMax(&s_max[0], tot_r);
		atomicMax(&s_max[1], tot_s);
	}
	const int abort = __syncthreads_or(g < G && (tot_r > cap_r || tot_s > cap_s));
	if (g < G) {
		out[CD_BASE_R + g] = (uint32_t)base_r;
		out[CD_BASE_S + g] = (uint32_t)base_s;
		if (g == me) {
			out[CD_RANGE_R] = 0;
			out[CD_RANGE_R + 1] = abort ? 0u : (uint32_t)tot_r;
			out[CD_RANGE_S] = 0;
			out[CD_RANGE_S + 1] = abort ? 0u : (uint32_t)tot_s;
		}
	}
	if (g == 0) {
		out[CD_ABORT] = abort ? 1u : 0u;
		out[CD_MAX_R] = (uint32_t)(s_max[0] > 0xFFFFFFFFull ? 0xFFFFFFFFull : s_max[0]);
		out[CD_MAX_S] = (uint32_t)(s_max[1] > 0xFFFFFFFFull ? 0xFFFFFFFFull : s_max[1]);
	}
}

extern "C" int hjb_cpra_bind(hjb_ctx *ctx, int gpu, int ngpus, void *const *peer_r_keys, void *const *peer_r_vals,
                             void *const *peer_s_keys, void *const *peer_s_vals, uint64_t r_capacity, uint64_t s_capacity)
{
	if (!ctx || !peer_r_keys || !peer_r_vals || !peer_s_keys || !peer_s_vals) return HJB_E_INVALID;
	const int gbits = log2_exact(ngpus);
	if (gbits < 1 || ngpus > 64 || gpu < 0 || gpu >= ngpus) return fail(ctx, HJB_E_INVALID, "ngpus must be a power of two in [2, 64], gpu in [0, ngpus)");
	if (r_capacity > 0xFFFFFFFFull || s_capacity > 0xFFFFFFFFull) return fail(ctx, HJB_E_INVALID, "capacity beyond 2^32-1 rows");
	CK(cudaSetDevice(ctx->device));
	int rc;
	if ((rc = cpra_dev_alloc(ctx))) return rc;
	void *const *cols[4] = {peer_r_keys, peer_r_vals, peer_s_keys, peer_s_vals};
	for (int c = 0; c < 4; ++c)
		for (int g = 0; g < ngpus; ++g) {
			if (!cols[c][g] || ((uintptr_t)cols[c][g] & 127)) return fail(ctx, HJB_E_INVALID, "peer columns must be 128-byte aligned device memory");
			ctx->bind_peer[c][g] = cols[c][g];
		}
	ctx->bind_gpu = gpu;
	ctx->bind_gpus = ngpus;
	ctx->bind_cap[0] = r_capacity;
	ctx->bind_cap[1] = s_capacity;
	ctx->step_state = 0;
	return HJB_OK;
}

extern "C" int hjb_cpra_count_async(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, uint64_t *counts_dev)
{
	if (!ctx || !counts_dev) return HJB_E_INVALID;
	if (!ctx->bind_gpus) return fail(ctx, HJB_E_INVALID, "hjb_cpra_bind must precede");
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	int rc;
	if ((rc = check_rel(ctx, R, true)) || (rc = check_rel(ctx, S, true))) return rc;
	CK(cudaSetDevice(ctx->device));
	timer_reset(ctx);
	uint32_t *off_dev[2];
	ctx->step_launches = 0;
	ctx->hot_pending = false;
	if ((rc = cpra_count_enqueue(ctx, R, S, ctx->bind_gpus, o, off_dev, &ctx->step_launches))) return rc;
	k_cpra_counts<<<1, 64, 0, ctx->stream>>>(off_dev[0], off_dev[1], ctx->bind_gpus, (unsigned long long *)counts_dev);
	ctx->step_launches += 1;
	CK(cudaGetLastError());
	ctx->pending_gpus = ctx->bind_gpus;
	ctx->step_state = 1;
	return HJB_OK;
}

extern "C" int hjb_cpra_scatter_async(hjb_ctx *ctx, const uint64_t *matrix_dev)
{
	if (!ctx || !matrix_dev) return HJB_E_INVALID;
	if (ctx->step_state != 1) return fail(ctx, HJB_E_INVALID, "hjb_cpra_count_async must precede");
	CK(cudaSetDevice(ctx->device));
	const int G = ctx->bind_gpus;
	int rc;
	k_cpra_bases<<<1, 64, 0, ctx->stream>>>((const unsigned long long *)matrix_dev, G, ctx->bind_gpu, ctx->bind_cap[0],
	                                        ctx->bind_cap[1], ctx->cpra_dev);
	ctx->step_launches += 1;
	if ((rc = cpra_scatter_one(ctx, 0, G, ctx->bind_peer[0], ctx->bind_peer[1], true, &ctx->step_launches))) return rc;
	if ((rc = cpra_scatter_one(ctx, 1, G, ctx->bind_peer[2], ctx->bind_peer[3], true, &ctx->step_launches))) return rc;
	CK(cudaGetLastError());
	ctx->pending_gpus = 0;
	ctx->step_state = 2;
	return HJB_OK;
}

extern "C" int hjb_cpra_join_async(hjb_ctx *ctx, const hjb_opts *opts, uint64_t r_expect, uint64_t s_expect)
{
	if (!ctx) return HJB_E_INVALID;
	if (ctx->step_state != 2) return fail(ctx, HJB_E_INVALID, "hjb_cpra_scatter_async must precede");
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	CK(cudaSetDevice(ctx->device));
	const int G = ctx->bind_gpus, me = ctx->bind_gpu;
	int rc;
	if (!ctx->step_phj && !(ctx->step_phj = (PhjState *)calloc(1, sizeof(PhjState)))) return HJB_E_NOMEM;
	PhjState &st = *ctx->step_phj;
	const uint64_t rc_cap = ctx->bind_cap[0], sc_cap = ctx->bind_cap[1];
	if (r_expect == 0 || r_expect > rc_cap) r_expect = rc_cap;
	if (s_expect == 0 || s_expect > sc_cap) s_expect = sc_cap;
	hjb_opts oo = *o;
	if (!oo.out_capacity) oo.out_capacity = sc_cap > rc_cap ? sc_cap : rc_cap;
	if ((rc = phj_setup(ctx, rc_cap, sc_cap, sc_cap, &oo, log2_exact(G), (uint32_t)me, &st, r_expect, s_expect))) return rc;
	ctx->step_opts = oo;
	cudaStream_t s = ctx->stream;
	CK(cudaEventRecord(ctx->ev[0], s));
	CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));
	const hjb_rel Rr = {(const uint32_t *)ctx->bind_peer[0][me], (const uint32_t *)ctx->bind_peer[1][me], rc_cap};
	const hjb_rel Sr = {(const uint32_t *)ctx->bind_peer[2][me], (const uint32_t *)ctx->bind_peer[3][me], sc_cap};
	if ((rc = phj_partition_side(ctx, &st, &Rr, true, &ctx->step_launches, ctx->cpra_dev + CD_RANGE_R))) return rc;
	if ((rc = phj_partition_side(ctx, &st, &Sr, false, &ctx->step_launches, ctx->cpra_dev + CD_RANGE_S))) return rc;
	CK(cudaEventRecord(ctx->ev[2], s));
	if ((rc = phj_launch_join(ctx, &st, &oo, &ctx->step_launches))) return rc;
	CK(cudaEventRecord(ctx->ev[3], s));
	CK(cudaGetLastError());
	ctx->step_state = 3;
	return HJB_OK;
}

static int hot_join_enqueue(hjb_ctx *ctx, uint32_t *launches);

extern "C" void *hjb_cpra_sums_dev(hjb_ctx *ctx) { return ctx ? (void *)(ctx->d_scalars + 1) : nullptr; }

extern "C" int hjb_cpra_finish(hjb_ctx *ctx, hjb_result *out, uint64_t received[2], uint64_t largest[2])
{
	if (!ctx || !out) return HJB_E_INVALID;
	if (ctx->step_state != 3) return fail(ctx, HJB_E_INVALID, "hjb_cpra_join_async must precede");
	ctx->step_state = 0;
	zero_result(out);
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	PhjState &st = *ctx->step_phj;
	const hjb_opts *o = &ctx->step_opts;
	int rc;
	CK(cudaMemcpyAsync(&ctx->h_small[128], ctx->cpra_dev + 128, 8 * 4, cudaMemcpyDeviceToHost, s));
	uint32_t launches = ctx->step_launches;
	for (int attempt = 0; attempt < 2; ++attempt) {
		if ((rc = read_scalars(ctx, out))) return rc;
		if (received) {
			received[0] = ctx->h_small[CD_RANGE_R + 1];
			received[1] = ctx->h_small[CD_RANGE_S + 1];
		}
		if (largest) {
			largest[0] = ctx->h_small[CD_MAX_R];
			largest[1] = ctx->h_small[CD_MAX_S];
		}
		if (ctx->h_small[CD_ABORT]) return fail(ctx, HJB_E_CAPACITY, "hjb_cpra_finish: an owner's receive buffer is too small; nothing was exchanged");
		if (ctx->h_scalars[7]) return fail(ctx, HJB_E_INVALID, "hjb_cpra_finish: a tuple does not hash into this owner's range");
		if (!o->materialize || out->count <= ctx->out_cap) break;
		if (attempt == 1) return fail(ctx, HJB_E_CUDA, "result overflow after regrow");
		if ((rc = grow_out(ctx, out->count))) return rc;
		CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));         // rerun the join phase only
		if ((rc = phj_launch_join(ctx, &st, o, &launches))) return rc;
		if (ctx->hot_pending && (rc = hot_join_enqueue(ctx, &launches))) return rc;
		CK(cudaEventRecord(ctx->ev[3], s));
	}
	ctx->hot_pending = false;
	float ms = 0;
	CK(cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]));
	out->seconds = ms * 1e-3;                               // the local join only
	CK(cudaEventElapsedTime(&out->phase_ms[0], ctx->ev[0], ctx->ev[2]));     // both sides' local passes
	CK(cudaEventElapsedTime(&out->phase_ms[4], ctx->ev[2], ctx->ev[3]));
	out->kernel_launches = launches;
	out->partitions = st.P;
	set_rows(ctx, out, o->materialize);
	ctx->launches += launches;
	return HJB_OK;
}

// ---- the STAGED exchange (stage.cu): stage A partitions the chunk locally by owner AND sub-partition, TMA copies push
// whole runs into the owners' columns, one local pass and the join follow.  Two radix passes over the data instead
// of the fused path's three (GPU-assign + two local ones), and the copies leave the SMs to the passes beside them.

constexpr int kStageMaxBits = 9;                  // fan-out of a pass of the tile-count scatter

// How the radix bits of one step are split: stage A takes abits (>= the owner bits), the local pass bbits.
// Every rank must call this with the same expected sizes.  Returns HJB_E_INVALID when two passes do not suffice
// (the caller then uses the fused path).
extern "C" int hjb_cpra_stage_plan(hjb_ctx *ctx, int ngpus, uint64_t r_expect, uint64_t s_expect, const hjb_opts *opts,
                                   int *abits, int *bbits, int *big_fill)
{
	if (!ctx || !abits || !bbits || !big_fill) return HJB_E_INVALID;
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	const int gbits = log2_exact(ngpus);
	if (gbits < 1 || ngpus > 64) return fail(ctx, HJB_E_INVALID, "ngpus must be a power of two in [2, 64]");
	Plan p;
	int rc;
	hjb_opts oo = *o;
	memset(oo.radix_bits, 0, sizeof oo.radix_bits);
	if ((rc = make_plan(ctx, r_expect ? r_expect : 1, s_expect, &oo, gbits, &p))) return rc;
	int total = gbits + p.total_bits;
	*big_fill = 0;
	// DIRECT tables (>= 16 bits in all): partitions may average 8192 build tuples when the join takes 12288-tuple fills
	if (total > 2 * kStageMaxBits && total - 1 <= 2 * kStageMaxBits && total - 1 >= 18) {
		total -= 1;
		*big_fill = 1;
	}
	if (total > 2 * kStageMaxBits) return fail(ctx, HJB_E_INVALID, "staged exchange: more radix bits than two passes take");
	int a = (total + 1) / 2;
	if (a < gbits) a = gbits;
	if (a < 2) a = 2;                              // the copy kernel's piece tables assume at least two sub-... owners x subs >= 4
	*abits = a;
	*bbits = total > a ? total - a : 1;            // always one local pass: it gathers a sub-partition's per-sender ranges
	return HJB_OK;
}

static int stage_dev_alloc(hjb_ctx *ctx)
{
	int rc;
	if ((rc = cpra_dev_alloc(ctx))) return rc;
	if (!ctx->stage_dev) CK(cudaMalloc(&ctx->stage_dev, SD_WORDS * 4));
	return HJB_OK;
}

// stage A's histogram + scan of both chunks; counts_dev[rel * 2^abits + digit] = this sender's tuples (uint64)
extern "C" int hjb_cpra_stage_count_async(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts, int abits,
                                          int nparts, uint64_t *counts_dev)
{
	if (!ctx || !counts_dev) return HJB_E_INVALID;
	if (!ctx->bind_gpus) return fail(ctx, HJB_E_INVALID, "hjb_cpra_bind must precede");
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	const int gbits = log2_exact(ctx->bind_gpus);
	if (abits < gbits || abits < 2 || abits > kStageMaxBits) return fail(ctx, HJB_E_INVALID, "stage A takes between max(2, owner bits) and 9 bits");
	if (nparts < 1 || nparts > kStageMaxParts || (nparts & (nparts - 1)) || nparts > (1 << (abits - gbits)))
		return fail(ctx, HJB_E_INVALID, "parts: a power of two, at most 8 and at most the sub-partitions per owner");
	int rc;
	if ((rc = check_rel(ctx, R, true)) || (rc = check_rel(ctx, S, true))) return rc;
	CK(cudaSetDevice(ctx->device));
	if ((rc = stage_dev_alloc(ctx))) return rc;
	timer_reset(ctx);
	ctx->step_launches = 0;
	ctx->hot_pending = false;
	const hjb_rel *rel[2] = {R, S};
	const uint32_t F = 1u << abits;
	size_t scratch[2], total = 0, cols[2];
	const int me = ctx->bind_gpu;
	bool inplace = true;
	for (int r = 0; r < 2; ++r) {
		cols[r] = rel[r]->tuples + 64 * (size_t)F + 64;         // every run may start up to 31 rows late and end up to 31 rows early
		inplace = inplace && ctx->bind_peer[2 * r][me] == ctx->recv_buf[2 * r] && ctx->bind_peer[2 * r + 1][me] == ctx->recv_buf[2 * r + 1] &&
		          cols[r] <= ctx->recv_stage_cap[r] && ctx->recv_stage_off[r] + ctx->recv_stage_cap[r] < 0xFFFFFFFFull;
	}
	static const int env_inplace = getenv("HJB_STAGE_INPLACE") ? atoi(getenv("HJB_STAGE_INPLACE")) : 1;
	inplace = inplace && env_inplace;
	for (int r = 0; r < 2; ++r) {
		uint32_t chunk, mi, tiles;
		scratch[r] = radix_scratch_bytes(rel[r]->tuples, 1, abits, &chunk, &mi, &tiles);
		total += scratch[r] + pad256((size_t)(F + 1) * 4) + (inplace ? 0 : 2 * pad256(cols[r] * 4));
	}
	ctx->stage_inplace = inplace;
	if ((rc = grow_device(ctx, &ctx->split_buf, &ctx->split_bytes, total + 4096))) return rc;
	cudaStream_t s = ctx->stream;
	Bump w = {ctx->split_buf, 0};
	for (int r = 0; r < 2; ++r) {
		RadixPassArgs &a = ctx->pending[r];
		memset(&a, 0, sizeof a);
		a.keys = rel[r]->keys; a.vals = rel[r]->vals;
		a.n = rel[r]->tuples;
		a.np = 1;
		a.factor = hjb_hash_factor(o->seed, 0);
		a.bits = abits;
		a.rshift = 32 - abits;
		a.child_off = w.take<uint32_t>(F + 1);
		if (inplace) {
			ctx->stage_k[r] = a.keys_out = (uint32_t *)ctx->recv_buf[2 * r];
			ctx->stage_v[r] = a.vals_out = (uint32_t *)ctx->recv_buf[2 * r + 1];
			ctx->stage_base[r] = (uint32_t)ctx->recv_stage_off[r];
		} else {
			ctx->stage_k[r] = a.keys_out = w.take<uint32_t>(cols[r]);
			ctx->stage_v[r] = a.vals_out = w.take<uint32_t>(cols[r]);
			ctx->stage_base[r] = 0;
		}
		a.shift = reinterpret_cast<const int32_t *>(ctx->stage_dev + (r ? SD_REL_S : SD_REL_R) + SD_SHIFT);
		radix_carve(a, w.take<char>(scratch[r]), true);
		if (a.n == 0) CK(cudaMemsetAsync(a.child_off, 0, (size_t)(F + 1) * 4, s));
		else ctx->step_launches += launch_radix_count(a, s, &ctx->timer);
		ctx->stage_state[r] = 1;
	}
	ctx->step_launches += launch_stage_counts(ctx->pending[0].child_off, ctx->pending[1].child_off, abits,
	                                          (unsigned long long *)counts_dev, s);
	CK(cudaGetLastError());
	ctx->stage_abits = abits;
	ctx->stage_parts = nparts;
	ctx->stage_copied[0] = ctx->stage_copied[1] = ctx->stage_done[0] = ctx->stage_done[1] = 0;
	ctx->step_state = 10;
	return HJB_OK;
}

// rel 0: the bases from the all-gathered matrix (G x 2 x 2^abits uint64), then stage A's scatter of R; rel 1: of S
extern "C" int hjb_cpra_stage_scatter_async(hjb_ctx *ctx, const uint64_t *matrix_dev, int rel)
{
	if (!ctx || !matrix_dev || rel < 0 || rel > 1) return HJB_E_INVALID;
	if (ctx->step_state != 10 || ctx->stage_state[rel] != 1 || (rel == 1 && ctx->stage_state[0] < 2))
		return fail(ctx, HJB_E_INVALID, "hjb_cpra_stage_count_async must precede; R is scattered before S");
	CK(cudaSetDevice(ctx->device));
	const int G = ctx->bind_gpus;
	cudaStream_t s = ctx->stream;
	if (rel == 0)
		ctx->step_launches += launch_stage_bases((const unsigned long long *)matrix_dev, G, ctx->bind_gpu, ctx->stage_abits, log2_exact(G),
		                                         ctx->bind_cap[0], ctx->bind_cap[1], ctx->pending[0].child_off, ctx->pending[1].child_off,
		                                         ctx->stage_base[0], ctx->stage_base[1], ctx->stage_inplace, ctx->stage_parts, ctx->stage_dev,
		                                         ctx->cpra_dev + CD_RANGE_R, s);
	if (rel == 0) {
		// The runs leave through the copy engines (cudaMemcpyAsync into the owners' columns: the reference's memcpy gather,
		// cpra2.cpp:1896-1904, literally), which needs their rows on the HOST: one synchronisation per step, here, where the
		// stream holds nothing but the counting kernels.  HJB_STAGE_COPY=tma: k_peer_copy reads them on the device (no
		// synchronisation, but the copies then take SMs from the passes beside them: 8 GPUs 14.1 vs 13.6 ms per step).
		const int env_ce = !(getenv("HJB_STAGE_COPY") && !strcmp(getenv("HJB_STAGE_COPY"), "tma"));       // read per step
		ctx->stage_copy_engine = env_ce;
		if (env_ce) {
			if (!ctx->h_stage) CK(cudaMallocHost(&ctx->h_stage, (2 * 1536 + 8) * 4));
			for (int r = 0; r < 2; ++r)           // SD_OWN_SRC, _LEN, _DST: 3 x [8 parts][64 owners]
				CK(cudaMemcpyAsync(ctx->h_stage + 1536 * r, ctx->stage_dev + (r ? SD_REL_S : SD_REL_R) + SD_OWN_SRC, 1536 * 4, cudaMemcpyDeviceToHost, s));
			CK(cudaMemcpyAsync(ctx->h_stage + 3072, ctx->cpra_dev + CD_ABORT, 4, cudaMemcpyDeviceToHost, s));
			CK(cudaStreamSynchronize(s));
		}
	}
	RadixPassArgs &a = ctx->pending[rel];
	if (a.n) ctx->step_launches += launch_radix_scatter(a, s, &ctx->timer, nullptr);
	CK(cudaGetLastError());
	ctx->stage_state[rel] = 2;
	return HJB_OK;
}

// the copies of part `part` of one relation's runs into the owners' columns, on `cuda_stream` (null: the context's stream)
// -- a side stream that waits for the scatter lets them cross NVLink beside the passes
extern "C" int hjb_cpra_stage_copy_async(hjb_ctx *ctx, int rel, int part, void *cuda_stream)
{
	if (!ctx || rel < 0 || rel > 1) return HJB_E_INVALID;
	if (ctx->step_state < 10 || ctx->step_state > 11 || ctx->stage_state[rel] != 2) return fail(ctx, HJB_E_INVALID, "hjb_cpra_stage_scatter_async must precede");
	if (part < 0 || part >= ctx->stage_parts || (ctx->stage_copied[rel] >> part & 1)) return fail(ctx, HJB_E_INVALID, "no such part, or copied already");
	CK(cudaSetDevice(ctx->device));
	const int G = ctx->bind_gpus;
	PeerCols pc;
	memset(&pc, 0, sizeof pc);
	for (int g = 0; g < G; ++g) {
		pc.k[g] = (uint32_t *)ctx->bind_peer[rel ? 2 : 0][g];
		pc.v[g] = (uint32_t *)ctx->bind_peer[rel ? 3 : 1][g];
	}
	if (ctx->stage_copy_engine) {
		cudaStream_t cs = cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream;
		const uint32_t *src = ctx->h_stage + 1536 * rel + 64 * part, *len = src + 512, *dst = src + 1024;
		if (!ctx->h_stage[3072])
			for (int i = 1; i <= G; ++i) {
				const int g = (ctx->bind_gpu + i) % G;           // every sender starts with its right-hand neighbour
				if ((g == ctx->bind_gpu && ctx->stage_inplace) || !len[g]) continue;
				CK(cudaMemcpyAsync(pc.k[g] + dst[g], ctx->stage_k[rel] + src[g], (size_t)len[g] * 4, cudaMemcpyDeviceToDevice, cs));
				CK(cudaMemcpyAsync(pc.v[g] + dst[g], ctx->stage_v[rel] + src[g], (size_t)len[g] * 4, cudaMemcpyDeviceToDevice, cs));
			}
	} else if (ctx->pending[rel].n)
		ctx->step_launches += launch_peer_copy(ctx->stage_k[rel], ctx->stage_v[rel], pc, ctx->stage_dev + (rel ? SD_REL_S : SD_REL_R) + 64 * part,
		                                       ctx->cpra_dev + CD_ABORT, log2_exact(G), ctx->bind_gpu, ctx->stage_inplace,
		                                       cuda_stream ? (cudaStream_t)cuda_stream : ctx->stream, cuda_stream ? nullptr : &ctx->timer);
	CK(cudaGetLastError());
	ctx->stage_copied[rel] |= 1u << part;
	return HJB_OK;
}

// One local pass over the sub-partitions [q0, q1) of what this GPU received of a relation (stage.cu: every sub-partition is
// the union of one range per sender) into the work columns, continuing at the rows where the range begins.
static int stage_local_pass(hjb_ctx *ctx, PhjState *st, int rel, uint32_t q0, uint32_t q1, int bbits)
{
	const int G = ctx->bind_gpus, me = ctx->bind_gpu;
	const uint32_t *desc = ctx->stage_dev + (rel ? SD_REL_S : SD_REL_R);
	RadixPassArgs a = {};
	a.keys = (const uint32_t *)ctx->bind_peer[rel ? 2 : 0][me];
	a.vals = (const uint32_t *)ctx->bind_peer[rel ? 3 : 1][me];
	a.keys_out = rel ? st->sbk[0] : st->rbk[0];
	a.vals_out = rel ? st->sbv[0] : st->rbv[0];
	a.n = ctx->bind_cap[rel];
	a.np = q1 - q0;
	a.parent_off = desc + SD_POFF + q0;
	a.seg = desc + SD_SEG + (size_t)q0 * G * 2;
	a.nseg = (uint32_t)G;
	a.out_base = desc + SD_POFF + q0;
	a.chunk_div = (uint32_t)ctx->stage_parts;
	a.child_off = (rel ? st->soff[0] : st->roff[0]) + ((size_t)q0 << bbits);
	a.factor = st->radix_factor;
	a.bits = bbits;
	a.rshift = 32 - ctx->stage_abits - bbits;
	radix_carve(a, st->scratch, true);
	ctx->step_launches += launch_radix_pass(a, ctx->stream, ctx->sms, &ctx->timer);
	return HJB_OK;
}

// rel 0: the local pass over part `part` of what this GPU received of R (call once every sender's copies of that part have
// landed); rel 1: the same for S, then the join of the part's partitions (R's part must have been passed).  After the last
// part of S hjb_cpra_finish follows.
extern "C" int hjb_cpra_stage_local_async(hjb_ctx *ctx, const hjb_opts *opts, int bbits, int big_fill, int rel, int part)
{
	if (!ctx || rel < 0 || rel > 1) return HJB_E_INVALID;
	if (ctx->step_state < 10 || ctx->step_state > 11) return fail(ctx, HJB_E_INVALID, "hjb_cpra_stage_copy_async must precede");
	if (part < 0 || part >= ctx->stage_parts || !(ctx->stage_copied[rel] >> part & 1) || (ctx->stage_done[rel] >> part & 1))
		return fail(ctx, HJB_E_INVALID, "that part has not been copied, or has been passed already");
	if (rel == 1 && !(ctx->stage_done[0] >> part & 1)) return fail(ctx, HJB_E_INVALID, "R's part is passed before S's");
	if (bbits < 1 || bbits > kStageMaxBits) return fail(ctx, HJB_E_INVALID, "the local pass takes between 1 and 9 bits");
	const hjb_opts *o = opts ? opts : &kDefaultOpts;
	CK(cudaSetDevice(ctx->device));
	const int G = ctx->bind_gpus, me = ctx->bind_gpu, gbits = log2_exact(G), pre = ctx->stage_abits - gbits;
	int rc;
	cudaStream_t s = ctx->stream;
	const uint64_t rc_cap = ctx->bind_cap[0], sc_cap = ctx->bind_cap[1];
	if (ctx->step_state == 10) {                   // the first part: plan, workspace, scalars
		if (rel != 0) return fail(ctx, HJB_E_INVALID, "R's part is passed before S's");
		if (!ctx->step_phj && !(ctx->step_phj = (PhjState *)calloc(1, sizeof(PhjState)))) return HJB_E_NOMEM;
		hjb_opts oo = *o;
		if (!oo.out_capacity) oo.out_capacity = sc_cap > rc_cap ? sc_cap : rc_cap;
		Plan p;
		memset(&p, 0, sizeof p);
		p.npass = 1;
		p.bits[0] = bbits;
		p.total_bits = bbits;
		if (32 - ctx->stage_abits - bbits > (big_fill ? 14 : 32)) return fail(ctx, HJB_E_INVALID, "12288-tuple fills need <= 14 hash bits below the partition id");
		if ((rc = phj_setup(ctx, rc_cap, sc_cap, sc_cap, &oo, gbits, (uint32_t)me, ctx->step_phj, 0, 0, &p, pre, (uint32_t)G,
		                    (uint32_t)ctx->stage_parts))) return rc;
		ctx->step_phj->big_fill = big_fill;
		ctx->step_opts = oo;
		CK(cudaEventRecord(ctx->ev[0], s));
		CK(cudaEventRecord(ctx->ev[2], s));
		CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, s));
		ctx->step_state = 11;
	}
	PhjState &st = *ctx->step_phj;
	if (st.plan.bits[0] != bbits || st.big_fill != big_fill) return fail(ctx, HJB_E_INVALID, "every part of a step takes the same plan");
	const uint32_t nsub = 1u << pre, q0 = stage_part_lo((uint32_t)part, (uint32_t)ctx->stage_parts, nsub),
	               q1 = stage_part_lo((uint32_t)part + 1, (uint32_t)ctx->stage_parts, nsub);
	if ((rc = stage_local_pass(ctx, &st, rel, q0, q1, bbits))) return rc;
	ctx->stage_done[rel] |= 1u << part;
	if (rel == 1) {
		// the part's partitions: both sides are in place
		st.pr.k = st.rbk[0]; st.pr.v = st.rbv[0]; st.pr.off = st.roff[0];
		st.ps.k = st.sbk[0]; st.ps.v = st.sbv[0]; st.ps.off = st.soff[0];
		if ((rc = phj_launch_join(ctx, &st, &ctx->step_opts, &ctx->step_launches, q0 << bbits, q1 << bbits))) return rc;
		CK(cudaEventRecord(ctx->ev[3], s));
	}
	CK(cudaGetLastError());
	const uint32_t all = (1u << ctx->stage_parts) - 1;
	if (ctx->stage_done[0] == all && ctx->stage_done[1] == all) {
		ctx->stage_state[0] = ctx->stage_state[1] = 0;
		ctx->pending_gpus = 0;
		ctx->step_state = 3;
	}
	return HJB_OK;
}

// ---- heavy-hitter handling (skew.cu): hot probe tuples stay with their sender, hot build tuples are replicated

extern "C" int hjb_cpra_split_hot(hjb_ctx *ctx, const hjb_rel *S, const uint32_t *hot_keys_dev, uint32_t n_hot, hjb_rel *cold,
                                  hjb_rel *hot)
{
	if (!ctx || !cold || !hot || (n_hot && !hot_keys_dev)) return HJB_E_INVALID;
	if (n_hot > kMaxHotKeys) return fail(ctx, HJB_E_INVALID, "more than 256 hot keys");
	int rc;
	if ((rc = check_rel(ctx, S, true))) return rc;
	CK(cudaSetDevice(ctx->device));
	const size_t col = pad256(S->tuples * 4);
	if ((rc = grow_device(ctx, &ctx->skew_buf, &ctx->skew_bytes, 4 * col + 256))) return rc;
	uint32_t *ck = (uint32_t *)ctx->skew_buf, *cv = (uint32_t *)(ctx->skew_buf + col);
	uint32_t *hk = (uint32_t *)(ctx->skew_buf + 2 * col), *hv = (uint32_t *)(ctx->skew_buf + 3 * col);
	cudaStream_t s = ctx->stream;
	ctx->launches += launch_split_hot(S->keys, S->vals, S->tuples, hot_keys_dev, n_hot, ck, cv, hk, hv, ctx->d_scalars + 10, s, ctx->sms);
	CK(cudaMemcpyAsync(ctx->h_scalars + 10, ctx->d_scalars + 10, 16, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	cold->keys = ck; cold->vals = cv; cold->tuples = ctx->h_scalars[10];
	hot->keys = hk; hot->vals = hv; hot->tuples = ctx->h_scalars[11];
	if (cold->tuples + hot->tuples != S->tuples) return fail(ctx, HJB_E_CUDA, "hjb_cpra_split_hot: the two parts do not add up");
	return HJB_OK;
}

extern "C" int hjb_cpra_select_hot(hjb_ctx *ctx, const hjb_rel *R, const uint32_t *hot_keys_dev, uint32_t n_hot,
                                   uint32_t *keys_out_dev, uint32_t *vals_out_dev, uint32_t capacity, uint64_t *found)
{
	if (!ctx || !found || (n_hot && !hot_keys_dev) || (capacity && (!keys_out_dev || !vals_out_dev))) return HJB_E_INVALID;
	if (n_hot > kMaxHotKeys) return fail(ctx, HJB_E_INVALID, "more than 256 hot keys");
	int rc;
	if ((rc = check_rel(ctx, R, true))) return rc;
	CK(cudaSetDevice(ctx->device));
	cudaStream_t s = ctx->stream;
	ctx->launches += launch_select_hot(R->keys, R->vals, R->tuples, hot_keys_dev, n_hot, keys_out_dev, vals_out_dev, capacity,
	                                   ctx->d_scalars + 12, s, ctx->sms);
	CK(cudaMemcpyAsync(ctx->h_scalars + 12, ctx->d_scalars + 12, 8, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	*found = ctx->h_scalars[12];              // may exceed capacity: the caller then gives up on the hot path for this key set
	return HJB_OK;
}

static int hot_join_enqueue(hjb_ctx *ctx, uint32_t *launches)
{
	const hjb_opts *o = &ctx->step_opts;
	*launches += launch_hot_join(ctx->hot_s.keys, ctx->hot_s.vals, ctx->hot_s.tuples, ctx->hot_r.keys, ctx->hot_r.vals,
	                             (uint32_t)ctx->hot_r.tuples, hjb_hash_factor(o->seed, 1), ctx->out_cols, ctx->out_cols + ctx->out_cap,
	                             ctx->out_cols + 2 * ctx->out_cap, o->materialize ? ctx->out_cap : 0, ctx->d_scalars, ctx->stream, ctx->sms);
	return HJB_OK;
}

extern "C" int hjb_cpra_hot_join(hjb_ctx *ctx, const hjb_rel *S_hot, const hjb_rel *R_hot)
{
	if (!ctx) return HJB_E_INVALID;
	if (ctx->step_state != 3) return fail(ctx, HJB_E_INVALID, "hjb_cpra_join_async must precede");
	int rc;
	if ((rc = check_rel(ctx, S_hot, true)) || (rc = check_rel(ctx, R_hot, true))) return rc;
	if (R_hot->tuples > kMaxHotBuild) return fail(ctx, HJB_E_INVALID, "more than 4096 build tuples with hot keys");
	CK(cudaSetDevice(ctx->device));
	ctx->hot_s = *S_hot;
	ctx->hot_r = *R_hot;
	ctx->hot_pending = true;
	rc = hot_join_enqueue(ctx, &ctx->step_launches);
	CK(cudaGetLastError());
	return rc;
}

// ---- the stream-ordered step from / to HOST memory: the chunk is copied in by the count call, the rows are
// copied out by the finish call (what a host application holds after fread, cpra2.cpp:2128-2136)

extern "C" int hjb_host_register(void *ptr, size_t bytes)
{
	if (!ptr || !bytes) return HJB_E_INVALID;
	return cudaHostRegister(ptr, bytes, cudaHostRegisterPortable) == cudaSuccess ? HJB_OK : HJB_E_CUDA;
}

extern "C" int hjb_host_unregister(void *ptr)
{
	if (!ptr) return HJB_E_INVALID;
	return cudaHostUnregister(ptr) == cudaSuccess ? HJB_OK : HJB_E_CUDA;
}

extern "C" int hjb_cpra_count_async_host(hjb_ctx *ctx, const hjb_rel *R, const hjb_rel *S, const hjb_opts *opts,
                                         uint64_t *counts_dev)
{
	if (!ctx || !counts_dev) return HJB_E_INVALID;
	int rc;
	if ((rc = check_rel(ctx, R, false)) || (rc = check_rel(ctx, S, false))) return rc;
	CK(cudaSetDevice(ctx->device));
	const size_t rb = pad256(R->tuples * 4), sb = pad256(S->tuples * 4);
	if ((rc = grow_device(ctx, &ctx->in_buf, &ctx->in_bytes, 2 * rb + 2 * sb + 256))) return rc;
	cudaStream_t s = ctx->stream;
	uint32_t *drk = (uint32_t *)ctx->in_buf, *drv = (uint32_t *)(ctx->in_buf + rb);
	uint32_t *dsk = (uint32_t *)(ctx->in_buf + 2 * rb), *dsv = (uint32_t *)(ctx->in_buf + 2 * rb + sb);
	CK(cudaEventRecord(ctx->ev[8], s));
	if (R->tuples) {
		CK(cudaMemcpyAsync(drk, R->keys, R->tuples * 4, cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(drv, R->vals, R->tuples * 4, cudaMemcpyHostToDevice, s));
	}
	if (S->tuples) {
		CK(cudaMemcpyAsync(dsk, S->keys, S->tuples * 4, cudaMemcpyHostToDevice, s));
		CK(cudaMemcpyAsync(dsv, S->vals, S->tuples * 4, cudaMemcpyHostToDevice, s));
	}
	CK(cudaEventRecord(ctx->ev[9], s));
	const hjb_rel dR = {drk, drv, R->tuples}, dS = {dsk, dsv, S->tuples};
	return hjb_cpra_count_async(ctx, &dR, &dS, opts, counts_dev);
}

extern "C" int hjb_cpra_finish_host(hjb_ctx *ctx, hjb_result *out, uint64_t received[2], uint64_t largest[2])
{
	int rc = hjb_cpra_finish(ctx, out, received, largest);
	if (rc) return rc;
	cudaStream_t s = ctx->stream;
	float h2d = 0, d2h = 0;
	CK(cudaEventElapsedTime(&h2d, ctx->ev[8], ctx->ev[9]));
	if (out->keys && out->count) {
		if ((rc = grow_host_rows(ctx, out->count))) return rc;
		CK(cudaEventRecord(ctx->ev[10], s));
		const uint32_t *cols[3] = {out->keys, out->outer_vals, out->inner_vals};
		for (int c = 0; c < 3; ++c)
			CK(cudaMemcpyAsync(ctx->h_rows + (size_t)c * ctx->h_rows_cap, cols[c], out->count * 4, cudaMemcpyDeviceToHost, s));
		CK(cudaEventRecord(ctx->ev[11], s));
		CK(cudaStreamSynchronize(s));
		CK(cudaEventElapsedTime(&d2h, ctx->ev[10], ctx->ev[11]));
		out->keys = ctx->h_rows;
		out->outer_vals = ctx->h_rows + ctx->h_rows_cap;
		out->inner_vals = ctx->h_rows + 2 * ctx->h_rows_cap;
	} else {
		out->keys = out->outer_vals = out->inner_vals = nullptr;
	}
	out->rows_on_device = 0;
	out->phase_ms[5] = h2d;
	out->phase_ms[6] = d2h;
	return HJB_OK;
}

// ------------------------------------------------------------------ kernel-level entry points

extern "C" int hjb_histogram(hjb_ctx *ctx, const uint32_t *keys, uint64_t size, uint32_t *counts, uint32_t factor,
                             int shift, int bits)
{
	if (!ctx || !counts || bits < 1 || bits > kMaxRadixBits || shift < 0 || shift + bits > 32)
		return fail(ctx, HJB_E_INVALID, "bad histogram arguments");
	if (size && (!keys || ((uintptr_t)keys & 15))) return fail(ctx, HJB_E_INVALID, "keys must be a 16-byte aligned device column");
	CK(cudaSetDevice(ctx->device));
	const uint32_t F = 1u << bits;
	int rc;
	if ((rc = grow_device(ctx, &ctx->ws, &ctx->ws_bytes, (size_t)F * 4 + 256))) return rc;
	ctx->launches += launch_histogram_only(keys, size, (uint32_t *)ctx->ws, factor, 32 - shift - bits, bits, ctx->stream, ctx->sms);
	CK(cudaMemcpyAsync(counts, ctx->ws, (size_t)F * 4, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaGetLastError());
	return HJB_OK;
}

extern "C" int hjb_partition_pass(hjb_ctx *ctx, const uint32_t *keys, const uint32_t *vals, uint64_t size,
                                  const uint32_t *parent_offsets, uint32_t *keys_out, uint32_t *vals_out,
                                  uint32_t *child_offsets, uint32_t factor, int shift, int bits)
{
	if (!ctx || !child_offsets || bits < 1 || bits > kMaxRadixBits || shift < 0 || shift + bits > 22)
		return fail(ctx, HJB_E_INVALID, "bad partition arguments");
	if (size > 0xFFFFFFFFull) return fail(ctx, HJB_E_INVALID, "more than 2^32-1 tuples");
	if (size && (!keys || !vals || !keys_out || !vals_out)) return fail(ctx, HJB_E_INVALID, "null column");
	if ((((uintptr_t)keys) | ((uintptr_t)vals) | ((uintptr_t)keys_out) | ((uintptr_t)vals_out)) & 15)
		return fail(ctx, HJB_E_INVALID, "columns must be 16-byte aligned");
	if (shift > 0 && !parent_offsets) return fail(ctx, HJB_E_INVALID, "parent_offsets required when shift > 0");
	CK(cudaSetDevice(ctx->device));
	const uint32_t np = 1u << shift, nc = np << bits;
	uint32_t chunk, mi, tiles;
	const size_t rs = radix_scratch_bytes(size, np, bits, &chunk, &mi, &tiles);
	int rc;
	if ((rc = grow_device(ctx, &ctx->ws, &ctx->ws_bytes, rs + pad256(((size_t)np + 1) * 4) + pad256(((size_t)nc + 1) * 4) + 4096))) return rc;
	Bump b = {ctx->ws, 0};
	uint32_t *d_parent = b.take<uint32_t>(np + 1);
	uint32_t *d_child = b.take<uint32_t>(nc + 1);
	cudaStream_t s = ctx->stream;
	if (size == 0) {
		memset(child_offsets, 0, ((size_t)nc + 1) * 4);
		return HJB_OK;
	}
	if (shift > 0) CK(cudaMemcpyAsync(d_parent, parent_offsets, ((size_t)np + 1) * 4, cudaMemcpyHostToDevice, s));
	RadixPassArgs a = {};
	a.keys = keys; a.vals = vals; a.keys_out = keys_out; a.vals_out = vals_out;
	a.n = size;
	a.np = np;
	a.parent_off = shift > 0 ? d_parent : nullptr;
	a.child_off = d_child;
	a.factor = factor;
	a.bits = bits;
	a.rshift = 32 - shift - bits;
	radix_carve(a, ctx->ws + b.off, true);
	ctx->launches += launch_radix_pass(a, s, ctx->sms);
	CK(cudaMemcpyAsync(child_offsets, d_child, ((size_t)nc + 1) * 4, cudaMemcpyDeviceToHost, s));
	CK(cudaStreamSynchronize(s));
	CK(cudaGetLastError());
	return HJB_OK;
}

extern "C" int hjb_npj_build(hjb_ctx *ctx, const uint32_t *keys, const uint32_t *vals, uint64_t size, uint64_t *table,
                             uint64_t buckets, uint32_t factor)
{
	if (!ctx || !table || buckets == 0 || buckets > 0xFFFFFFFFull || size >= buckets * 4)
		return fail(ctx, HJB_E_INVALID, "bad build arguments (need size < 4*buckets)");
	if (size && (!keys || !vals || ((((uintptr_t)keys) | ((uintptr_t)vals)) & 15)))
		return fail(ctx, HJB_E_INVALID, "columns must be 16-byte aligned device memory");
	CK(cudaSetDevice(ctx->device));
	NpjArgs a;
	memset(&a, 0, sizeof a);
	a.rk = keys; a.rv = vals; a.nr = size;
	a.table = table; a.buckets = buckets; a.factor = factor;
	a.scalars = ctx->d_scalars;
	CK(cudaMemsetAsync(ctx->d_scalars, 0, 16 * 8, ctx->stream));
	ctx->launches += launch_npj_build(a, ctx->stream, ctx->sms);
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaGetLastError());
	return HJB_OK;
}

// ------------------------------------------------------------------ generator, checksums, files

extern "C" int hjb_generate(hjb_ctx *ctx, const hjb_gen *g, uint32_t *keys_dev, uint32_t *vals_dev)
{
	if (!ctx || !g) return HJB_E_INVALID;
	if (g->tuples && (!keys_dev || !vals_dev)) return fail(ctx, HJB_E_INVALID, "null column");
	if (g->kind < 0 || g->kind > 2) return fail(ctx, HJB_E_INVALID, "kind must be 0, 1 or 2");
	if (g->total == 0 || g->first + g->tuples > g->total) return fail(ctx, HJB_E_INVALID, "first + tuples exceeds total");
	if (g->domain == 0 || g->domain > 0xFFFFFFFEull) return fail(ctx, HJB_E_INVALID, "domain must be in [1, 2^32-2]");
	if (g->kind == 0 && g->total > 0xFFFFFFFEull) return fail(ctx, HJB_E_INVALID, "at most 2^32-2 unique keys");
	if (g->kind == 2 && (g->selectivity < 0.0 || g->selectivity > 1.0 || g->theta < 0.0))
		return fail(ctx, HJB_E_INVALID, "bad theta / selectivity");
	CK(cudaSetDevice(ctx->device));
	ctx->launches += launch_generate(*g, keys_dev, vals_dev, ctx->stream);
	CK(cudaGetLastError());
	return HJB_OK;
}

extern "C" int hjb_column_sum(hjb_ctx *ctx, const uint32_t *col_dev, uint64_t size, uint64_t *sum)
{
	if (!ctx || !sum || (size && !col_dev)) return HJB_E_INVALID;
	CK(cudaSetDevice(ctx->device));
	ctx->launches += launch_column_sum(col_dev, size, ctx->d_scalars + 8, ctx->stream, ctx->sms);
	CK(cudaMemcpyAsync(ctx->h_scalars + 8, ctx->d_scalars + 8, 8, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaGetLastError());
	*sum = ctx->h_scalars[8];
	return HJB_OK;
}

extern "C" int hjb_rows_fingerprint(hjb_ctx *ctx, const uint32_t *keys_dev, const uint32_t *outer_vals_dev,
                                    const uint32_t *inner_vals_dev, uint64_t rows, uint64_t fp[2])
{
	if (!ctx || !fp || (rows && (!keys_dev || !outer_vals_dev || !inner_vals_dev))) return HJB_E_INVALID;
	CK(cudaSetDevice(ctx->device));
	ctx->launches += launch_rows_fingerprint(keys_dev, outer_vals_dev, inner_vals_dev, rows, ctx->d_scalars + 8, ctx->stream, ctx->sms);
	CK(cudaMemcpyAsync(ctx->h_scalars + 8, ctx->d_scalars + 8, 16, cudaMemcpyDeviceToHost, ctx->stream));
	CK(cudaStreamSynchronize(ctx->stream));
	CK(cudaGetLastError());
	fp[0] = ctx->h_scalars[8];
	fp[1] = ctx->h_scalars[9];
	return HJB_OK;
}

static int rel_path(char *buf, size_t cap, const char *dir, int outer, char col, uint64_t n)
{
	const int w = snprintf(buf, cap, "%s/%c%c_%llu.txt", dir && *dir ? dir : ".", outer ? 'o' : 'i', col,
	                       (unsigned long long)n);
	return w > 0 && (size_t)w < cap ? 0 : 1;
}

extern "C" int hjb_relation_write(const char *dir, int outer, uint64_t tuples, const uint32_t *keys, const uint32_t *vals)
{
	if (tuples && (!keys || !vals)) return HJB_E_INVALID;
	char path[4096];
	const uint32_t *cols[2] = {keys, vals};
	const char names[2] = {'k', 'v'};
	for (int c = 0; c < 2; ++c) {
		if (rel_path(path, sizeof path, dir, outer, names[c], tuples)) return HJB_E_INVALID;
		FILE *f = fopen(path, "wb");
		if (!f) return HJB_E_IO;
		const size_t w = fwrite(cols[c], 4, tuples, f);
		if (fclose(f) != 0 || w != tuples) return HJB_E_IO;
	}
	return HJB_OK;
}

extern "C" int hjb_relation_read(const char *dir, int outer, uint64_t tuples, uint32_t *keys, uint32_t *vals)
{
	if (tuples && (!keys || !vals)) return HJB_E_INVALID;
	char path[4096];
	uint32_t *cols[2] = {keys, vals};
	const char names[2] = {'k', 'v'};
	for (int c = 0; c < 2; ++c) {
		if (rel_path(path, sizeof path, dir, outer, names[c], tuples)) return HJB_E_INVALID;
		FILE *f = fopen(path, "rb");
		if (!f) return HJB_E_IO;
		// the reference never checks fopen/fread (npj.cpp:1031-1039); a short or long file is an error here
		fseek(f, 0, SEEK_END);
		const long long bytes = ftell(f);
		fseek(f, 0, SEEK_SET);
		size_t r = 0;
		if (bytes == (long long)(tuples * 4)) r = fread(cols[c], 4, tuples, f);
		fclose(f);
		if (bytes != (long long)(tuples * 4) || r != tuples) return HJB_E_IO;
	}
	return HJB_OK;
}
