/* hj_oracle.c -- CPU ORACLE for the hash-join hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement (no SIMD, no copy) of what the reference computes; every function
 * cites the reference file:line it follows.  See hj_oracle.h for who may call this and how
 * it is pinned.  Nothing under hash_join_codes_knl_b200/ links against it.
 */
#define _GNU_SOURCE
#include "hj_oracle.h"
#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------ utilities */

static double now_seconds(void)
{
	struct timespec t;
	clock_gettime(CLOCK_MONOTONIC, &t);
	return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

static void *xalloc64(size_t bytes)
{
	void *p = NULL;
	if (bytes == 0) bytes = 64;
	return posix_memalign(&p, 64, bytes) ? NULL : p;
}

/* MT19937, restating rand32_init / rand32_next (npj.cpp:138-175). */
void hjo_rand32_seed(hjo_rand32 *st, uint32_t seed)
{
	uint32_t *n = st->num;
	n[0] = seed;
	for (size_t i = 0; i != 623; ++i)
		n[i + 1] = 0x6c078965u * (n[i] ^ (n[i] >> 30));
	/* npj.cpp:144 stops at 623, so n[624] and the "+ i + 1" of textbook MT are absent */
	n[624] = 0;
	st->index = 624;
}

uint32_t hjo_rand32_next(hjo_rand32 *st)
{
	uint32_t y, *n = st->num;
	if (st->index == 624) {
		size_t i = 0;
		do {
			y = (n[i] & 0x80000000u) + (n[i + 1] & 0x7fffffffu);
			n[i] = n[i + 397] ^ (y >> 1);
			n[i] ^= 0x9908b0dfu & (0u - (y & 1u));
		} while (++i != 227);
		n[624] = n[0];
		do {
			y = (n[i] & 0x80000000u) + (n[i + 1] & 0x7fffffffu);
			n[i] = n[i - 227] ^ (y >> 1);
			n[i] ^= 0x9908b0dfu & (0u - (y & 1u));
		} while (++i != 624);
		st->index = 0;
	}
	y = n[st->index++];
	y ^= (y >> 11);
	y ^= (y << 7) & 0x9d2c5680u;
	y ^= (y << 15) & 0xefc60000u;
	y ^= (y >> 18);
	return y;
}

/* multiplicative hash + multiply-shift range reduction (npj.cpp:200-201) */
uint32_t hjo_hash(uint32_t key, uint32_t factor, uint64_t n)
{
	uint64_t h = (uint32_t)(key * factor);
	return (uint32_t)((h * n) >> 32);
}

int hjo_odd_prime(uint64_t x) /* cpra2.cpp:281-289 */
{
	for (uint64_t d = 3; d * d <= x; d += 2)
		if (x % d == 0) return 0;
	return 1;
}

size_t hjo_thread_beg(size_t size, size_t alignment, size_t thread, size_t threads)
{
	size_t part = (size / threads) & ~(alignment - 1);
	return part * thread;
}

size_t hjo_thread_end(size_t size, size_t alignment, size_t thread, size_t threads)
{
	size_t part = (size / threads) & ~(alignment - 1);
	if (thread + 1 == threads) return size;
	return part * (thread + 1);
}

/* ------------------------------------------------------------------ generator */

void hjo_shuffle(uint32_t *data, size_t size, hjo_rand32 *gen) /* cpra2.cpp:1530-1542 */
{
	for (size_t i = 0; i != size; ++i) {
		uint64_t j = hjo_rand32_next(gen);
		j *= size - i;
		j >>= 32;
		j += i;
		uint32_t t = data[i];
		data[i] = data[j];
		data[j] = t;
	}
}

void hjo_unique(uint32_t *keys, size_t size, uint32_t *table, size_t buckets,
                uint32_t factor, uint32_t empty, hjo_rand32 *gen) /* cpra2.cpp:1544-1570 */
{
	size_t i = 0;
	while (i != size) {
		uint32_t key;
		do {
			key = hjo_rand32_next(gen);
		} while (key == empty);
		size_t h = hjo_hash(key, factor, buckets);
		uint32_t tab = table[h];
		while (tab != key) {
			if (tab == empty) { /* single-threaded: the CAS always succeeds */
				table[h] = key;
				keys[i++] = key;
				break;
			}
			if (++h == buckets) h = 0;
			tab = table[h];
		}
	}
}

/* generate_data_for_join (cpra2.cpp:1578-1696) with the `threads` workers run one after the
 * other.  Worker t seeds its MT19937 with seed+t (npj.cpp:1049 gives each thread its own
 * rand()); factors come from one more MT19937 stream instead of libc rand() (cpra2.cpp:2069-2071). */
int hjo_generate(size_t inner_tuples, size_t outer_tuples, double selectivity,
                 int threads, uint32_t seed,
                 uint32_t *inner_keys, uint32_t *inner_vals,
                 uint32_t *outer_keys, uint32_t *outer_vals,
                 uint32_t *inner_factor_out, uint32_t *outer_factor_out)
{
	if (threads < 1 || selectivity < 0.0 || selectivity > 1.0) return 1;
	size_t T = (size_t)threads;
	size_t inner_distinct = inner_tuples < outer_tuples ? inner_tuples : outer_tuples;
	size_t outer_distinct = inner_distinct;                         /* cpra2.cpp:2024-2026 */
	size_t join_distinct = (size_t)((double)inner_distinct * selectivity); /* write.cpp:1689 */
	size_t distinct = inner_distinct + outer_distinct - join_distinct;
	size_t buckets = distinct * 2 + 1;                              /* cpra2.cpp:2086-2088 */
	while (!hjo_odd_prime(buckets)) buckets += 2;
	hjo_rand32 fgen;
	hjo_rand32_seed(&fgen, seed ^ 0x9e3779b9u);
	uint32_t unique_factor = hjo_rand32_next(&fgen) | 1u;
	uint32_t inner_factor = hjo_rand32_next(&fgen) | 1u;
	uint32_t outer_factor = hjo_rand32_next(&fgen) | 1u;
	uint32_t *uniq = (uint32_t *)malloc((distinct ? distinct : 1) * sizeof(uint32_t));
	uint32_t *table = (uint32_t *)calloc(buckets, sizeof(uint32_t));
	hjo_rand32 *gens = (hjo_rand32 *)malloc(T * sizeof(hjo_rand32));
	if (!uniq || !table || !gens) { free(uniq); free(table); free(gens); return 2; }
	for (size_t t = 0; t != T; ++t) {
		hjo_rand32_seed(&gens[t], seed + (uint32_t)t);
		size_t db = hjo_thread_beg(distinct, 1, t, T), de = hjo_thread_end(distinct, 1, t, T);
		hjo_unique(&uniq[db], de - db, table, buckets, unique_factor, 0, &gens[t]);
	}
	free(table);
	const uint32_t *inner_unique = uniq;
	const uint32_t *outer_unique = &uniq[inner_distinct - join_distinct];
	for (size_t t = 0; t != T; ++t) {
		size_t ib = hjo_thread_beg(inner_tuples, 16, t, T), ie = hjo_thread_end(inner_tuples, 16, t, T);
		size_t u = hjo_thread_beg(inner_distinct, 16, t, T), ue = hjo_thread_end(inner_distinct, 16, t, T);
		for (size_t i = ib; i != ie; ++i) {
			if (u != ue) inner_keys[i] = inner_unique[u++];
			else {
				uint64_t r = hjo_rand32_next(&gens[t]);
				inner_keys[i] = inner_unique[(r * inner_distinct) >> 32];
			}
		}
		size_t ob = hjo_thread_beg(outer_tuples, 16, t, T), oe = hjo_thread_end(outer_tuples, 16, t, T);
		u = hjo_thread_beg(outer_distinct, 16, t, T);
		ue = hjo_thread_end(outer_distinct, 16, t, T);
		for (size_t o = ob; o != oe; ++o) {
			if (u != ue) outer_keys[o] = outer_unique[u++];
			else {
				uint64_t r = hjo_rand32_next(&gens[t]);
				outer_keys[o] = outer_unique[(r * outer_distinct) >> 32];
			}
		}
	}
	free(uniq);
	hjo_shuffle(inner_keys, inner_tuples, &gens[0]);                /* cpra2.cpp:1654-1660 */
	hjo_shuffle(outer_keys, outer_tuples, &gens[0]);
	for (size_t i = 0; i != inner_tuples; ++i) inner_vals[i] = inner_keys[i] * inner_factor;
	for (size_t o = 0; o != outer_tuples; ++o) outer_vals[o] = outer_keys[o] * outer_factor;
	free(gens);
	if (inner_factor_out) *inner_factor_out = inner_factor;
	if (outer_factor_out) *outer_factor_out = outer_factor;
	return 0;
}

/* raw little-endian uint32[n], keys and payloads in separate files named
 * <prefix>k_<n>.txt / <prefix>v_<n>.txt (write.cpp:1824-1865; read npj.cpp:1013-1039) */
static int rel_path(char *buf, size_t cap, const char *dir, const char *prefix, char col, size_t n)
{
	int w = snprintf(buf, cap, "%s/%s%c_%zu.txt", dir && *dir ? dir : ".", prefix, col, n);
	return w > 0 && (size_t)w < cap ? 0 : 1;
}

int hjo_relation_write(const char *dir, const char *prefix, size_t tuples,
                       const uint32_t *keys, const uint32_t *vals)
{
	char path[4096];
	const uint32_t *cols[2] = { keys, vals };
	const char names[2] = { 'k', 'v' };
	for (int c = 0; c != 2; ++c) {
		if (rel_path(path, sizeof path, dir, prefix, names[c], tuples)) return 1;
		FILE *f = fopen(path, "wb");
		if (!f) return 2;
		size_t w = fwrite(cols[c], sizeof(uint32_t), tuples, f);
		if (fclose(f) != 0 || w != tuples) return 3;
	}
	return 0;
}

int hjo_relation_read(const char *dir, const char *prefix, size_t tuples,
                      uint32_t *keys, uint32_t *vals)
{
	char path[4096];
	uint32_t *cols[2] = { keys, vals };
	const char names[2] = { 'k', 'v' };
	for (int c = 0; c != 2; ++c) {
		if (rel_path(path, sizeof path, dir, prefix, names[c], tuples)) return 1;
		FILE *f = fopen(path, "rb");
		if (!f) return 2;
		size_t r = fread(cols[c], sizeof(uint32_t), tuples, f);
		fclose(f);
		if (r != tuples) return 3;
	}
	return 0;
}

/* ------------------------------------------------------------------ result sink */

typedef struct sink {
	uint64_t count, sum_key, sum_outer, sum_inner;
	uint32_t *keys, *outer_vals, *inner_vals;
	size_t cap;
	int materialize;
	int oom;
} sink_t;

static inline void sink_emit(sink_t *s, uint32_t key, uint32_t oval, uint32_t ival)
{
	if (s->materialize) {
		if (s->count == s->cap) {
			size_t ncap = s->cap ? s->cap * 2 : 4096;
			uint32_t *k = (uint32_t *)realloc(s->keys, ncap * sizeof(uint32_t));
			uint32_t *o = (uint32_t *)realloc(s->outer_vals, ncap * sizeof(uint32_t));
			uint32_t *i = (uint32_t *)realloc(s->inner_vals, ncap * sizeof(uint32_t));
			if (k) s->keys = k;
			if (o) s->outer_vals = o;
			if (i) s->inner_vals = i;
			if (!k || !o || !i) { s->oom = 1; s->materialize = 0; }
			else s->cap = ncap;
		}
		if (s->materialize) {
			s->keys[s->count] = key;
			s->outer_vals[s->count] = oval;
			s->inner_vals[s->count] = ival;
		}
	}
	s->count++;
	s->sum_key += key;
	s->sum_outer += oval;
	s->sum_inner += ival;
}

static int sinks_merge(sink_t *sinks, size_t n, int materialize, hjo_result *out)
{
	memset(out, 0, sizeof *out);
	int oom = 0;
	for (size_t t = 0; t != n; ++t) {
		out->count += sinks[t].count;
		out->sum_key += sinks[t].sum_key;
		out->sum_outer += sinks[t].sum_outer;
		out->sum_inner += sinks[t].sum_inner;
		oom |= sinks[t].oom;
	}
	if (materialize && !oom) {
		size_t m = out->count ? out->count : 1;
		out->keys = (uint32_t *)malloc(m * sizeof(uint32_t));
		out->outer_vals = (uint32_t *)malloc(m * sizeof(uint32_t));
		out->inner_vals = (uint32_t *)malloc(m * sizeof(uint32_t));
		if (!out->keys || !out->outer_vals || !out->inner_vals) oom = 1;
		else {
			size_t o = 0; /* the reference closes gaps between per-thread blocks (npj.cpp:475-514) */
			for (size_t t = 0; t != n; ++t) {
				memcpy(&out->keys[o], sinks[t].keys, sinks[t].count * sizeof(uint32_t));
				memcpy(&out->outer_vals[o], sinks[t].outer_vals, sinks[t].count * sizeof(uint32_t));
				memcpy(&out->inner_vals[o], sinks[t].inner_vals, sinks[t].count * sizeof(uint32_t));
				o += sinks[t].count;
			}
		}
	}
	for (size_t t = 0; t != n; ++t) {
		free(sinks[t].keys);
		free(sinks[t].outer_vals);
		free(sinks[t].inner_vals);
	}
	if (oom) { hjo_result_free(out); return 2; }
	return 0;
}

void hjo_result_free(hjo_result *r)
{
	if (!r) return;
	free(r->keys);
	free(r->outer_vals);
	free(r->inner_vals);
	r->keys = r->outer_vals = r->inner_vals = NULL;
}

static int has_zero_key(const uint32_t *k, size_t n)
{
	for (size_t i = 0; i != n; ++i)
		if (k[i] == 0) return 1;
	return 0;
}

/* ------------------------------------------------------------------ NPJ */

/* build (npj.cpp:190-212): linear probing, slot = payload<<32 | key, claimed by CAS when the
 * slot's low word is `empty`.  The DDR/MCDRAM `ratio` split (npj.cpp:195,202) is one table here. */
void hjo_npj_build(const uint32_t *keys, const uint32_t *vals, size_t size,
                   volatile uint64_t *table, size_t buckets, uint32_t factor, uint32_t empty)
{
	for (size_t i = 0; i != size; ++i) {
		uint32_t key = keys[i];
		uint64_t pair = ((uint64_t)vals[i] << 32) | key;
		uint64_t h = hjo_hash(key, factor, buckets);
		uint64_t tab = table[h];
		while (empty != (uint32_t)tab ||
		       !__sync_bool_compare_and_swap(&table[h], tab, pair)) {
			if (++h == buckets) h = 0;
			tab = table[h];
		}
	}
}

/* scalar probe (npj.cpp:412-445): walk the chain to the first empty slot, emit every equal key */
static void npj_probe(const uint32_t *keys, const uint32_t *vals, size_t size,
                      const uint64_t *table, size_t buckets, uint32_t factor, uint32_t empty,
                      sink_t *out)
{
	for (size_t i = 0; i != size; ++i) {
		uint32_t key = keys[i], val = vals[i];
		uint64_t h = hjo_hash(key, factor, buckets);
		uint64_t tab = table[h];
		while (empty != (uint32_t)tab) {
			if (key == (uint32_t)tab) sink_emit(out, key, val, (uint32_t)(tab >> 32));
			if (++h == buckets) h = 0;
			tab = table[h];
		}
	}
}

typedef struct join_shared {
	const uint32_t *rk, *rv, *sk, *sv;
	size_t nr, ns;
	int threads;
	uint32_t seed;
	pthread_barrier_t barrier;
	sink_t *sinks;
	double *seconds;
	/* npj */
	uint64_t *table;
	size_t buckets;
	uint32_t factor;
	/* phj / cpra */
	uint32_t *r_keys[2], *r_vals[2], *s_keys[2], *s_vals[2];
	uint32_t **r_counts, **s_counts;       /* [thread] -> counts (published pointers) */
	uint32_t **r_base_k, **r_base_v, **s_base_k, **s_base_v; /* cpra: per-thread chunk bases */
	uint32_t thread_factor;
	size_t num_partitions;
} join_shared;

typedef struct join_thread {
	join_shared *sh;
	int thread;
	pthread_t id;
} join_thread;

/* run (npj.cpp:769-927): timed region = table init + build + probe (+ close_gaps) */
static void *npj_run(void *arg)
{
	join_thread *jt = (join_thread *)arg;
	join_shared *d = jt->sh;
	size_t t = (size_t)jt->thread, T = (size_t)d->threads;
	size_t ib = hjo_thread_beg(d->nr, 16, t, T), ie = hjo_thread_end(d->nr, 16, t, T);
	size_t ob = hjo_thread_beg(d->ns, 16, t, T), oe = hjo_thread_end(d->ns, 16, t, T);
	pthread_barrier_wait(&d->barrier);
	double t0 = now_seconds();
	size_t tb = hjo_thread_beg(d->buckets, 1, t, T), te = hjo_thread_end(d->buckets, 1, t, T);
	for (size_t i = tb; i != te; ++i) d->table[i] = 0;               /* set, npj.cpp:366-380 */
	pthread_barrier_wait(&d->barrier);
	hjo_npj_build(&d->rk[ib], &d->rv[ib], ie - ib, d->table, d->buckets, d->factor, 0);
	pthread_barrier_wait(&d->barrier);
	npj_probe(&d->sk[ob], &d->sv[ob], oe - ob, d->table, d->buckets, d->factor, 0, &d->sinks[t]);
	pthread_barrier_wait(&d->barrier);
	d->seconds[t] = now_seconds() - t0;
	return NULL;
}

static int run_threads(join_shared *sh, void *(*fn)(void *), int materialize, hjo_result *out)
{
	size_t T = (size_t)sh->threads;
	join_thread *jt = (join_thread *)calloc(T, sizeof *jt);
	sh->sinks = (sink_t *)calloc(T, sizeof(sink_t));
	sh->seconds = (double *)calloc(T, sizeof(double));
	if (!jt || !sh->sinks || !sh->seconds) return 2;
	for (size_t t = 0; t != T; ++t) sh->sinks[t].materialize = materialize;
	pthread_barrier_init(&sh->barrier, NULL, (unsigned)T);
	for (size_t t = 0; t != T; ++t) {
		jt[t].sh = sh;
		jt[t].thread = (int)t;
		pthread_create(&jt[t].id, NULL, fn, &jt[t]);
	}
	for (size_t t = 0; t != T; ++t) pthread_join(jt[t].id, NULL);
	pthread_barrier_destroy(&sh->barrier);
	int rc = sinks_merge(sh->sinks, T, materialize, out);
	double mx = 0;
	for (size_t t = 0; t != T; ++t) if (sh->seconds[t] > mx) mx = sh->seconds[t];
	out->seconds = mx;                                               /* npj.cpp:1114: max over threads */
	free(jt);
	free(sh->sinks);
	free(sh->seconds);
	return rc;
}

int hjo_npj(const uint32_t *rk, const uint32_t *rv, size_t nr,
            const uint32_t *sk, const uint32_t *sv, size_t ns,
            int threads, uint32_t seed, int materialize, hjo_result *out)
{
	if (!out || threads < 1) return 1;
	if (has_zero_key(rk, nr) || has_zero_key(sk, ns)) return 3;
	join_shared sh;
	memset(&sh, 0, sizeof sh);
	sh.rk = rk; sh.rv = rv; sh.nr = nr; sh.sk = sk; sh.sv = sv; sh.ns = ns;
	sh.threads = threads; sh.seed = seed;
	sh.buckets = (size_t)((double)nr / 0.90);                       /* npj.cpp:944-946 */
	if (sh.buckets < nr + 1) sh.buckets = nr + 1;                    /* keep one empty slot so probes stop */
	hjo_rand32 g;
	hjo_rand32_seed(&g, seed);
	sh.factor = hjo_rand32_next(&g) | 1u;                            /* npj.cpp:975-977 */
	sh.table = (uint64_t *)xalloc64(sh.buckets * sizeof(uint64_t));
	if (!sh.table) return 2;
	int rc = run_threads(&sh, npj_run, materialize, out);
	free(sh.table);
	return rc;
}

/* ------------------------------------------------------------------ radix partitioning */

void hjo_histogram(const uint32_t *keys, size_t size, uint32_t *counts,
                   uint32_t factor, size_t partitions) /* histogram_s, cpra2.cpp:730-741 */
{
	for (size_t p = 0; p != partitions; ++p) counts[p] = 0;
	for (size_t i = 0; i != size; ++i) counts[hjo_hash(keys[i], factor, partitions)]++;
}

/* partition_s + flush (cpra2.cpp:742-796, 711-729).  The reference stages 16 tuples per
 * partition and flushes whole cache lines; the bytes that land are those of a stable
 * scatter, which is what this states. */
void hjo_partition(const uint32_t *keys, const uint32_t *vals, size_t size,
                   const uint32_t *counts, uint32_t *keys_out, uint32_t *vals_out,
                   uint32_t factor, size_t partitions)
{
	size_t *offsets = (size_t *)malloc((partitions ? partitions : 1) * sizeof(size_t));
	size_t i = 0;
	for (size_t p = 0; p != partitions; ++p) { offsets[p] = i; i += counts[p]; }
	for (i = 0; i != size; ++i) {
		size_t o = offsets[hjo_hash(keys[i], factor, partitions)]++;
		keys_out[o] = keys[i];
		vals_out[o] = vals[i];
	}
	free(offsets);
}

/* scalar interleave (cpra2.cpp:1426-1440 / phj.cpp:1441-1455): offsets[p] = global start of
 * partition p + tuples of p owned by lower threads; aggr_counts[p] = size of p; returns total */
size_t hjo_interleave(uint32_t **counts, uint32_t *offsets, uint32_t *aggr_counts,
                      size_t partitions, size_t thread, size_t threads)
{
	size_t i = 0;
	for (size_t p = 0; p != partitions; ++p) {
		size_t s = 0, t = 0;
		for (; t != thread; ++t) s += counts[t][p];
		offsets[p] = (uint32_t)(i + s);
		for (; t != threads; ++t) s += counts[t][p];
		aggr_counts[p] = (uint32_t)s;
		i += s;
	}
	return i;
}

/* fan-out planner (cpra2.cpp:1757-1772, phj.cpp:1791-1808) */
size_t hjo_plan_fanout(size_t partitions, size_t fanout[6])
{
	size_t passes = 0, p;
	if (partitions > 1000000) passes = 4;
	else if (partitions > 20000) passes = 3;
	else if (partitions > 400) passes = 2;
	else if (partitions > 10) passes = 1;
	for (p = 0; p != passes; ++p) fanout[p] = (size_t)pow((double)partitions, 1.0 / (double)passes);
	fanout[p] = 1;
	if (passes) {
		size_t product = 1;
		for (p = 0; p != passes - 1; ++p) product *= fanout[p];
		fanout[p] = partitions / product;
	}
	return passes;
}

/* ------------------------------------------------------------------ per-partition join */

/* build_s (cpra2.cpp:640-666): double hashing, h1 = h(k,f0,B), step h2 = h(k,f1,B-1)+1 */
void hjo_dh_build(const uint32_t *keys, const uint32_t *vals, size_t size,
                  uint64_t *table, size_t buckets, const uint32_t factor[2], uint32_t empty)
{
	for (size_t i = 0; i != buckets; ++i) table[i] = empty;
	for (size_t i = 0; i != size; ++i) {
		uint32_t k = keys[i];
		uint64_t p = ((uint64_t)vals[i] << 32) | k;
		uint64_t h1 = hjo_hash(k, factor[0], buckets);
		if (empty != (uint32_t)table[h1]) {
			uint64_t h2 = (uint64_t)hjo_hash(k, factor[1], buckets - 1) + 1;
			do {
				h1 += h2;
				if (h1 >= buckets) h1 -= buckets;
			} while (empty != (uint32_t)table[h1]);
		}
		table[h1] = p;
	}
}

/* probe_s (cpra2.cpp:668-710) */
static void dh_probe(const uint32_t *keys, const uint32_t *vals, size_t size,
                     const uint64_t *table, size_t buckets, const uint32_t factor[2],
                     uint32_t empty, sink_t *out)
{
	for (size_t i = 0; i != size; ++i) {
		uint32_t k = keys[i], v = vals[i];
		uint64_t h1 = hjo_hash(k, factor[0], buckets);
		uint64_t t = table[h1];
		if (empty != (uint32_t)t) {
			uint64_t h2 = (uint64_t)hjo_hash(k, factor[1], buckets - 1) + 1;
			do {
				if (k == (uint32_t)t) sink_emit(out, k, v, (uint32_t)(t >> 32));
				h1 += h2;
				if (h1 >= buckets) h1 -= buckets;
				t = table[h1];
			} while (empty != (uint32_t)t);
		}
	}
}

/* table sizing policy of the cache-resident join loop (phj.cpp:1899-1909, cpra2.cpp:1920-1933) */
typedef struct table_state {
	uint64_t *table;
	size_t max_buckets;
} table_state;

static size_t table_prepare(table_state *ts, size_t size, double inverse_load, int reuse)
{
	size_t buckets = (size_t)((double)size * inverse_load);
	if (buckets < 3) buckets = 3;               /* double hashing needs buckets-1 >= 1 and a free slot */
	if (!reuse) ts->max_buckets = 0;            /* cpra2.cpp:1920 resets max_buckets per partition */
	if (buckets > ts->max_buckets) {
		for (buckets |= 1; !hjo_odd_prime(buckets); buckets += 2);
		ts->max_buckets = buckets;
		ts->table = (uint64_t *)realloc(ts->table, buckets * sizeof(uint64_t));
	} else if ((double)buckets * 1.2 > (double)ts->max_buckets) {
		for (buckets |= 1; !hjo_odd_prime(buckets); buckets += 2);
	} else buckets = ts->max_buckets;
	return buckets;
}

static void join_factors(hjo_rand32 *gen, uint32_t f[2]) /* phj.cpp:1873-1876 */
{
	do {
		f[0] = hjo_rand32_next(gen) | 1u;
		f[1] = hjo_rand32_next(gen) | 1u;
	} while (((f[0] - f[1]) & 3) == 0);
}

static void swap_ptr(uint32_t **x, uint32_t **y) { uint32_t *t = *x; *x = *y; *y = t; }

/* local multi-pass partitioning of [beg,end) (phj.cpp:1809-1863, cpra2.cpp:1773-1827).
 * On return *counts_io holds the final per-partition counts and in/out are swapped so that
 * `*_in` holds the partitioned data.  Returns the number of final partitions. */
static size_t local_passes(hjo_rand32 *gen, const size_t fanout[6],
                           uint32_t **rk_in, uint32_t **rv_in, uint32_t **rk_out, uint32_t **rv_out,
                           uint32_t **sk_in, uint32_t **sv_in, uint32_t **sk_out, uint32_t **sv_out,
                           size_t r_beg, size_t r_size, size_t s_beg, size_t s_size,
                           uint32_t **r_counts_io, uint32_t **s_counts_io, uint32_t *factor_1st)
{
	uint32_t *rc = (uint32_t *)malloc(sizeof(uint32_t)), *sc = (uint32_t *)malloc(sizeof(uint32_t));
	rc[0] = (uint32_t)r_size;
	sc[0] = (uint32_t)s_size;
	size_t partitions = 1;
	for (size_t f = 0; fanout[f] != 1; ++f) {
		size_t fo = fanout[f];
		uint32_t *rn = (uint32_t *)malloc(partitions * fo * sizeof(uint32_t));
		uint32_t *sn = (uint32_t *)malloc(partitions * fo * sizeof(uint32_t));
		uint32_t factor = hjo_rand32_next(gen) | 1u;
		if (f == 0 && factor_1st) *factor_1st = factor;
		size_t i = r_beg, o = s_beg;
		for (size_t p = 0; p != partitions; ++p) {
			hjo_histogram(&(*rk_in)[i], rc[p], &rn[p * fo], factor, fo);
			hjo_partition(&(*rk_in)[i], &(*rv_in)[i], rc[p], &rn[p * fo], &(*rk_out)[i], &(*rv_out)[i], factor, fo);
			i += rc[p];
			hjo_histogram(&(*sk_in)[o], sc[p], &sn[p * fo], factor, fo);
			hjo_partition(&(*sk_in)[o], &(*sv_in)[o], sc[p], &sn[p * fo], &(*sk_out)[o], &(*sv_out)[o], factor, fo);
			o += sc[p];
		}
		free(rc);
		free(sc);
		rc = rn;
		sc = sn;
		partitions *= fo;
		swap_ptr(rk_in, rk_out); swap_ptr(rv_in, rv_out);
		swap_ptr(sk_in, sk_out); swap_ptr(sv_in, sv_out);
	}
	*r_counts_io = rc;
	*s_counts_io = sc;
	return partitions;
}

/* ------------------------------------------------------------------ PHJ */

/* run_hj (phj.cpp:1646-1949) with the join phase the shipped file comments out
 * (phj.cpp:1869-1924) and without the DDR/MCDRAM `ratio` split (original form:
 * write.cpp:782-910). */
static void *phj_run(void *arg)
{
	join_thread *jt = (join_thread *)arg;
	join_shared *d = jt->sh;
	size_t t = (size_t)jt->thread, T = (size_t)d->threads;
	uint32_t *rk_in = d->r_keys[0], *rk_out = d->r_keys[1], *rv_in = d->r_vals[0], *rv_out = d->r_vals[1];
	uint32_t *sk_in = d->s_keys[0], *sk_out = d->s_keys[1], *sv_in = d->s_vals[0], *sv_out = d->s_vals[1];
	size_t ib = hjo_thread_beg(d->nr, 16, t, T), ie = hjo_thread_end(d->nr, 16, t, T);
	size_t ob = hjo_thread_beg(d->ns, 16, t, T), oe = hjo_thread_end(d->ns, 16, t, T);
	hjo_rand32 gen;
	hjo_rand32_seed(&gen, d->seed + (uint32_t)t);
	pthread_barrier_wait(&d->barrier);
	double t0 = now_seconds();
	if (T > 1) {                                                     /* shared pass, fan-out = threads */
		size_t P = T;
		uint32_t *cnt = (uint32_t *)malloc(P * 6 * sizeof(uint32_t));
		uint32_t *r_off = &cnt[0], *s_off = &cnt[P], *r_cnt = &cnt[2 * P], *s_cnt = &cnt[3 * P];
		uint32_t *r_agg = &cnt[4 * P], *s_agg = &cnt[5 * P];
		d->r_counts[t] = r_cnt;
		d->s_counts[t] = s_cnt;
		hjo_histogram(&rk_in[ib], ie - ib, r_cnt, d->thread_factor, P);
		hjo_histogram(&sk_in[ob], oe - ob, s_cnt, d->thread_factor, P);
		pthread_barrier_wait(&d->barrier);
		hjo_interleave(d->r_counts, r_off, r_agg, P, t, T);
		hjo_interleave(d->s_counts, s_off, s_agg, P, t, T);
		for (size_t i = ib; i != ie; ++i) {                          /* partition_shared + flush_shared */
			size_t o = r_off[hjo_hash(rk_in[i], d->thread_factor, P)]++;
			rk_out[o] = rk_in[i];
			rv_out[o] = rv_in[i];
		}
		for (size_t i = ob; i != oe; ++i) {
			size_t o = s_off[hjo_hash(sk_in[i], d->thread_factor, P)]++;
			sk_out[o] = sk_in[i];
			sv_out[o] = sv_in[i];
		}
		swap_ptr(&rk_in, &rk_out); swap_ptr(&rv_in, &rv_out);
		swap_ptr(&sk_in, &sk_out); swap_ptr(&sv_in, &sv_out);
		ib = ob = 0;
		for (size_t q = 0; q != t; ++q) { ib += r_agg[q]; ob += s_agg[q]; }
		ie = ib + r_agg[t];
		oe = ob + s_agg[t];
		pthread_barrier_wait(&d->barrier);
		free(cnt);
	}
	size_t fanout[6];
	hjo_plan_fanout((ie - ib) / 6400, fanout);                       /* hash_table_limit, phj.cpp:1977 */
	uint32_t *r_counts, *s_counts, factor_1st = 0;
	size_t partitions = local_passes(&gen, fanout, &rk_in, &rv_in, &rk_out, &rv_out,
	                                 &sk_in, &sv_in, &sk_out, &sv_out,
	                                 ib, ie - ib, ob, oe - ob, &r_counts, &s_counts, &factor_1st);
	uint32_t factors[2];
	join_factors(&gen, factors);
	table_state ts = { NULL, 0 };
	size_t i = ib, o = ob;
	for (size_t p = 0; p != partitions; ++p) {
		uint32_t empty = 0;
		size_t buckets = table_prepare(&ts, r_counts[p], 1.0 / 0.4, 1);
		hjo_dh_build(&rk_in[i], &rv_in[i], r_counts[p], ts.table, buckets, factors, empty);
		i += r_counts[p];
		dh_probe(&sk_in[o], &sv_in[o], s_counts[p], ts.table, buckets, factors, empty, &d->sinks[t]);
		o += s_counts[p];
	}
	free(ts.table);
	free(r_counts);
	free(s_counts);
	pthread_barrier_wait(&d->barrier);
	d->seconds[t] = now_seconds() - t0;
	return NULL;
}

static int partitioned_join(void *(*fn)(void *), size_t num_partitions,
                            const uint32_t *rk, const uint32_t *rv, size_t nr,
                            const uint32_t *sk, const uint32_t *sv, size_t ns,
                            int threads, uint32_t seed, int materialize, hjo_result *out)
{
	if (!out || threads < 1) return 1;
	if (has_zero_key(rk, nr) || has_zero_key(sk, ns)) return 3;
	if (nr >> 32 || ns >> 32) return 4;                              /* uint32 counts/offsets, phj.cpp:1722-1727 */
	join_shared sh;
	memset(&sh, 0, sizeof sh);
	sh.nr = nr; sh.ns = ns; sh.threads = threads; sh.seed = seed;
	sh.num_partitions = num_partitions;
	hjo_rand32 g;
	hjo_rand32_seed(&g, seed ^ 0x5bd1e995u);
	sh.thread_factor = hjo_rand32_next(&g) | 1u;
	size_t T = (size_t)threads;
	int rc = 2;
	for (int b = 0; b != 2; ++b) {
		sh.r_keys[b] = (uint32_t *)xalloc64(nr * sizeof(uint32_t));
		sh.r_vals[b] = (uint32_t *)xalloc64(nr * sizeof(uint32_t));
		sh.s_keys[b] = (uint32_t *)xalloc64(ns * sizeof(uint32_t));
		sh.s_vals[b] = (uint32_t *)xalloc64(ns * sizeof(uint32_t));
	}
	sh.r_counts = (uint32_t **)calloc(T, sizeof(uint32_t *));
	sh.s_counts = (uint32_t **)calloc(T, sizeof(uint32_t *));
	sh.r_base_k = (uint32_t **)calloc(T, sizeof(uint32_t *));
	sh.r_base_v = (uint32_t **)calloc(T, sizeof(uint32_t *));
	sh.s_base_k = (uint32_t **)calloc(T, sizeof(uint32_t *));
	sh.s_base_v = (uint32_t **)calloc(T, sizeof(uint32_t *));
	if (sh.r_keys[1] && sh.r_vals[1] && sh.s_keys[1] && sh.s_vals[1] && sh.r_keys[0] && sh.r_vals[0] &&
	    sh.s_keys[0] && sh.s_vals[0] && sh.r_counts && sh.s_counts && sh.r_base_k && sh.r_base_v &&
	    sh.s_base_k && sh.s_base_v) {
		memcpy(sh.r_keys[0], rk, nr * sizeof(uint32_t));
		memcpy(sh.r_vals[0], rv, nr * sizeof(uint32_t));
		memcpy(sh.s_keys[0], sk, ns * sizeof(uint32_t));
		memcpy(sh.s_vals[0], sv, ns * sizeof(uint32_t));
		rc = run_threads(&sh, fn, materialize, out);
	}
	for (int b = 0; b != 2; ++b) {
		free(sh.r_keys[b]); free(sh.r_vals[b]); free(sh.s_keys[b]); free(sh.s_vals[b]);
	}
	free(sh.r_counts); free(sh.s_counts);
	free(sh.r_base_k); free(sh.r_base_v); free(sh.s_base_k); free(sh.s_base_v);
	return rc;
}

int hjo_phj(const uint32_t *rk, const uint32_t *rv, size_t nr,
            const uint32_t *sk, const uint32_t *sv, size_t ns,
            int threads, uint32_t seed, int materialize, hjo_result *out)
{
	return partitioned_join(phj_run, 0, rk, rv, nr, sk, sv, ns, threads, seed, materialize, out);
}

/* ------------------------------------------------------------------ CPRA */

/* run_hj (cpra2.cpp:1697-1986): every thread radix-partitions ITS OWN chunk into
 * NUM_PARTITIONS = 4096 (64 x 64) with factors common to all threads (same seed,
 * cpra2.cpp:2153), then thread t gathers the `threads` pieces of each partition it owns
 * (cpra2.cpp:1868-1906) and joins them.  Unlike the reference every thread's staged tail is
 * emitted (the reference only flushes the last partition's owner, cpra2.cpp:1969 -- a bug,
 * SURVEY.md appendix B #6). */
static void *cpra_run(void *arg)
{
	join_thread *jt = (join_thread *)arg;
	join_shared *d = jt->sh;
	size_t t = (size_t)jt->thread, T = (size_t)d->threads;
	uint32_t *rk_in = d->r_keys[0], *rk_out = d->r_keys[1], *rv_in = d->r_vals[0], *rv_out = d->r_vals[1];
	uint32_t *sk_in = d->s_keys[0], *sk_out = d->s_keys[1], *sv_in = d->s_vals[0], *sv_out = d->s_vals[1];
	size_t ib = hjo_thread_beg(d->nr, 16, t, T), ie = hjo_thread_end(d->nr, 16, t, T);
	size_t ob = hjo_thread_beg(d->ns, 16, t, T), oe = hjo_thread_end(d->ns, 16, t, T);
	hjo_rand32 gen;
	hjo_rand32_seed(&gen, d->seed);                                  /* common seed */
	pthread_barrier_wait(&d->barrier);
	double t0 = now_seconds();
	size_t fanout[6];
	hjo_plan_fanout(d->num_partitions, fanout);
	uint32_t *r_counts, *s_counts;
	size_t partitions = local_passes(&gen, fanout, &rk_in, &rv_in, &rk_out, &rv_out,
	                                 &sk_in, &sv_in, &sk_out, &sv_out,
	                                 ib, ie - ib, ob, oe - ob, &r_counts, &s_counts, NULL);
	d->r_base_k[t] = &rk_in[ib]; d->r_base_v[t] = &rv_in[ib];
	d->s_base_k[t] = &sk_in[ob]; d->s_base_v[t] = &sv_in[ob];
	d->r_counts[t] = r_counts;
	d->s_counts[t] = s_counts;
	pthread_barrier_wait(&d->barrier);
	uint32_t factors[2];
	join_factors(&gen, factors);
	size_t par_start = (partitions / T) * t, par_end = (partitions / T) * (t + 1);
	if (par_end > partitions || t == T - 1) par_end = partitions;    /* cpra2.cpp:1868-1872 */
	size_t *r_off = (size_t *)calloc(T, sizeof(size_t)), *s_off = (size_t *)calloc(T, sizeof(size_t));
	for (size_t q = 0; q != T; ++q)
		for (size_t p = 0; p != par_start; ++p) {
			r_off[q] += d->r_counts[q][p];
			s_off[q] += d->s_counts[q][p];
		}
	table_state ts = { NULL, 0 };
	for (size_t p = par_start; p != par_end; ++p) {
		size_t rsize = 0, ssize = 0;
		for (size_t q = 0; q != T; ++q) { rsize += d->r_counts[q][p]; ssize += d->s_counts[q][p]; }
		uint32_t *gk = (uint32_t *)xalloc64(rsize * sizeof(uint32_t));
		uint32_t *gv = (uint32_t *)xalloc64(rsize * sizeof(uint32_t));
		size_t w = 0;
		for (size_t q = 0; q != T; ++q) {                            /* the "copy" step, cpra2.cpp:1896-1904 */
			size_t c = d->r_counts[q][p];
			memcpy(&gk[w], &d->r_base_k[q][r_off[q]], c * sizeof(uint32_t));
			memcpy(&gv[w], &d->r_base_v[q][r_off[q]], c * sizeof(uint32_t));
			r_off[q] += c;
			w += c;
		}
		size_t buckets = table_prepare(&ts, rsize, 1.0 / 0.4, 0);
		hjo_dh_build(gk, gv, rsize, ts.table, buckets, factors, 0);
		free(gk);
		free(gv);
		gk = (uint32_t *)xalloc64(ssize * sizeof(uint32_t));
		gv = (uint32_t *)xalloc64(ssize * sizeof(uint32_t));
		w = 0;
		for (size_t q = 0; q != T; ++q) {
			size_t c = d->s_counts[q][p];
			memcpy(&gk[w], &d->s_base_k[q][s_off[q]], c * sizeof(uint32_t));
			memcpy(&gv[w], &d->s_base_v[q][s_off[q]], c * sizeof(uint32_t));
			s_off[q] += c;
			w += c;
		}
		dh_probe(gk, gv, ssize, ts.table, buckets, factors, 0, &d->sinks[t]);
		free(gk);
		free(gv);
	}
	free(ts.table);
	free(r_off);
	free(s_off);
	pthread_barrier_wait(&d->barrier);
	d->seconds[t] = now_seconds() - t0;
	pthread_barrier_wait(&d->barrier);
	free(r_counts);
	free(s_counts);
	return NULL;
}

int hjo_cpra(const uint32_t *rk, const uint32_t *rv, size_t nr,
             const uint32_t *sk, const uint32_t *sv, size_t ns,
             int threads, uint32_t seed, int materialize, hjo_result *out)
{
	return partitioned_join(cpra_run, 4096 /* cpra2.cpp:2023 */, rk, rv, nr, sk, sv, ns,
	                        threads, seed, materialize, out);
}
