/* icc_compat.h -- lets g++ compile the reference's ICC/KNC-dialect AVX-512 sources
 * (/root/reference/{npj,phj,cpra2}.cpp) where they lie, without editing them.
 *
 * TEST INFRASTRUCTURE ONLY.  Force-included (-include) by oracle/Makefile when it builds
 * oracle/_ref/.  Nothing in the product (hash_join_codes_knl_b200/, include/) sees it.
 *
 * Each mapping below names the ICC-only intrinsic the reference uses and the AVX-512F
 * equivalent g++ 13 provides.  Semantics were checked against Intel's intrinsics guide
 * descriptions of the KNC-compat forms; the pinned-oracle tests (tests/test_oracle_ref.py)
 * then check the compiled reference functions against an independent join.
 */
#pragma once
#include <string>
#include <iostream>
#include <stdio.h>
#include <string.h>
#include <stdint.h>
#include <immintrin.h>

/* main() of npj.cpp / cpra2.cpp uses these without declaring them (npj.cpp:1031-1034,
 * cpra2.cpp:2128-2131). */
static FILE *f_inner_keys, *f_inner_vals, *f_outer_keys, *f_outer_vals;

/* swap adjacent 32-bit lanes inside each 64-bit pair: {c,d,a,b} <- {d,c,b,a} */
#define _MM_SWIZ_REG_CDAB 0xB1
#define _mm512_swizzle_epi32(v, pattern) _mm512_shuffle_epi32((v), (_MM_PERM_ENUM)(pattern))

/* permute the four 128-bit lanes by a 2-bit-per-lane selector */
#define _mm512_permute4f128_epi32(v, sel) _mm512_shuffle_i32x4((v), (v), (sel))

/* KNC name for the full 16-lane variable permute (index first, data second) */
#define _mm512_permutevar_epi32(idx, v) _mm512_permutexvar_epi32((idx), (v))

/* "lo" gathers/scatters: use the LOW eight 32-bit indices of a 512-bit index vector to
 * address eight 64-bit elements */
#define _mm512_i32logather_epi64(idx, base, scale) \
	_mm512_i32gather_epi64(_mm512_castsi512_si256(idx), (const long long *)(base), (scale))
#define _mm512_mask_i32logather_epi64(src, k, idx, base, scale) \
	_mm512_mask_i32gather_epi64((src), (k), _mm512_castsi512_si256(idx), \
	                            (const long long *)(base), (scale))
#define _mm512_i32loscatter_epi64(base, idx, v, scale) \
	_mm512_i32scatter_epi64((long long *)(base), _mm512_castsi512_si256(idx), (v), (scale))
#define _mm512_mask_i32loscatter_epi64(base, k, idx, v, scale) \
	_mm512_mask_i32scatter_epi64((long long *)(base), (k), _mm512_castsi512_si256(idx), \
	                             (v), (scale))

/* mask helpers: concatenate two 16-bit masks into the low 32 bits of a 64-bit integer,
 * population count, trailing-zero count (64 when the input is zero) */
static inline uint64_t hjref_kconcatlo_64(__mmask16 hi, __mmask16 lo)
{
	return ((uint64_t)(uint16_t)hi << 16) | (uint64_t)(uint16_t)lo;
}
#define _mm512_kconcatlo_64(hi, lo) hjref_kconcatlo_64((hi), (lo))
#define _mm_countbits_64(x) ((size_t)__builtin_popcountll((unsigned long long)(x)))
static inline size_t hjref_tzcnt_64(uint64_t x) { return x ? (size_t)__builtin_ctzll(x) : 64; }
#define _mm_tzcnt_64(x) hjref_tzcnt_64((uint64_t)(x))

/* ICC accepts uint32_t* where the ps load/store/stream intrinsics want float* / void* */
static inline __m512 hjref_load_ps(const void *p) { return _mm512_load_ps(p); }
static inline void hjref_store_ps(void *p, __m512 v) { _mm512_store_ps(p, v); }
static inline void hjref_stream_ps(void *p, __m512 v) { _mm512_stream_ps((float *)p, v); }
static inline __m512 hjref_mask_loadu_ps(__m512 s, __mmask16 k, const void *p)
{
	return _mm512_mask_loadu_ps(s, k, p);
}
#undef _mm512_load_ps
#undef _mm512_store_ps
#undef _mm512_stream_ps
#undef _mm512_mask_loadu_ps
#define _mm512_load_ps(p) hjref_load_ps((const void *)(p))
#define _mm512_store_ps(p, v) hjref_store_ps((void *)(p), (v))
#define _mm512_stream_ps(p, v) hjref_stream_ps((void *)(p), (v))
#define _mm512_mask_loadu_ps(s, k, p) hjref_mask_loadu_ps((s), (k), (const void *)(p))
