// gen.cu -- deterministic relation generator on the device and column checksums
// (reference: generate_data_for_join, cpra2.cpp:1578-1696 -- unique non-zero 32-bit keys,
// foreign keys "every key once, the rest uniform", a shuffle, payload = key * odd factor;
// write.cpp:1685-1689 selectivity; its Zipf knob is inert, write.cpp:1553-1571, so the skewed
// kind here is new).  Counter-based instead of MT19937 + hash-set + Fisher-Yates: element j of
// a column depends on (seed, j) only, so 2^31-tuple relations are generated in place, sharded
// over GPUs, and regenerated identically anywhere (hash_join_codes_knl_b200/datagen.py is the
// numpy mirror of the integer kinds).
#include "hj_device.cuh"
#include "hj_internal.h"
#include <math.h>

namespace hjb {

// bijection on 32-bit integers with mix32(0) == 0 (multiply-xorshift rounds)
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t x)
{
	x ^= x >> 16;
	x *= 0x7feb352du;
	x ^= x >> 15;
	x *= 0x846ca68bu;
	x ^= x >> 16;
	return x;
}

// rank r in [0, 2^32 - 1) -> distinct non-zero key
__host__ __device__ __forceinline__ uint32_t key_of_rank(uint32_t r, uint32_t seed)
{
	const uint32_t salt = mix32(seed * 2u + 1u);
	const uint32_t k = mix32((r + 1u) ^ salt);
	// exactly one input maps to 0 (the bijection's preimage of 0); input 0 ^ salt is never
	// used by a rank (r + 1 == 0 is out of range), so its image replaces the zero
	return k ? k : mix32(salt) ? mix32(salt) : 1u;
}

// bijection on [0, total): three multiply-add-xorshift rounds on the enclosing power of two,
// cycle-walked back into range
__host__ __device__ __forceinline__ uint64_t permute_index(uint64_t x, uint64_t total, uint32_t seed)
{
	int bits = 1;
	while ((1ull << bits) < total) ++bits;
	const uint64_t mask = (1ull << bits) - 1;
	const int sh = (bits + 1) / 2;
	const uint64_t a0 = ((uint64_t)mix32(seed ^ 0x11111111u) << 1) | 1, c0 = mix32(seed ^ 0x22222222u);
	const uint64_t a1 = ((uint64_t)mix32(seed ^ 0x33333333u) << 1) | 1, c1 = mix32(seed ^ 0x44444444u);
	const uint64_t a2 = ((uint64_t)mix32(seed ^ 0x55555555u) << 1) | 1, c2 = mix32(seed ^ 0x66666666u);
	do {
		x = (x * a0 + c0) & mask;
		x ^= x >> sh;
		x = (x * a1 + c1) & mask;
		x ^= x >> sh;
		x = (x * a2 + c2) & mask;
		x ^= x >> sh;
	} while (x >= total);
	return x;
}

__global__ void k_generate(hjb_gen g, uint32_t *__restrict__ keys, uint32_t *__restrict__ vals)
{
	const double log_n1 = log((double)g.domain + 1.0);
	const uint64_t miss_domain = g.domain < (0xFFFFFFFEull - g.domain) ? g.domain : (0xFFFFFFFEull - g.domain);
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < g.tuples; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t j = g.first + i;
		uint32_t rank;
		if (g.kind == 0) {
			// unique keys: a seeded permutation of ranks [0, total)
			rank = (uint32_t)permute_index(j, g.total, g.order_seed);
		} else if (g.kind == 1) {
			// foreign keys: the first `domain` permuted positions cover every rank once, the
			// rest are uniform picks (cpra2.cpp:1639-1646), and the permutation is the shuffle
			const uint64_t t = permute_index(j, g.total, g.order_seed);
			if (t < g.domain) rank = (uint32_t)t;
			else rank = __umulhi(mix32((uint32_t)t ^ mix32((uint32_t)(t >> 32) + g.order_seed)), (uint32_t)g.domain);
		} else {
			// skewed: rank = floor((N+1)^u) - 1, the continuous inversion of a theta = 1 Zipf law
			// (theta != 1: power-law inversion); a `selectivity` fraction of tuples hit the build
			// side, the others take ranks past it
			const uint32_t h1 = mix32((uint32_t)j ^ mix32(g.order_seed + 0x9e3779b9u + (uint32_t)(j >> 32)));
			const uint32_t h2 = mix32(h1 ^ 0x85ebca6bu), h3 = mix32(h2 + 0xc2b2ae35u);
			const double u = ((double)h1 * 4294967296.0 + (double)h2 + 0.5) / 18446744073709551616.0;
			const bool hit = (double)h3 < g.selectivity * 4294967296.0;
			const double nn = hit ? (double)g.domain : (double)miss_domain;
			double r;
			if (fabs(g.theta - 1.0) < 1e-9) r = exp(u * (hit ? log_n1 : log(nn + 1.0))) - 1.0;
			else {
				const double e = 1.0 - g.theta;
				r = pow(u * (pow(nn + 1.0, e) - 1.0) + 1.0, 1.0 / e) - 1.0;
			}
			uint64_t ri = (uint64_t)r;
			if (ri >= (uint64_t)nn) ri = (uint64_t)nn - 1;
			rank = hit ? (uint32_t)ri : (uint32_t)(g.domain + ri);
		}
		const uint32_t key = key_of_rank(rank, g.seed);
		keys[i] = key;
		vals[i] = key * g.payload_factor;
	}
}

int launch_generate(const hjb_gen &g, uint32_t *keys, uint32_t *vals, cudaStream_t s)
{
	if (g.tuples == 0) return 0;
	uint64_t grid = (g.tuples + 255) / 256;
	if (grid > 148 * 32) grid = 148 * 32;
	k_generate<<<(uint32_t)grid, 256, 0, s>>>(g, keys, vals);
	return 1;
}

__global__ void __launch_bounds__(256)
k_column_sum(const uint32_t *__restrict__ col, uint64_t n, unsigned long long *__restrict__ out)
{
	__shared__ uint64_t s_part[8];
	uint64_t acc = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
		acc += col[i];
	acc = warp_sum_u64(acc);
	if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
	__syncthreads();
	if (threadIdx.x == 0) {
		uint64_t t = 0;
		for (int w = 0; w < 8; ++w) t += s_part[w];
		atomicAdd(out, (unsigned long long)t);
	}
}

// order-independent fingerprint of result rows (hjb_rows_fingerprint): out[0] += z, out[1] ^= z
__device__ __forceinline__ uint64_t splitmix64(uint64_t z)
{
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
k_rows_fingerprint(const uint32_t *__restrict__ k, const uint32_t *__restrict__ o, const uint32_t *__restrict__ iv, uint64_t n,
                   unsigned long long *__restrict__ out)
{
	__shared__ uint64_t s_sum[8], s_xor[8];
	uint64_t sum = 0, x = 0;
	for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
		const uint64_t z = splitmix64(((uint64_t)k[i] | ((uint64_t)o[i] << 32)) ^ splitmix64(iv[i]));
		sum += z;
		x ^= z;
	}
	sum = warp_sum_u64(sum);
#pragma unroll
	for (int off = 16; off; off >>= 1) x ^= __shfl_xor_sync(kFullMask, x, off);
	if ((threadIdx.x & 31) == 0) {
		s_sum[threadIdx.x >> 5] = sum;
		s_xor[threadIdx.x >> 5] = x;
	}
	__syncthreads();
	if (threadIdx.x == 0) {
		uint64_t a = 0, b = 0;
		for (int w = 0; w < 8; ++w) {
			a += s_sum[w];
			b ^= s_xor[w];
		}
		atomicAdd(&out[0], (unsigned long long)a);
		atomicXor(&out[1], (unsigned long long)b);
	}
}

int launch_rows_fingerprint(const uint32_t *k, const uint32_t *o, const uint32_t *iv, uint64_t n, unsigned long long *out_dev,
                            cudaStream_t s, int sms)
{
	cudaMemsetAsync(out_dev, 0, 16, s);
	uint64_t grid = (n + 255) / 256;
	if (grid > (uint64_t)sms * 8) grid = (uint64_t)sms * 8;
	if (grid == 0) grid = 1;
	k_rows_fingerprint<<<(uint32_t)grid, 256, 0, s>>>(k, o, iv, n, out_dev);
	return 1;
}

int launch_column_sum(const uint32_t *col, uint64_t n, unsigned long long *out_dev, cudaStream_t s, int sms)
{
	cudaMemsetAsync(out_dev, 0, 8, s);
	uint64_t grid = (n + 255) / 256;
	if (grid > (uint64_t)sms * 8) grid = (uint64_t)sms * 8;
	if (grid == 0) grid = 1;
	k_column_sum<<<(uint32_t)grid, 256, 0, s>>>(col, n, out_dev);
	return 1;
}

}  // namespace hjb
