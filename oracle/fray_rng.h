/*
 * fray_rng.h -- the counter-based random-number CONTRACT shared by every CPU-side checker.
 *
 * TEST INFRASTRUCTURE (oracle). The product (fray_b200/csrc/rng.cuh) carries its own, independently
 * written implementation of this same contract; tests/test_rng.py checks the two draw for draw.
 *
 * Why it exists: the reference draws from std::mt19937 instances keyed by thread id and hands
 * buckets to threads dynamically (src/random_generator.cpp:82-131, src/main.cpp:333-336), so its
 * streams are not reproducible even against itself. Same-seed parity therefore needs a generator
 * that is a pure function of (seed, pixel, sample, draw index) on BOTH sides (SURVEY.md section 0,
 * Appendix A.30).
 *
 * Contract
 *   block(seed; pixel, sample, branch; j) = Philox4x32-10(counter = {j, pixel, sample, branch},
 *                                                        key     = {seed, 0x46524159})
 *   draw i of a stream is word (i & 3) of block (i >> 2).
 *   pixel  = y * frameWidth + x,  sample = index of the sample inside the pixel (0..spp-1),
 *   branch = 0 for the stream that RendMT::entry would use for that sample (src/main.cpp:348-358);
 *            a Whitted secondary ray (reflection / refraction / glossy sample k) owns the stream
 *            branch' = fray_rng_child(branch, draws consumed by the parent at the spawn, k).
 *   Mappings of the reference's Random methods (src/random_generator.cpp:41-80):
 *     randfloat()  : 1 draw  -> (x >> 8) * 2^-24                      in [0,1)
 *     randdouble() : 2 draws -> (((hi << 32) | lo) >> 11) * 2^-53, lo drawn first, in [0,1)
 *     randint(a,b) : 1 draw  -> a + ((x * (b - a + 1)) >> 32)
 *     unitDiscSample: angle = randdouble()*2*PI, rad = sqrt(randdouble()), (sin*rad, cos*rad)
 *   The number of draws each reference call consumes equals what libstdc++ 13 consumes for
 *   mt19937, so the program-order draw layout is the reference's.
 */
#ifndef FRAY_RNG_H
#define FRAY_RNG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRAY_RNG_KEY1 0x46524159u /* "FRAY" */

static inline void fray_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	/* Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32, 10 rounds */
	const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int r = 0; r < 10; r++) {
		uint64_t p0 = (uint64_t) M0 * c0;
		uint64_t p1 = (uint64_t) M1 * c2;
		uint32_t n0 = (uint32_t) (p1 >> 32) ^ c1 ^ k0;
		uint32_t n1 = (uint32_t) p1;
		uint32_t n2 = (uint32_t) (p0 >> 32) ^ c3 ^ k1;
		uint32_t n3 = (uint32_t) p0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += W0; k1 += W1;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint32_t fray_rng_mix(uint32_t x)
{
	/* 32-bit finaliser (two multiply-xorshift rounds) */
	x ^= x >> 16; x *= 0x7FEB352Du;
	x ^= x >> 15; x *= 0x846CA68Bu;
	x ^= x >> 16;
	return x;
}

/* stream id of the k-th secondary ray spawned after the parent consumed `draws` numbers */
static inline uint32_t fray_rng_child(uint32_t branch, uint32_t draws, uint32_t k)
{
	return fray_rng_mix(branch ^ fray_rng_mix(draws * 0x9E3779B9u + k + 1u)) | 1u; /* never 0 */
}

typedef struct FrayRng {
	uint32_t seed, pixel, sample, branch;
	uint32_t count;      /* draws consumed so far */
	uint32_t cache[4];   /* block (count-1) >> 2, valid if count & 3 */
} FrayRng;

static inline void fray_rng_init(FrayRng* r, uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t branch)
{
	r->seed = seed; r->pixel = pixel; r->sample = sample; r->branch = branch; r->count = 0;
}

static inline uint32_t fray_rng_next(FrayRng* r)
{
	uint32_t lane = r->count & 3u;
	if (lane == 0) {
		uint32_t ctr[4] = { r->count >> 2, r->pixel, r->sample, r->branch };
		uint32_t key[2] = { r->seed, FRAY_RNG_KEY1 };
		fray_philox4x32_10(ctr, key, r->cache);
	}
	r->count++;
	return r->cache[lane];
}

static inline void fray_rng_skip(FrayRng* r, uint32_t n)
{
	uint32_t target = r->count + n;
	r->count = target & ~3u;              /* re-materialise the block that holds draw `target` */
	if (target & 3u) { (void) fray_rng_next(r); r->count = target; }
}

static inline float fray_rng_float(FrayRng* r) { return (float) (fray_rng_next(r) >> 8) * (1.0f / 16777216.0f); }

/* ONE draw per double (32 random bits): a Lambert path segment then consumes 8 draws -- two Philox blocks -- instead of 12 */
static inline double fray_rng_double(FrayRng* r)
{
	return (double) fray_rng_next(r) * (1.0 / 4294967296.0);
}

static inline int fray_rng_int(FrayRng* r, int a, int b)
{
	uint64_t n = (uint64_t) ((int64_t) b - (int64_t) a + 1);
	return a + (int) (((uint64_t) fray_rng_next(r) * n) >> 32);
}

#ifdef __cplusplus
}
#endif

#endif
