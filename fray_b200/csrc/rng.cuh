// rng.cuh -- counter-based per-pixel random numbers (product implementation).
//
// Replaces the reference's thread-id keyed std::mt19937 table (/root/reference/src/random_generator.cpp:31-131),
// whose streams are not reproducible. A stream is a pure function of (seed, pixel, sample, branch) and the
// index of the draw, so the GPU, the host and the CPU checkers see identical sample sequences no matter how
// work is scheduled. The stream layout and the mapping of Random::randfloat / randdouble / randint /
// unitDiscSample (src/random_generator.cpp:41-80) onto 32-bit draws is specified in DESIGN.md "RNG contract";
// tests/test_rng.py checks this file draw for draw against the independently written oracle/fray_rng.h.
//
// Usable from host code too (the .fray parser's randfloat()/randint() macros, src/scene.cpp:609-653).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FRAY_HD __host__ __device__ __forceinline__
// large, rarely executed or multiply-referenced routines: one out-of-line copy per kernel keeps code size and
// compile time in check (CSG evaluation, the direct-lighting loop shared by all Whitted shader levels)
#define FRAY_HD_COLD __host__ __device__ __noinline__
// hot routines: inlined in the fast-precision unit; the parity unit (render_fp64.cu) trades speed for compile time
#if defined(FRAY_PARITY_UNIT)
#define FRAY_HD_HOT __host__ __device__ __noinline__
#else
#define FRAY_HD_HOT __host__ __device__ __forceinline__
#endif
#else
#define FRAY_HD inline
#define FRAY_HD_COLD inline
#define FRAY_HD_HOT inline
#endif

namespace fray {

struct Philox4 { uint32_t x, y, z, w; };

FRAY_HD uint32_t mulhi32(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
	return __umulhi(a, b);
#else
	return (uint32_t) (((uint64_t) a * b) >> 32);
#endif
}

// Philox-4x32 with 10 rounds (Salmon, Moraes, Dror, Shaw, SC'11); key schedule bumps after every round.
FRAY_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1)
{
#if defined(__CUDACC__)
#pragma unroll
#endif
	for (int round = 0; round < 10; round++) {
		const uint32_t lo0 = 0xD2511F53u * c0, hi0 = mulhi32(0xD2511F53u, c0);
		const uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = mulhi32(0xCD9E8D57u, c2);
		c0 = hi1 ^ c1 ^ k0;
		c1 = lo1;
		c2 = hi0 ^ c3 ^ k1;
		c3 = lo0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	return Philox4{ c0, c1, c2, c3 };
}

FRAY_HD uint32_t rngMix(uint32_t v)
{
	v = (v ^ (v >> 16)) * 0x7FEB352Du;
	v = (v ^ (v >> 15)) * 0x846CA68Bu;
	return v ^ (v >> 16);
}

// stream id of the k-th secondary ray a Whitted shader spawns after `draws` numbers were consumed
FRAY_HD uint32_t rngChildBranch(uint32_t branch, uint32_t draws, uint32_t k)
{
	return rngMix(branch ^ rngMix(draws * 0x9E3779B9u + k + 1u)) | 1u;
}

// Ten rounds with the round keys read from a table (kernel parameters: a constant-bank operand of the LOP3, which saves
// the nine key additions per block of the seed-based form). keys[i] = seed + i * 0x9E3779B9; the second key word is constant.
FRAY_HD Philox4 philox4x32_10_keyed(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const uint32_t* keys)
{
	uint32_t k1 = 0x46524159u;
#if defined(__CUDACC__)
#pragma unroll
#endif
	for (int round = 0; round < 10; round++) {
		const uint32_t lo0 = 0xD2511F53u * c0, hi0 = mulhi32(0xD2511F53u, c0);
		const uint32_t lo1 = 0xCD9E8D57u * c2, hi1 = mulhi32(0xCD9E8D57u, c2);
		c0 = hi1 ^ c1 ^ keys[round];
		c1 = lo1;
		c2 = hi0 ^ c3 ^ k1;
		c3 = lo0;
		k1 += 0xBB67AE85u;
	}
	return Philox4{ c0, c1, c2, c3 };
}

FRAY_HD void philoxRoundKeys(uint32_t seed, uint32_t* keys)
{
	for (int i = 0; i < 10; i++) keys[i] = seed + (uint32_t) i * 0x9E3779B9u;
}

// How a stream generates its blocks: inline from the seed (host code) or inline with a round-key table (the render kernels:
// `keys` points at RenderParams::roundKeys, i.e. into the kernel's constant bank).
enum { FRAY_RNG_INLINE = 0, FRAY_RNG_KEYED = 2 };

template <int MODE> struct RngT {
	uint32_t seed, pixel, sample, branch;
	uint32_t count; // draws consumed
	Philox4 blk;    // block ((count - 1) >> 2) when count & 3
	const uint32_t* keys = nullptr; // FRAY_RNG_KEYED only

	FRAY_HD void init(uint32_t seed_, uint32_t pixel_, uint32_t sample_, uint32_t branch_)
	{
		seed = seed_; pixel = pixel_; sample = sample_; branch = branch_; count = 0;
	}
	// another stream of the same (seed, pixel, sample)
	FRAY_HD void initBranch(const RngT& parent, uint32_t branch_)
	{
		init(parent.seed, parent.pixel, parent.sample, branch_);
		keys = parent.keys;
	}
	FRAY_HD void ensure(uint32_t) {} // blocks are generated on demand
	FRAY_HD void refill()
	{
		if (MODE == FRAY_RNG_KEYED) blk = philox4x32_10_keyed(count >> 2, pixel, sample, branch, keys);
		else blk = philox4x32_10(count >> 2, pixel, sample, branch, seed, 0x46524159u);
	}
	FRAY_HD uint32_t next()
	{
		const uint32_t lane = count & 3u;
		if (lane == 0) refill();
		count++;
		return lane == 0 ? blk.x : (lane == 1 ? blk.y : (lane == 2 ? blk.z : blk.w));
	}
	// advance without generating (the discarded first spawnRay of pathtrace, src/main.cpp:219-224)
	FRAY_HD void skip(uint32_t n)
	{
		const uint32_t target = count + n;
		count = target & ~3u;
		if (target & 3u) refill();
		count = target;
	}
	FRAY_HD float randfloat() { return (float) (next() >> 8) * (1.0f / 16777216.0f); }
	// ONE draw per double (32 random bits), see DESIGN.md "RNG contract"
	FRAY_HD double randdouble() { return (double) next() * (1.0 / 4294967296.0); }
	// single-precision view of randdouble(): the same draw, its top 24 bits
	FRAY_HD float randdoubleAsFloat() { return (float) (next() >> 8) * (1.0f / 16777216.0f); }
	FRAY_HD int randint(int a, int b)
	{
		const uint32_t n = (uint32_t) (b - a + 1);
		return a + (int) mulhi32(next(), n);
	}
};

#if defined(__CUDACC__)
// The same streams for the path-tracing kernels, generated a few blocks ahead into a per-thread ring of 16 words in shared
// memory (layout [word][thread]: conflict-free whatever position each lane is at). A path segment consumes up to 8 draws
// (src/main.cpp:118-169, 219-236), always at data-dependent positions of the stream; with the blocks in registers every
// draw paid for a refill test and a four-way select (a tenth of the kernel's instructions). Here ensure(n) runs the ten
// rounds in ONE loop per call site (two trips for a Lambert segment) and a draw is an address computation and an LDS.
// Word p of the stream lives in ring[p & 15]; generating block g overwrites block g - 4, so ensure(n) requires n <= 12.
#define FRAY_RNG_RING_WORDS 16
struct RngRing {
	uint32_t pixel, sample, branch;
	uint32_t count;   // draws consumed
	uint32_t gen;     // blocks generated
	uint32_t base;    // shared-memory address of this thread's word 0
	const uint32_t* keys;

	__device__ __forceinline__ void attach(uint32_t sharedAddr, const uint32_t* keys_) { base = sharedAddr; keys = keys_; }
	__device__ __forceinline__ void init(uint32_t, uint32_t pixel_, uint32_t sample_, uint32_t branch_)
	{
		pixel = pixel_; sample = sample_; branch = branch_; count = 0; gen = 0;
	}
	// make the next n draws available (n <= 12)
	__device__ __forceinline__ void ensure(uint32_t n)
	{
		const uint32_t need = (count + n + 3u) >> 2;
		while (gen < need) {
			const Philox4 b = philox4x32_10_keyed(gen, pixel, sample, branch, keys);
			const uint32_t a = base + ((gen & 3u) << 2) * (blockDimRing() * 4u);
			asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(b.x));
			asm volatile("st.shared.u32 [%0], %1;" :: "r"(a + blockDimRing() * 4u), "r"(b.y));
			asm volatile("st.shared.u32 [%0], %1;" :: "r"(a + blockDimRing() * 8u), "r"(b.z));
			asm volatile("st.shared.u32 [%0], %1;" :: "r"(a + blockDimRing() * 12u), "r"(b.w));
			gen++;
		}
	}
	static __device__ __forceinline__ uint32_t blockDimRing() { return 128u; } // threads per CTA of the render kernels
	__device__ __forceinline__ uint32_t next()
	{
		uint32_t v;
		asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(base + (count & 15u) * (blockDimRing() * 4u)));
		count++;
		return v;
	}
	__device__ __forceinline__ void skip(uint32_t n) { count += n; }
	__device__ __forceinline__ float randfloat() { return (float) (next() >> 8) * (1.0f / 16777216.0f); }
	__device__ __forceinline__ double randdouble() { return (double) next() * (1.0 / 4294967296.0); }
	__device__ __forceinline__ float randdoubleAsFloat() { return (float) (next() >> 8) * (1.0f / 16777216.0f); }
	__device__ __forceinline__ int randint(int a, int b)
	{
		const uint32_t n = (uint32_t) (b - a + 1);
		return a + (int) mulhi32(next(), n);
	}
};
#endif

typedef RngT<FRAY_RNG_INLINE> Rng;

} // namespace fray
