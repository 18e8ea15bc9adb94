// wave_kernels.cuh -- the sm_100a kernels of the wavefront Whitted integrator (stage bodies: wave.cuh).
//
// One frame = for every level of the ray tree ("wave" w = Ray::depth, 0 .. maxTraceDepth, plus one for the right eye of a
// stereo pair): traceKernel -> shadeKernel -> shadowKernel, then resolveWaveKernel. All queues live in HBM:
//
//   rays of wave w   SoA, 56 bytes: {origin, pixel} {direction, meta} {weight, stream} {draws, sample}; wave 0 has no queue,
//                    its rays are regenerated from the sample index (camera ray + stream state) by both TRACE and SHADE
//   hits             24 bytes per ray: {t, l2, l3, node} {triangle, flat-table index}
//   lit records      56 bytes (+ 32 for Phong): one per Lambert / Phong evaluation, at a FIXED place: record k of ray i of the
//                    wave is slot i * litPerRay + k (litPerRay = the most any shader of the scene can produce), unused slots
//                    are marked. No atomics, and the 32 records a SHADOW warp works on are the hits of 32 neighbouring rays
//                    -- one 8x4 pixel tile in wave 0 -- which shoot at the same light sample at the same time. (Records
//                    appended through the sub-queues below came out interleaved in chunks of a few from tiles far apart:
//                    21 of 32 lanes busy in the shadow pass of boxed instead of 28.)
//   accumulators     3 x int64 per pixel, 2^-32 fixed point (wave.cuh)
//
// SHADE appends to the queues through warp-aggregated atomics (one atomicAdd per warp per push site, slots handed out by
// lane rank). A queue is FRAY_WAVE_REGIONS sub-queues with a counter each, 128 bytes apart, and a warp always appends to the
// sub-queue of its own number: atomics on ONE address retire at ~3 ns each on this chip, which for 8 M pushes per frame
// (260 k warp-level atomics) was most of the frame; spread over 64 addresses they disappear. Consumers see a dense index
// space again through the prefix sums of the sub-queue counts (64 words, scanned once per CTA into shared memory).
// (Measured and dropped: binning the queued rays into "touches the world box of a KD mesh" / "does not", so that a TRACE warp
// holds only tree walkers or none. hw9/dragon, where 11 of 32 lanes are busy in the walk of the 10 M glossy rays, did not move:
// 6.57 -> 6.60 ms. The walkers themselves diverge -- random directions off the floor, trees of very different depth along them.)
// Persistent grids; work is handed out dynamically in groups of 32 consecutive entries, one warp-level atomicAdd per group,
// on FRAY_WAVE_STRIPES interleaved counters for the same reason (stripe s owns the groups s, s + 16, s + 32, ...; a warp works
// on its own stripe until that is used up, then helps the next one). Static round-robin was 25-30 % slower here: the cost of a
// ray varies by 20x with what it hits, and the sum over a warp's 55 groups varies accordingly. The counts of the next wave
// never travel to the host: the kernels of all waves are enqueued back to back and read their counts from device memory.
#pragma once
#include "render_kernels.cuh"
#include "wave.cuh"

namespace fray {

#define FRAY_WAVE_MAX 16     // waves per frame the counters are laid out for (maxTraceDepth + 2 must fit)
#define FRAY_WAVE_REGIONS 64 // sub-queues per queue (power of two)
#define FRAY_WAVE_CTR_STRIDE 32 // words between two sub-queue counters (128 bytes: another L2 line, another slice)
#ifndef FRAY_WAVE_SHADE_CTAS
#define FRAY_WAVE_SHADE_CTAS 7 // resident CTAs per SM of the shade pass (72 registers): forest 4K -2 % against 6, which beat 4, 5 and 8
#endif
#define FRAY_WAVE_STRIPES 16    // interleaved work counters per launch (power of two)

// indices into WaveParams::ctr
enum {
	FRAY_WCTR_RAYS = 0,                                                                       // [w][region]: rays queued for wave w (w >= 1)
	FRAY_WCTR_WORK = FRAY_WAVE_MAX * FRAY_WAVE_REGIONS * FRAY_WAVE_CTR_STRIDE,                // [w][stage][stripe]: groups handed out by a launch
	FRAY_WCTR_OVERFLOW = FRAY_WCTR_WORK + FRAY_WAVE_MAX * 3 * FRAY_WAVE_STRIPES * FRAY_WAVE_CTR_STRIDE, // a sub-queue was full: the frame is incomplete, the host grows the queues and renders it again
	FRAY_WCTR_COUNT = FRAY_WCTR_OVERFLOW + 32
};

struct WaveParams {
	RenderParams rp;     // frame geometry, sample range, seed, round keys, statistics counters, output
	int wave;
	unsigned numPrimary; // ownedTiles * 32 * (s1 - s0): entries of wave 0
	int numSamples;      // s1 - s0
	float invNumSamples;
	unsigned rayCap, litCap; // entries per queue: FRAY_WAVE_REGIONS sub-queues of rayCap / FRAY_WAVE_REGIONS (litCap / ...) each
	unsigned stackOffset; // bytes from the start of dynamic shared memory to the KD stacks
	float4* rayO;        // [2][rayCap] each: rays of odd / even waves
	float4* rayD;
	float4* rayW;
	uint2* rayC;
	float4* hitA;        // [max(numPrimary, rayCap)]
	int2* hitB;
	int direct;          // single-owner frame (one wave, one sample per pixel): no accumulators, see waveAccumulate
	int litPerRay;       // lit-record slots per ray
	float4* litA;        // [litCap] {ip, pixel}; pixel < 0: unused slot
	float4* litB;        // {n, meta}
	float4* litC;        // {diffuse, stream}
	uint2* litD;         // {draws, sample}
	float4* litE;        // Phong: {specular, exponent}
	float4* litF;        // Phong: {ray direction, -}
	long long* acc;      // [height][width][3]
	unsigned* ctr;       // FRAY_WCTR_*
};

#if defined(FRAY_WAVE_IMPL) // the kernels themselves: render_wave.cu only (fray_gpu.cu needs the parameter block and the launcher)
// Queue entries are written once and read once, each by one thread: streaming accesses (evict-first), so that they do not push
// the KD nodes and triangle records -- which every warp walks again and again -- out of the L1 and the L2.
#if defined(FRAY_WAVE_NO_STREAM)
template <typename T> __device__ __forceinline__ T qld(const T* p) { return *p; }
template <typename T> __device__ __forceinline__ void qst(T* p, const T& v) { *p = v; }
#else
template <typename T> __device__ __forceinline__ T qld(const T* p) { return __ldcs(p); }
template <typename T> __device__ __forceinline__ void qst(T* p, const T& v) { __stcs(p, v); }
#endif
// KD short stack in shared memory: entry i of this thread at column[i * 128] (conflict-free: consecutive lanes, consecutive words)
struct KdStoreShared {
	uint2* column;
	__device__ __forceinline__ void put(unsigned i, int n, float t) { column[i * 128u] = make_uint2((unsigned) n, __float_as_uint(t)); }
	__device__ __forceinline__ void get(unsigned i, int& n, float& t) const
	{
		const uint2 e = column[i * 128u];
		n = (int) e.x;
		t = __uint_as_float(e.y);
	}
};
typedef KdShortStack<KdStoreShared, FRAY_KD_SHORT> KdStackShared;

__device__ __forceinline__ KdStackShared waveStack(const WaveParams& p)
{
	extern __shared__ float4 flatSmem[];
	KdStackShared s;
	s.st.column = reinterpret_cast<uint2*>(reinterpret_cast<char*>(flatSmem) + p.stackOffset) + threadIdx.x;
	s.reset();
	return s;
}

// one slot of a queue for every lane that calls this together (warp-aggregated atomic)
__device__ __forceinline__ unsigned wavePush(unsigned* counter)
{
	const unsigned mask = __activemask();
	const unsigned lane = threadIdx.x & 31u;
	const int leader = __ffs(mask) - 1;
	unsigned base = 0;
	if ((int) lane == leader) base = atomicAdd(counter, (unsigned) __popc(mask));
	base = __shfl_sync(mask, base, leader);
	return base + (unsigned) __popc(mask & ((1u << lane) - 1u));
}

// the sub-queue this warp appends to
__device__ __forceinline__ unsigned waveRegion() { return (blockIdx.x * 4u + (threadIdx.x >> 5)) & (unsigned) (FRAY_WAVE_REGIONS - 1); }

// Dense view of a queue: prefix sums of its sub-queue counts in shared memory (prefix[FRAY_WAVE_REGIONS] = total). Called by all
// threads of the CTA; `counters` = the queue's counter block, `cap` = entries per sub-queue.
__device__ __forceinline__ void waveScanQueue(const unsigned* counters, unsigned cap, unsigned* prefix)
{
	static_assert(FRAY_WAVE_REGIONS <= 128, "one thread of the 128 per counter");
	__shared__ unsigned counts[FRAY_WAVE_REGIONS];
	if (threadIdx.x < FRAY_WAVE_REGIONS) counts[threadIdx.x] = min(counters[threadIdx.x * FRAY_WAVE_CTR_STRIDE], cap); // all loads in flight at once
	__syncthreads();
	if (threadIdx.x == 0) {
		unsigned sum = 0;
		for (int r = 0; r < FRAY_WAVE_REGIONS; r++) {
			prefix[r] = sum;
			sum += counts[r];
		}
		prefix[FRAY_WAVE_REGIONS] = sum;
	}
	__syncthreads();
}

// dense index -> slot in the queue's storage
__device__ __forceinline__ unsigned waveSlot(const unsigned* prefix, unsigned cap, unsigned dense)
{
	unsigned r = 0;
#pragma unroll
	for (int step = FRAY_WAVE_REGIONS / 2; step > 0; step >>= 1)
		if (prefix[r + step] <= dense) r += step;
	return r * cap + (dense - prefix[r]);
}

// meta word of rays and lit records: origin node + 1 (16 bits) | depth (8 bits) | eye (2 bits) | Phong flag
__device__ __forceinline__ unsigned waveMeta(int origin, int depth, int eye, int phong) { return (unsigned) (origin + 1) | ((unsigned) depth << 16) | ((unsigned) eye << 24) | ((unsigned) phong << 26); }

// A frame with ONE wave (no shader reflects or refracts, no stereo) and ONE sample per pixel has one owner per pixel in each
// pass: SHADE stores what it knows (`waveStoreDirect`), SHADOW adds the light loops of the same ray (`waveAddDirect`, a plain
// read-modify-write). That saves the accumulator clear (200 MB at 3840x2160), the resolve pass and ~6 64-bit atomics per pixel.
__device__ __forceinline__ void waveStoreDirect(const WaveParams& p, int pixel, const Col& c)
{
	const float scale = p.rp.sumOnly ? 1.0f : 1.0f / (float) p.rp.spp;
	float* o = p.rp.out + 3 * (size_t) pixel;
	o[0] = c.r * scale; o[1] = c.g * scale; o[2] = c.b * scale;
}
__device__ __forceinline__ void waveAddDirect(const WaveParams& p, int pixel, const Col& c)
{
	const float scale = p.rp.sumOnly ? 1.0f : 1.0f / (float) p.rp.spp;
	float* o = p.rp.out + 3 * (size_t) pixel;
	o[0] += c.r * scale; o[1] += c.g * scale; o[2] += c.b * scale;
}

__device__ __forceinline__ void waveAccumulate(const WaveParams& p, int pixel, const Col& c)
{
	unsigned long long* a = reinterpret_cast<unsigned long long*>(p.acc) + 3 * (size_t) pixel;
	if (c.r != 0.0f) atomicAdd(a, (unsigned long long) waveFixed(c.r));
	if (c.g != 0.0f) atomicAdd(a + 1, (unsigned long long) waveFixed(c.g));
	if (c.b != 0.0f) atomicAdd(a + 2, (unsigned long long) waveFixed(c.b));
}

#define FRAY_FUSED_LIT 2 // lit records a ray may park for the fused shade + shadow kernel (scenes whose shaders produce more use three passes)

struct WaveSink {
	const DScene<float>& sc;
	const WaveParams& p;
	unsigned litBase; // first lit-record slot of the ray being shaded
	int litUsed;
	Col known;        // direct frames: what SHADE knows of the ray's radiance, stored once by flush()
	WaveLit* parked;  // fused kernel: the ray's lit records stay with the thread (nullptr: they go to the queue)
	__device__ __forceinline__ void add(const WaveRay& r, const Col& c)
	{
		if (p.direct) known = known + c;
		else waveAccumulate(p, r.pixel, waveEyeColor(sc, r.eye, c));
	}
	__device__ __forceinline__ void flush(int pixel)
	{
		if (p.direct) waveStoreDirect(p, pixel, known);
	}
	__device__ __forceinline__ void ray(const WaveRay& c)
	{
		const unsigned region = waveRegion(), cap = p.rayCap / FRAY_WAVE_REGIONS;
		const unsigned slot = wavePush(p.ctr + FRAY_WCTR_RAYS + ((p.wave + 1) * FRAY_WAVE_REGIONS + region) * FRAY_WAVE_CTR_STRIDE);
		if (slot >= cap) { p.ctr[FRAY_WCTR_OVERFLOW] = 1; return; }
		const size_t i = (size_t) ((p.wave + 1) & 1) * p.rayCap + (size_t) region * cap + slot;
		qst(p.rayO + i, make_float4(c.start.x, c.start.y, c.start.z, __int_as_float(c.pixel)));
		qst(p.rayD + i, make_float4(c.dir.x, c.dir.y, c.dir.z, __uint_as_float(waveMeta(c.origin, c.depth, c.eye, 0))));
		qst(p.rayW + i, make_float4(c.weight.r, c.weight.g, c.weight.b, __uint_as_float(c.branch)));
		qst(p.rayC + i, make_uint2(c.count, (unsigned) c.sample));
	}
	__device__ __forceinline__ void lit(const WaveLit& L)
	{
		if (litUsed >= p.litPerRay) { p.ctr[FRAY_WCTR_OVERFLOW] = 2; return; } // cannot happen: litPerRay is the scene's maximum
		if (parked) { parked[litUsed++] = L; return; }
		const unsigned slot = litBase + (unsigned) litUsed++;
		qst(p.litA + slot, make_float4(L.ip.x, L.ip.y, L.ip.z, __int_as_float(L.pixel)));
		qst(p.litB + slot, make_float4(L.n.x, L.n.y, L.n.z, __uint_as_float(waveMeta(L.origin, 0, L.eye, L.phong))));
		qst(p.litC + slot, make_float4(L.diffuse.r, L.diffuse.g, L.diffuse.b, __uint_as_float(L.branch)));
		qst(p.litD + slot, make_uint2(L.count, (unsigned) L.sample));
		if (L.phong) {
			qst(p.litE + slot, make_float4(L.specular.r, L.specular.g, L.specular.b, L.exponent));
			qst(p.litF + slot, make_float4(L.rayDir.x, L.rayDir.y, L.rayDir.z, 0.0f));
		}
	}
	// start the lit-record slots of ray i / close them (mark what was not used)
	__device__ __forceinline__ void begin(unsigned i) { litBase = i * (unsigned) p.litPerRay; litUsed = 0; known = Col(0, 0, 0); }
	__device__ __forceinline__ void end()
	{
		if (parked) return;
		for (int k = litUsed; k < p.litPerRay; k++) qst(p.litA + litBase + (unsigned) k, make_float4(0.0f, 0.0f, 0.0f, __int_as_float(-1)));
	}
};

// entry i of wave 0 -> pixel and sample; false for the slots of border tiles that lie outside the image
__device__ __forceinline__ bool wavePrimarySlot(const WaveParams& p, unsigned i, int& px, int& py, int& s)
{
	unsigned j, k;
	divmodSmall(i >> 5, (unsigned) p.numSamples, p.invNumSamples, p.rp.exactDiv != 0, j, k);
	s = p.rp.s0 + (int) k;
	return slotPixel(p.rp, (j << 5) | (i & 31u), px, py);
}

__device__ __forceinline__ void waveLoadRay(const WaveParams& p, unsigned slot, WaveRay& r)
{
	const size_t k = (size_t) (p.wave & 1) * p.rayCap + slot;
	const float4 o = qld(p.rayO + k), d = qld(p.rayD + k), w = qld(p.rayW + k);
	const uint2 c = qld(p.rayC + k);
	const unsigned meta = __float_as_uint(d.w);
	r.start = V3<float>(o.x, o.y, o.z);
	r.dir = V3<float>(d.x, d.y, d.z);
	r.weight = Col(w.x, w.y, w.z);
	r.pixel = __float_as_int(o.w);
	r.sample = (int) c.y;
	r.depth = (int) ((meta >> 16) & 0xffu);
	r.origin = (int) (meta & 0xffffu) - 1;
	r.eye = (int) ((meta >> 24) & 3u);
	r.branch = __float_as_uint(w.w);
	r.count = c.x;
}

// hands out the groups of 32 consecutive entries of a launch (see the head of this file)
struct WaveWork {
	unsigned* counters;
	unsigned long long numGroups;
	unsigned stripe, tried;
	__device__ __forceinline__ void init(const WaveParams& p, int stage, unsigned long long entries)
	{
		counters = p.ctr + FRAY_WCTR_WORK + (p.wave * 3 + stage) * FRAY_WAVE_STRIPES * FRAY_WAVE_CTR_STRIDE;
		numGroups = (entries + 31ull) >> 5;
		stripe = (blockIdx.x * 4u + (threadIdx.x >> 5)) & (unsigned) (FRAY_WAVE_STRIPES - 1);
		tried = numGroups == 0 ? (unsigned) FRAY_WAVE_STRIPES : 0u; // an empty wave: nothing to ask for
	}
	// all lanes of the warp call this together
	__device__ __forceinline__ bool next(unsigned long long& group)
	{
		const unsigned lane = threadIdx.x & 31u;
		while (tried < (unsigned) FRAY_WAVE_STRIPES) {
			unsigned k = 0;
			if (lane == 0) k = atomicAdd(counters + stripe * FRAY_WAVE_CTR_STRIDE, 1u);
			k = __shfl_sync(0xffffffffu, k, 0);
			group = (unsigned long long) k * FRAY_WAVE_STRIPES + stripe;
			if (group < numGroups) return true;
			// this stripe is used up: one look at all of them (a lane each) instead of finding out one atomic at a time
			unsigned seen = 0xffffffffu;
			if (lane < (unsigned) FRAY_WAVE_STRIPES) seen = *reinterpret_cast<volatile unsigned*>(counters + lane * FRAY_WAVE_CTR_STRIDE);
			const unsigned avail = __ballot_sync(0xffffffffu, lane < (unsigned) FRAY_WAVE_STRIPES && (unsigned long long) seen * FRAY_WAVE_STRIPES + lane < numGroups);
			if (avail == 0) break;
			const unsigned rotated = ((avail >> stripe) | (avail << (FRAY_WAVE_STRIPES - stripe))) & ((1u << FRAY_WAVE_STRIPES) - 1u);
			stripe = (stripe + (unsigned) __ffs(rotated) - 1u) & (unsigned) (FRAY_WAVE_STRIPES - 1);
			tried++; // (bounds the loop; stripes only ever run out)
		}
		tried = (unsigned) FRAY_WAVE_STRIPES;
		return false;
	}
};

// entries of this wave's ray queue (wave 0: the primary samples); fills `prefix` for waves >= 1
__device__ __forceinline__ unsigned waveRayCount(const WaveParams& p, unsigned* prefix)
{
	if (p.wave == 0) return p.numPrimary;
	waveScanQueue(p.ctr + FRAY_WCTR_RAYS + p.wave * FRAY_WAVE_REGIONS * FRAY_WAVE_CTR_STRIDE, p.rayCap / FRAY_WAVE_REGIONS, prefix);
	return prefix[FRAY_WAVE_REGIONS];
}

// ---- TRACE -----------------------------------------------------------------------------------------------------------------
#ifndef FRAY_WAVE_TRACE_CTAS
#define FRAY_WAVE_TRACE_CTAS 8
#endif
template <int F>
__global__ void __launch_bounds__(128, FRAY_WAVE_TRACE_CTAS) waveTraceKernel(const DScene<float> sc, const WaveParams p)
{
	const FlatTab ft = stageFlat<float, F>(sc);
	KdStackShared stk = waveStack(p);
	__shared__ unsigned prefix[FRAY_WAVE_REGIONS + 1];
	const unsigned n = waveRayCount(p, prefix);
	const bool randomOffsets = sc.cam.dof || sc.gi;
	unsigned traced = 0;
	WaveWork work;
	work.init(p, 0, n);
	for (unsigned long long group; work.next(group);) {
		const unsigned i = (unsigned) group * 32u + (threadIdx.x & 31u);
		if (i >= n) continue;
		Ray<float> ray;
		int origin = -1;
		if (p.wave == 0) {
			int px, py, s;
			if (!wavePrimarySlot(p, i, px, py, s)) continue;
			WaveRay l, r;
			bool stereo;
			wavePrimary<F>(sc, p.rp.roundKeys, p.rp.seed, px, py, p.rp.width, s, randomOffsets, l, r, stereo);
			ray.start = l.start;
			ray.dir = l.dir;
		} else {
			const size_t k = (size_t) (p.wave & 1) * p.rayCap + waveSlot(prefix, p.rayCap / FRAY_WAVE_REGIONS, i);
			const float4 o = qld(p.rayO + k), d = qld(p.rayD + k);
			ray.start = V3<float>(o.x, o.y, o.z);
			ray.dir = V3<float>(d.x, d.y, d.z);
			origin = (int) (__float_as_uint(d.w) & 0xffffu) - 1;
		}
		WaveHit wh;
		waveClosest<F>(sc, ft, ray, origin, stk, wh);
		qst(p.hitA + i, make_float4(wh.t, wh.l2, wh.l3, __int_as_float(wh.node)));
		qst(p.hitB + i, make_int2(wh.tri, wh.flat));
		traced++;
	}
	unsigned long long total = traced;
	for (int m = 16; m > 0; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
	if ((threadIdx.x & 31u) == 0 && total) atomicAdd(p.rp.counters + 0, total);
}

// ---- SHADE -----------------------------------------------------------------------------------------------------------------
// FUSED: scenes where a lit record means a shadow ray or two (point lights) and a ray produces at most FRAY_FUSED_LIT records.
// The records never leave the thread: when the ray is shaded its light loops run right here (the very code of the shadow pass,
// summed in the same order, delivered by the same operations: the frame is bit-identical to the three-pass frame) and there is
// no shadow pass. At 3840x2160 forest's records are 0.93 GB of DRAM traffic per frame, written once to be read once -- and
// that round trip turned out CHEAPER than running the walks inside the shade kernel (80 registers, 6 resident CTAs, 6400 + 2800
// instructions): forest 4K 1.51 -> 1.66 ms, with AA 6.93 -> 7.51 ms, dragon 4.74 -> 4.78 ms. Kept as a measured alternative
// (FRAY_GPU_FUSE=1); three passes are the default.
template <int F, bool FUSED>
__global__ void __launch_bounds__(128, FRAY_WAVE_SHADE_CTAS) waveShadeKernel(const DScene<float> sc, const WaveParams p)
{
	const FlatTab ft = stageFlat<float, F>(sc);
	__shared__ unsigned prefix[FRAY_WAVE_REGIONS + 1];
	const unsigned n = waveRayCount(p, prefix);
	const bool randomOffsets = sc.cam.dof || sc.gi;
	WaveLit parkedLit[FUSED ? FRAY_FUSED_LIT : 1];
	WaveSink sink{ sc, p, 0u, 0, Col(0, 0, 0), FUSED ? parkedLit : nullptr };
	KdStackShared stk;
	if constexpr (FUSED) stk = waveStack(p);
	unsigned traced = 0;
	unsigned primaries = 0;
	WaveWork work;
	work.init(p, 1, n);
	for (unsigned long long group; work.next(group);) {
		const unsigned i = (unsigned) group * 32u + (threadIdx.x & 31u);
		if (i >= n) continue;
		WaveRay r, right;
		bool stereo = false;
		sink.begin(i);
		if (p.wave == 0) {
			int px, py, s;
			if (!wavePrimarySlot(p, i, px, py, s)) { sink.end(); continue; }
			wavePrimary<F>(sc, p.rp.roundKeys, p.rp.seed, px, py, p.rp.width, s, randomOffsets, r, right, stereo);
			primaries += stereo ? 2u : 1u;
		} else {
			waveLoadRay(p, waveSlot(prefix, p.rayCap / FRAY_WAVE_REGIONS, i), r);
		}
		const float4 ha = qld(p.hitA + i);
		const int2 hb = qld(p.hitB + i);
		WaveHit wh;
		wh.t = ha.x; wh.l2 = ha.y; wh.l3 = ha.z; wh.node = __float_as_int(ha.w); wh.tri = hb.x; wh.flat = hb.y;
		uint32_t count = r.count;
		waveShade<F>(sc, ft, r, wh, p.rp.roundKeys, p.rp.seed, count, sink);
		sink.end();
		sink.flush(r.pixel);
		if constexpr (FUSED) {
			if (sink.litUsed > 0) { // what waveShadowKernel does with the ray's records
				Col sum(0, 0, 0);
				for (int k = 0; k < sink.litUsed; k++)
					sum = sum + waveEyeColor(sc, parkedLit[k].eye, waveLightLoop<F>(sc, ft, parkedLit[k], p.rp.roundKeys, p.rp.seed, stk, traced));
				if (p.direct) waveAddDirect(p, r.pixel, sum);
				else waveAccumulate(p, r.pixel, sum);
			}
		}
		if (p.wave == 0 && stereo) { // the right eye goes on in the stream where the left eye's light loops stopped
			right.count = count;
			sink.ray(right);
		}
	}
	unsigned long long total = primaries, shadows = traced;
	for (int m = 16; m > 0; m >>= 1) {
		total += __shfl_xor_sync(0xffffffffu, total, m);
		shadows += __shfl_xor_sync(0xffffffffu, shadows, m);
	}
	if ((threadIdx.x & 31u) == 0 && total) atomicAdd(p.rp.counters + 1, total);
	if (FUSED && (threadIdx.x & 31u) == 0 && shadows) {
		atomicAdd(p.rp.counters + 0, shadows);
		atomicAdd(p.rp.counters + 2, shadows);
	}
}

// ---- SHADOW ----------------------------------------------------------------------------------------------------------------
// one thread per lit record: its whole light loop (wave.cuh, waveLightLoop), one term delivered to the pixel's accumulator.
// Compiled for CTAS resident CTAs per SM: 8 (64 registers, ~260 bytes of spills inside the light loop) where a record means one
// or two shadow rays and occupancy hides the record fetch from DRAM (forest and dragon lose 2-4 % with 6), and
// FRAY_WAVE_SHADOW_HEAVY_CTAS = 3 (168 registers, nothing spilled) where it means many (area lights): 32 walks per record out of
// the L1 are arithmetic and L1 latency, which twelve warps per SM cover, and every spilled value is reloaded 32 times.
#ifndef FRAY_WAVE_SHADOW_HEAVY_CTAS
#define FRAY_WAVE_SHADOW_HEAVY_CTAS 3 // resident CTAs per SM the shadow pass of area-light scenes is compiled for: boxed 2.00 / 1.89 / 1.77 / 1.68 / 1.66 / 1.63 ms with 7 / 6 / 5 / 4 / 3 / 2
#endif
#ifndef FRAY_WAVE_SHADOW_MANY
#define FRAY_WAVE_SHADOW_MANY 8 // shadow rays per lit record from which the 6-CTA build is used
#endif
template <int F, int CTAS>
__global__ void __launch_bounds__(128, CTAS) waveShadowKernel(const DScene<float> sc, const WaveParams p)
{
	const FlatTab ft = stageFlat<float, F>(sc);
	KdStackShared stk = waveStack(p);
	__shared__ unsigned prefix[FRAY_WAVE_REGIONS + 1];
	const unsigned n = waveRayCount(p, prefix); // one thread per ray of the wave: the light loops of all its lit records
	const unsigned lane = threadIdx.x & 31u;
	unsigned traced = 0;
	WaveWork work;
	work.init(p, 2, n);
	for (unsigned long long group; work.next(group);) {
		const unsigned ray = (unsigned) group * 32u + lane;
		if (ray >= n) continue;
		Col sum(0, 0, 0);
		int pixel = -1;
		for (int k = 0; k < p.litPerRay; k++) {
			const size_t slot = (size_t) ray * (unsigned) p.litPerRay + (unsigned) k;
			const float4 a = qld(p.litA + slot);
			if (__float_as_int(a.w) < 0) break; // unused slot: a ray fills its slots from the front
			const float4 nb = qld(p.litB + slot), dc = qld(p.litC + slot);
			const uint2 cs = qld(p.litD + slot);
			const unsigned meta = __float_as_uint(nb.w);
			WaveLit L;
			L.ip = V3<float>(a.x, a.y, a.z);
			L.n = V3<float>(nb.x, nb.y, nb.z);
			L.diffuse = Col(dc.x, dc.y, dc.z);
			L.phong = (int) ((meta >> 26) & 1u);
			L.pixel = pixel = __float_as_int(a.w);
			L.sample = (int) cs.y;
			L.origin = (int) (meta & 0xffffu) - 1;
			L.eye = (int) ((meta >> 24) & 3u);
			L.branch = __float_as_uint(dc.w);
			L.count = cs.x;
			if (L.phong) {
				const float4 sp = qld(p.litE + slot), rd = qld(p.litF + slot);
				L.specular = Col(sp.x, sp.y, sp.z);
				L.exponent = sp.w;
				L.rayDir = V3<float>(rd.x, rd.y, rd.z);
			} else {
				L.specular = Col(0, 0, 0);
				L.exponent = 1;
				L.rayDir = V3<float>(0, 0, 1);
			}
			sum = sum + waveEyeColor(sc, L.eye, waveLightLoop<F>(sc, ft, L, p.rp.roundKeys, p.rp.seed, stk, traced));
		}
		if (pixel < 0) continue;
		if (p.direct) waveAddDirect(p, pixel, sum);
		else waveAccumulate(p, pixel, sum);
	}
	unsigned long long total = traced;
	for (int m = 16; m > 0; m >>= 1) total += __shfl_xor_sync(0xffffffffu, total, m);
	if (lane == 0 && total) {
		atomicAdd(p.rp.counters + 0, total);
		atomicAdd(p.rp.counters + 2, total);
	}
}

// accumulators -> frame: `avg / samplesPerPixel` (src/main.cpp:360) or the plain sum (FRAY_FRAME_SUM), owned pixels only
__global__ void resolveWaveKernel(const WaveParams p)
{
	const unsigned slots = (unsigned) p.rp.numOwnedTiles * 32u;
	for (unsigned slot = blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += gridDim.x * blockDim.x) {
		int px, py;
		if (!slotPixel(p.rp, slot, px, py)) continue;
		const size_t pix = (size_t) py * p.rp.width + px;
		const long long* a = p.acc + 3 * pix;
		float* o = p.rp.out + 3 * pix;
		for (int k = 0; k < 3; k++) {
			const float v = (float) ((double) a[k] * (1.0 / 4294967296.0));
			o[k] = p.rp.sumOnly ? v : v / (float) p.rp.spp;
		}
	}
}

#endif // FRAY_WAVE_IMPL

// ---- launch ----------------------------------------------------------------------------------------------------------------
// The kernel variants of the wavefront path: scenes with a generic node loop and no CSG (CSG evaluation keeps 6 KB of hit
// lists per thread: those scenes stay on the megakernel).
struct WaveVariants {
	static constexpr int kPlain = Variants<float>::mask(3);
	static constexpr int kTextured = Variants<float>::mask(4);
	static bool covers(int features) { return (features & FRAY_F_NODES) && !(features & FRAY_F_CSG) && ((features & ~kPlain) == 0 || (features & ~kTextured) == 0); }
};

struct WaveLaunch {
	int numSMs;
	int waves;          // waves to enqueue
	cudaStream_t stream;
	int occTrace, occShade, occShadow; // resident CTAs per SM (0: query)
	bool fused;         // out: the light loops ran inside the shade pass (two launches per wave)
};

cudaError_t launchWaveFrame(const DScene<float>& sc, WaveParams p, int features, WaveLaunch& cfg);

#if defined(FRAY_WAVE_IMPL)
template <int F> cudaError_t launchWaveFrameT(const DScene<float>& sc, WaveParams p, WaveLaunch& cfg)
{
	const size_t flat = (F & FRAY_F_FLAT) ? ((flatSmemBytes(sc) + 15) & ~(size_t) 15) : 0;
	const size_t stack = (size_t) FRAY_KD_SHORT * 128 * sizeof(uint2);
	p.stackOffset = (unsigned) flat;
	// shade + shadow in one kernel where a record means few shadow rays: OFF unless FRAY_GPU_FUSE=1 (measured, see waveShadeKernel)
	static const bool fuse = getenv("FRAY_GPU_FUSE") != nullptr;
	const bool fused = cfg.fused = fuse && sc.numLights > 0 && sc.lightSamples < FRAY_WAVE_SHADOW_MANY && p.litPerRay <= FRAY_FUSED_LIT;
	if (cfg.occTrace <= 0) {
		cudaFuncSetAttribute(waveTraceKernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (flat + stack));
		cudaFuncSetAttribute(waveShadowKernel<F, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (flat + stack));
		cudaFuncSetAttribute(waveShadowKernel<F, FRAY_WAVE_SHADOW_HEAVY_CTAS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (flat + stack));
		cudaFuncSetAttribute(waveShadeKernel<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) flat);
		cudaFuncSetAttribute(waveShadeKernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) (flat + stack));
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.occTrace, waveTraceKernel<F>, 128, flat + stack);
		if (fused) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.occShade, waveShadeKernel<F, true>, 128, flat + stack);
		else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.occShade, waveShadeKernel<F, false>, 128, flat);
		if (sc.lightSamples >= FRAY_WAVE_SHADOW_MANY) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.occShadow, waveShadowKernel<F, FRAY_WAVE_SHADOW_HEAVY_CTAS>, 128, flat + stack);
		else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&cfg.occShadow, waveShadowKernel<F, 8>, 128, flat + stack);
		if (cfg.occTrace < 1 || cfg.occShade < 1 || cfg.occShadow < 1) return cudaErrorLaunchOutOfResources;
	}
	for (int w = 0; w < cfg.waves; w++) {
		p.wave = w;
		waveTraceKernel<F><<<cfg.numSMs * cfg.occTrace, 128, flat + stack, cfg.stream>>>(sc, p);
		if (fused) {
			waveShadeKernel<F, true><<<cfg.numSMs * cfg.occShade, 128, flat + stack, cfg.stream>>>(sc, p);
			continue;
		}
		waveShadeKernel<F, false><<<cfg.numSMs * cfg.occShade, 128, flat, cfg.stream>>>(sc, p);
		if (sc.numLights > 0) {
			if (sc.lightSamples >= FRAY_WAVE_SHADOW_MANY) waveShadowKernel<F, FRAY_WAVE_SHADOW_HEAVY_CTAS><<<cfg.numSMs * cfg.occShadow, 128, flat + stack, cfg.stream>>>(sc, p);
			else waveShadowKernel<F, 8><<<cfg.numSMs * cfg.occShadow, 128, flat + stack, cfg.stream>>>(sc, p);
		}
	}
	if (!p.direct) resolveWaveKernel<<<cfg.numSMs * 4, 256, 0, cfg.stream>>>(p);
	return cudaGetLastError();
}

#endif

} // namespace fray
