// render_kernels.cuh -- the sm_100a kernels that replace RendMT::entry (/root/reference/src/main.cpp:323-371).
//
// Scheduling model ("persistent lanes, two-level work queue, path regeneration"):
//   * the frame is cut into 8x4 pixel tiles (row-major over the image); a call owns the tiles t with
//     t % taskStride == taskOffset (multi-GPU tile split) and a sample range [s0, s1) of every owned pixel (sample split);
//   * the samples of a pixel are cut into CHUNKS of consecutive samples. A WORK ITEM is (owned pixel, chunk); item ids are
//     laid out so that 32 consecutive ids are the 32 pixels of one tile for one chunk (coherent primary rays). Chunk sizes
//     DESCEND (guided self-scheduling, fray_gpu.cu: buildChunks): most of the frame runs in items of 8 paths -- little
//     bookkeeping, few chunk sums -- and the queue ends in items of one path, so that the lanes run dry within one path of each
//     other (~35 us), not one item. The leading chunks of equal size (the "bulk") are handed out tile by tile, all chunks of a
//     tile in a row -- what runs at any one time covers a small part of the screen, which the textures and meshes of zaphod
//     want (1.40 ms against 1.47 ms chunk by chunk) -- and the descending rest chunk by chunk, the smallest last;
//   * persistent CTAs (grid = SMs x resident CTAs). Every LANE owns one item at a time and is a state machine: whenever
//     its path (GI) or ray tree (Whitted) is finished it starts the next sample of its item, and when the item is used up it
//     writes the chunk's sum and takes the next item. Items come from a warp-level pool (a contiguous id range in uniform
//     registers; lanes that need one agree by __ballot_sync + popc) which lane 0 refills 32 ids at a time with ONE atomicAdd
//     on the global counter. No lane ever waits for another lane's path, pixel or tile; the only drain is at the very end
//     of the kernel, and it is bounded by one item (C is chosen so that a lane sees >= ~32 items per call). A refill decodes
//     its batch -- one tile, one chunk -- with all 32 lanes and parks the items in shared memory, so taking one is two LDS;
//   * a chunk's sum is accumulated in sample order by one lane; chunk sums go to a scratch buffer and combineKernel adds
//     them per pixel in chunk order (with one chunk per pixel the lane writes the pixel directly). A pixel is therefore a pure
//     function of (scene, seed, sample range, C): no float atomics, run-to-run bit-identical.
// The per-ray work itself is core.cuh.
#pragma once
#include <cuda_runtime.h>

#include "core.cuh"

namespace fray {

#define FRAY_TILE_W 8
#define FRAY_TILE_H 4
#define FRAY_POOL_BATCH 32 // item ids fetched per atomic
#define FRAY_MAX_CHUNKS 96  // chunks per pixel (entries of RenderParams::chunkStart)

struct RenderParams {
	int width, height;
	int spp;          // samples per pixel of the whole frame (divisor)
	int s0, s1;       // sample range rendered by this call
	uint32_t seed;
	uint32_t roundKeys[10]; // philoxRoundKeys(seed)
	int sumOnly;      // FRAY_FRAME_SUM
	int numChunks;    // chunks the sample range of a pixel is cut into
	int numBulk;      // the leading chunks of equal size, handed out tile by tile; the rest chunk by chunk (see below)
	unsigned int bulkItems; // numBulk * numOwnedTiles * 32
	float invNumBulk;
	unsigned chunkStart[FRAY_MAX_CHUNKS + 1]; // chunk k = samples s0 + [chunkStart[k], chunkStart[k + 1]) (descending sizes, see the head of this file)
	int tilesX;       // tiles per image row
	float invTilesX, invNumOwnedTiles; // reciprocals for divmodSmall
	int exactDiv;     // ranges too large for the float estimate: use integer division
	int numOwnedTiles;
	int taskStride, taskOffset; // owned tile j is tile j * taskStride + taskOffset of the frame
	unsigned int totalItems;    // numChunks * numOwnedTiles * 32
	float* out;       // [height][width][3]
	float* scratch;   // numChunks > 1: [chunk][owned pixel slot][3]
	unsigned long long* counters; // rays, primary, shadow
	unsigned int* workCounter;
	int* errorFlag;
};

// x / d and x % d for x < 2^22 without the ~25-instruction integer division: float reciprocal estimate (off by at most one
// in that range) and one correction step. The hosts sets invd = 1 / d and falls back to exact division for larger ranges.
__device__ __forceinline__ void divmodSmall(unsigned x, unsigned d, float invd, bool exact, unsigned& q, unsigned& r)
{
	if (exact) {
		q = x / d;
		r = x - q * d;
		return;
	}
	q = (unsigned) (__uint2float_rz(x) * invd);
	int rem = (int) (x - q * d);
	if (rem < 0) { q--; rem += (int) d; }
	else if (rem >= (int) d) { q++; rem -= (int) d; }
	r = (unsigned) rem;
}

// owned pixel slot (owned tile j, position in tile) -> pixel; false outside the image (border tiles)
__device__ __forceinline__ bool slotPixel(const RenderParams& p, unsigned slot, int& px, int& py)
{
	const unsigned j = slot >> 5, within = slot & 31u;
	const unsigned t = j * (unsigned) p.taskStride + (unsigned) p.taskOffset;
	unsigned ty, tx;
	divmodSmall(t, (unsigned) p.tilesX, p.invTilesX, p.exactDiv != 0, ty, tx);
	px = (int) tx * FRAY_TILE_W + (int) (within % FRAY_TILE_W);
	py = (int) ty * FRAY_TILE_H + (int) (within / FRAY_TILE_W);
	return px < p.width && py < p.height;
}

// Copies the flat polygon table (flat.cuh) into dynamic shared memory: 128-bit loads, once per CTA. The records of a scene
// are read millions of times per CTA (every ray walks all of them), always by all lanes at once at the same address.
template <typename R, int F> __device__ __forceinline__ FlatTab stageFlat(const DScene<R>& sc)
{
	extern __shared__ float4 flatSmem[];
	FlatTab ft;
	// layout: records (incl. shadow sets) | spheres | hexahedra | two-sided records | FlatInfo (records, lights, spheres, hexahedron faces, two-sided records)
	const int nVec = FRAY_FLAT_POLY_VEC * (sc.numFlatTotal + sc.numFlat2) + sc.numFlatSpheres + FRAY_HEX_VEC * sc.numFlatHex;
	ft.polys = flatSmem;
	ft.spheres = flatSmem + FRAY_FLAT_POLY_VEC * sc.numFlatTotal;
	ft.hexes = ft.spheres + sc.numFlatSpheres;
	ft.polys2 = ft.hexes + FRAY_HEX_VEC * sc.numFlatHex;
	ft.info = reinterpret_cast<const FlatInfo*>(flatSmem + nVec);
	if (F & FRAY_F_FLAT) {
		const int nInfo = (int) (sizeof(FlatInfo) / sizeof(float4)) * sc.numFlatInfo;
		const float4* gi = reinterpret_cast<const float4*>(sc.flatInfo);
		for (int i = threadIdx.x; i < nVec; i += blockDim.x) flatSmem[i] = sc.flatPolys[i];
		for (int i = threadIdx.x; i < nInfo; i += blockDim.x) flatSmem[nVec + i] = gi[i];
		__syncthreads();
	}
	return ft;
}

// resident CTAs per SM the Whitted kernels with a generic node loop (KD walks: latency-bound) are compiled for
#ifndef FRAY_GI_FLAT_CTAS
#define FRAY_GI_FLAT_CTAS 7 // the untextured table-only path tracers (cornell_box, smallpt)
#endif
#ifndef FRAY_WHITTED_FLAT_CTAS
#define FRAY_WHITTED_FLAT_CTAS 4 // the table-only Whitted kernel (zaphod)
#endif
#ifndef FRAY_WHITTED_KD_CTAS
#define FRAY_WHITTED_KD_CTAS 6
#endif

template <bool GI> struct RenderRng {
	typedef RngT<FRAY_RNG_KEYED> type;
	static __device__ __forceinline__ void attach(type& rng, uint32_t*, const uint32_t* keys) { rng.keys = keys; }
};
template <> struct RenderRng<true> {
	typedef RngRing type;
	static __device__ __forceinline__ void attach(type& rng, uint32_t* ring, const uint32_t* keys)
	{
		rng.attach((uint32_t) __cvta_generic_to_shared(ring + threadIdx.x), keys);
	}
};

template <typename R, bool GI, int F>
__global__ void __launch_bounds__(128, Num<R>::kExact ? 1 : (GI ? (((F & (FRAY_F_NODES | FRAY_F_TEX)) == 0) ? FRAY_GI_FLAT_CTAS : 6) : ((F & FRAY_F_NODES) ? FRAY_WHITTED_KD_CTAS : FRAY_WHITTED_FLAT_CTAS))) renderKernel(const DScene<R> sc, const RenderParams p)
{
	const FlatTab ft = stageFlat<R, F>(sc);
	const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
	const unsigned ltMask = (1u << lane) - 1u;
	__shared__ uint2 stagedItem[4][32]; // per warp, the decoded items of its pool batch: {px | py << 16 (all ones: outside the image), output index}
	__shared__ int stagedCur[4], stagedEnd[4]; // the samples of the batch's chunk
	const bool randomOffsets = sc.cam.dof || sc.gi;
	const bool stereo = (F & FRAY_F_LENS) && sc.cam.stereoSep > 0;

	RayCounters cnt = { 0, 0, 0 };
	WhittedState<R> ws; // only touched by the Whitted instantiation
	ws.sp = 0;
	ws.overflow = 0;
	ws.rootPending = false;

	// warp-level item pool [poolNext, poolEnd): identical in all lanes
	unsigned poolNext = 0, poolEnd = 0;
	bool exhausted = false; // the global counter ran past totalItems (warp-uniform)

	// lane state
	bool hasItem = false, active = false;
	int px = 0, py = 0;
	unsigned outIndex = 0; // where the chunk sum goes: pixel index (one chunk) or scratch index
	int cur = 0, end = 0;  // next sample / end of the chunk
	Col accum(0, 0, 0);    // sum over the finished samples of the chunk
	Col eyeCol(0, 0, 0);   // radiance of the path / ray tree in flight
	// stream of the sample in flight (branch 0): blocks a few draws ahead in a shared-memory ring for path tracing, on demand
	// in registers for Whitted (most of its rays draw nothing)
	__shared__ uint32_t rngRingSmem[GI ? FRAY_RNG_RING_WORDS * 128 : 1];
	typename RenderRng<GI>::type rng;
	RenderRng<GI>::attach(rng, rngRingSmem, p.roundKeys);
	PathState<R> ps;
	Ray<R> rightEye;       // stereo: the second ray is generated up front (src/main.cpp:307-308) and traced afterwards
	int eye = 0;

	for (;;) {
		// ---- lanes without a running sample: next sample of the item, or next item --------------------------------------
		if (!active && hasItem && cur == end) { // chunk finished: write its sum
			float* o;
			if (p.numChunks == 1) {
				o = p.out + 3 * (size_t) outIndex;
				if (!p.sumOnly) accum = accum / (float) p.spp; // `avg / samplesPerPixel`, src/main.cpp:360
			} else {
				o = p.scratch + 3 * (size_t) outIndex;
			}
			o[0] = accum.r; o[1] = accum.g; o[2] = accum.b;
			hasItem = false;
		}
		bool want = !active && !hasItem;
		unsigned wanters = __ballot_sync(0xffffffffu, want);
		while (wanters != 0 && !exhausted) {
			if (poolNext == poolEnd) { // refill: one atomic per FRAY_POOL_BATCH items per warp
				unsigned base = 0;
				if (lane == 0) base = atomicAdd(p.workCounter, (unsigned) FRAY_POOL_BATCH);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (base >= p.totalItems) { exhausted = true; break; }
				poolNext = base;
				poolEnd = base + (unsigned) FRAY_POOL_BATCH; // totalItems is a multiple of the batch
				// The batch is one (chunk, owned tile) pair: id = (chunk * numOwnedTiles + ownedTile) * 32 + within. Every lane decodes
				// item base + lane, once per 32 items and with all lanes busy, and parks it in shared memory for whoever takes it.
				unsigned chunk, j;
				if (base < p.bulkItems) { // bulk: id = (ownedTile * numBulk + chunk) * 32 + within
					divmodSmall(base >> 5, (unsigned) p.numBulk, p.invNumBulk, p.exactDiv != 0, j, chunk);
				} else {                  // rest: id = bulkItems + ((chunk - numBulk) * numOwnedTiles + ownedTile) * 32 + within
					divmodSmall((base - p.bulkItems) >> 5, (unsigned) p.numOwnedTiles, p.invNumOwnedTiles, p.exactDiv != 0, chunk, j);
					chunk += (unsigned) p.numBulk;
				}
				const unsigned slot = (j << 5) | lane;
				int qx, qy;
				uint2 st;
				st.x = slotPixel(p, slot, qx, qy) ? ((unsigned) qx | ((unsigned) qy << 16)) : 0xffffffffu;
				st.y = p.numChunks == 1 ? (unsigned) (qy * p.width + qx) : chunk * ((unsigned) p.numOwnedTiles * 32u) + slot;
				__syncwarp();
				stagedItem[warp][lane] = st;
				if (lane == 0) {
					stagedCur[warp] = p.s0 + (int) p.chunkStart[chunk];
					stagedEnd[warp] = p.s0 + (int) p.chunkStart[chunk + 1];
				}
				__syncwarp();
			}
			const unsigned avail = poolEnd - poolNext;
			const unsigned rank = __popc(wanters & ltMask);
			if (want && rank < avail) {
				const uint2 st = stagedItem[warp][(poolNext + rank) & 31u];
				if (st.x != 0xffffffffu) {
					hasItem = true;
					px = (int) (st.x & 0xffffu);
					py = (int) (st.x >> 16);
					cur = stagedCur[warp];
					end = stagedEnd[warp];
					accum = Col(0, 0, 0);
					outIndex = st.y;
				}
				want = false; // a slot outside the image is simply dropped; the lane asks again next round
			}
			poolNext += min(avail, (unsigned) __popc(wanters));
			wanters = __ballot_sync(0xffffffffu, want && !hasItem);
		}
		if (!active && hasItem && cur < end) { // start sample `cur`
			const int s = cur++;
			rng.init(p.seed, (uint32_t) (py * p.width + px), (uint32_t) s, 0);
			rng.ensure((F & FRAY_F_LENS) ? 6 : 2); // pixel offset (2 draws), one thin-lens sample (2 draws) per eye
			float ox, oy;
			sampleOffset(randomOffsets, s, rng, ox, oy);
			const R fx = (R) ((float) px + ox), fy = (R) ((float) py + oy);
			Ray<R> first = cameraRay(sc.cam, rng, fx, fy, stereo ? 1 : 0, (F & FRAY_F_LENS) != 0);
			if (stereo) rightEye = cameraRay(sc.cam, rng, fx, fy, 2, true);
			eye = 0;
			eyeCol = Col(0, 0, 0);
			cnt.primary++;
			if (GI) {
				ps.start = first.start; ps.dir = first.dir; ps.mult = Col(1, 1, 1); ps.depth = 0; ps.flags = 0; ps.origin = -1;
			} else {
				ws.setRoot(first.start, first.dir);
			}
			active = true;
		}
		if (exhausted && __all_sync(0xffffffffu, !active && !hasItem)) break;

		// ---- one path segment / one ray of the Whitted tree ---------------------------------------------------------------
		if (active) {
			bool finished;
			if constexpr (GI) {
				finished = !pathSegment<R, F>(sc, ft, ps, rng, eyeCol, cnt);
			} else {
				whittedPop<R, F>(sc, ft, rng, ws, eyeCol, cnt);
				finished = ws.done();
			}
			if (finished) {
				if (stereo) {
					if (sc.saturation != 1) eyeCol = adjustSaturation(eyeCol, sc.saturation);
					eyeCol = eyeCol * (eye == 0 ? loadCol(sc.cam.leftMask) : loadCol(sc.cam.rightMask));
				}
				accum = accum + eyeCol;
				if (stereo && eye == 0) {
					eye = 1;
					eyeCol = Col(0, 0, 0);
					cnt.primary++;
					if (GI) {
						ps.start = rightEye.start; ps.dir = rightEye.dir; ps.mult = Col(1, 1, 1); ps.depth = 0; ps.flags = 0; ps.origin = -1;
					} else {
						ws.setRoot(rightEye.start, rightEye.dir);
					}
				} else {
					active = false;
				}
			}
		}
	}

	// statistics: one atomic per warp
	unsigned long long nRays = cnt.rays, nPrimary = cnt.primary, nShadow = cnt.shadow;
	for (int m = 16; m > 0; m >>= 1) {
		nRays += __shfl_xor_sync(0xffffffffu, nRays, m);
		nPrimary += __shfl_xor_sync(0xffffffffu, nPrimary, m);
		nShadow += __shfl_xor_sync(0xffffffffu, nShadow, m);
	}
	if (lane == 0) {
		atomicAdd(p.counters + 0, nRays);
		atomicAdd(p.counters + 1, nPrimary);
		atomicAdd(p.counters + 2, nShadow);
		if (ws.overflow) atomicExch(p.errorFlag, 1);
	}
}

// FRAY_RENDER_AOV: primary hit ids of the un-jittered pin-hole ray through (x, y)
template <typename R, int F>
__global__ void __launch_bounds__(128) aovKernel(const DScene<R> sc, const RenderParams p)
{
	const FlatTab ft = stageFlat<R, F>(sc);
	const unsigned slots = (unsigned) p.numOwnedTiles * 32u;
	for (unsigned slot = blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += gridDim.x * blockDim.x) {
		int px, py;
		if (!slotPixel(p, slot, px, py)) continue;
		const Ray<R> ray = screenRay(sc.cam, (R) px, (R) py, 0);
		int node, light;
		Hit<R> h;
		closestHit<R, F>(sc, ft, ray, node, light, h);
		float* o = p.out + 3 * ((size_t) py * p.width + px);
		o[0] = light >= 0 ? (float) (-2 - light) : (float) node;
		o[1] = (light < 0 && node >= 0 && h.tri >= 0) ? (float) (h.tri - sc.meshes[h.mesh].firstTri) : -1.0f;
		o[2] = (float) h.dist;
	}
}

// FRAY_RENDER_PREPASS: the 16x16 preview of render(), src/main.cpp:376-391 -- one sample through the centre pixel of every
// 16x16 square (raytraceSinglePixel(cx, cy), no sub-pixel offset), painted over the whole square. One thread per square.
#define FRAY_PREPASS_SQUARE 16
template <typename R, int F>
__global__ void __launch_bounds__(128) prepassKernel(const DScene<R> sc, const RenderParams p)
{
	const FlatTab ft = stageFlat<R, F>(sc);
	const int sx = (p.width + FRAY_PREPASS_SQUARE - 1) / FRAY_PREPASS_SQUARE, sy = (p.height + FRAY_PREPASS_SQUARE - 1) / FRAY_PREPASS_SQUARE;
	WhittedState<R> ws;
	ws.sp = 0;
	ws.overflow = 0;
	ws.rootPending = false;
	RayCounters cnt = { 0, 0, 0 };
	for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < sx * sy; q += gridDim.x * blockDim.x) {
		const int x0 = (q % sx) * FRAY_PREPASS_SQUARE, y0 = (q / sx) * FRAY_PREPASS_SQUARE;
		const int x1 = min(p.width, x0 + FRAY_PREPASS_SQUARE), y1 = min(p.height, y0 + FRAY_PREPASS_SQUARE);
		const int cx = (x0 + x1) / 2, cy = (y0 + y1) / 2;
		const Col c = renderSample<R, F>(sc, ft, p.seed, cx, cy, p.width, 0, &ws, cnt, true);
		for (int y = y0; y < y1; y++)
			for (int x = x0; x < x1; x++) {
				float* o = p.out + 3 * ((size_t) y * p.width + x);
				o[0] = c.r; o[1] = c.g; o[2] = c.b;
			}
	}
	atomicAdd(p.counters + 0, (unsigned long long) cnt.rays);
	atomicAdd(p.counters + 1, (unsigned long long) cnt.primary);
	atomicAdd(p.counters + 2, (unsigned long long) cnt.shadow);
	if (ws.overflow) atomicExch(p.errorFlag, 1);
}

struct LaunchConfig {
	int gridBlocks; // persistent grid
	cudaStream_t stream;
};

template <typename R> inline size_t flatSmemBytes(const DScene<R>& sc)
{
	return ((size_t) (sc.numFlatTotal + sc.numFlat2) * FRAY_FLAT_POLY_VEC + sc.numFlatSpheres + (size_t) sc.numFlatHex * FRAY_HEX_VEC) * sizeof(float4) + (size_t) sc.numFlatInfo * sizeof(FlatInfo);
}

// one launch of the render (or AOV) kernel for precision R; defined in render_fp32.cu / render_fp64.cu
template <typename R> cudaError_t launchRender(const DScene<R>& sc, const RenderParams& p, int features, int mode, const LaunchConfig& cfg);
template <typename R> int renderOccupancy(const DScene<R>& sc, int features, bool gi); // resident CTAs of 128 threads per SM

template <typename R, int I> struct VariantDispatch {
	static constexpr int F = Variants<R>::mask(I);
	static cudaError_t launch(const DScene<R>& sc, const RenderParams& p, int need, int mode, const LaunchConfig& cfg)
	{
		if ((need & ~F) != 0) return VariantDispatch<R, I + 1>::launch(sc, p, need, mode, cfg);
		const size_t smem = (F & FRAY_F_FLAT) ? flatSmemBytes(sc) : 0;
		if (mode == FRAY_RENDER_AOV) aovKernel<R, F><<<cfg.gridBlocks, 128, smem, cfg.stream>>>(sc, p);
		else if (mode == FRAY_RENDER_PREPASS) prepassKernel<R, F><<<cfg.gridBlocks, 128, smem, cfg.stream>>>(sc, p);
		else if (sc.gi) renderKernel<R, true, F><<<cfg.gridBlocks, 128, smem, cfg.stream>>>(sc, p);
		else renderKernel<R, false, F><<<cfg.gridBlocks, 128, smem, cfg.stream>>>(sc, p);
		return cudaGetLastError();
	}
	static int occupancy(const DScene<R>& sc, int need, bool gi)
	{
		if ((need & ~F) != 0) return VariantDispatch<R, I + 1>::occupancy(sc, need, gi);
		const size_t smem = (F & FRAY_F_FLAT) ? flatSmemBytes(sc) : 0;
		int n = 0;
		if (gi) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, renderKernel<R, true, F>, 128, smem);
		else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, renderKernel<R, false, F>, 128, smem);
		return n;
	}
};
template <typename R> struct VariantDispatch<R, Variants<R>::count> {
	static cudaError_t launch(const DScene<R>&, const RenderParams&, int, int, const LaunchConfig&) { return cudaErrorInvalidValue; }
	static int occupancy(const DScene<R>&, int, bool) { return 0; }
};

#define FRAY_DEFINE_LAUNCHERS(R)                                                                                              \
	template <> cudaError_t launchRender<R>(const DScene<R>& sc, const RenderParams& p, int features, int mode, const LaunchConfig& cfg) \
	{                                                                                                                         \
		return VariantDispatch<R, 0>::launch(sc, p, features, mode, cfg);                                                     \
	}                                                                                                                         \
	template <> int renderOccupancy<R>(const DScene<R>& sc, int features, bool gi)                                            \
	{                                                                                                                         \
		return VariantDispatch<R, 0>::occupancy(sc, features, gi);                                                            \
	}

} // namespace fray
