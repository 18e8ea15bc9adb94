// render_kernels.cuh -- the sm_100a kernels that replace RendMT::entry (/root/reference/src/main.cpp:323-371).
//
// Scheduling model ("persistent warps with path regeneration"):
//   * the frame is cut into the reference's 48x48 buckets (src/sdl.cpp:243-262); a call owns a subset of them
//     (multi-GPU tile split) and a sample range [s0, s1) of every owned pixel (multi-GPU sample split);
//   * a WARP TASK is a small pixel tile of one bucket: 32/G pixels, G lanes per pixel (G = 1..32, a power of two chosen
//     from the number of samples). Persistent CTAs (grid = SMs x resident CTAs) pull warp tasks from one global
//     counter, one atomic per task (warp-aggregated: lane 0 fetches, __shfl broadcasts);
//   * inside a task every lane runs a state machine: whenever its path (GI) or ray tree (Whitted) is finished it takes
//     the next unstarted sample of its pixel - the lanes of a pixel agree on who takes which sample with one
//     __ballot_sync + popc per iteration - so lanes stay busy although paths end after different numbers of bounces
//     (divergence per bounce is bounded by one segment instead of one whole path);
//   * per-lane partial sums are combined with a fixed-order __shfl_xor tree, so a pixel is a pure function of
//     (scene, seed, sample range): no float atomics, run-to-run bit-identical.
// The per-ray work itself is core.cuh.
#pragma once
#include <cuda_runtime.h>

#include "core.cuh"

namespace fray {

struct RenderParams {
	int width, height;
	int spp;          // samples per pixel of the whole frame (divisor)
	int s0, s1;       // sample range rendered by this call
	uint32_t seed;
	int sumOnly;      // FRAY_FRAME_SUM
	int lanesPerPixel;// G
	int tileW, tileH; // pixel tile of one warp task, tileW * tileH * G == 32
	int numBuckets;   // owned buckets
	const int4* buckets; // x0, y0, w, h
	int totalTasks;   // warp tasks of the whole frame
	int taskStride, taskOffset; // this call owns the tasks t with t % taskStride == taskOffset (multi-GPU tile split)
	float* out;       // [height][width][3]
	unsigned long long* counters; // rays, primary, shadow
	unsigned int* workCounter;
	int* errorFlag;
};

#define FRAY_BUCKET 48

__device__ __forceinline__ Col shflXorCol(const Col& c, int m)
{
	return Col(__shfl_xor_sync(0xffffffffu, c.r, m), __shfl_xor_sync(0xffffffffu, c.g, m), __shfl_xor_sync(0xffffffffu, c.b, m));
}

// Copies the flat polygon table (flat.cuh) into dynamic shared memory: 128-bit loads, once per CTA. The records of a scene
// are read millions of times per CTA (every ray walks all of them), always by all lanes at once at the same address.
template <typename R, int F> __device__ __forceinline__ FlatTab stageFlat(const DScene<R>& sc)
{
	extern __shared__ float4 flatSmem[];
	FlatTab ft;
	ft.polys = flatSmem;
	ft.info = reinterpret_cast<const FlatInfo*>(flatSmem + FRAY_FLAT_POLY_VEC * sc.numFlatAll);
	if (F & FRAY_F_FLAT) {
		const int nPoly = FRAY_FLAT_POLY_VEC * sc.numFlatAll, nInfo = (int) (sizeof(FlatInfo) / sizeof(float4)) * sc.numFlatAll;
		const float4* gi = reinterpret_cast<const float4*>(sc.flatInfo);
		for (int i = threadIdx.x; i < nPoly; i += blockDim.x) flatSmem[i] = sc.flatPolys[i];
		for (int i = threadIdx.x; i < nInfo; i += blockDim.x) flatSmem[nPoly + i] = gi[i];
		__syncthreads();
	}
	return ft;
}

template <typename R, bool GI, int F>
__global__ void __launch_bounds__(128, (F == FRAY_F_FLAT && GI) ? 6 : 1) renderKernel(const DScene<R> sc, const RenderParams p)
{
	const FlatTab ft = stageFlat<R, F>(sc);
	const unsigned lane = threadIdx.x & 31u;
	const int G = p.lanesPerPixel;
	const unsigned groupMask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (lane & ~(unsigned) (G - 1));
	const unsigned ltMask = (1u << lane) - 1u;
	const int pixInWarp = (int) lane / G;
	const int tilesX = FRAY_BUCKET / p.tileW, tilesPerBucket = tilesX * (FRAY_BUCKET / p.tileH);
	const bool randomOffsets = sc.cam.dof || sc.gi;
	const bool stereo = sc.cam.stereoSep > 0;

	RayCounters cnt = { 0, 0, 0 };
	WhittedState<R> ws; // only touched by the Whitted instantiation
	ws.sp = 0;
	ws.overflow = 0;

	for (;;) {
		unsigned task = 0;
		if (lane == 0) task = atomicAdd(p.workCounter, 1u);
		task = __shfl_sync(0xffffffffu, task, 0) * (unsigned) p.taskStride + (unsigned) p.taskOffset;
		if (task >= (unsigned) p.totalTasks) break;

		const int4 bk = p.buckets[task / tilesPerBucket];
		const int tile = task % tilesPerBucket;
		const int lx = (tile % tilesX) * p.tileW + pixInWarp % p.tileW;
		const int ly = (tile / tilesX) * p.tileH + pixInWarp / p.tileW;
		const bool valid = lx < bk.z && ly < bk.w;
		const int px = bk.x + lx, py = bk.y + ly;

		enum { IDLE, ACTIVE, DONE };
		int state = valid ? IDLE : DONE;
		int next = p.s0;     // first unstarted sample of this lane's pixel (identical in all lanes of the group)
		Col accum(0, 0, 0);  // sum over the samples this lane finished
		Col eyeCol(0, 0, 0); // radiance of the path / ray tree in flight
		Rng rng;             // stream of the sample in flight (branch 0)
		PathState<R> ps;
		Ray<R> rightEye;     // stereo: the second ray is generated up front (src/main.cpp:307-308) and traced afterwards
		int eye = 0;

		for (;;) {
			const bool need = state == IDLE;
			const unsigned ballot = __ballot_sync(0xffffffffu, need);
			if (need) {
				const int s = next + __popc(ballot & groupMask & ltMask);
				if (s < p.s1) {
					rng.init(p.seed, (uint32_t) (py * p.width + px), (uint32_t) s, 0);
					float ox, oy;
					sampleOffset(randomOffsets, s, rng, ox, oy);
					const R fx = (R) ((float) px + ox), fy = (R) ((float) py + oy);
					Ray<R> first = cameraRay(sc.cam, rng, fx, fy, stereo ? 1 : 0);
					if (stereo) rightEye = cameraRay(sc.cam, rng, fx, fy, 2);
					eye = 0;
					eyeCol = Col(0, 0, 0);
					cnt.primary++;
					if (GI) {
						ps.start = first.start; ps.dir = first.dir; ps.mult = Col(1, 1, 1); ps.depth = 0; ps.flags = 0;
					} else {
						RayTask<R> root;
						root.start = first.start; root.dir = first.dir; root.weight = Col(1, 1, 1);
						root.depth = 0; root.branch = 0; root.count = 0;
						ws.stack[0] = root;
						ws.sp = 1;
					}
					state = ACTIVE;
				} else {
					state = DONE;
				}
			}
			next += __popc(ballot & groupMask);
			if (__all_sync(0xffffffffu, state == DONE)) break;

			if (state == ACTIVE) {
				bool finished;
				if (GI) {
					finished = !pathSegment<R, F>(sc, ft, ps, rng, eyeCol, cnt);
				} else {
					const RayTask<R> t = ws.stack[--ws.sp];
					if (t.branch == 0) {
						whittedStep<R, F>(sc, ft, t, rng, ws, eyeCol, cnt); // primary invocation: the sample's own stream
					} else {
						Rng child;
						child.init(p.seed, rng.pixel, rng.sample, t.branch);
						child.skip(t.count);
						whittedStep<R, F>(sc, ft, t, child, ws, eyeCol, cnt);
					}
					finished = ws.sp == 0;
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
							ps.start = rightEye.start; ps.dir = rightEye.dir; ps.mult = Col(1, 1, 1); ps.depth = 0; ps.flags = 0;
						} else {
							RayTask<R> root;
							root.start = rightEye.start; root.dir = rightEye.dir; root.weight = Col(1, 1, 1);
							root.depth = 0; root.branch = 0; root.count = 0;
							ws.stack[0] = root;
							ws.sp = 1;
						}
					} else {
						state = IDLE;
					}
				}
			}
		}

		// fixed-order tree over the G lanes of the pixel
		for (int m = G >> 1; m > 0; m >>= 1) accum = accum + shflXorCol(accum, m);
		if (valid && (lane & (unsigned) (G - 1)) == 0) {
			if (!p.sumOnly) accum = accum / (float) p.spp; // `avg / samplesPerPixel`, src/main.cpp:360
			float* o = p.out + 3 * ((size_t) py * p.width + px);
			o[0] = accum.r; o[1] = accum.g; o[2] = accum.b;
		}
	}

	// statistics: one atomic per warp
	for (int m = 16; m > 0; m >>= 1) {
		cnt.rays += __shfl_xor_sync(0xffffffffu, cnt.rays, m);
		cnt.primary += __shfl_xor_sync(0xffffffffu, cnt.primary, m);
		cnt.shadow += __shfl_xor_sync(0xffffffffu, cnt.shadow, m);
	}
	if (lane == 0) {
		atomicAdd(p.counters + 0, cnt.rays);
		atomicAdd(p.counters + 1, cnt.primary);
		atomicAdd(p.counters + 2, cnt.shadow);
		if (ws.overflow) atomicExch(p.errorFlag, 1);
	}
}

// FRAY_RENDER_AOV: primary hit ids of the un-jittered pin-hole ray through (x, y)
template <typename R, int F>
__global__ void __launch_bounds__(128) aovKernel(const DScene<R> sc, const RenderParams p)
{
	const FlatTab ft = stageFlat<R, F>(sc);
	const int total = p.numBuckets * FRAY_BUCKET * FRAY_BUCKET;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
		if ((i / 32) % p.taskStride != p.taskOffset) continue; // shards own interleaved runs of 32 pixels
		const int4 bk = p.buckets[i / (FRAY_BUCKET * FRAY_BUCKET)];
		const int k = i % (FRAY_BUCKET * FRAY_BUCKET);
		const int lx = k % FRAY_BUCKET, ly = k / FRAY_BUCKET;
		if (lx >= bk.z || ly >= bk.w) continue;
		const int px = bk.x + lx, py = bk.y + ly;
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

struct LaunchConfig {
	int gridBlocks; // persistent grid
	cudaStream_t stream;
};

template <typename R> inline size_t flatSmemBytes(const DScene<R>& sc)
{
	return (size_t) sc.numFlatAll * (FRAY_FLAT_POLY_VEC * sizeof(float4) + sizeof(FlatInfo));
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
