// fray_gpu.cu -- context management and the C ABI of libfray_gpu.so (include/fray_gpu.h).
//
// The context owns every device allocation: the scene image (one blob, see scene_image.h), the chunk-sum scratch buffer of the
// current call, the work counter / ray counters, a device framebuffer and a pinned host staging buffer for
// fray_gpu_render(). One CUDA stream per context; kernel time is measured with CUDA events on that stream.
// There is no CPU code path: if CUDA is unavailable every entry point fails with FRAY_GPU_ENODEVICE.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "render_kernels.cuh"
#include "wave_kernels.cuh"
#include "scene_image.h"

using namespace fray;

static thread_local std::string g_lastError;

static int fail(int code, const std::string& msg)
{
	g_lastError = msg;
	return code;
}

#define CUDA_TRY(expr)                                                                                   \
	do {                                                                                                 \
		cudaError_t e_ = (expr);                                                                         \
		if (e_ != cudaSuccess) return fail(FRAY_GPU_ECUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
	} while (0)

struct FrayGpuCtx {
	int device = 0;
	int precision = FRAY_GPU_FP32;
	int features = 0;
	int width = 0, height = 0;
	int defaultSpp = 1;
	int numSMs = 0;
	SceneImage<float> img32;
	SceneImage<double> img64;
	DScene<float> sc32;
	DScene<double> sc64;
	void* dBlob = nullptr;
	cudaStream_t stream = nullptr;
	cudaEvent_t evStart = nullptr, evStop = nullptr;
	float* dScratch = nullptr;   // chunk sums of the current call (numChunks > 1)
	size_t scratchFloats = 0;
	unsigned long long* dCounters = nullptr; // 3 counters
	unsigned int* dWork = nullptr;
	int* dError = nullptr;
	float* dShared = nullptr;    // frame exported to other processes (fray_gpu_frame_export)
	float* dFrame = nullptr;     // own framebuffer for fray_gpu_render
	float* hStaging = nullptr;   // pinned
	int occGI = -1, occWhitted = -1;
	// wavefront Whitted path (wave_kernels.cuh): queues, accumulators, counters; grown on demand
	void* dWave = nullptr;          // one allocation, carved up by waveLayout()
	size_t waveBytes = 0;
	unsigned waveRayCap = 0, waveLitCap = 0, waveHitCap = 0;
	unsigned* dWaveCtr = nullptr;
	WaveLaunch waveCfg = { 0, 0, nullptr, 0, 0, 0 };
	bool waveLast = false;          // the last render went through the wavefront path
	FrayGpuFrame lastFrame = {};
	float* lastOut = nullptr;
	bool lastTimed = false;
	double chunkPixels = 0;         // > 0: cut the samples into the chunks a frame of this many pixels gets (fray_gpu_multi_render: bit-identical tile shares)
	int waveFan = 1;                // secondary rays one hit can spawn (glossy samples), for the first guess of the queue size
	bool waveSecondary = false;     // some node's shader reflects or refracts: the ray tree is deeper than the camera rays
	int waveLitPerRay = 1;          // Lambert / Phong evaluations one hit can ask for (through Layered shaders)
	bool pendingStats = false;
	int launches = 0;
	cudaStream_t lastStream = nullptr;
};

// d_rgb[i] = d_sum[i] / spp  (`avg / samplesPerPixel`, src/main.cpp:360)
__global__ void resolveKernel(const float* __restrict__ sum, float* __restrict__ rgb, size_t n, float spp)
{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) rgb[i] = sum[i] / spp;
}

// numChunks > 1: pixel = sum of its chunk sums in chunk order (then / spp unless FRAY_FRAME_SUM)
__global__ void combineKernel(const RenderParams p)
{
	const unsigned slots = (unsigned) p.numOwnedTiles * 32u;
	for (unsigned slot = blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += gridDim.x * blockDim.x) {
		int px, py;
		if (!slotPixel(p, slot, px, py)) continue;
		const float* c = p.scratch + 3 * (size_t) slot; // chunk-major: consecutive threads read consecutive triples
		const size_t stride = 3 * (size_t) slots;
		Col sum(0, 0, 0);
		for (int k = 0; k < p.numChunks; k++, c += stride) sum = sum + Col(c[0], c[1], c[2]);
		if (!p.sumOnly) sum = sum / (float) p.spp;
		float* o = p.out + 3 * ((size_t) py * p.width + px);
		o[0] = sum.r; o[1] = sum.g; o[2] = sum.b;
	}
}

// the pixels of the owned tiles, device frame -> (mapped host) frame: fray_gpu_render_to_host for the wavefront path, whose passes
// read the frame back while they build it
__global__ void gatherOwnedKernel(const RenderParams p, const float* __restrict__ src)
{
	const unsigned slots = (unsigned) p.numOwnedTiles * 32u;
	for (unsigned slot = blockIdx.x * blockDim.x + threadIdx.x; slot < slots; slot += gridDim.x * blockDim.x) {
		int px, py;
		if (!slotPixel(p, slot, px, py)) continue;
		const size_t i = 3 * ((size_t) py * p.width + px);
		p.out[i] = src[i]; p.out[i + 1] = src[i + 1]; p.out[i + 2] = src[i + 2];
	}
}

// ---- roofline denominators ---------------------------------------------------------------------------
__global__ void ffmaPeakKernel(float* sink, int iters, float a, float b)
{
	float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
	for (int i = 0; i < iters; i++) {
#pragma unroll
		for (int k = 0; k < 16; k++) {
			x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
			x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
		}
	}
	const float r = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
	if (r == 123.456f) sink[0] = r; // never true: keeps the chains alive
}

__global__ void l2ReadKernel(const float4* __restrict__ buf, size_t n, int passes, float* sink)
{
	float acc = 0;
	for (int p = 0; p < passes; p++)
		for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
			const float4 v = __ldcg(buf + i); // cache at L2 only
			acc += v.x + v.y + v.z + v.w;
		}
	if (acc == 123.456f) sink[0] = acc;
}

template <typename R> static void convertCamera(DCamera<R>& d, const FrayGpuCamera& cam)
{
	for (int i = 0; i < 3; i++) {
		d.pos[i] = (R) cam.pos[i]; d.topLeft[i] = (R) cam.top_left[i]; d.topRight[i] = (R) cam.top_right[i];
		d.bottomLeft[i] = (R) cam.bottom_left[i]; d.front[i] = (R) cam.front[i]; d.up[i] = (R) cam.up[i]; d.right[i] = (R) cam.right[i];
		d.leftMask[i] = cam.left_mask[i]; d.rightMask[i] = cam.right_mask[i];
	}
	for (int i = 0; i < 3; i++) {
		d.du[i] = (R) ((cam.top_right[i] - cam.top_left[i]) / cam.w);
		d.dv[i] = (R) ((cam.bottom_left[i] - cam.top_left[i]) / cam.h);
	}
	d.w = (R) cam.w; d.h = (R) cam.h; d.aperture = (R) cam.aperture_size; d.focalDist = (R) cam.focal_plane_dist;
	d.stereoSep = (R) cam.stereo_separation; d.dof = cam.dof;
}

extern "C" {

uint32_t fray_gpu_abi_version(void) { return FRAY_GPU_ABI_VERSION; }

const char* fray_gpu_last_error(void) { return g_lastError.c_str(); }

int fray_gpu_device_count(void)
{
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

int fray_gpu_samples_per_pixel(const FrayGpuScene* scene) { return scene ? samplesPerPixel(*scene) : 0; }

void fray_gpu_destroy(FrayGpuCtx* c)
{
	if (!c) return;
	cudaSetDevice(c->device);
	if (c->stream) cudaStreamSynchronize(c->stream);
	cudaFree(c->dBlob);
	cudaFree(c->dScratch);
	cudaFree(c->dCounters);
	cudaFree(c->dWork);
	cudaFree(c->dError);
	cudaFree(c->dFrame);
	cudaFree(c->dShared);
	cudaFree(c->dWave);
	cudaFree(c->dWaveCtr);
	if (c->hStaging) cudaFreeHost(c->hStaging);
	if (c->evStart) cudaEventDestroy(c->evStart);
	if (c->evStop) cudaEventDestroy(c->evStop);
	if (c->stream) cudaStreamDestroy(c->stream);
	delete c;
}

int fray_gpu_create(const FrayGpuScene* scene, int device, int precision, FrayGpuCtx** out)
{
	if (!scene || !out) return fail(FRAY_GPU_EINVAL, "null argument");
	*out = nullptr;
	if (precision != FRAY_GPU_FP32 && precision != FRAY_GPU_FP64) return fail(FRAY_GPU_EINVAL, "precision must be FRAY_GPU_FP32 or FRAY_GPU_FP64");
	int ndev = 0;
	cudaError_t e = cudaGetDeviceCount(&ndev);
	if (e != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(FRAY_GPU_ENODEVICE, std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
		                                    "); libfray_gpu has no CPU fallback");
	}
	if (device < 0 || device >= ndev) return fail(FRAY_GPU_EINVAL, "device ordinal out of range");

	FrayGpuCtx* c = new FrayGpuCtx;
	c->device = device;
	c->precision = precision;
	c->width = scene->settings.frame_width;
	c->height = scene->settings.frame_height;
	c->defaultSpp = samplesPerPixel(*scene);
	std::string err;
	const bool ok = precision == FRAY_GPU_FP32 ? c->img32.build(*scene, err) : c->img64.build(*scene, err);
	if (!ok) {
		delete c;
		return fail(FRAY_GPU_EINVAL, err);
	}
	c->features = precision == FRAY_GPU_FP32 ? c->img32.features : c->img64.features;
	for (int i = 0; i < scene->num_shaders; i++)
		if (scene->shaders[i].type == FRAY_SHADER_REFL && !scene->shaders[i].pure_reflection) c->waveFan = std::max(c->waveFan, scene->shaders[i].num_samples);
	{
		// lit evaluations per hit: Lambert / Phong leaves of a node's shader tree
		std::function<int(int, int)> litLeaves = [&](int si, int depth) -> int {
			if (si < 0 || si >= scene->num_shaders || depth > 4) return 0;
			const FrayGpuShader& sh = scene->shaders[si];
			if (sh.type == FRAY_SHADER_LAMBERT || sh.type == FRAY_SHADER_PHONG) return 1;
			int n = 0;
			if (sh.type == FRAY_SHADER_LAYERED)
				for (int k = 0; k < sh.num_layers && sh.first_layer + k < scene->num_layers; k++) n += litLeaves(scene->layers[sh.first_layer + k].shader, depth + 1);
			return n;
		};
		for (int i = 0; i < scene->num_nodes; i++) c->waveLitPerRay = std::max(c->waveLitPerRay, litLeaves(scene->nodes[i].shader, 0));
		std::vector<int> todo;
		std::vector<char> seen(scene->num_shaders, 0);
		for (int i = 0; i < scene->num_nodes; i++) todo.push_back(scene->nodes[i].shader);
		while (!todo.empty()) {
			const int si = todo.back();
			todo.pop_back();
			if (si < 0 || si >= scene->num_shaders || seen[si]) continue;
			seen[si] = 1;
			const FrayGpuShader& sh = scene->shaders[si];
			if (sh.type == FRAY_SHADER_REFL || sh.type == FRAY_SHADER_REFR) c->waveSecondary = true;
			if (sh.type == FRAY_SHADER_LAYERED)
				for (int k = 0; k < sh.num_layers && sh.first_layer + k < scene->num_layers; k++) todo.push_back(scene->layers[sh.first_layer + k].shader);
		}
	}
	const std::vector<unsigned char>& blob = precision == FRAY_GPU_FP32 ? c->img32.blob : c->img64.blob;

#define CREATE_TRY(expr)                                                                   \
	do {                                                                                   \
		cudaError_t e_ = (expr);                                                           \
		if (e_ != cudaSuccess) {                                                           \
			std::string m_ = std::string(#expr) + ": " + cudaGetErrorString(e_);           \
			fray_gpu_destroy(c);                                                           \
			return fail(FRAY_GPU_ECUDA, m_);                                               \
		}                                                                                  \
	} while (0)
	CREATE_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CREATE_TRY(cudaGetDeviceProperties(&prop, device));
	c->numSMs = prop.multiProcessorCount;
	CREATE_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
	CREATE_TRY(cudaEventCreate(&c->evStart));
	CREATE_TRY(cudaEventCreate(&c->evStop));
	CREATE_TRY(cudaMalloc(&c->dBlob, blob.size()));
	CREATE_TRY(cudaMemcpyAsync(c->dBlob, blob.data(), blob.size(), cudaMemcpyHostToDevice, c->stream));
	CREATE_TRY(cudaMalloc(&c->dCounters, 3 * sizeof(unsigned long long)));
	CREATE_TRY(cudaMalloc(&c->dWork, sizeof(unsigned int)));
	CREATE_TRY(cudaMalloc(&c->dError, sizeof(int)));
	CREATE_TRY(cudaMemsetAsync(c->dError, 0, sizeof(int), c->stream));
	const size_t frameBytes = (size_t) c->width * c->height * 3 * sizeof(float);
	CREATE_TRY(cudaMalloc(&c->dFrame, frameBytes));
	CREATE_TRY(cudaMallocHost(&c->hStaging, frameBytes));
	CREATE_TRY(cudaStreamSynchronize(c->stream));
#undef CREATE_TRY
	if (precision == FRAY_GPU_FP32) c->sc32 = c->img32.bind(c->dBlob);
	else c->sc64 = c->img64.bind(c->dBlob);
	*out = c;
	return FRAY_GPU_OK;
}

int fray_gpu_update_camera(FrayGpuCtx* c, const FrayGpuCamera* cam)
{
	if (!c || !cam) return fail(FRAY_GPU_EINVAL, "null argument");
	// the camera travels as a kernel parameter (constant bank), so a per-frame update costs no device copy
	if (c->precision == FRAY_GPU_FP32) convertCamera(c->sc32.cam, *cam); else convertCamera(c->sc64.cam, *cam);
	if ((cam->dof || cam->stereo_separation > 0) && !(c->features & FRAY_F_LENS)) { // the new camera needs a kernel variant with a lens
		c->features |= FRAY_F_LENS;
		c->occGI = c->occWhitted = -1;
	}
	if ((int) cam->w != c->width || (int) cam->h != c->height) return fail(FRAY_GPU_EINVAL, "frame size cannot change without re-creating the context");
	return FRAY_GPU_OK;
}

static int pow2Floor(int v)
{
	int p = 1;
	while (p * 2 <= v) p *= 2;
	return p;
}

// samples per work item (render_kernels.cuh) for a call that owns `ownedTiles` tiles and renders `samples` samples of each pixel
// How the samples [0, samples) of a pixel are cut into chunks (render_kernels.cuh: a work item is one chunk of one pixel, and the
// queue hands out all items of chunk 0 first). Guided self-scheduling: one chunk index gives every lane pixels / lanes items,
// and a chunk is a third of the paths a lane still has to do -- at most 8 (beyond that the bookkeeping per item and the chunk
// sums are noise already), at least 1 -- so most of the frame runs in items of 8 paths and the queue ends in items of one: the
// lanes run dry within one path (~35 us on a busy SM) of each other, where uniform chunks of 8 drained for up to eight.
// One GPU, the headline frame (160000 pixels, 256 paths, 132608 lanes): 8 x 30, 6, 4, 2, 1, 1, 1, 1 = 37 chunks; an eighth of
// its samples: 8, 8, 6, 4, 2, 1, 1, 1, 1 = 9 chunks where uniform chunks of 1 needed 32 (a quarter of the chunk sums to write
// and add). `starts` gets numChunks + 1 entries. When that makes more than FRAY_MAX_CHUNKS chunks, or more than 192 MB of chunk
// sums (one RGB triple per pixel and chunk), all sizes are doubled until it does not (smallpt, 1024 paths: 32 ... 24, 4, 4).
static void buildChunks(double pixels, int samples, double lanes, std::vector<int>& starts)
{
	const double perLane = std::max(pixels, 1.0) / std::max(lanes, 1.0);
	const int maxChunks = (int) std::min((double) FRAY_MAX_CHUNKS, std::max(1.0, 192e6 / (std::max(pixels, 1.0) * 12.0)));
	const char* e = getenv("FRAY_GPU_CHUNK"); // experiments only: uniform chunks of that many samples
	const int uniform = e ? std::max(1, atoi(e)) : 0;
	for (int unit = 1;; unit *= 2) {
		starts.assign(1, 0);
		for (int done = 0; done < samples;) {
			const int left = samples - done;
			int len = unit * std::max(1, std::min(8, (int) (left * perLane / (3.0 * unit))));
			if (uniform) len = uniform * unit;
			starts.push_back(done += std::min(left, len));
		}
		if ((int) starts.size() - 1 <= maxChunks || unit >= samples) return;
	}
}

// ---- wavefront Whitted path ------------------------------------------------------------------------------------------------
// Fast precision, Whitted integrator, a generic node loop (KD meshes / analytic primitives) and no CSG: wave_kernels.cuh.
// FRAY_GPU_NO_WAVE=1 keeps such scenes on the megakernel (A/B timing).
static bool waveEligible(const FrayGpuCtx* c, const FrayGpuFrame* f)
{
	if (c->precision != FRAY_GPU_FP32 || c->sc32.gi || f->mode != FRAY_RENDER_BEAUTY) return false;
	if (!WaveVariants::covers(c->features)) return false;
	if (c->sc32.maxTraceDepth + 3 > FRAY_WAVE_MAX || c->sc32.numNodes > 65000) return false;
	static const bool off = getenv("FRAY_GPU_NO_WAVE") != nullptr;
	return !off;
}

static int waveAllocate(FrayGpuCtx* c, unsigned numPrimary)
{
	// first guess: every primary hit spawns `fan` secondary rays (glossy samples), capped; SHADE raises a flag when a queue is
	// full, waveFinish() then doubles the queues and renders the frame again
	if (c->waveRayCap == 0) {
		double guess = c->waveSecondary ? (double) numPrimary * std::min(std::max(c->waveFan, 1), 8) + (1 << 20) : (double) (64 * 1024);
		if (c->sc32.cam.stereoSep > 0) guess += (double) numPrimary; // the right eyes travel through the queue of wave 1
		c->waveRayCap = (unsigned) std::min(guess, 6.0e8);
	}
	c->waveRayCap = (c->waveRayCap + FRAY_WAVE_REGIONS - 1) / FRAY_WAVE_REGIONS * FRAY_WAVE_REGIONS; // FRAY_WAVE_REGIONS equal sub-queues
	const unsigned hitCap = std::max(numPrimary, c->waveRayCap);
	if ((double) hitCap * c->waveLitPerRay >= 4.0e9) return fail(FRAY_GPU_ENOMEM, "the lit records of this frame do not fit the wavefront queues");
	const unsigned litCap = hitCap * (unsigned) c->waveLitPerRay; // a fixed place for every record: ray i, evaluation k -> slot i * litPerRay + k
	const size_t pixels = (size_t) c->width * c->height;
	const size_t need = (size_t) c->waveRayCap * 2 * (3 * sizeof(float4) + sizeof(uint2)) + (size_t) hitCap * (sizeof(float4) + sizeof(int2)) +
	                    (size_t) litCap * (5 * sizeof(float4) + sizeof(uint2)) + pixels * 3 * sizeof(long long) + 4096;
	if (need > c->waveBytes || hitCap > c->waveHitCap || litCap > c->waveLitCap) {
		cudaFree(c->dWave);
		c->dWave = nullptr;
		c->waveBytes = 0;
		CUDA_TRY(cudaMalloc(&c->dWave, need));
		c->waveBytes = need;
		c->waveHitCap = hitCap;
		c->waveLitCap = litCap;
	}
	if (!c->dWaveCtr) CUDA_TRY(cudaMalloc(&c->dWaveCtr, FRAY_WCTR_COUNT * sizeof(unsigned)));
	return FRAY_GPU_OK;
}

static void waveLayout(const FrayGpuCtx* c, WaveParams& w)
{
	char* b = (char*) c->dWave;
	auto take = [&](size_t bytes) { char* r = b; b += (bytes + 255) & ~(size_t) 255; return r; };
	const size_t rays = (size_t) c->waveRayCap * 2;
	w.rayO = (float4*) take(rays * sizeof(float4));
	w.rayD = (float4*) take(rays * sizeof(float4));
	w.rayW = (float4*) take(rays * sizeof(float4));
	w.rayC = (uint2*) take(rays * sizeof(uint2));
	w.hitA = (float4*) take((size_t) c->waveHitCap * sizeof(float4));
	w.hitB = (int2*) take((size_t) c->waveHitCap * sizeof(int2));
	w.litA = (float4*) take((size_t) c->waveLitCap * sizeof(float4));
	w.litB = (float4*) take((size_t) c->waveLitCap * sizeof(float4));
	w.litC = (float4*) take((size_t) c->waveLitCap * sizeof(float4));
	w.litD = (uint2*) take((size_t) c->waveLitCap * sizeof(uint2));
	w.litE = (float4*) take((size_t) c->waveLitCap * sizeof(float4));
	w.litF = (float4*) take((size_t) c->waveLitCap * sizeof(float4));
	w.acc = (long long*) take((size_t) c->width * c->height * 3 * sizeof(long long));
	w.rayCap = c->waveRayCap;
	w.litCap = c->waveLitCap;
	w.litPerRay = c->waveLitPerRay;
	w.ctr = c->dWaveCtr;
}

static int renderWave(FrayGpuCtx* c, const RenderParams& rp, float* dOut, cudaStream_t stream, bool timed)
{
	WaveParams w;
	memset(&w, 0, sizeof(w));
	w.rp = rp;
	w.rp.out = dOut;
	w.numSamples = rp.s1 - rp.s0;
	w.invNumSamples = 1.0f / (float) w.numSamples;
	const double prim = (double) rp.numOwnedTiles * 32.0 * w.numSamples;
	if (prim >= 4.0e9) return fail(FRAY_GPU_EINVAL, "more than 2^32 primary samples in one call");
	w.numPrimary = (unsigned) prim;
	w.rp.exactDiv = (prim / 32.0 >= 4e6 || rp.exactDiv) ? 1 : 0;
	int rc = waveAllocate(c, w.numPrimary);
	if (rc != FRAY_GPU_OK) return rc;
	waveLayout(c, w);
	const bool stereo = c->sc32.cam.stereoSep > 0;
	c->waveCfg.numSMs = c->numSMs;
	c->waveCfg.stream = stream;
	// the right eye of a stereo pair is queued by the left eye's SHADE (it continues the left eye's random stream): one more wave
	c->waveCfg.waves = std::min(FRAY_WAVE_MAX, (c->waveSecondary ? c->sc32.maxTraceDepth + 1 : 1) + (stereo ? 1 : 0));
	if (c->waveCfg.waves < 1) c->waveCfg.waves = 1;
	// one wave and one sample per pixel: every pixel has a single owner in each pass, no accumulators (wave_kernels.cuh)
	w.direct = (c->waveCfg.waves == 1 && w.numSamples == 1 && !getenv("FRAY_GPU_WAVE_ACCUMULATE")) ? 1 : 0;
	CUDA_TRY(cudaMemsetAsync(c->dWaveCtr, 0, FRAY_WCTR_COUNT * sizeof(unsigned), stream));
	if (timed) CUDA_TRY(cudaEventRecord(c->evStart, stream));
	if (!w.direct) CUDA_TRY(cudaMemsetAsync(w.acc, 0, (size_t) c->width * c->height * 3 * sizeof(long long), stream));
	cudaError_t e = launchWaveFrame(c->sc32, w, c->features, c->waveCfg);
	if (e != cudaSuccess) return fail(FRAY_GPU_ECUDA, std::string("wavefront kernel launch: ") + cudaGetErrorString(e));
	c->launches = c->waveCfg.waves * ((c->sc32.numLights > 0 && !c->waveCfg.fused) ? 3 : 2) + (w.direct ? 0 : 1);
	if (timed) CUDA_TRY(cudaEventRecord(c->evStop, stream));
	c->pendingStats = timed;
	c->lastStream = stream;
	c->waveLast = true;
	return FRAY_GPU_OK;
}

static int renderInto(FrayGpuCtx* c, const FrayGpuFrame* f, float* dOut, cudaStream_t stream, bool timed);

// After a wavefront frame: wait for it and, if a queue overflowed, double the queues and render the frame again.
static int waveFinish(FrayGpuCtx* c)
{
	for (int attempt = 0; c->waveLast && attempt < 12; attempt++) {
		CUDA_TRY(cudaStreamSynchronize(c->lastStream ? c->lastStream : c->stream));
		unsigned overflow = 0;
		CUDA_TRY(cudaMemcpy(&overflow, c->dWaveCtr + FRAY_WCTR_OVERFLOW, sizeof(unsigned), cudaMemcpyDeviceToHost));
		if (!overflow) return FRAY_GPU_OK;
		if (c->waveRayCap >= 600000000u) return fail(FRAY_GPU_ENOMEM, "the ray tree of this frame does not fit the wavefront queues");
		c->waveRayCap = (unsigned) std::min(2.0 * c->waveRayCap, 6.0e8);
		const FrayGpuFrame f = c->lastFrame;
		int rc = renderInto(c, &f, c->lastOut, c->lastStream, c->lastTimed);
		if (rc != FRAY_GPU_OK) return rc;
	}
	return FRAY_GPU_OK;
}

static int renderInto(FrayGpuCtx* c, const FrayGpuFrame* f, float* dOut, cudaStream_t stream, bool timed)
{
	if (!c || !f || !dOut) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaSetDevice(c->device));
	const bool gi = c->precision == FRAY_GPU_FP32 ? c->sc32.gi : c->sc64.gi;
	const bool randomOffsets = gi || (c->precision == FRAY_GPU_FP32 ? c->sc32.cam.dof : c->sc64.cam.dof);
	const int spp = f->spp > 0 ? f->spp : c->defaultSpp;
	int s0 = f->sample_begin, s1 = f->sample_end;
	if (s0 == 0 && s1 == 0 && !(f->flags & FRAY_FRAME_SAMPLE_RANGE)) s1 = spp;
	if (s0 < 0 || s1 < s0 || s1 > spp) return fail(FRAY_GPU_EINVAL, "sample range outside [0, spp]");
	if (!randomOffsets && spp > 5) return fail(FRAY_GPU_EINVAL, "more than 5 samples need dof or gi (fixed AA table has 5 entries, src/main.cpp:55-61)");
	if (f->mode != FRAY_RENDER_BEAUTY && f->mode != FRAY_RENDER_AOV && f->mode != FRAY_RENDER_PREPASS) return fail(FRAY_GPU_EINVAL, "unknown render mode");
	if (f->mode == FRAY_RENDER_PREPASS && f->bucket_count > 1) return fail(FRAY_GPU_EINVAL, "the prepass is not split into buckets");
	const int bcount = f->bucket_count > 0 ? f->bucket_count : 1;
	const int brank = f->bucket_count > 0 ? f->bucket_rank : 0;
	if (brank < 0 || brank >= bcount) return fail(FRAY_GPU_EINVAL, "bucket_rank outside [0, bucket_count)");

	// work decomposition (render_kernels.cuh): 8x4 pixel tiles, tile t belongs to share t % bcount; samples in chunks of C
	const int tilesX = (c->width + FRAY_TILE_W - 1) / FRAY_TILE_W, tilesY = (c->height + FRAY_TILE_H - 1) / FRAY_TILE_H;
	const int totalTiles = tilesX * tilesY;
	const int ownedTiles = totalTiles > brank ? (totalTiles - brank + bcount - 1) / bcount : 0;

	int& occ = gi ? c->occGI : c->occWhitted;
	if (occ < 0) {
		occ = c->precision == FRAY_GPU_FP32 ? renderOccupancy<float>(c->sc32, c->features, gi) : renderOccupancy<double>(c->sc64, c->features, gi);
		if (occ < 1) occ = 1;
	}
	LaunchConfig cfg;
	cfg.stream = stream;
	cfg.gridBlocks = c->numSMs * occ; // persistent: every resident CTA slot of the chip, exactly once
	if (f->mode == FRAY_RENDER_AOV) cfg.gridBlocks = c->numSMs * 8;
	if (f->mode == FRAY_RENDER_PREPASS) cfg.gridBlocks = std::max(1, std::min(c->numSMs * 2, ((c->width + 15) / 16) * ((c->height + 15) / 16) / 128 + 1));

	RenderParams p;
	memset(&p, 0, sizeof(p));
	p.width = c->width; p.height = c->height; p.spp = spp; p.s0 = s0; p.s1 = s1; p.seed = f->seed;
	philoxRoundKeys(p.seed, p.roundKeys);
	p.sumOnly = (f->flags & FRAY_FRAME_SUM) ? 1 : 0;
	// chunk size: every lane of the grid should see >= ~16 items, so that the drain at the end of the kernel (lanes running dry
	// while the last items finish; an item of C paths takes C x ~35 us on a busy SM) is a few percent of the call even when 8 GPUs
	// share a frame. Measured on a 1/8 share of the headline frame (tools/share_time.py, 1.10 ms of work): C = 1 1.229 ms,
	// C = 2 1.204 ms, C = 4 1.233 ms, C = 8 1.289 ms. The scratch buffer (one RGB sum per pixel and chunk) is kept below 192 MB
	const int samples = std::max(1, s1 - s0);
	std::vector<int> starts;
	buildChunks(c->chunkPixels > 0 ? c->chunkPixels : (double) ownedTiles * 32.0, samples, (double) cfg.gridBlocks * 128.0, starts);
	p.numChunks = (int) starts.size() - 1;
	for (int k = 0; k <= p.numChunks; k++) p.chunkStart[k] = (unsigned) starts[k];
	p.numBulk = 1; // the leading run of equally long chunks goes tile by tile, the descending rest chunk by chunk
	while (p.numBulk < p.numChunks && starts[p.numBulk + 1] - starts[p.numBulk] == starts[1] - starts[0]) p.numBulk++;
	if (p.numBulk == p.numChunks && p.numChunks > 1) p.numBulk--; // (uniform chunks: keep the decode of the rest in use)
	p.bulkItems = (unsigned) p.numBulk * (unsigned) ownedTiles * 32u;
	p.invNumBulk = 1.0f / (float) p.numBulk;
	p.tilesX = tilesX;
	p.numOwnedTiles = ownedTiles;
	p.taskStride = bcount;
	p.taskOffset = brank;
	p.totalItems = (unsigned) ownedTiles * 32u * (unsigned) p.numChunks;
	p.invTilesX = 1.0f / (float) tilesX;
	p.invNumOwnedTiles = 1.0f / (float) std::max(ownedTiles, 1);
	p.exactDiv = ((double) totalTiles >= 4e6 || (double) ownedTiles * p.numChunks >= 4e6) ? 1 : 0; // beyond 2^22: integer division
	p.out = dOut;
	p.counters = c->dCounters;
	p.workCounter = c->dWork;
	p.errorFlag = c->dError;
	if (p.numChunks > 1) {
		const size_t need = (size_t) ownedTiles * 32 * p.numChunks * 3;
		if (need > c->scratchFloats) {
			cudaFree(c->dScratch);
			c->dScratch = nullptr;
			c->scratchFloats = 0;
			CUDA_TRY(cudaMalloc(&c->dScratch, need * sizeof(float)));
			c->scratchFloats = need;
		}
		p.scratch = c->dScratch;
	}

	if (bcount > 1 && !(f->flags & FRAY_FRAME_OWNED_ONLY)) CUDA_TRY(cudaMemsetAsync(dOut, 0, (size_t) c->width * c->height * 3 * sizeof(float), stream));
	CUDA_TRY(cudaMemsetAsync(c->dCounters, 0, 3 * sizeof(unsigned long long), stream));
	CUDA_TRY(cudaMemsetAsync(c->dWork, 0, sizeof(unsigned int), stream));

	c->launches = 0;
	c->waveLast = false;
	c->lastFrame = *f;
	c->lastOut = dOut;
	c->lastTimed = timed;
	if (waveEligible(c, f) && ownedTiles > 0 && s1 > s0) return renderWave(c, p, dOut, stream, timed);
	if (timed) CUDA_TRY(cudaEventRecord(c->evStart, stream));
	if (ownedTiles > 0 && (s1 > s0 || f->mode != FRAY_RENDER_BEAUTY)) {
		cudaError_t e = c->precision == FRAY_GPU_FP32 ? launchRender<float>(c->sc32, p, c->features, f->mode, cfg)
		                                                : launchRender<double>(c->sc64, p, c->features, f->mode, cfg);
		if (e != cudaSuccess) return fail(FRAY_GPU_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e));
		c->launches = 1;
		if (p.numChunks > 1 && f->mode == FRAY_RENDER_BEAUTY) {
			combineKernel<<<c->numSMs * 4, 256, 0, stream>>>(p);
			CUDA_TRY(cudaGetLastError());
			c->launches = 2;
		}
	} else if (bcount == 1) {
		CUDA_TRY(cudaMemsetAsync(dOut, 0, (size_t) c->width * c->height * 3 * sizeof(float), stream));
	}
	if (timed) CUDA_TRY(cudaEventRecord(c->evStop, stream));
	c->pendingStats = timed;
	c->lastStream = stream;
	return FRAY_GPU_OK;
}

static int fetchStats(FrayGpuCtx* c, FrayGpuStats* stats)
{
	int wrc = waveFinish(c);
	if (wrc != FRAY_GPU_OK) return wrc;
	CUDA_TRY(cudaStreamSynchronize(c->lastStream ? c->lastStream : c->stream));
	unsigned long long h[3] = { 0, 0, 0 };
	int err = 0;
	CUDA_TRY(cudaMemcpy(h, c->dCounters, sizeof(h), cudaMemcpyDeviceToHost));
	CUDA_TRY(cudaMemcpy(&err, c->dError, sizeof(int), cudaMemcpyDeviceToHost));
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		stats->rays = h[0];
		stats->primary_rays = h[1];
		stats->shadow_rays = h[2];
		stats->kernel_launches = c->launches;
		if (c->pendingStats) {
			float ms = 0;
			CUDA_TRY(cudaEventElapsedTime(&ms, c->evStart, c->evStop));
			stats->device_ms = ms;
		}
	}
	if (err) {
		cudaMemset(c->dError, 0, sizeof(int));
		return fail(FRAY_GPU_EUNSUPPORTED, "device ray-task stack overflow (too many pending secondary rays)");
	}
	return FRAY_GPU_OK;
}

int fray_gpu_render(FrayGpuCtx* c, const FrayGpuFrame* frame, float* rgb_out, FrayGpuStats* stats)
{
	if (!c || !frame || !rgb_out) return fail(FRAY_GPU_EINVAL, "null argument");
	int rc = renderInto(c, frame, c->dFrame, c->stream, true);
	if (rc != FRAY_GPU_OK) return rc;
	rc = waveFinish(c); // wavefront frames: a full queue means the frame is rendered again before it is copied out
	if (rc != FRAY_GPU_OK) return rc;
	const size_t bytes = (size_t) c->width * c->height * 3 * sizeof(float);
	// straight into the caller's buffer when it is page-locked (e.g. a pinned torch tensor), else through our pinned staging buffer
	cudaPointerAttributes attr;
	bool pinned = cudaPointerGetAttributes(&attr, rgb_out) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (pinned) {
		CUDA_TRY(cudaMemcpyAsync(rgb_out, c->dFrame, bytes, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
	} else {
		CUDA_TRY(cudaMemcpyAsync(c->hStaging, c->dFrame, bytes, cudaMemcpyDeviceToHost, c->stream));
		CUDA_TRY(cudaStreamSynchronize(c->stream));
		memcpy(rgb_out, c->hStaging, bytes);
	}
	return fetchStats(c, stats);
}

int fray_gpu_render_device(FrayGpuCtx* c, const FrayGpuFrame* frame, void* d_rgb, void* cuda_stream)
{
	if (!c) return fail(FRAY_GPU_EINVAL, "null argument");
	return renderInto(c, frame, (float*) d_rgb, cuda_stream ? (cudaStream_t) cuda_stream : c->stream, true);
}

int fray_gpu_frame_export(FrayGpuCtx* c, void** d_frame, unsigned char handle[64])
{
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "the ABI carries the IPC handle as 64 opaque bytes");
	if (!c || !d_frame || !handle) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaSetDevice(c->device));
	if (!c->dShared) { // TWO frames back to back (see include/fray_gpu.h: double buffering)
		const size_t bytes = 2 * (size_t) c->width * c->height * 3 * sizeof(float);
		CUDA_TRY(cudaMalloc(&c->dShared, bytes));
		CUDA_TRY(cudaMemset(c->dShared, 0, bytes));
	}
	cudaIpcMemHandle_t h;
	CUDA_TRY(cudaIpcGetMemHandle(&h, c->dShared));
	memcpy(handle, &h, sizeof(h));
	*d_frame = c->dShared;
	return FRAY_GPU_OK;
}

int fray_gpu_frame_import(FrayGpuCtx* c, const unsigned char handle[64], void** d_frame)
{
	if (!c || !d_frame || !handle) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaSetDevice(c->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, sizeof(h));
	void* p = nullptr;
	CUDA_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
	*d_frame = p;
	return FRAY_GPU_OK;
}

int fray_gpu_frame_close(FrayGpuCtx* c, void* d_frame)
{
	if (!c || !d_frame) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaSetDevice(c->device));
	CUDA_TRY(cudaIpcCloseMemHandle(d_frame));
	return FRAY_GPU_OK;
}

// device-visible address of a page-locked host buffer (nullptr if `host` is ordinary pageable memory)
static float* mappedHostPointer(const void* host)
{
	cudaPointerAttributes attr;
	if (cudaPointerGetAttributes(&attr, host) != cudaSuccess || attr.type != cudaMemoryTypeHost || !attr.devicePointer) {
		cudaGetLastError();
		return nullptr;
	}
	return (float*) attr.devicePointer;
}

int fray_gpu_resolve_to_host(FrayGpuCtx* c, const void* d_sum, float* pinned_rgb, int32_t spp, void* cuda_stream)
{
	if (!c || !d_sum || !pinned_rgb || spp < 1) return fail(FRAY_GPU_EINVAL, "bad argument");
	CUDA_TRY(cudaSetDevice(c->device));
	float* dev = mappedHostPointer(pinned_rgb);
	if (!dev) return fail(FRAY_GPU_EINVAL, "fray_gpu_resolve_to_host needs page-locked (pinned) host memory");
	const size_t n = (size_t) c->width * c->height * 3;
	cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : c->stream;
	resolveKernel<<<c->numSMs * 4, 256, 0, st>>>((const float*) d_sum, dev, n, (float) spp);
	CUDA_TRY(cudaGetLastError());
	return FRAY_GPU_OK;
}

int fray_gpu_render_to_host(FrayGpuCtx* c, const FrayGpuFrame* frame, float* pinned_rgb, void* cuda_stream)
{
	if (!c || !frame || !pinned_rgb) return fail(FRAY_GPU_EINVAL, "null argument");
	if (frame->flags & FRAY_FRAME_SUM) return fail(FRAY_GPU_EINVAL, "fray_gpu_render_to_host delivers finished pixels, not sums");
	CUDA_TRY(cudaSetDevice(c->device));
	float* dev = mappedHostPointer(pinned_rgb);
	if (!dev) return fail(FRAY_GPU_EINVAL, "fray_gpu_render_to_host needs page-locked (pinned) host memory");
	cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : c->stream;
	FrayGpuFrame f = *frame;
	f.flags |= FRAY_FRAME_OWNED_ONLY;
	if (!waveEligible(c, &f)) return renderInto(c, &f, dev, st, true); // the kernels' own final stores go over PCIe
	// wavefront frames are built up in device memory (the shadow pass adds to what the shade pass stored): copy the owned tiles out
	int rc = renderInto(c, &f, c->dFrame, st, true);
	if (rc != FRAY_GPU_OK) return rc;
	rc = waveFinish(c); // a full queue means the frame is rendered again before it leaves
	if (rc != FRAY_GPU_OK) return rc;
	const int bcount = f.bucket_count > 0 ? f.bucket_count : 1, brank = f.bucket_count > 0 ? f.bucket_rank : 0;
	const int tilesX = (c->width + FRAY_TILE_W - 1) / FRAY_TILE_W, tilesY = (c->height + FRAY_TILE_H - 1) / FRAY_TILE_H;
	const int totalTiles = tilesX * tilesY;
	RenderParams p;
	memset(&p, 0, sizeof(p));
	p.width = c->width; p.height = c->height;
	p.tilesX = tilesX;
	p.invTilesX = 1.0f / (float) tilesX;
	p.exactDiv = (double) totalTiles >= 4e6 ? 1 : 0;
	p.numOwnedTiles = totalTiles > brank ? (totalTiles - brank + bcount - 1) / bcount : 0;
	p.taskStride = bcount;
	p.taskOffset = brank;
	p.out = dev;
	gatherOwnedKernel<<<c->numSMs * 4, 256, 0, st>>>(p, c->dFrame);
	CUDA_TRY(cudaGetLastError());
	return FRAY_GPU_OK;
}

int fray_gpu_host_register(void* host, size_t bytes)
{
	if (!host || !bytes) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaHostRegister(host, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped));
	return FRAY_GPU_OK;
}

int fray_gpu_host_unregister(void* host)
{
	if (!host) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaHostUnregister(host));
	return FRAY_GPU_OK;
}

int fray_gpu_resolve_device(FrayGpuCtx* c, const void* d_sum, void* d_rgb, int32_t spp, void* cuda_stream)
{
	if (!c || !d_sum || !d_rgb || spp < 1) return fail(FRAY_GPU_EINVAL, "bad argument");
	CUDA_TRY(cudaSetDevice(c->device));
	const size_t n = (size_t) c->width * c->height * 3;
	cudaStream_t st = cuda_stream ? (cudaStream_t) cuda_stream : c->stream;
	resolveKernel<<<c->numSMs * 4, 256, 0, st>>>((const float*) d_sum, (float*) d_rgb, n, (float) spp);
	CUDA_TRY(cudaGetLastError());
	return FRAY_GPU_OK;
}

// ---- random-stream self-test ---------------------------------------------------------------------------
struct RngProbeParams {
	uint32_t seed, pixel, sample, branch;
	uint32_t roundKeys[10];
	int mode, n;
	uint32_t* out;
	unsigned char* drawn;
};

// 128 threads as in the render kernels (the ring layout depends on it); thread 77 runs the probe, the others run decoy
// streams through the same ring so that a layout mistake would show
__global__ void rngProbeKernel(const RngProbeParams p)
{
	__shared__ uint32_t ring[FRAY_RNG_RING_WORDS * 128];
	const bool probe = threadIdx.x == 77;
	const uint32_t pixel = probe ? p.pixel : p.pixel + 1000u + threadIdx.x;
	int i = 0;
	auto put = [&](uint32_t v) { if (probe) { p.out[i] = v; p.drawn[i] = 1; } i++; };
	if (p.mode == 0) {
		RngT<FRAY_RNG_KEYED> rng;
		rng.keys = p.roundKeys;
		rng.init(p.seed, pixel, p.sample, p.branch);
		while (i < p.n) put(rng.next());
		return;
	}
	RngRing rng;
	rng.attach((uint32_t) __cvta_generic_to_shared(ring + threadIdx.x), p.roundKeys);
	rng.init(p.seed, pixel, p.sample, p.branch);
	if (p.mode == 1) {
		while (i < p.n) {
			const int group = min(12, p.n - i);
			rng.ensure((uint32_t) group);
			for (int k = 0; k < group; k++) put(rng.next());
		}
		return;
	}
	rng.ensure(2);
	for (int k = 0; k < 2 && i < p.n; k++) put(rng.next());
	while (i + 8 <= p.n) {
		rng.ensure(8);
		rng.skip(2); i += 2;
		for (int k = 0; k < 6; k++) put(rng.next());
	}
}

int fray_gpu_rng_probe(int device, uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t branch, int mode, int n, uint32_t* out, unsigned char* drawn)
{
	if (!out || !drawn || n < 0 || mode < 0 || mode > 2) return fail(FRAY_GPU_EINVAL, "bad argument");
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(FRAY_GPU_ENODEVICE, "no CUDA device available");
	}
	if (device < 0 || device >= ndev) return fail(FRAY_GPU_EINVAL, "device ordinal out of range");
	CUDA_TRY(cudaSetDevice(device));
	memset(out, 0, sizeof(uint32_t) * (size_t) n);
	memset(drawn, 0, (size_t) n);
	if (n == 0) return FRAY_GPU_OK;
	RngProbeParams p;
	p.seed = seed; p.pixel = pixel; p.sample = sample; p.branch = branch; p.mode = mode; p.n = n;
	philoxRoundKeys(seed, p.roundKeys);
	CUDA_TRY(cudaMalloc(&p.out, sizeof(uint32_t) * (size_t) n));
	CUDA_TRY(cudaMalloc(&p.drawn, (size_t) n));
	CUDA_TRY(cudaMemset(p.out, 0, sizeof(uint32_t) * (size_t) n));
	CUDA_TRY(cudaMemset(p.drawn, 0, (size_t) n));
	rngProbeKernel<<<1, 128>>>(p);
	CUDA_TRY(cudaGetLastError());
	CUDA_TRY(cudaMemcpy(out, p.out, sizeof(uint32_t) * (size_t) n, cudaMemcpyDeviceToHost));
	CUDA_TRY(cudaMemcpy(drawn, p.drawn, (size_t) n, cudaMemcpyDeviceToHost));
	cudaFree(p.out);
	cudaFree(p.drawn);
	return FRAY_GPU_OK;
}

int fray_gpu_measure_peaks(int device, double ms, double* fp32_tflops, double* l2_gbs)
{
	int ndev = 0;
	if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
		cudaGetLastError();
		return fail(FRAY_GPU_ENODEVICE, "no CUDA device available");
	}
	if (device < 0 || device >= ndev) return fail(FRAY_GPU_EINVAL, "device ordinal out of range");
	CUDA_TRY(cudaSetDevice(device));
	cudaDeviceProp prop;
	CUDA_TRY(cudaGetDeviceProperties(&prop, device));
	cudaEvent_t e0, e1;
	CUDA_TRY(cudaEventCreate(&e0));
	CUDA_TRY(cudaEventCreate(&e1));
	float* sink = nullptr;
	CUDA_TRY(cudaMalloc(&sink, 256));
	float t = 0;
	if (fp32_tflops) {
		const int blocks = prop.multiProcessorCount * 8, threads = 256;
		int iters = 2000;
		for (int round = 0; round < 3; round++) { // calibrate to ~ms
			CUDA_TRY(cudaEventRecord(e0));
			ffmaPeakKernel<<<blocks, threads>>>(sink, iters, 1.0000001f, 1e-9f);
			CUDA_TRY(cudaEventRecord(e1));
			CUDA_TRY(cudaEventSynchronize(e1));
			CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
			if (round < 2 && t > 0) iters = (int) std::min(2e6, std::max(100.0, iters * ms / t));
		}
		*fp32_tflops = 2.0 * 8 * 16 * (double) iters * blocks * threads / (t * 1e-3) / 1e12;
	}
	if (l2_gbs) {
		const size_t bytes = 32u << 20, n = bytes / sizeof(float4);
		float4* buf = nullptr;
		CUDA_TRY(cudaMalloc(&buf, bytes));
		CUDA_TRY(cudaMemset(buf, 0, bytes));
		int passes = 20;
		for (int round = 0; round < 3; round++) {
			CUDA_TRY(cudaEventRecord(e0));
			l2ReadKernel<<<prop.multiProcessorCount * 8, 256>>>(buf, n, passes, sink);
			CUDA_TRY(cudaEventRecord(e1));
			CUDA_TRY(cudaEventSynchronize(e1));
			CUDA_TRY(cudaEventElapsedTime(&t, e0, e1));
			if (round < 2 && t > 0) passes = (int) std::min(5000.0, std::max(4.0, passes * ms / t));
		}
		*l2_gbs = (double) bytes * passes / (t * 1e-3) / 1e9;
		cudaFree(buf);
	}
	cudaFree(sink);
	cudaEventDestroy(e0);
	cudaEventDestroy(e1);
	CUDA_TRY(cudaGetLastError());
	return FRAY_GPU_OK;
}

int fray_gpu_sync(FrayGpuCtx* c, FrayGpuStats* stats)
{
	if (!c) return fail(FRAY_GPU_EINVAL, "null argument");
	CUDA_TRY(cudaSetDevice(c->device));
	return fetchStats(c, stats);
}


// ---- several GPUs, one process ---------------------------------------------------------------------------------------------
// The reference spreads a frame over its host threads with one call, pool.run(&worker, numThreads) (src/main.cpp:402-404); this
// is that call for the GPUs of a node. One context per device, a host thread per context (enqueueing a share costs ~25 us of
// API calls; in sequence that would be a fifth of an 8-GPU frame), device 0 owns the frame:
//   tile split    share d = the 8x4 tiles t with t % n == d, all samples; its kernels store the finished pixels straight into
//                 device 0's frame through peer access (NVLink), every share with the samples-per-item the single GPU would
//                 use, so the frame is bit-identical to the single-GPU frame;
//   sample split  share d = samples [d*spp/n, (d+1)*spp/n) of every pixel, written as partial sums into slot d of a buffer on
//                 device 0 (peer stores again); device 0 adds the slots in order and divides. Equal to the single-GPU frame up
//                 to the FP32 summation order of the path-traced kernels, exactly for the wavefront (integer accumulators).
// No NCCL here: within one process, events order "share written" before "device 0 reads". (The one-process-per-GPU form with
// an NCCL reduce is fray_b200/dist.py.)
struct FrayGpuMulti {
	std::vector<FrayGpuCtx*> ctx;
	int width = 0, height = 0;
	float* dPartials = nullptr; // device of ctx[0]: [n][height][width][3]
	std::vector<cudaEvent_t> done;
	// the job of the current frame, one entry per context
	std::vector<FrayGpuFrame> frames;
	std::vector<float*> outs;
	std::vector<int> rcs;
	std::vector<std::string> errs;
	std::vector<FrayGpuStats> stats;
	// worker threads
	std::vector<std::thread> workers;
	std::mutex m;
	std::condition_variable cvGo, cvDone;
	unsigned long long generation = 0;
	int enqueued = 0, finished = 0;
	bool quit = false;
};

__global__ void sumPartialsKernel(const float* __restrict__ partials, int n, size_t frameFloats, float* __restrict__ out, float divisor)
{
	for (size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x; i < frameFloats; i += (size_t) gridDim.x * blockDim.x) {
		float sum = 0;
		for (int d = 0; d < n; d++) sum += partials[(size_t) d * frameFloats + i]; // in share order: deterministic
		out[i] = sum / divisor;
	}
}

static void multiWorker(FrayGpuMulti* mg, int d)
{
	unsigned long long seen = 0;
	for (;;) {
		{
			std::unique_lock<std::mutex> lk(mg->m);
			mg->cvGo.wait(lk, [&] { return mg->quit || mg->generation != seen; });
			if (mg->quit) return;
			seen = mg->generation;
		}
		FrayGpuCtx* c = mg->ctx[d];
		int rc = renderInto(c, &mg->frames[d], mg->outs[d], c->stream, true);
		if (rc == FRAY_GPU_OK && c->waveLast) rc = waveFinish(c); // a full wavefront queue: the share is rendered again before anyone reads it
		if (rc == FRAY_GPU_OK && cudaEventRecord(mg->done[d], c->stream) != cudaSuccess) rc = fail(FRAY_GPU_ECUDA, "cudaEventRecord failed");
		if (rc != FRAY_GPU_OK) mg->errs[d] = g_lastError;
		mg->rcs[d] = rc;
		{
			std::lock_guard<std::mutex> lk(mg->m);
			mg->enqueued++;
		}
		mg->cvDone.notify_all();
		if (rc == FRAY_GPU_OK) { // the share's statistics, while device 0 already assembles the frame
			rc = fetchStats(c, &mg->stats[d]);
			if (rc != FRAY_GPU_OK) { mg->errs[d] = g_lastError; mg->rcs[d] = rc; }
		}
		{
			std::lock_guard<std::mutex> lk(mg->m);
			mg->finished++;
		}
		mg->cvDone.notify_all();
	}
}

void fray_gpu_multi_destroy(FrayGpuMulti* mg)
{
	if (!mg) return;
	{
		std::lock_guard<std::mutex> lk(mg->m);
		mg->quit = true;
	}
	mg->cvGo.notify_all();
	for (std::thread& t: mg->workers) t.join();
	for (size_t d = 0; d < mg->ctx.size(); d++) {
		if (mg->ctx[d]) {
			cudaSetDevice(mg->ctx[d]->device);
			if (d < mg->done.size() && mg->done[d]) cudaEventDestroy(mg->done[d]);
		}
	}
	if (!mg->ctx.empty() && mg->ctx[0]) {
		cudaSetDevice(mg->ctx[0]->device);
		cudaFree(mg->dPartials);
	}
	for (FrayGpuCtx* c: mg->ctx) fray_gpu_destroy(c);
	delete mg;
}

int fray_gpu_multi_create(const FrayGpuScene* scene, int n_devices, const int* devices, int precision, FrayGpuMulti** out)
{
	if (!scene || !out || n_devices < 1 || n_devices > 64) return fail(FRAY_GPU_EINVAL, "bad argument");
	*out = nullptr;
	FrayGpuMulti* mg = new FrayGpuMulti;
	mg->width = scene->settings.frame_width;
	mg->height = scene->settings.frame_height;
	for (int d = 0; d < n_devices; d++) {
		FrayGpuCtx* c = nullptr;
		const int rc = fray_gpu_create(scene, devices ? devices[d] : d, precision, &c);
		if (rc != FRAY_GPU_OK) {
			const std::string msg = g_lastError;
			fray_gpu_multi_destroy(mg);
			return fail(rc, msg);
		}
		mg->ctx.push_back(c);
	}
	const int dev0 = mg->ctx[0]->device;
	for (int d = 0; d < n_devices; d++) {
		FrayGpuCtx* c = mg->ctx[d];
		cudaSetDevice(c->device);
		cudaEvent_t ev = nullptr;
		if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
			fray_gpu_multi_destroy(mg);
			return fail(FRAY_GPU_ECUDA, "cudaEventCreate failed");
		}
		mg->done.push_back(ev);
		if (c->device != dev0) { // this device's kernels store into device 0's memory
			int can = 0;
			cudaDeviceCanAccessPeer(&can, c->device, dev0);
			const cudaError_t e = can ? cudaDeviceEnablePeerAccess(dev0, 0) : cudaErrorPeerAccessUnsupported;
			if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
				cudaGetLastError();
				fray_gpu_multi_destroy(mg);
				return fail(FRAY_GPU_EUNSUPPORTED, "no peer access from device " + std::to_string(c->device) + " to device " + std::to_string(dev0));
			}
			cudaGetLastError();
		}
	}
	mg->frames.resize(n_devices);
	mg->outs.resize(n_devices);
	mg->rcs.assign(n_devices, 0);
	mg->errs.resize(n_devices);
	mg->stats.resize(n_devices);
	for (int d = 0; d < n_devices; d++) mg->workers.emplace_back(multiWorker, mg, d);
	*out = mg;
	return FRAY_GPU_OK;
}

int fray_gpu_multi_device_count(const FrayGpuMulti* mg) { return mg ? (int) mg->ctx.size() : 0; }

int fray_gpu_multi_update_camera(FrayGpuMulti* mg, const FrayGpuCamera* cam)
{
	if (!mg || !cam) return fail(FRAY_GPU_EINVAL, "null argument");
	for (FrayGpuCtx* c: mg->ctx) {
		const int rc = fray_gpu_update_camera(c, cam);
		if (rc != FRAY_GPU_OK) return rc;
	}
	return FRAY_GPU_OK;
}

int fray_gpu_multi_render(FrayGpuMulti* mg, const FrayGpuFrame* frame, int split, float* rgb_out, FrayGpuStats* stats)
{
	if (!mg || !frame || !rgb_out) return fail(FRAY_GPU_EINVAL, "null argument");
	if (frame->bucket_count > 0 || (frame->flags & (FRAY_FRAME_OWNED_ONLY | FRAY_FRAME_SAMPLE_RANGE)) || frame->sample_begin != 0 || frame->sample_end != 0)
		return fail(FRAY_GPU_EINVAL, "fray_gpu_multi_render splits the frame itself: pass a whole frame");
	const int n = (int) mg->ctx.size();
	FrayGpuCtx* c0 = mg->ctx[0];
	if (frame->mode == FRAY_RENDER_PREPASS) return fray_gpu_render(c0, frame, rgb_out, stats); // a few thousand rays: one GPU
	const int spp = frame->spp > 0 ? frame->spp : c0->defaultSpp;
	const bool gi = c0->precision == FRAY_GPU_FP32 ? c0->sc32.gi : c0->sc64.gi;
	if (split == FRAY_GPU_SPLIT_AUTO) split = (frame->mode == FRAY_RENDER_BEAUTY && spp >= 8 * n) ? FRAY_GPU_SPLIT_SAMPLES : FRAY_GPU_SPLIT_TILES;
	if (split != FRAY_GPU_SPLIT_TILES && split != FRAY_GPU_SPLIT_SAMPLES) return fail(FRAY_GPU_EINVAL, "unknown split");
	if (split == FRAY_GPU_SPLIT_SAMPLES && (spp < n || frame->mode != FRAY_RENDER_BEAUTY)) split = FRAY_GPU_SPLIT_TILES;
	const size_t frameFloats = (size_t) mg->width * mg->height * 3;
	CUDA_TRY(cudaSetDevice(c0->device));
	if (split == FRAY_GPU_SPLIT_SAMPLES && !mg->dPartials) CUDA_TRY(cudaMalloc(&mg->dPartials, frameFloats * sizeof(float) * n));
	// tile shares with the chunks of the whole frame on one GPU: every pixel is then summed exactly as there
	double chunkPixels = 0;
	if (split == FRAY_GPU_SPLIT_TILES && n > 1 && frame->mode == FRAY_RENDER_BEAUTY && !(frame->flags & FRAY_GPU_MULTI_FAST)) {
		const int tilesX = (mg->width + FRAY_TILE_W - 1) / FRAY_TILE_W, tilesY = (mg->height + FRAY_TILE_H - 1) / FRAY_TILE_H;
		chunkPixels = (double) tilesX * tilesY * 32.0;
	}
	for (int d = 0; d < n; d++) {
		FrayGpuFrame f = *frame;
		f.flags &= ~FRAY_GPU_MULTI_FAST;
		f.spp = spp;
		if (split == FRAY_GPU_SPLIT_TILES) {
			f.bucket_rank = d;
			f.bucket_count = n;
			if (n > 1) f.flags |= FRAY_FRAME_OWNED_ONLY;
			mg->outs[d] = c0->dFrame;
		} else {
			f.sample_begin = (int) ((long long) d * spp / n);
			f.sample_end = (int) ((long long) (d + 1) * spp / n);
			f.flags |= FRAY_FRAME_SUM | FRAY_FRAME_SAMPLE_RANGE;
			mg->outs[d] = mg->dPartials + (size_t) d * frameFloats;
		}
		mg->frames[d] = f;
		mg->ctx[d]->chunkPixels = chunkPixels;
	}
	{
		std::lock_guard<std::mutex> lk(mg->m);
		mg->enqueued = mg->finished = 0;
		mg->generation++;
	}
	mg->cvGo.notify_all();
	{
		std::unique_lock<std::mutex> lk(mg->m);
		mg->cvDone.wait(lk, [&] { return mg->enqueued == n; });
	}
	int rc = FRAY_GPU_OK;
	for (int d = 0; d < n && rc == FRAY_GPU_OK; d++)
		if (mg->rcs[d] != FRAY_GPU_OK) rc = fail(mg->rcs[d], "share " + std::to_string(d) + ": " + mg->errs[d]);
	if (rc == FRAY_GPU_OK) {
		// device 0: wait for every share, assemble, copy out
		cudaSetDevice(c0->device);
		cudaError_t e = cudaSuccess;
		for (int d = 0; d < n && e == cudaSuccess; d++) e = cudaStreamWaitEvent(c0->stream, mg->done[d], 0);
		const size_t bytes = frameFloats * sizeof(float);
		float* mapped = mappedHostPointer(rgb_out); // a page-locked caller buffer: the GPU writes it directly
		const bool pinned = mapped != nullptr;
		if (!mapped) mapped = mappedHostPointer(c0->hStaging);
		bool copied = false;
		if (e == cudaSuccess && split == FRAY_GPU_SPLIT_SAMPLES) {
			// add the shares and hand the frame over in one pass: the sums are stored straight into host memory over PCIe
			// (no frame on device 0, no separate device-to-host copy)
			float* target = mapped ? mapped : c0->dFrame;
			sumPartialsKernel<<<c0->numSMs * 4, 256, 0, c0->stream>>>(mg->dPartials, n, frameFloats, target, (frame->flags & FRAY_FRAME_SUM) ? 1.0f : (float) spp);
			e = cudaGetLastError();
			copied = mapped != nullptr;
		}
		if (e == cudaSuccess && !copied) e = cudaMemcpyAsync(pinned ? rgb_out : c0->hStaging, c0->dFrame, bytes, cudaMemcpyDeviceToHost, c0->stream);
		if (e == cudaSuccess) e = cudaStreamSynchronize(c0->stream);
		if (e == cudaSuccess && !pinned) memcpy(rgb_out, c0->hStaging, bytes);
		if (e != cudaSuccess) rc = fail(FRAY_GPU_ECUDA, std::string("assembling the frame: ") + cudaGetErrorString(e));
	}
	{
		std::unique_lock<std::mutex> lk(mg->m);
		mg->cvDone.wait(lk, [&] { return mg->finished == n; });
	}
	for (int d = 0; d < n; d++) mg->ctx[d]->chunkPixels = 0;
	for (int d = 0; d < n && rc == FRAY_GPU_OK; d++)
		if (mg->rcs[d] != FRAY_GPU_OK) rc = fail(mg->rcs[d], "share " + std::to_string(d) + ": " + mg->errs[d]);
	if (rc == FRAY_GPU_OK && stats) {
		memset(stats, 0, sizeof(*stats));
		for (int d = 0; d < n; d++) {
			stats->rays += mg->stats[d].rays;
			stats->primary_rays += mg->stats[d].primary_rays;
			stats->shadow_rays += mg->stats[d].shadow_rays;
			stats->kernel_launches += mg->stats[d].kernel_launches;
			stats->device_ms = std::max(stats->device_ms, mg->stats[d].device_ms); // the slowest share
		}
		if (split == FRAY_GPU_SPLIT_SAMPLES) stats->kernel_launches += 1;
	}
	return rc;
}

} // extern "C"
