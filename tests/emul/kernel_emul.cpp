// kernel_emul.cpp -- TEST-ONLY host compilation of the per-ray device code (fray_b200/csrc/core.cuh).
//
// This container has no GPU. To debug the device algorithms (iterative path tracer, Whitted task stack, stack-based KD
// descent, CSG lists, FP32 epsilon policy) before spending GPU minutes, the templated __host__ __device__ core is also
// compiled here with g++ and driven by a plain loop over pixels and samples. It is NOT part of the product:
// libfray_gpu.so neither links nor loads this file, the C ABI (include/fray_gpu.h) has no path to it, and the
// `-m gpu` tests, smoke() and bench.py never use it. tests/test_emul_vs_oracle.py uses it as an early warning only.
#include <atomic>
#include <thread>
#include <vector>
#include <string>

#include "../../fray_b200/csrc/scene_image.h"
#include "../../fray_b200/csrc/wave.cuh"

using namespace fray;

#if defined(FRAY_DEBUG_TRACE)
namespace fray { thread_local int g_frayTrace = 0; }
static int tracePixel(int which) { const char* e = getenv(which ? "FRAY_TRACE_Y" : "FRAY_TRACE_X"); return e ? atoi(e) : -1; }
#endif

template <typename R, int F>
static void renderRows(const DScene<R>& sc, const FlatTab& ft, const FrayGpuFrame& fr, int W, int H, int spp, int s0, int s1, float* out,
                       std::atomic<int>& nextRow, RayCounters& total)
{
	WhittedState<R>* ws = new WhittedState<R>;
	ws->overflow = 0;
	ws->rootPending = false;
	ws->sp = 0;
	RayCounters cnt = { 0, 0, 0 };
	const int bcount = fr.bucket_count > 0 ? fr.bucket_count : 1, brank = fr.bucket_count > 0 ? fr.bucket_rank : 0;
	for (;;) {
		int y = nextRow++;
		if (y >= H) break;
		for (int x = 0; x < W; x++) {
			float* o = out + 3 * ((size_t) y * W + x);
			const int BW = (W - 1) / 48 + 1, bx = x / 48, by = y / 48;
			const int bucket = by * BW + ((by % 2 == 0) ? bx : (BW - 1 - bx));
			if (bucket % bcount != brank) { o[0] = o[1] = o[2] = 0; continue; }
			if (fr.mode == FRAY_RENDER_AOV) {
				Ray<R> ray = screenRay(sc.cam, (R) x, (R) y, 0);
				int node, light;
				Hit<R> h;
				closestHit<R, F>(sc, ft, ray, node, light, h);
				o[0] = light >= 0 ? (float) (-2 - light) : (float) node;
				o[1] = (light < 0 && node >= 0 && h.tri >= 0) ? (float) (h.tri - sc.meshes[h.mesh].firstTri) : -1.0f;
				o[2] = (float) h.dist;
				continue;
			}
			Col sum(0, 0, 0);
#if defined(FRAY_DEBUG_TRACE)
			g_frayTrace = (x == tracePixel(0) && y == tracePixel(1));
			if (g_frayTrace) printf("pixel %d %d (%s)\n", x, y, Num<R>::kExact ? "fp64" : "fp32");
#endif
			for (int i = s0; i < s1; i++) sum = sum + renderSample<R, F>(sc, ft, fr.seed, x, y, W, i, ws, cnt);
			if (!(fr.flags & FRAY_FRAME_SUM)) sum = sum / (float) spp;
			o[0] = sum.r; o[1] = sum.g; o[2] = sum.b;
		}
	}
	static std::atomic_flag lock = ATOMIC_FLAG_INIT;
	while (lock.test_and_set()) {}
	total.rays += cnt.rays; total.primary += cnt.primary; total.shadow += cnt.shadow;
	lock.clear();
	delete ws;
}

template <typename R>
static int renderT(const FrayGpuScene* scene, const FrayGpuFrame* fr, float* out, FrayGpuStats* stats, int threads)
{
	SceneImage<R> img;
	std::string err;
	if (!img.build(*scene, err)) { fprintf(stderr, "emul: %s\n", err.c_str()); return -1; }
	DScene<R> sc = img.bind(img.blob.data());
	const int W = scene->settings.frame_width, H = scene->settings.frame_height;
	const int spp = fr->spp > 0 ? fr->spp : samplesPerPixel(*scene);
	int s0 = fr->sample_begin, s1 = fr->sample_end;
	if (s0 == 0 && s1 == 0 && !(fr->flags & FRAY_FRAME_SAMPLE_RANGE)) s1 = spp;
	if (threads < 1) threads = (int) std::thread::hardware_concurrency();
	std::atomic<int> nextRow(0);
	RayCounters total = { 0, 0, 0 };
	std::vector<std::thread> pool;
	FlatTab ft;
	ft.polys = sc.flatPolys;
	ft.info = sc.flatInfo;
	ft.spheres = sc.flatPolys + FRAY_FLAT_POLY_VEC * sc.numFlatTotal;
	ft.hexes = ft.spheres + sc.numFlatSpheres;
	ft.polys2 = ft.hexes + FRAY_HEX_VEC * sc.numFlatHex;
	const int need = img.features;
	auto run = [&]() { // the same variant selection as VariantDispatch in render_kernels.cuh
		if (Variants<R>::count > 0 && (need & ~Variants<R>::mask(0)) == 0) renderRows<R, Variants<R>::mask(0)>(sc, ft, *fr, W, H, spp, s0, s1, out, nextRow, total);
		else if (Variants<R>::count > 1 && (need & ~Variants<R>::mask(1)) == 0) renderRows<R, Variants<R>::mask(1)>(sc, ft, *fr, W, H, spp, s0, s1, out, nextRow, total);
		else if (Variants<R>::count > 2 && (need & ~Variants<R>::mask(2)) == 0) renderRows<R, Variants<R>::mask(2)>(sc, ft, *fr, W, H, spp, s0, s1, out, nextRow, total);
		else if (Variants<R>::count > 3 && (need & ~Variants<R>::mask(3)) == 0) renderRows<R, Variants<R>::mask(3)>(sc, ft, *fr, W, H, spp, s0, s1, out, nextRow, total);
		else if (Variants<R>::count > 4 && (need & ~Variants<R>::mask(4)) == 0) renderRows<R, Variants<R>::mask(4)>(sc, ft, *fr, W, H, spp, s0, s1, out, nextRow, total);
		else renderRows<R, Variants<R>::mask(5)>(sc, ft, *fr, W, H, spp, s0, s1, out, nextRow, total);
	};
	for (int t = 1; t < threads; t++) pool.emplace_back(run);
	run();
	for (auto& t: pool) t.join();
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		stats->rays = total.rays; stats->primary_rays = total.primary; stats->shadow_rays = total.shadow;
	}
	return 0;
}

// ---- the wavefront Whitted integrator (fray_b200/csrc/wave.cuh) with std::vector queues ------------------------------------
struct HostSink {
	const DScene<float>* sc;
	std::vector<WaveRay>* next;
	std::vector<WaveLit>* lits;
	long long* acc;
	void add(const WaveRay& r, Col c)
	{
		c = waveEyeColor(*sc, r.eye, c);
		acc[3 * (size_t) r.pixel] += waveFixed(c.r); acc[3 * (size_t) r.pixel + 1] += waveFixed(c.g); acc[3 * (size_t) r.pixel + 2] += waveFixed(c.b);
	}
	void ray(const WaveRay& c) { next->push_back(c); }
	void lit(const WaveLit& L) { lits->push_back(L); }
};

template <int F, int D>
static void renderWaveT(const DScene<float>& sc, const FlatTab& ft, const FrayGpuFrame& fr, int W, int H, int spp, int s0, int s1, float* out, RayCounters& cnt)
{
	typedef KdShortStack<KdStoreArray<D>, D> Stack;
	Stack stk;
	stk.reset();
	uint32_t keys[10];
	philoxRoundKeys(fr.seed, keys);
	const bool randomOffsets = sc.cam.dof || sc.gi;
	const int bcount = fr.bucket_count > 0 ? fr.bucket_count : 1, brank = fr.bucket_count > 0 ? fr.bucket_rank : 0;
	std::vector<long long> acc((size_t) W * H * 3, 0);
	std::vector<WaveRay> cur, next, rightEyes;
	std::vector<WaveLit> lits;
	std::vector<char> owned((size_t) W * H, 0);
	for (int y = 0; y < H; y++)
		for (int x = 0; x < W; x++) {
			const int BW = (W - 1) / 48 + 1, bx = x / 48, by = y / 48;
			const int bucket = by * BW + ((by % 2 == 0) ? bx : (BW - 1 - bx));
			if (bucket % bcount != brank) continue;
			owned[(size_t) y * W + x] = 1;
			for (int s = s0; s < s1; s++) {
				WaveRay l, r;
				bool stereo;
				wavePrimary<F>(sc, keys, fr.seed, x, y, W, s, randomOffsets, l, r, stereo);
				cur.push_back(l);
				if (stereo) rightEyes.push_back(r);
				cnt.primary += stereo ? 2 : 1;
			}
		}
	const bool stereo = !rightEyes.empty();
	HostSink sink{ &sc, &next, &lits, acc.data() };
	for (int wave = 0; !cur.empty(); wave++) {
		std::vector<WaveHit> hits(cur.size());
		for (size_t i = 0; i < cur.size(); i++) {
			Ray<float> ray;
			ray.start = cur[i].start; ray.dir = cur[i].dir;
			waveClosest<F>(sc, ft, ray, cur[i].origin, stk, hits[i]);
			cnt.rays++;
		}
		next.clear();
		lits.clear();
		for (size_t i = 0; i < cur.size(); i++) {
			uint32_t count = cur[i].count;
			waveShade<F>(sc, ft, cur[i], hits[i], keys, fr.seed, count, sink);
			if (wave == 0 && stereo) { // the right eye goes on in the stream where the left eye's light loops stopped
				rightEyes[i].count = count;
				sink.ray(rightEyes[i]);
			}
		}
		for (const WaveLit& L: lits) {
			WaveRay tag;
			tag.pixel = L.pixel; tag.eye = L.eye;
			unsigned shot = 0;
			sink.add(tag, waveLightLoop<F>(sc, ft, L, keys, fr.seed, stk, shot));
			cnt.rays += shot;
			cnt.shadow += shot;
		}
		cur.swap(next);
	}
	for (size_t p = 0; p < (size_t) W * H; p++)
		for (int k = 0; k < 3; k++) {
			float v = owned[p] ? (float) ((double) acc[3 * p + k] / 4294967296.0) : 0.0f;
			if (!(fr.flags & FRAY_FRAME_SUM)) v /= (float) spp;
			out[3 * p + k] = v;
		}
}

// precision code 2: fast precision through the wavefront stages; kdShort: entries of the KD short stack (2 forces restarts)
static int renderWave(const FrayGpuScene* scene, const FrayGpuFrame* fr, float* out, FrayGpuStats* stats, int kdShort)
{
	SceneImage<float> img;
	std::string err;
	if (!img.build(*scene, err)) { fprintf(stderr, "emul: %s\n", err.c_str()); return -1; }
	if (img.offsets.gi || (img.features & FRAY_F_CSG) || fr->mode != FRAY_RENDER_BEAUTY) return -2; // not what the wavefront path renders
	DScene<float> sc = img.bind(img.blob.data());
	const int W = scene->settings.frame_width, H = scene->settings.frame_height;
	const int spp = fr->spp > 0 ? fr->spp : samplesPerPixel(*scene);
	int s0 = fr->sample_begin, s1 = fr->sample_end;
	if (s0 == 0 && s1 == 0 && !(fr->flags & FRAY_FRAME_SAMPLE_RANGE)) s1 = spp;
	FlatTab ft;
	ft.polys = sc.flatPolys;
	ft.info = sc.flatInfo;
	ft.spheres = sc.flatPolys + FRAY_FLAT_POLY_VEC * sc.numFlatTotal;
	ft.hexes = ft.spheres + sc.numFlatSpheres;
	ft.polys2 = ft.hexes + FRAY_HEX_VEC * sc.numFlatHex;
	RayCounters cnt = { 0, 0, 0 };
	constexpr int kAll = Variants<float>::mask(4) | FRAY_F_TWOSIDED; // everything but CSG
	if (kdShort == 2) renderWaveT<kAll, 2>(sc, ft, *fr, W, H, spp, s0, s1, out, cnt);
	else renderWaveT<kAll, FRAY_KD_SHORT>(sc, ft, *fr, W, H, spp, s0, s1, out, cnt);
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		stats->rays = cnt.rays; stats->primary_rays = cnt.primary; stats->shadow_rays = cnt.shadow;
	}
	return 0;
}

extern "C" int fray_emul_render_wave(const FrayGpuScene* scene, const FrayGpuFrame* fr, float* out, FrayGpuStats* stats, int kdShort)
{
	return renderWave(scene, fr, out, stats, kdShort);
}

extern "C" int fray_emul_render(const FrayGpuScene* scene, const FrayGpuFrame* fr, float* out, FrayGpuStats* stats, int precision, int threads)
{
	return precision == FRAY_GPU_FP64 ? renderT<double>(scene, fr, out, stats, threads) : renderT<float>(scene, fr, out, stats, threads);
}
