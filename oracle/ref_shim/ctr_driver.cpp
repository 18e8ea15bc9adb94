/*
 * ctr_driver.cpp -- frame driver that runs the UNMODIFIED reference renderer with the
 * counter-based RNG contract (oracle/fray_rng.h). TEST INFRASTRUCTURE ONLY.
 *
 * What is the reference's and what is ours:
 *   reference (compiled from /root/reference/src, untouched): scene parser, Scene::beginRender /
 *     beginFrame, Camera, Node/geometry intersection, KD build + traversal, shaders, lights,
 *     environment, raytrace(), pathtrace(), raytraceSinglePixel()  (src/main.cpp:64-321).
 *   ours: this file replaces RendMT::entry + render() (src/main.cpp:323-405) with the same
 *     per-pixel sample loop, except that (a) the RNG stream is re-keyed to (pixel, sample)
 *     before each sample, (b) rows are split statically over std::threads, (c) no prepass and
 *     no display. main.cpp's own main() is renamed at link time (objcopy, see oracle/Makefile).
 *
 * Usage: fray_ref_ctr scene.fray out.f32 [--seed N] [--threads N] [--aov out.aov]
 *   out.f32 : int32 w, int32 h, float32 rgb[h][w][3]
 *   out.aov : int32 w, int32 h, then per pixel {int32 node, float64 dist} for the un-jittered
 *             pin-hole ray through (x, y)  (node = index in scene.nodes, -1 = miss,
 *             -2-k = light k hit first)
 * Prints "Render took %.3fs" like src/main.cpp:520 (render loop only).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <thread>
#include <vector>
#include <atomic>
#include <chrono>
#include <algorithm>

#include "sdl.h"
#include "color.h"
#include "vector.h"
#include "camera.h"
#include "geometry.h"
#include "scene.h"
#include "lights.h"
#include "random_generator.h"
#include "cxxptl-sdl.h"

extern Color vfb[VFB_MAX_SIZE][VFB_MAX_SIZE];               // src/main.cpp:53
Color raytraceSinglePixel(double x, double y, Random& rnd); // src/main.cpp:304
void fray_ctr_rekey(unsigned pixel, unsigned sample);       // ctr_random.cpp

static const double kOffsets[5][2] = { {0, 0}, {0.6, 0}, {0.3, 0.3}, {0, 0.6}, {0.6, 0.6} }; // src/main.cpp:55-61

int main(int argc, char** argv)
{
	if (argc < 3) {
		fprintf(stderr, "Usage: fray_ref_ctr scene.fray out.f32 [--seed N] [--threads N] [--aov file]\n");
		return -1;
	}
	unsigned seed = 42;
	int threads = (int) std::thread::hardware_concurrency();
	const char* aovFile = nullptr;
	for (int i = 3; i < argc; i++) {
		if (!strcmp(argv[i], "--seed") && i + 1 < argc) seed = (unsigned) strtoul(argv[++i], 0, 10);
		else if (!strcmp(argv[i], "--threads") && i + 1 < argc) threads = atoi(argv[++i]);
		else if (!strcmp(argv[i], "--aov") && i + 1 < argc) aovFile = argv[++i];
	}
	if (threads < 1) threads = 1;

	initRandom(seed);
	if (!scene.parseScene(argv[1])) return -3;
	initGraphics(scene.settings.frameWidth, scene.settings.frameHeight, false);
	scene.beginRender();
	scene.beginFrame();

	const int W = frameWidth(), H = frameHeight();
	// samples per pixel, src/main.cpp:395-400
	int spp = 5;
	if (!scene.settings.wantAA) spp = 1;
	if (scene.camera->dof) spp = std::max(spp, scene.camera->numDOFSamples);
	if (scene.settings.gi) spp = std::max(spp, scene.settings.numPaths);
	const bool randomOffsets = scene.camera->dof || scene.settings.gi;

	auto t0 = std::chrono::steady_clock::now();
	std::atomic<int> nextRow(0);
	auto worker = [&]() {
		Random& rnd = getRandomGen();
		for (;;) {
			int y = nextRow++;
			if (y >= H) return;
			for (int x = 0; x < W; x++) {
				Color avg(0, 0, 0);
				for (int i = 0; i < spp; i++) {
					fray_ctr_rekey((unsigned) (y * W + x), (unsigned) i);
					float ox, oy;
					if (randomOffsets) { ox = rnd.randfloat(); oy = rnd.randfloat(); }
					else { ox = kOffsets[i][0]; oy = kOffsets[i][1]; }
					avg += raytraceSinglePixel(x + ox, y + oy, rnd);
				}
				vfb[y][x] = avg / spp;
			}
		}
	};
	std::vector<std::thread> pool;
	for (int t = 0; t < threads; t++) pool.emplace_back(worker);
	for (auto& t: pool) t.join();
	double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	printf("Render took %.3fs (%d threads, %dx%d, %d spp)\n", sec, threads, W, H, spp);

	FILE* f = fopen(argv[2], "wb");
	if (!f) { fprintf(stderr, "cannot write %s\n", argv[2]); return -4; }
	int32_t wh[2] = { W, H };
	fwrite(wh, sizeof(wh), 1, f);
	for (int y = 0; y < H; y++) fwrite(&vfb[y][0], sizeof(Color), W, f);
	fclose(f);

	if (aovFile) {
		f = fopen(aovFile, "wb");
		if (!f) { fprintf(stderr, "cannot write %s\n", aovFile); return -4; }
		fwrite(wh, sizeof(wh), 1, f);
		for (int y = 0; y < H; y++)
			for (int x = 0; x < W; x++) {
				// the closest-hit loops of raytrace(), src/main.cpp:250-271
				Ray ray = scene.camera->getScreenRay(x, y);
				int32_t hitNode = -1;
				double best = 1e99;
				for (int n = 0; n < (int) scene.nodes.size(); n++) {
					IntersectionInfo info;
					if (scene.nodes[n]->intersect(ray, info) && info.dist < best) { best = info.dist; hitNode = n; }
				}
				for (int l = 0; l < (int) scene.lights.size(); l++) {
					IntersectionInfo info;
					if (scene.lights[l]->intersect(ray, info) && info.dist < best) { best = info.dist; hitNode = -2 - l; }
				}
				fwrite(&hitNode, 4, 1, f);
				fwrite(&best, 8, 1, f);
			}
		fclose(f);
	}
	return 0;
}
