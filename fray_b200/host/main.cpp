// main.cpp -- the `fray [--gpu] scene.fray` entry point (host side of the B200 back end).
//
// Mirrors main() of the reference (/root/reference/src/main.cpp:494-530): parse the command line (exit -1 on a usage
// error, :414-424), check the scene file (exit -2, :498-501), parse it (exit -3, :503-506), beginRender(), render one frame
// and report "Render took %.2fs" (:517-520). Where the reference opens an SDL window and waits for F12 to write a
// screenshot (src/sdl.cpp:101-160), this tool is headless and writes the frame to --out (BMP or EXR, the two formats
// Bitmap::saveImage knows, src/bitmap.cpp:197-284). The frame itself is produced by libfray_gpu.so through the C ABI of
// include/fray_gpu.h -- loaded with dlopen so that the host layer has no link-time CUDA dependency. There is no CPU
// renderer in this build: without --gpu (or without a CUDA device) the tool says so and fails.
//
//   fray --gpu [--fp64] [--out frame.bmp] [--seed N] [--spp N] [--device D | --devices N [--split auto|tiles|samples]]
//        [--frames K] [--move dx,dz,dyaw,dpitch] [--prepass] [--bucket-rank R --bucket-count C] [--samples A:B] [--aov] [-v] scene.fray
//
// --devices N renders every frame on GPUs 0 .. N-1 of this node from this one process (fray_gpu_multi_render: what
// pool.run(&worker, numThreads) is to the reference's host threads, src/main.cpp:402-404). Tile shares are stored straight
// into GPU 0's frame over NVLink and give the single-GPU frame bit for bit; sample shares are added on GPU 0.
//
// --prepass: before the frame, the 16x16-pixel preview of the reference's render() (src/main.cpp:378-391: one sample through
// the centre of every 16x16 square, painted over the square) is rendered and, with --out, written as <out>.prepass.<ext>;
// "wantPrepass on" in the scene file's GlobalSettings asks for it as well, as in the reference.
//
// --frames K --move ... is the headless form of the reference's interactive loop (mainloop, src/main.cpp:437-491): before
// every frame after the first the camera is moved and turned by the given amounts (Camera::move / Camera::rotate,
// src/camera.cpp:95-106), re-prepared with beginFrame() and sent to the GPU with fray_gpu_update_camera() -- no scene upload.
// With several frames, "%d" in --out is replaced by the frame number.
//
// --bucket-* and --samples render one shard of the frame (tile split / sample split, include/fray_gpu.h FrayGpuFrame) so
// that an outer launcher can spread a frame over several GPUs; fray_b200/dist.py does that with one process per GPU and an
// NCCL reduce.
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <vector>

#include "../../include/fray_host.h"

namespace {

struct GpuApi {
	void* handle = nullptr;
	uint32_t (*abi_version)(void) = nullptr;
	int (*device_count)(void) = nullptr;
	int (*create)(const FrayGpuScene*, int, int, FrayGpuCtx**) = nullptr;
	int (*render)(FrayGpuCtx*, const FrayGpuFrame*, float*, FrayGpuStats*) = nullptr;
	int (*update_camera)(FrayGpuCtx*, const FrayGpuCamera*) = nullptr;
	void (*destroy)(FrayGpuCtx*) = nullptr;
	int (*multi_create)(const FrayGpuScene*, int, const int*, int, FrayGpuMulti**) = nullptr;
	int (*multi_render)(FrayGpuMulti*, const FrayGpuFrame*, int, float*, FrayGpuStats*) = nullptr;
	int (*multi_update_camera)(FrayGpuMulti*, const FrayGpuCamera*) = nullptr;
	void (*multi_destroy)(FrayGpuMulti*) = nullptr;
	const char* (*last_error)(void) = nullptr;
};

std::string exeDir()
{
	char buf[4096];
	ssize_t n = readlink("/proc/self/exe", buf, sizeof(buf) - 1);
	if (n <= 0) return ".";
	buf[n] = 0;
	char* slash = strrchr(buf, '/');
	if (slash) *slash = 0;
	return buf;
}

bool loadGpu(GpuApi& api, std::string& err)
{
	const std::string beside = exeDir() + "/libfray_gpu.so";
	api.handle = dlopen(beside.c_str(), RTLD_NOW | RTLD_LOCAL);
	if (!api.handle) api.handle = dlopen("libfray_gpu.so", RTLD_NOW | RTLD_LOCAL);
	if (!api.handle) {
		err = dlerror();
		return false;
	}
#define FRAY_SYM(member, name)                                              \
	api.member = reinterpret_cast<decltype(api.member)>(dlsym(api.handle, name)); \
	if (!api.member) { err = std::string("missing symbol ") + name; return false; }
	FRAY_SYM(abi_version, "fray_gpu_abi_version");
	FRAY_SYM(device_count, "fray_gpu_device_count");
	FRAY_SYM(create, "fray_gpu_create");
	FRAY_SYM(render, "fray_gpu_render");
	FRAY_SYM(update_camera, "fray_gpu_update_camera");
	FRAY_SYM(destroy, "fray_gpu_destroy");
	FRAY_SYM(multi_create, "fray_gpu_multi_create");
	FRAY_SYM(multi_render, "fray_gpu_multi_render");
	FRAY_SYM(multi_update_camera, "fray_gpu_multi_update_camera");
	FRAY_SYM(multi_destroy, "fray_gpu_multi_destroy");
	FRAY_SYM(last_error, "fray_gpu_last_error");
#undef FRAY_SYM
	if (api.abi_version() != FRAY_GPU_ABI_VERSION) {
		err = "libfray_gpu.so has a different ABI version";
		return false;
	}
	return true;
}

void usage() { fprintf(stderr, "Usage: fray --gpu [--fp64] [--out file.bmp|.exr] [--seed N] [--spp N] [--device D | --devices N [--split auto|tiles|samples]]\n"
                               "            [--frames K] [--move dx,dz,dyaw,dpitch] [--prepass] [--bucket-rank R --bucket-count C] [--samples A:B] [--aov] [-v] scene.fray\n"); }

} // namespace

int main(int argc, char** argv)
{
	bool gpu = false, verbose = false, aov = false;
	int precision = FRAY_GPU_FP32, device = 0, frames = 1, devices = 0, split = FRAY_GPU_SPLIT_AUTO;
	bool prepass = false;
	FrayGpuFrame frame;
	memset(&frame, 0, sizeof(frame));
	frame.seed = 42; // initRandom(42), src/main.cpp:502
	double move[4] = { 0, 0, 0, 0 }; // per frame: dx, dz (Camera::move), dyaw, dpitch (Camera::rotate)
	bool moving = false;
	std::string out, sceneFile = "data/boxed.fray"; // default scene of the reference, src/main.cpp:51
	for (int i = 1; i < argc; i++) {
		const std::string a = argv[i];
		auto next = [&](const char* what) -> const char* {
			if (i + 1 >= argc) { fprintf(stderr, "fray: %s needs a value\n", what); usage(); exit(-1); }
			return argv[++i];
		};
		if (a == "--gpu") gpu = true;
		else if (a == "--fp64") precision = FRAY_GPU_FP64;
		else if (a == "-v") verbose = true;
		else if (a == "--aov") aov = true;
		else if (a == "--out") out = next("--out");
		else if (a == "--seed") frame.seed = (uint32_t) strtoul(next("--seed"), nullptr, 10);
		else if (a == "--spp") frame.spp = atoi(next("--spp"));
		else if (a == "--device") device = atoi(next("--device"));
		else if (a == "--devices") devices = atoi(next("--devices"));
		else if (a == "--prepass") prepass = true;
		else if (a == "--split") {
			const std::string v = next("--split");
			if (v == "auto") split = FRAY_GPU_SPLIT_AUTO;
			else if (v == "tiles") split = FRAY_GPU_SPLIT_TILES;
			else if (v == "samples") split = FRAY_GPU_SPLIT_SAMPLES;
			else { usage(); return -1; }
		}
		else if (a == "--frames") frames = atoi(next("--frames"));
		else if (a == "--move") {
			if (sscanf(next("--move"), "%lf,%lf,%lf,%lf", &move[0], &move[1], &move[2], &move[3]) != 4) { usage(); return -1; }
			moving = true;
		}
		else if (a == "--bucket-rank") frame.bucket_rank = atoi(next("--bucket-rank"));
		else if (a == "--bucket-count") frame.bucket_count = atoi(next("--bucket-count"));
		else if (a == "--samples") {
			if (sscanf(next("--samples"), "%d:%d", &frame.sample_begin, &frame.sample_end) != 2) { usage(); return -1; }
			frame.flags |= FRAY_FRAME_SAMPLE_RANGE; // literal: A == B is an empty share
		} else if (!a.empty() && a[0] == '-') { usage(); return -1; }
		else sceneFile = a;
	}
	if (!gpu) {
		fprintf(stderr, "fray: this build contains only the CUDA back end; run it as `fray --gpu %s`\n", sceneFile.c_str());
		return -1;
	}
	struct stat st;
	if (stat(sceneFile.c_str(), &st) != 0) {
		fprintf(stderr, "The specified scene file does not exist: %s", sceneFile.c_str());
		return -2;
	}
	fray_host_set_verbose(verbose ? 1 : 0);
	FrayHostScene* scene = fray_host_load_scene(sceneFile.c_str());
	if (!scene) {
		fprintf(stderr, "%s\n", fray_host_last_error());
		return -3;
	}
	const FrayGpuScene* flat = fray_host_flat_scene(scene);
	const int W = flat->settings.frame_width, H = flat->settings.frame_height;

	GpuApi api;
	std::string err;
	if (!loadGpu(api, err)) {
		fprintf(stderr, "fray: cannot load the CUDA back end: %s\n", err.c_str());
		return -4;
	}
	// one GPU (fray_gpu_create / fray_gpu_render) or several driven from this process (fray_gpu_multi_*)
	FrayGpuCtx* ctx = nullptr;
	FrayGpuMulti* multi = nullptr;
	if (devices > 0) {
		if (frame.bucket_count > 0 || (frame.flags & FRAY_FRAME_SAMPLE_RANGE)) {
			fprintf(stderr, "fray: --devices splits the frame itself; it cannot be combined with --bucket-* or --samples\n");
			return -1;
		}
		if (api.multi_create(flat, devices, nullptr, precision, &multi) != FRAY_GPU_OK) {
			fprintf(stderr, "fray: %s\n", api.last_error());
			return -4;
		}
	} else if (api.create(flat, device, precision, &ctx) != FRAY_GPU_OK) {
		fprintf(stderr, "fray: %s\n", api.last_error());
		return -4;
	}
	auto destroy = [&]() {
		if (ctx) api.destroy(ctx);
		if (multi) api.multi_destroy(multi);
	};
	auto renderFrame = [&](const FrayGpuFrame& fr, float* rgb, FrayGpuStats* st) {
		return multi ? api.multi_render(multi, &fr, split, rgb, st) : api.render(ctx, &fr, rgb, st);
	};
	if (aov) frame.mode = FRAY_RENDER_AOV;
	std::vector<float> rgb((size_t) W * H * 3);
	FrayGpuStats stats;
	memset(&stats, 0, sizeof(stats));
	int rc = 0;
	auto outputName = [&](int f, const char* infix) {
		std::string name = out;
		const size_t pos = name.find("%d");
		if (pos != std::string::npos) name.replace(pos, 2, std::to_string(f));
		if (infix) {
			const size_t dot = name.rfind('.');
			name.insert(dot == std::string::npos ? name.size() : dot, infix);
		}
		return name;
	};
	for (int f = 0; f < frames; f++) {
		if (f > 0 && moving) {
			FrayGpuCamera cam;
			if (fray_host_move_camera(scene, move[0], move[1], move[2], move[3], &cam) != 0 ||
			    (multi ? api.multi_update_camera(multi, &cam) : api.update_camera(ctx, &cam)) != FRAY_GPU_OK) {
				fprintf(stderr, "fray: camera update failed: %s\n", api.last_error());
				destroy();
				return -5;
			}
		}
		const auto t0 = std::chrono::steady_clock::now();
		if (prepass && !aov) { // the 16x16 preview of render(), src/main.cpp:376-391; always on one GPU (a few thousand rays)
			FrayGpuFrame pre = frame;
			pre.mode = FRAY_RENDER_PREPASS;
			pre.bucket_rank = pre.bucket_count = 0;
			FrayGpuStats ps;
			const int prc = multi ? api.multi_render(multi, &pre, FRAY_GPU_SPLIT_TILES, rgb.data(), &ps) : api.render(ctx, &pre, rgb.data(), &ps);
			if (prc != FRAY_GPU_OK) {
				fprintf(stderr, "fray: prepass: %s\n", api.last_error());
				destroy();
				return -5;
			}
			if (verbose) printf("  prepass: %llu rays\n", (unsigned long long) ps.rays);
			if (!out.empty() && (f == frames - 1 || out.find("%d") != std::string::npos) &&
			    fray_host_save_image(outputName(f, ".prepass").c_str(), rgb.data(), W, H) != 0) {
				fprintf(stderr, "fray: cannot write the prepass image: %s\n", fray_host_last_error());
				rc = -6;
			}
		}
		if (renderFrame(frame, rgb.data(), &stats) != FRAY_GPU_OK) {
			fprintf(stderr, "fray: %s\n", api.last_error());
			destroy();
			return -5;
		}
		const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("Render took %.2fs\n", sec); // src/main.cpp:519
		if (verbose)
			printf("  %dx%d, %llu rays (%llu primary, %llu shadow), device %.3f ms, %.1f Mrays/s%s\n", W, H, (unsigned long long) stats.rays,
			       (unsigned long long) stats.primary_rays, (unsigned long long) stats.shadow_rays, stats.device_ms,
			       stats.device_ms > 0 ? stats.rays / stats.device_ms / 1e3 : 0.0, multi ? (" on " + std::to_string(devices) + " GPUs").c_str() : "");
		if (!out.empty() && (f == frames - 1 || out.find("%d") != std::string::npos)) {
			const std::string name = outputName(f, nullptr);
			if (fray_host_save_image(name.c_str(), rgb.data(), W, H) != 0) {
				fprintf(stderr, "fray: cannot write %s: %s\n", name.c_str(), fray_host_last_error());
				rc = -6;
			}
		}
	}
	destroy();
	fray_host_free_scene(scene);
	if (rc == 0) printf("Exited cleanly\n");
	return rc;
}
