// capi.cpp -- C ABI of the host scene layer (include/fray_host.h).
#include <cstdlib>
#include <cstring>

#include "scene.h"
#include "fray_host.h"
#include "../csrc/rng.cuh"

using namespace fray;

struct FrayHostScene {
	Scene scene;
	FlatScene flat;
};

static thread_local std::string g_error;

extern "C" {

const char* fray_host_last_error(void) { return g_error.c_str(); }

void fray_host_set_verbose(int verbose) { g_verbose = verbose; }

FrayHostScene* fray_host_load_scene(const char* path)
{
	FrayHostScene* h = new FrayHostScene;
	if (!h->scene.parseScene(path)) {
		g_error = h->scene.lastError.empty() ? std::string("cannot parse ") + path : h->scene.lastError;
		delete h;
		return nullptr;
	}
	h->scene.beginRender();
	h->scene.beginFrame();
	if (!flatten(h->scene, h->flat)) {
		g_error = h->scene.lastError;
		delete h;
		return nullptr;
	}
	return h;
}

void fray_host_free_scene(FrayHostScene* h) { delete h; }

const FrayGpuScene* fray_host_flat_scene(const FrayHostScene* h) { return h ? &h->flat.view : nullptr; }

int fray_host_samples_per_pixel(const FrayHostScene* h) { return h ? h->scene.samplesPerPixel() : 0; }

static void refreshFrameState(FrayHostScene* h)
{
	h->scene.beginFrame();
	FrayGpuSettings& st = h->flat.view.settings;
	const GlobalSettings& gs = h->scene.settings;
	st.frame_width = gs.frameWidth;
	st.frame_height = gs.frameHeight;
	st.max_trace_depth = gs.maxTraceDepth;
	st.gi = gs.gi;
	st.num_paths = gs.numPaths;
	st.want_aa = gs.wantAA;
	flattenCamera(h->scene, h->flat.view.camera);
}

int fray_host_set_int(FrayHostScene* h, const char* key, int value)
{
	if (!h || !key) return -1;
	GlobalSettings& gs = h->scene.settings;
	Camera& cam = *h->scene.camera;
	if (!strcmp(key, "frameWidth")) gs.frameWidth = value;
	else if (!strcmp(key, "frameHeight")) gs.frameHeight = value;
	else if (!strcmp(key, "pathsPerPixel")) gs.numPaths = value;
	else if (!strcmp(key, "maxTraceDepth")) gs.maxTraceDepth = value;
	else if (!strcmp(key, "wantAA")) gs.wantAA = value != 0;
	else if (!strcmp(key, "gi")) gs.gi = value != 0;
	else if (!strcmp(key, "dof")) cam.dof = value != 0;
	else if (!strcmp(key, "numSamples")) cam.numDOFSamples = value;
	else return -1;
	refreshFrameState(h);
	return 0;
}

int fray_host_move_camera(FrayHostScene* h, double dx, double dz, double dyaw, double dpitch, FrayGpuCamera* out)
{
	if (!h) return -1;
	h->scene.camera->move(dx, dz);
	h->scene.camera->rotate(dyaw, dpitch);
	refreshFrameState(h);
	if (out) *out = h->flat.view.camera;
	return 0;
}

int fray_host_mesh_stats(const FrayHostScene* h, int mesh_index, int* nodes, int* leaf_refs, int* max_depth, int* triangles)
{
	if (!h) return -1;
	int k = 0;
	for (Geometry* g: h->scene.geometries) {
		Mesh* m = dynamic_cast<Mesh*>(g);
		if (!m) continue;
		if (k++ == mesh_index) {
			if (nodes) *nodes = (int) m->kdNodes.size();
			if (leaf_refs) *leaf_refs = (int) m->leafRefs.size();
			if (max_depth) *max_depth = m->maxTreeDepth;
			if (triangles) *triangles = (int) m->triangles.size();
			return 0;
		}
	}
	return -1;
}

int fray_host_save_image(const char* path, const float* rgb, int width, int height)
{
	Bitmap bmp;
	bmp.generateEmptyImage(width, height);
	if (!bmp.isOK()) return -1;
	for (size_t i = 0; i < (size_t) width * height; i++) bmp.data[i] = Color(rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);
	return bmp.saveImage(path) ? 0 : -1;
}

int fray_host_load_image(const char* path, float** rgb, int* width, int* height)
{
	Bitmap bmp;
	if (!bmp.loadImage(path) || !bmp.isOK()) return -1;
	float* p = (float*) malloc(sizeof(float) * 3 * bmp.data.size());
	if (!p) return -1;
	for (size_t i = 0; i < bmp.data.size(); i++) {
		p[3 * i] = bmp.data[i].r;
		p[3 * i + 1] = bmp.data[i].g;
		p[3 * i + 2] = bmp.data[i].b;
	}
	*rgb = p;
	*width = bmp.width;
	*height = bmp.height;
	return 0;
}

void fray_host_free_pixels(float* rgb) { free(rgb); }

void fray_host_rng_draws(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t branch, int n, uint32_t* out)
{
	Rng r;
	r.init(seed, pixel, sample, branch);
	for (int i = 0; i < n; i++) out[i] = r.next();
}

uint32_t fray_host_rng_child(uint32_t branch, uint32_t draws, uint32_t k) { return rngChildBranch(branch, draws, k); }

} // extern "C"
