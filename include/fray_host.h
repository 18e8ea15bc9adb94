/*
 * fray_host.h -- C ABI of the host scene layer (libfray_host.so): the part of fray that stays on the CPU.
 *
 * Mirrors the start-up sequence of the reference's main() (/root/reference/src/main.cpp:494-530):
 *   scene.parseScene(file)  ->  scene.beginRender()  ->  [per frame] scene.beginFrame()
 * and exposes the prepared scene as the flattened tables of fray_gpu.h, which is what
 * fray_gpu_create() consumes. Also exports the image writers behind the reference's F12 screenshots
 * (src/sdl.cpp:101-140, src/bitmap.cpp:197-284). Plain C: opaque handle, pointers and sizes.
 */
#ifndef FRAY_HOST_H
#define FRAY_HOST_H

#include "fray_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct FrayHostScene FrayHostScene;

/* Parse `path`, run beginRender() + beginFrame() and flatten. Returns NULL on failure (fray_host_last_error()). */
FrayHostScene* fray_host_load_scene(const char* path);
void fray_host_free_scene(FrayHostScene* scene);
const char* fray_host_last_error(void);

/* 1: print the reference's progress lines ("Mesh loaded, N triangles", KD statistics, unknown-property warnings); default 0 */
void fray_host_set_verbose(int verbose);

/* The flattened scene; valid until the next fray_host_* call that mutates the scene, or free. */
const FrayGpuScene* fray_host_flat_scene(const FrayHostScene* scene);

/* samples per pixel by the rule of src/main.cpp:395-400 */
int fray_host_samples_per_pixel(const FrayHostScene* scene);

/* Override GlobalSettings / Camera properties after parsing (what the benchmark configs change):
 * key in {"frameWidth","frameHeight","pathsPerPixel","maxTraceDepth","wantAA","gi","dof","numSamples","seed"...};
 * re-runs beginFrame() and re-flattens the per-frame state. Returns 0, or -1 for an unknown key. */
int fray_host_set_int(FrayHostScene* scene, const char* key, int value);

/* Interactive camera (Camera::move / Camera::rotate, src/camera.cpp:95-106) followed by beginFrame();
 * writes the refreshed camera for fray_gpu_update_camera(). */
int fray_host_move_camera(FrayHostScene* scene, double dx, double dz, double dyaw, double dpitch, FrayGpuCamera* out);

/* KD statistics of mesh `mesh_index` (order of FrayGpuScene.meshes): nodes, leaf references, max depth. */
int fray_host_mesh_stats(const FrayHostScene* scene, int mesh_index, int* nodes, int* leaf_refs, int* max_depth, int* triangles);

/* Image IO: rgb is float[h][w][3], top-left origin. `path` ends in .bmp or .exr. 0 on success. */
int fray_host_save_image(const char* path, const float* rgb, int width, int height);
/* Load .bmp / .exr into a malloc'ed float[h][w][3]; caller frees with fray_host_free_pixels(). */
int fray_host_load_image(const char* path, float** rgb, int* width, int* height);
void fray_host_free_pixels(float* rgb);

/* The product's counter-based RNG (fray_b200/csrc/rng.cuh, the code the kernels run) evaluated on the host: the first
 * `n` 32-bit draws of stream (seed, pixel, sample, branch), and the branch id of a Whitted secondary ray.
 * Exposed so that tests can hold it against the contract stated in DESIGN.md draw for draw. */
void fray_host_rng_draws(uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t branch, int n, uint32_t* out);
uint32_t fray_host_rng_child(uint32_t branch, uint32_t draws, uint32_t k);

#ifdef __cplusplus
}
#endif

#endif
