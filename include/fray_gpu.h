/*
 * fray_gpu.h -- C ABI of the B200 render back end (libfray_gpu.so).
 *
 * The reference has no plugin/FFI layer. The seam this library drops into is the body of render()
 * in /root/reference/src/main.cpp:373-405: "given the fully prepared global Scene (after
 * Scene::beginRender + Scene::beginFrame, src/scene.cpp:757-781) and the sample count
 * (src/main.cpp:395-400), fill vfb[0..H)[0..W) (src/main.cpp:53) with linear float RGB", which the
 * reference does with pool.run(&RendMT) (src/main.cpp:402-404, worker at :323-371).
 *
 * Everything crossing the boundary is plain C: PODs, pointers and sizes. The host (the reference's own
 * Scene object model, or fray_b200/host which mirrors it) FLATTENS its scene into the tables below;
 * fray_gpu_create() copies them to the GPU (converting to the device layout: SoA float4/double2
 * records, 16-byte KD nodes, packed texel pool) and the host keeps ownership of its own objects.
 * Geometry is FP64 on this side of the boundary exactly like the reference (struct Vector,
 * src/vector.h:30-34); colour is FP32 (struct Color, src/color.h:37-43).
 *
 * Thread-compatibility: one context is used by one thread at a time. All functions return 0 on success
 * and a negative FRAY_GPU_E* code on failure; fray_gpu_last_error() gives the message (thread-local).
 * There is NO CPU fallback: without a CUDA device fray_gpu_create() fails with FRAY_GPU_ENODEVICE.
 */
#ifndef FRAY_GPU_H
#define FRAY_GPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FRAY_GPU_ABI_VERSION 1u

/* error codes */
#define FRAY_GPU_OK          0
#define FRAY_GPU_EINVAL     -1  /* malformed scene / frame description */
#define FRAY_GPU_ENODEVICE  -2  /* no usable CUDA device (no CPU fallback exists) */
#define FRAY_GPU_ECUDA      -3  /* a CUDA runtime call failed */
#define FRAY_GPU_ENOMEM     -4
#define FRAY_GPU_EUNSUPPORTED -5 /* e.g. CSG nesting deeper than FRAY_GPU_MAX_CSG_DEPTH */

/* ---- scene tables ------------------------------------------------------------------------- */

/* struct Transform, src/matrix.h:72-98. m and inv are row-major: m[3*i+j] == Matrix::m[i][j];
 * points/directions are ROW vectors multiplied from the left (operator*, src/matrix.h:53-60). */
typedef struct FrayGpuTransform {
	double offset[3];
	double m[9];
	double inv[9];
} FrayGpuTransform;

/* Geometry classes, src/geometry.h:54-154 and src/mesh.h:55-100 */
enum {
	FRAY_GEOM_PLANE = 0,     /* p = {height, limit}            src/geometry.cpp:30-50  */
	FRAY_GEOM_SPHERE = 1,    /* p = {O.x, O.y, O.z, R}         src/geometry.cpp:52-83  */
	FRAY_GEOM_CUBE = 2,      /* p = {O.x, O.y, O.z, halfSide}  src/geometry.cpp:85-137 */
	FRAY_GEOM_CSG_PLUS = 3,  /* left | right                   src/geometry.cpp:139-194 */
	FRAY_GEOM_CSG_AND = 4,   /* left & right */
	FRAY_GEOM_CSG_MINUS = 5, /* left & !right */
	FRAY_GEOM_MESH = 6       /* mesh = index into meshes[]     src/mesh.cpp:144-165 */
};
#define FRAY_GPU_MAX_CSG_DEPTH 2 /* CsgOp whose operand is a CsgOp whose operands are not */

typedef struct FrayGpuGeometry {
	int32_t type;
	int32_t mesh;
	int32_t left, right; /* CSG operands: indices into geometries[] */
	double p[4];
} FrayGpuGeometry;

/* struct Node, src/geometry.h:158-177 (only nodes WITH a shader, src/scene.cpp:563-568, in file order) */
typedef struct FrayGpuNode {
	int32_t geometry; /* index into geometries[] */
	int32_t shader;   /* index into shaders[] */
	int32_t bump;     /* index into textures[] of a FRAY_TEX_BUMP texture, or -1 (src/main.cpp:82-90) */
	int32_t reserved;
	FrayGpuTransform T;
} FrayGpuNode;

/* class Mesh after beginRender(), src/mesh.h:55-100, src/mesh.cpp:67-94 */
enum {
	FRAY_MESH_FACETED = 1,       /* faceted || normals.empty()  (src/mesh.cpp:70, :112) */
	FRAY_MESH_BACKFACE_CULL = 2, /* src/mesh.cpp:106 */
	FRAY_MESH_HAS_NORMALS = 4,
	FRAY_MESH_HAS_UVS = 8        /* !uvs.empty()  (src/mesh.cpp:123) */
};

typedef struct FrayGpuMesh {
	int32_t flags;
	int32_t first_vertex, num_vertices; /* ranges into vertices[] (xyz doubles); slot 0 of every  */
	int32_t first_normal, num_normals;  /* range is the reference's dummy element so OBJ's 1-based */
	int32_t first_uv, num_uvs;          /* indices are used raw (src/mesh.cpp:209-211)             */
	int32_t first_triangle, num_triangles;
	int32_t kd_root;                    /* index RELATIVE to first_kd_node, -1 = brute force (src/mesh.cpp:85,154-162) */
	int32_t first_kd_node, num_kd_nodes;
	int32_t first_leaf_ref, num_leaf_refs;
	int32_t reserved;
	double bbox_min[3], bbox_max[3];    /* BBox over vertices incl. the dummy (0,0,0)  (src/mesh.cpp:74-79) */
} FrayGpuMesh;

/* struct KDTreeNode, src/mesh.h:35-53, flattened. Children of an inner node are adjacent. */
typedef struct FrayGpuKdNode {
	int32_t axis;  /* 0,1,2 = inner node split axis; 3 = leaf (Axis::AXIS_NONE) */
	int32_t a;     /* inner: index of children[0] relative to mesh.first_kd_node (children[1] = a+1)
	                  leaf : first entry in leaf_refs[] relative to mesh.first_leaf_ref */
	int32_t b;     /* leaf: number of triangle references */
	int32_t reserved;
	double split;  /* inner: splitPos */
} FrayGpuKdNode;

/* Shader classes, src/shading.h:109-255 */
enum {
	FRAY_SHADER_CONST = 0,   /* src/shading.cpp:35-38 */
	FRAY_SHADER_LAMBERT = 1, /* src/shading.cpp:48-99 */
	FRAY_SHADER_PHONG = 2,   /* src/shading.cpp:101-144 */
	FRAY_SHADER_REFL = 3,    /* src/shading.cpp:160-227 */
	FRAY_SHADER_REFR = 4,    /* src/shading.cpp:238-299 */
	FRAY_SHADER_LAYERED = 5  /* src/shading.cpp:357-367 */
};

typedef struct FrayGpuShader {
	int32_t type;
	int32_t texture;          /* diffuseTex: index into textures[] or -1 */
	int32_t first_layer, num_layers; /* LAYERED: range into layers[], bottom layer first */
	int32_t num_samples;      /* REFL: numSamples (glossy, at ray depth 0) */
	int32_t pure_reflection;  /* REFL: glossiness == 1.0 (src/shading.h:197-201) */
	float color[3];           /* CONST / LAMBERT / PHONG */
	float specular_color[3];  /* PHONG */
	float mult[3];            /* REFL / REFR multiplier colour */
	float reserved;
	double exponent;          /* PHONG specularExponent */
	double specular_multiplier;
	double deflection_scaling;/* REFL: pow(10, 2 - 4*glossiness) */
	double ior;               /* REFR */
} FrayGpuShader;

typedef struct FrayGpuLayer { /* Layered::Layer, src/shading.h:238-243 */
	int32_t shader;
	int32_t texture;          /* opacity texture or -1 */
	float opacity[3];
	float reserved;
} FrayGpuLayer;

/* Texture classes, src/shading.h:33-107, 227-237 */
enum {
	FRAY_TEX_CHECKER = 0, /* src/shading.cpp:40-46 */
	FRAY_TEX_BITMAP = 1,  /* src/shading.cpp:147-158; scaling already inverted (src/shading.h:66-67) */
	FRAY_TEX_BUMP = 2,    /* src/shading.cpp:387-418; bitmap already differentiated (src/bitmap.cpp:300-315) */
	FRAY_TEX_FRESNEL = 3  /* src/shading.cpp:369-385 */
};

typedef struct FrayGpuTexture {
	int32_t type;
	int32_t bitmap;       /* index into bitmaps[] or -1 */
	float color1[3];
	float color2[3];
	double scaling;
	double ior;           /* FRESNEL */
	double bump_intensity;/* BUMP strength */
} FrayGpuTexture;

typedef struct FrayGpuBitmap { /* class Bitmap, src/bitmap.h:30-60: row-major float RGB, origin top-left */
	int32_t width, height;
	int64_t first_texel;  /* index of texel (0,0) in texels[] counted in RGB triplets */
} FrayGpuBitmap;

/* Light classes, src/lights.h:32-99 (state after beginFrame, src/lights.cpp:37-46) */
enum { FRAY_LIGHT_POINT = 0, FRAY_LIGHT_RECT = 1 };

typedef struct FrayGpuLight {
	int32_t type;
	int32_t x_subd, y_subd;
	int32_t reserved;
	float color[3];
	float power;
	double pos[3];        /* POINT */
	FrayGpuTransform T;   /* RECT */
	double center[3];     /* RECT: T.transformPoint(0,0,0) */
	double area;          /* RECT: float(width)*float(height) widened (src/lights.cpp:43-45) */
} FrayGpuLight;

/* class Camera after beginFrame(), src/camera.h:37-86, src/camera.cpp:34-57 */
typedef struct FrayGpuCamera {
	double pos[3];
	double top_left[3], top_right[3], bottom_left[3];
	double front[3], up[3], right[3];
	double w, h;              /* frameWidth(), frameHeight() as doubles */
	double aperture_size;     /* 1 / fNumber (src/camera.cpp:56) */
	double focal_plane_dist;
	double stereo_separation;
	float left_mask[3], right_mask[3];
	int32_t dof;
	int32_t num_dof_samples;
} FrayGpuCamera;

/* struct GlobalSettings, src/scene.h:252-278 (only what the render loop reads) */
typedef struct FrayGpuSettings {
	int32_t frame_width, frame_height; /* 1 .. 65535 each (the reference stops at VFB_MAX_SIZE = 3000, src/constants.h:27) */
	int32_t max_trace_depth;
	int32_t gi;
	int32_t num_paths;
	int32_t want_aa;
	float ambient[3];
	float saturation;
} FrayGpuSettings;

typedef struct FrayGpuScene {
	uint32_t abi_version; /* FRAY_GPU_ABI_VERSION */
	uint32_t reserved;
	FrayGpuSettings settings;
	FrayGpuCamera camera;

	int32_t num_nodes;      const FrayGpuNode* nodes;
	int32_t num_geometries; const FrayGpuGeometry* geometries;
	int32_t num_meshes;     const FrayGpuMesh* meshes;
	int32_t num_shaders;    const FrayGpuShader* shaders;
	int32_t num_layers;     const FrayGpuLayer* layers;
	int32_t num_textures;   const FrayGpuTexture* textures;
	int32_t num_bitmaps;    const FrayGpuBitmap* bitmaps;
	int32_t num_lights;     const FrayGpuLight* lights;

	/* mesh pools (struct Triangle, src/triangle.h:30-41, split into SoA planes) */
	int64_t num_vertices;   const double* vertices;  /* xyz */
	int64_t num_normals;    const double* normals;   /* xyz */
	int64_t num_uvs;        const double* uvs;       /* xyz (z = 0) */
	int64_t num_triangles;
	const int32_t* tri_v;   /* 3 per triangle, relative to mesh.first_vertex */
	const int32_t* tri_n;   /* 3 per triangle, relative to mesh.first_normal */
	const int32_t* tri_t;   /* 3 per triangle, relative to mesh.first_uv */
	const double* tri_gnormal; /* xyz, normalised AB^AC */
	const double* tri_dndx;    /* xyz */
	const double* tri_dndy;    /* xyz */
	const double* tri_ab;      /* xyz */
	const double* tri_ac;      /* xyz */
	const double* tri_abxac;   /* xyz, un-normalised */
	int64_t num_kd_nodes;   const FrayGpuKdNode* kd_nodes;
	int64_t num_leaf_refs;  const int32_t* leaf_refs; /* triangle indices relative to mesh.first_triangle */

	int64_t num_texels;     const float* texels;     /* RGB triplets */

	int32_t has_environment;   /* CubemapEnvironment present (src/environment.h:52-77) */
	int32_t env_bitmaps[6];    /* negx,negy,negz,posx,posy,posz (src/environment.h:27-34); -1 = missing */
	int32_t reserved2;
} FrayGpuScene;

/* ---- rendering ---------------------------------------------------------------------------- */

enum {
	FRAY_GPU_FP32 = 0, /* device arithmetic in float with magnitude-scaled ray offsets (fast path) */
	FRAY_GPU_FP64 = 1  /* device arithmetic in double, the reference's literal epsilons (parity path) */
};

enum {
	FRAY_RENDER_BEAUTY = 0, /* rgb_out: float[h][w][3] */
	FRAY_RENDER_AOV = 1,    /* rgb_out: float[h][w][3] = {node index (-1 miss, -2-k light k), triangle index or -1,
	                           world distance} of the un-jittered pin-hole ray through (x, y) */
	FRAY_RENDER_PREPASS = 2 /* rgb_out: the 16x16 preview the reference paints before a frame (render(), src/main.cpp:376-391):
	                           one sample through the centre pixel of every 16x16 square, the square filled with its colour */
};

#define FRAY_FRAME_SUM 1u   /* write per-pixel SUMS over the rendered samples instead of sum / spp */
#define FRAY_FRAME_OWNED_ONLY 2u /* tile split: write only the pixels of this call's share and leave all others untouched (the
                                 default clears them so that partial frames can be summed). For shares that write straight into
                                 ONE frame, e.g. rank 0's frame mapped into every GPU through fray_gpu_frame_import(). */

#define FRAY_FRAME_SAMPLE_RANGE 4u /* [sample_begin, sample_end) is taken literally, so that begin == end is an EMPTY share
                                 (zeros / nothing written). Without this flag 0,0 keeps meaning "all samples". Launchers that
                                 compute ranges (r * spp / world) set it. */

/* Samples per pixel by the reference's rule (src/main.cpp:395-400). */
int fray_gpu_samples_per_pixel(const FrayGpuScene* scene);

typedef struct FrayGpuFrame {
	int32_t spp;            /* total samples per pixel of the frame; 0 = fray_gpu_samples_per_pixel() */
	uint32_t seed;          /* initRandom() seed, 42 in src/main.cpp:502 */
	int32_t sample_begin;   /* this call renders samples [sample_begin, sample_end) of every owned pixel; */
	int32_t sample_end;     /*   0,0 = all, unless FRAY_FRAME_SAMPLE_RANGE is set */
	int32_t bucket_rank;    /* this call owns share `bucket_rank` of `bucket_count` of the image: the frame is cut into  */
	int32_t bucket_count;   /*   8x4 pixel tiles (row-major) and tile t belongs to share t % bucket_count -- the         */
	                        /*   multi-GPU form of the reference's bucket list (src/sdl.cpp:243-262); 0,0 = all          */
	int32_t mode;           /* FRAY_RENDER_* */
	uint32_t flags;         /* FRAY_FRAME_* */
} FrayGpuFrame;

typedef struct FrayGpuStats {
	uint64_t rays;          /* closest-hit (raytrace/pathtrace past its cut-off) + any-hit (visible) queries */
	uint64_t primary_rays;  /* pixel samples started */
	uint64_t shadow_rays;   /* any-hit queries */
	double device_ms;       /* CUDA-event time of the kernels of this call */
	int32_t kernel_launches;
	int32_t reserved;
} FrayGpuStats;

typedef struct FrayGpuCtx FrayGpuCtx;

uint32_t fray_gpu_abi_version(void);
int fray_gpu_device_count(void);

/* Upload `scene` to CUDA device `device` (ordinal). `precision` is FRAY_GPU_FP32 or FRAY_GPU_FP64. */
int fray_gpu_create(const FrayGpuScene* scene, int device, int precision, FrayGpuCtx** out);

/* Per-frame camera change (interactive loop, src/main.cpp:437-491). */
int fray_gpu_update_camera(FrayGpuCtx* ctx, const FrayGpuCamera* camera);

/* Render into a HOST buffer of width*height*3 floats (row-major, origin top-left like vfb[y][x]).
 * Timed region of `stats->device_ms` covers the kernels only; the call returns after the D2H copy. */
int fray_gpu_render(FrayGpuCtx* ctx, const FrayGpuFrame* frame, float* rgb_out, FrayGpuStats* stats);

/* Render into DEVICE memory (width*height*3 floats on the context's device), asynchronously on
 * `cuda_stream` (a cudaStream_t, NULL = the context's own stream). Pixels not owned by this call are
 * written as 0 so that partial frames from several GPUs can be summed (ncclReduce). Stats are valid
 * after fray_gpu_sync(). */
int fray_gpu_render_device(FrayGpuCtx* ctx, const FrayGpuFrame* frame, void* d_rgb, void* cuda_stream);

/* Peer-to-peer frames (multi-GPU tile split without a reduction). fray_gpu_frame_export() allocates TWO frames of
 * width*height*3 floats back to back on the context's device (owned by the context; the second starts width*height*3 floats
 * after the returned address -- alternate between them from frame to frame, so that the shares of frame k+1 never land in memory
 * the owner is still reading frame k from) and returns the address and an opaque 64-byte handle
 * (a cudaIpcMemHandle_t); another PROCESS that drives another GPU of the same node passes the handle to
 * fray_gpu_frame_import() and gets an address through which its kernels write into that frame over NVLink (pass it as d_rgb to
 * fray_gpu_render_device with FRAY_FRAME_OWNED_ONLY). The importer calls fray_gpu_frame_close() when done. The caller
 * orders "all shares are written" before "the frame is read" (e.g. a barrier on the streams involved). */
int fray_gpu_frame_export(FrayGpuCtx* ctx, void** d_frame, unsigned char handle[64]);
int fray_gpu_frame_import(FrayGpuCtx* ctx, const unsigned char handle[64], void** d_frame);
int fray_gpu_frame_close(FrayGpuCtx* ctx, void* d_frame);

/* d_rgb[i] = d_sum[i] / spp on the device (the `avg / samplesPerPixel` of src/main.cpp:360). */
int fray_gpu_resolve_device(FrayGpuCtx* ctx, const void* d_sum, void* d_rgb, int32_t spp, void* cuda_stream);

/* The same, but the frame goes straight to the HOST: pinned_rgb must be page-locked host memory (cudaMallocHost /
 * cudaHostRegister, e.g. a pinned torch tensor); the kernel stores the quotients into it over PCIe, so no separate
 * device-to-host copy follows. Asynchronous on `cuda_stream`; FRAY_GPU_EINVAL for pageable memory. */
int fray_gpu_resolve_to_host(FrayGpuCtx* ctx, const void* d_sum, float* pinned_rgb, int32_t spp, void* cuda_stream);

/* Tile split with the frame wanted in HOST memory: renders the share `frame` describes (bucket_rank / bucket_count; all
 * samples) and stores the finished pixels of ITS OWN tiles -- nothing else -- into the page-locked host frame `pinned_rgb`
 * (width*height*3 floats), asynchronously on `cuda_stream`. Every GPU sends its tiles over its own PCIe link; the shares of N
 * GPUs -- N threads of one process, or N rank processes that map one shared-memory frame and page-lock it with
 * fray_gpu_host_register() -- assemble the frame without a reduction, a peer copy or a device-to-host copy of the whole frame.
 * (This is what every worker thread of the reference does with its buckets and the shared `vfb`, src/main.cpp:331-370.) The
 * caller waits for all shares (fray_gpu_sync on each) before it reads the frame. FRAY_GPU_EINVAL for pageable memory. */
int fray_gpu_render_to_host(FrayGpuCtx* ctx, const FrayGpuFrame* frame, float* pinned_rgb, void* cuda_stream);

/* Page-lock `bytes` of host memory that something else allocated (a POSIX shared-memory mapping, a memory-mapped file, a
 * buffer of the host application) and map it into the CUDA address space, so that kernels can store into it
 * (fray_gpu_render_to_host, fray_gpu_resolve_to_host). Undo with fray_gpu_host_unregister() before the memory is unmapped. */
int fray_gpu_host_register(void* host, size_t bytes);
int fray_gpu_host_unregister(void* host);

/* Wait for the context's outstanding work and fetch the statistics of the last render. */
int fray_gpu_sync(FrayGpuCtx* ctx, FrayGpuStats* stats);

void fray_gpu_destroy(FrayGpuCtx* ctx);

/* ---- several GPUs of one node, one process -------------------------------------------------- */
/* What pool.run(&worker, scene.settings.numThreads) is to the reference's host threads (src/main.cpp:402-404,
 * src/cxxptl-sdl.cpp:285-318): one call renders the frame on `n_devices` GPUs. The scene is uploaded to each of
 * `devices[0..n_devices)` (NULL = ordinals 0 .. n_devices-1; an ordinal may repeat); devices[0] owns the frame and every
 * other device needs peer access to it (FRAY_GPU_EUNSUPPORTED otherwise). */
typedef struct FrayGpuMulti FrayGpuMulti;

enum {
	FRAY_GPU_SPLIT_AUTO = 0,   /* samples while every GPU keeps >= 8 samples per pixel, else tiles */
	FRAY_GPU_SPLIT_TILES = 1,  /* share d = the 8x4 pixel tiles t with t % n == d (the bucket list of src/sdl.cpp:243-262 dealt round
	                              robin), stored straight into device 0's frame over NVLink; bit-identical to the single-GPU frame */
	FRAY_GPU_SPLIT_SAMPLES = 2 /* share d = samples [d*spp/n, (d+1)*spp/n) of every pixel; partial sums are stored into device 0
	                              and added there in share order (equal to the single-GPU frame up to FP32 summation order) */
};
#define FRAY_GPU_MULTI_FAST 0x100u /* frame flag, tile split only: let every share choose its own samples-per-item (a few
                                      percent faster on small shares; the frame then equals the single-GPU frame only up to
                                      FP32 summation order) */

int fray_gpu_multi_create(const FrayGpuScene* scene, int n_devices, const int* devices, int precision, FrayGpuMulti** out);
int fray_gpu_multi_device_count(const FrayGpuMulti* multi);
int fray_gpu_multi_update_camera(FrayGpuMulti* multi, const FrayGpuCamera* camera);
/* `frame` describes the WHOLE frame (no bucket / sample range); `split` is FRAY_GPU_SPLIT_*. Host buffer as fray_gpu_render().
 * stats: rays summed over the shares, device_ms of the slowest share. */
int fray_gpu_multi_render(FrayGpuMulti* multi, const FrayGpuFrame* frame, int split, float* rgb_out, FrayGpuStats* stats);
void fray_gpu_multi_destroy(FrayGpuMulti* multi);

/* Roofline denominators measured on the spot (bench.py): sustained FP32 FFMA rate of `device` in TFLOP/s (8 independent
 * FMA chains per thread, all SMs, ~`ms` milliseconds) and L2-resident read bandwidth in GB/s (a 32 MiB buffer read
 * repeatedly with 128-bit loads). Either output may be NULL. */
int fray_gpu_measure_peaks(int device, double ms, double* fp32_tflops, double* l2_gbs);

/* Self-test of the random streams as the kernels generate them (tests/test_rng.py compares with the CPU statement of the
 * contract, draw for draw). Runs stream (seed, pixel, sample, branch) on `device` in one of the device forms:
 *   mode 0  on-demand blocks with round keys from the kernel parameters (the Whitted kernels),
 *   mode 1  the shared-memory ring of the path-tracing kernels, drawn sequentially in groups of up to 12,
 *   mode 2  the ring under the path tracer's own pattern: 2 draws for the pixel offset, then per path segment 2 draws
 *           skipped and 6 drawn (the discarded first spawnRay, explicitLightSample, hemisphereSample; src/main.cpp:92-169, 219-236).
 * out[i] receives draw i of the stream and drawn[i] = 1 for every position i < n that was drawn (all of them in modes 0, 1). */
int fray_gpu_rng_probe(int device, uint32_t seed, uint32_t pixel, uint32_t sample, uint32_t branch, int mode, int n, uint32_t* out, unsigned char* drawn);

const char* fray_gpu_last_error(void);

#ifdef __cplusplus
}
#endif

#endif /* FRAY_GPU_H */
