// ref_gpu_main.cpp -- the reference-side binding of INTEGRATION.md, compiled. TEST INFRASTRUCTURE.
//
// Built by oracle/Makefile into oracle/_ref/fray_ref_gpu from the UNMODIFIED reference sources (all of /root/reference/src
// except main.cpp's main(), renamed at link time as for fray_ref_ctr) plus this file, which is what a maintainer of the
// reference would add to call libfray_gpu.so:
//   * FlatBuilder: walks the reference's own prepared `Scene scene` (src/scene.h:280-299) -- textures, shaders, geometries,
//     nodes, lights, camera, environment, in file order -- and fills the tables of include/fray_gpu.h;
//   * the replacement of render()'s body (src/main.cpp:373-405): fray_gpu_create once, fray_gpu_render per frame, vfb filled.
// The maintainer's version would reach the members through friend declarations or the getInterface() idiom
// (src/scene.h:138); this test target reads the reference's headers with `private` / `protected` opened instead, which leaves
// the reference's files untouched (the object layout does not depend on access specifiers).
//
//   fray_ref_gpu scene.fray --dump out.flat                     flatten only (no GPU needed)
//   fray_ref_gpu scene.fray --render out.f32 [--fp64] [--lib path/libfray_gpu.so]
//                                                               render through the C ABI; out.f32 = int32 w, h, float rgb[h][w][3]
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <math.h>
#include <algorithm>
#include <functional>
#include <map>
#include <random>
#include <string>
#include <vector>
#include <SDL/SDL.h>
#include <SDL/SDL_thread.h>
#include <SDL/SDL_mutex.h>

// every standard header the reference's headers pull in is included above, so the three defines only touch the reference's own
// declarations: `class X { members...` (private by default) reads as `struct X {` (and `enum class` as `enum struct`)
#define private public
#define protected public
#define class struct
#include "sdl.h"
#include "color.h"
#include "vector.h"
#include "matrix.h"
#include "camera.h"
#include "geometry.h"
#include "mesh.h"
#include "shading.h"
#include "lights.h"
#include "environment.h"
#include "bitmap.h"
#include "scene.h"
#include "random_generator.h"
#undef private
#undef protected
#undef class

#include "fray_gpu.h"
#include "flat_dump.h"

extern Color vfb[VFB_MAX_SIZE][VFB_MAX_SIZE]; // src/main.cpp:53

namespace {

void put3(double* d, const Vector& v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; }
void put3(float* d, const Color& c) { d[0] = c.r; d[1] = c.g; d[2] = c.b; }
void push3(std::vector<double>& v, const Vector& p) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); }

void putTransform(FrayGpuTransform& out, const Transform& T) // struct Transform, src/matrix.h:72-98
{
	put3(out.offset, T.offset);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) {
			out.m[3 * i + j] = T.m.m[i][j];
			out.inv[3 * i + j] = T.invM.m[i][j];
		}
}

void flattenCamera(const Camera& cam, FrayGpuCamera& c) // class Camera after beginFrame(), src/camera.h:37-86, src/camera.cpp:34-57
{
	memset(&c, 0, sizeof(c));
	put3(c.pos, cam.pos);
	put3(c.top_left, cam.topLeft);
	put3(c.top_right, cam.topRight);
	put3(c.bottom_left, cam.bottomLeft);
	put3(c.front, cam.frontDir);
	put3(c.up, cam.upDir);
	put3(c.right, cam.rightDir);
	c.w = cam.w;
	c.h = cam.h;
	c.aperture_size = cam.apertureSize;
	c.focal_plane_dist = cam.focalPlaneDist;
	c.stereo_separation = cam.stereoSeparation;
	put3(c.left_mask, cam.leftMask);
	put3(c.right_mask, cam.rightMask);
	c.dof = cam.dof;
	c.num_dof_samples = cam.numDOFSamples;
}

struct FlatBuilder {
	FrayGpuScene view;
	std::vector<FrayGpuNode> nodes;
	std::vector<FrayGpuGeometry> geometries;
	std::vector<FrayGpuMesh> meshes;
	std::vector<FrayGpuShader> shaders;
	std::vector<FrayGpuLayer> layers;
	std::vector<FrayGpuTexture> textures;
	std::vector<FrayGpuBitmap> bitmaps;
	std::vector<FrayGpuLight> lights;
	std::vector<double> vertices, normals, uvs, tri_gnormal, tri_dndx, tri_dndy, tri_ab, tri_ac, tri_abxac;
	std::vector<int32_t> tri_v, tri_n, tri_t, leaf_refs;
	std::vector<FrayGpuKdNode> kd_nodes;
	std::vector<float> texels;
	std::map<const Geometry*, int> geomIdx;
	std::map<const Shader*, int> shaderIdx;
	std::map<const Texture*, int> texIdx;
	std::string error;

	int addBitmap(const Bitmap& bmp) // class Bitmap, src/bitmap.h:30-60
	{
		FrayGpuBitmap b;
		memset(&b, 0, sizeof(b));
		b.width = bmp.width;
		b.height = bmp.height;
		b.first_texel = (int64_t) texels.size() / 3;
		for (int i = 0; i < bmp.width * bmp.height; i++) {
			texels.push_back(bmp.data[i].r);
			texels.push_back(bmp.data[i].g);
			texels.push_back(bmp.data[i].b);
		}
		bitmaps.push_back(b);
		return (int) bitmaps.size() - 1;
	}
	int texture(const Texture* t)
	{
		if (!t) return -1;
		auto it = texIdx.find(t);
		return it == texIdx.end() ? -1 : it->second;
	}
	// struct KDTreeNode (src/mesh.h:35-53): a node's two children are adjacent, allocated when the node is visited
	void addKd(const KDTreeNode* node, size_t idx, size_t firstNode, size_t firstRef)
	{
		FrayGpuKdNode n;
		memset(&n, 0, sizeof(n));
		if (node->isLeafNode()) {
			n.axis = 3;
			n.a = (int32_t) (leaf_refs.size() - firstRef);
			n.b = (int32_t) node->triangles->size();
			leaf_refs.insert(leaf_refs.end(), node->triangles->begin(), node->triangles->end());
			kd_nodes[idx] = n;
			return;
		}
		const size_t child = kd_nodes.size();
		kd_nodes.push_back(n);
		kd_nodes.push_back(n);
		n.axis = (int32_t) node->axis;
		n.a = (int32_t) (child - firstNode);
		n.split = node->splitPos;
		kd_nodes[idx] = n;
		addKd(&node->children[0], child, firstNode, firstRef);
		addKd(&node->children[1], child + 1, firstNode, firstRef);
	}
	int addMesh(const Mesh& m) // class Mesh after beginRender(), src/mesh.h:55-100, src/mesh.cpp:67-94
	{
		FrayGpuMesh fm;
		memset(&fm, 0, sizeof(fm));
		fm.flags = (m.faceted || m.normals.empty() ? FRAY_MESH_FACETED : 0) | (m.backfaceCulling ? FRAY_MESH_BACKFACE_CULL : 0) |
		           (!m.normals.empty() ? FRAY_MESH_HAS_NORMALS : 0) | (!m.uvs.empty() ? FRAY_MESH_HAS_UVS : 0);
		fm.first_vertex = (int32_t) (vertices.size() / 3);
		fm.num_vertices = (int32_t) m.vertices.size();
		for (const Vector& v: m.vertices) push3(vertices, v);
		fm.first_normal = (int32_t) (normals.size() / 3);
		fm.num_normals = (int32_t) m.normals.size();
		for (const Vector& v: m.normals) push3(normals, v);
		fm.first_uv = (int32_t) (uvs.size() / 3);
		fm.num_uvs = (int32_t) m.uvs.size();
		for (const Vector& v: m.uvs) push3(uvs, v);
		fm.first_triangle = (int32_t) (tri_v.size() / 3);
		fm.num_triangles = (int32_t) m.triangles.size();
		for (const Triangle& t: m.triangles) {
			for (int k = 0; k < 3; k++) {
				tri_v.push_back(t.v[k]);
				tri_n.push_back(t.n[k]);
				tri_t.push_back(t.t[k]);
			}
			push3(tri_gnormal, t.gnormal);
			push3(tri_dndx, t.dNdx);
			push3(tri_dndy, t.dNdy);
			push3(tri_ab, t.AB);
			push3(tri_ac, t.AC);
			push3(tri_abxac, t.ABcrossAC);
		}
		fm.first_kd_node = (int32_t) kd_nodes.size();
		fm.first_leaf_ref = (int32_t) leaf_refs.size();
		fm.kd_root = -1;
		if (m.kdRoot) {
			const size_t firstNode = kd_nodes.size(), firstRef = leaf_refs.size();
			FrayGpuKdNode zero;
			memset(&zero, 0, sizeof(zero));
			kd_nodes.push_back(zero);
			addKd(m.kdRoot, firstNode, firstNode, firstRef);
			fm.kd_root = 0;
		}
		fm.num_kd_nodes = (int32_t) kd_nodes.size() - fm.first_kd_node;
		fm.num_leaf_refs = (int32_t) leaf_refs.size() - fm.first_leaf_ref;
		put3(fm.bbox_min, m.bbox.vmin);
		put3(fm.bbox_max, m.bbox.vmax);
		meshes.push_back(fm);
		return (int) meshes.size() - 1;
	}
	int csgDepth(const Geometry* g)
	{
		const CsgOp* op = dynamic_cast<const CsgOp*>(g);
		if (!op) return 0;
		return 1 + std::max(csgDepth(op->left), csgDepth(op->right));
	}

	bool build(Scene& sc)
	{
		memset(&view, 0, sizeof(view));
		const GlobalSettings& st = sc.settings; // struct GlobalSettings, src/scene.h:252-278
		view.settings.frame_width = st.frameWidth;
		view.settings.frame_height = st.frameHeight;
		view.settings.max_trace_depth = st.maxTraceDepth;
		view.settings.gi = st.gi;
		view.settings.num_paths = st.numPaths;
		view.settings.want_aa = st.wantAA;
		put3(view.settings.ambient, st.ambientLight);
		view.settings.saturation = st.saturation;
		flattenCamera(*sc.camera, view.camera);

		for (Texture* t: sc.textures) { // src/shading.h:33-107, 227-237
			FrayGpuTexture ft;
			memset(&ft, 0, sizeof(ft));
			ft.bitmap = -1;
			ft.scaling = 1;
			if (auto* c = dynamic_cast<CheckerTexture*>(t)) {
				ft.type = FRAY_TEX_CHECKER;
				put3(ft.color1, c->color1);
				put3(ft.color2, c->color2);
				ft.scaling = c->scaling;
			} else if (auto* b = dynamic_cast<BitmapTexture*>(t)) {
				ft.type = FRAY_TEX_BITMAP;
				ft.scaling = b->scaling;
				ft.bitmap = addBitmap(b->bmp);
			} else if (auto* bm = dynamic_cast<BumpTexture*>(t)) {
				ft.type = FRAY_TEX_BUMP;
				ft.scaling = bm->scaling;
				ft.bump_intensity = bm->bumpIntensity;
				ft.bitmap = addBitmap(bm->bumpTex);
			} else if (auto* f = dynamic_cast<FresnelTexture*>(t)) {
				ft.type = FRAY_TEX_FRESNEL;
				ft.ior = f->ior;
			} else {
				error = "unknown texture class";
				return false;
			}
			texIdx[t] = (int) textures.size();
			textures.push_back(ft);
		}

		for (size_t i = 0; i < sc.shaders.size(); i++) shaderIdx[sc.shaders[i]] = (int) i;
		for (Shader* s: sc.shaders) { // src/shading.h:109-255
			FrayGpuShader fs;
			memset(&fs, 0, sizeof(fs));
			fs.texture = texture(s->diffuseTex);
			if (auto* c = dynamic_cast<ConstantShader*>(s)) {
				fs.type = FRAY_SHADER_CONST;
				put3(fs.color, c->color);
			} else if (auto* l = dynamic_cast<Lambert*>(s)) {
				fs.type = FRAY_SHADER_LAMBERT;
				put3(fs.color, l->color);
			} else if (auto* p = dynamic_cast<Phong*>(s)) {
				fs.type = FRAY_SHADER_PHONG;
				put3(fs.color, p->color);
				put3(fs.specular_color, p->specularColor);
				fs.exponent = p->exponent;
				fs.specular_multiplier = p->specularMultiplier;
			} else if (auto* r = dynamic_cast<Reflection*>(s)) {
				fs.type = FRAY_SHADER_REFL;
				put3(fs.mult, r->mult);
				fs.num_samples = r->numSamples;
				fs.pure_reflection = r->pureReflection;
				fs.deflection_scaling = r->deflectionScaling;
			} else if (auto* rf = dynamic_cast<Refraction*>(s)) {
				fs.type = FRAY_SHADER_REFR;
				put3(fs.mult, rf->mult);
				fs.ior = rf->ior;
			} else if (auto* ly = dynamic_cast<Layered*>(s)) {
				fs.type = FRAY_SHADER_LAYERED;
				fs.first_layer = (int32_t) layers.size();
				fs.num_layers = ly->numLayers;
				for (int k = 0; k < ly->numLayers; k++) {
					FrayGpuLayer fl;
					memset(&fl, 0, sizeof(fl));
					fl.shader = shaderIdx.at(ly->layers[k].shader);
					fl.texture = texture(ly->layers[k].texture);
					put3(fl.opacity, ly->layers[k].opacity);
					layers.push_back(fl);
				}
			} else {
				error = "unknown shader class";
				return false;
			}
			shaders.push_back(fs);
		}

		for (size_t i = 0; i < sc.geometries.size(); i++) geomIdx[sc.geometries[i]] = (int) i;
		for (Geometry* g: sc.geometries) { // src/geometry.h:54-154, src/mesh.h:55-100
			FrayGpuGeometry fg;
			memset(&fg, 0, sizeof(fg));
			fg.mesh = fg.left = fg.right = -1;
			if (auto* p = dynamic_cast<Plane*>(g)) {
				fg.type = FRAY_GEOM_PLANE;
				fg.p[0] = p->height;
				fg.p[1] = p->limit;
			} else if (auto* s = dynamic_cast<Sphere*>(g)) {
				fg.type = FRAY_GEOM_SPHERE;
				fg.p[0] = s->O.x; fg.p[1] = s->O.y; fg.p[2] = s->O.z; fg.p[3] = s->R;
			} else if (auto* c = dynamic_cast<Cube*>(g)) {
				fg.type = FRAY_GEOM_CUBE;
				fg.p[0] = c->O.x; fg.p[1] = c->O.y; fg.p[2] = c->O.z; fg.p[3] = c->halfSide;
			} else if (auto* op = dynamic_cast<CsgOp*>(g)) {
				fg.type = dynamic_cast<CsgPlus*>(g) ? FRAY_GEOM_CSG_PLUS : (dynamic_cast<CsgIntersect*>(g) ? FRAY_GEOM_CSG_AND : FRAY_GEOM_CSG_MINUS);
				if (csgDepth(op) > FRAY_GPU_MAX_CSG_DEPTH) {
					error = "CSG nesting deeper than the GPU back end supports";
					return false;
				}
				fg.left = geomIdx.at(op->left);
				fg.right = geomIdx.at(op->right);
			} else if (auto* m = dynamic_cast<Mesh*>(g)) {
				fg.type = FRAY_GEOM_MESH;
				fg.mesh = addMesh(*m);
			} else {
				error = "unknown geometry class";
				return false;
			}
			geometries.push_back(fg);
		}

		for (Node* n: sc.nodes) { // struct Node, src/geometry.h:158-177 (only nodes WITH a shader are in scene.nodes, src/scene.cpp:563-568)
			FrayGpuNode fn;
			memset(&fn, 0, sizeof(fn));
			if (!n->geometry) {
				error = "node without geometry";
				return false;
			}
			fn.geometry = geomIdx.at(n->geometry);
			fn.shader = shaderIdx.at(n->shader);
			fn.bump = (n->bump && n->bump->getInterface(BumpMapperInterface::ID)) ? texture(n->bump) : -1; // src/main.cpp:82-90
			putTransform(fn.T, n->T);
			nodes.push_back(fn);
		}

		for (Light* l: sc.lights) { // src/lights.h:32-99 (state after beginFrame, src/lights.cpp:37-46)
			FrayGpuLight fl;
			memset(&fl, 0, sizeof(fl));
			put3(fl.color, l->color);
			fl.power = l->power;
			fl.x_subd = fl.y_subd = 1;
			Transform identity;
			putTransform(fl.T, identity);
			if (auto* p = dynamic_cast<PointLight*>(l)) {
				fl.type = FRAY_LIGHT_POINT;
				put3(fl.pos, p->pos);
			} else if (auto* r = dynamic_cast<RectLight*>(l)) {
				fl.type = FRAY_LIGHT_RECT;
				fl.x_subd = r->xSubd;
				fl.y_subd = r->ySubd;
				putTransform(fl.T, r->T);
				put3(fl.center, r->center);
				fl.area = r->area;
			} else {
				error = "unknown light class";
				return false;
			}
			lights.push_back(fl);
		}

		view.has_environment = 0;
		for (int i = 0; i < 6; i++) view.env_bitmaps[i] = -1;
		if (auto* env = dynamic_cast<CubemapEnvironment*>(sc.environment)) { // src/environment.h:52-77
			view.has_environment = 1;
			for (int i = 0; i < 6; i++)
				if (env->maps[i] && env->maps[i]->isOK()) view.env_bitmaps[i] = addBitmap(*env->maps[i]);
		}

		view.abi_version = FRAY_GPU_ABI_VERSION;
		view.num_nodes = (int32_t) nodes.size();           view.nodes = nodes.data();
		view.num_geometries = (int32_t) geometries.size(); view.geometries = geometries.data();
		view.num_meshes = (int32_t) meshes.size();         view.meshes = meshes.data();
		view.num_shaders = (int32_t) shaders.size();       view.shaders = shaders.data();
		view.num_layers = (int32_t) layers.size();         view.layers = layers.data();
		view.num_textures = (int32_t) textures.size();     view.textures = textures.data();
		view.num_bitmaps = (int32_t) bitmaps.size();       view.bitmaps = bitmaps.data();
		view.num_lights = (int32_t) lights.size();         view.lights = lights.data();
		view.num_vertices = (int64_t) vertices.size() / 3; view.vertices = vertices.data();
		view.num_normals = (int64_t) normals.size() / 3;   view.normals = normals.data();
		view.num_uvs = (int64_t) uvs.size() / 3;           view.uvs = uvs.data();
		view.num_triangles = (int64_t) tri_v.size() / 3;
		view.tri_v = tri_v.data(); view.tri_n = tri_n.data(); view.tri_t = tri_t.data();
		view.tri_gnormal = tri_gnormal.data(); view.tri_dndx = tri_dndx.data(); view.tri_dndy = tri_dndy.data();
		view.tri_ab = tri_ab.data(); view.tri_ac = tri_ac.data(); view.tri_abxac = tri_abxac.data();
		view.num_kd_nodes = (int64_t) kd_nodes.size();     view.kd_nodes = kd_nodes.data();
		view.num_leaf_refs = (int64_t) leaf_refs.size();   view.leaf_refs = leaf_refs.data();
		view.num_texels = (int64_t) texels.size() / 3;     view.texels = texels.data();
		return true;
	}
};

// the C ABI, bound at run time (the reference links nothing of this repository)
struct GpuApi {
	void* handle = nullptr;
	int (*create)(const FrayGpuScene*, int, int, FrayGpuCtx**) = nullptr;
	int (*update_camera)(FrayGpuCtx*, const FrayGpuCamera*) = nullptr;
	int (*render)(FrayGpuCtx*, const FrayGpuFrame*, float*, FrayGpuStats*) = nullptr;
	void (*destroy)(FrayGpuCtx*) = nullptr;
	const char* (*last_error)(void) = nullptr;
	bool load(const char* path)
	{
		handle = dlopen(path, RTLD_NOW | RTLD_LOCAL);
		if (!handle) { fprintf(stderr, "fray_ref_gpu: %s\n", dlerror()); return false; }
		create = (decltype(create)) dlsym(handle, "fray_gpu_create");
		update_camera = (decltype(update_camera)) dlsym(handle, "fray_gpu_update_camera");
		render = (decltype(render)) dlsym(handle, "fray_gpu_render");
		destroy = (decltype(destroy)) dlsym(handle, "fray_gpu_destroy");
		last_error = (decltype(last_error)) dlsym(handle, "fray_gpu_last_error");
		return create && update_camera && render && destroy && last_error;
	}
};

GpuApi gpuApi;
FrayGpuCtx* gpu = nullptr;

// render(), src/main.cpp:373-405, with everything below scene.beginFrame() replaced by the C ABI
bool renderOnGpu()
{
	scene.beginFrame();
	FrayGpuCamera cam;
	flattenCamera(*scene.camera, cam);
	if (gpuApi.update_camera(gpu, &cam) != 0) { fprintf(stderr, "%s\n", gpuApi.last_error()); return false; }
	static std::vector<float> rgb;
	rgb.resize((size_t) frameWidth() * frameHeight() * 3);
	FrayGpuFrame f;
	memset(&f, 0, sizeof(f));
	f.seed = 42; // initRandom(42), src/main.cpp:502; spp 0 = the rule of src/main.cpp:395-400
	FrayGpuStats st;
	if (gpuApi.render(gpu, &f, rgb.data(), &st) != 0) { fprintf(stderr, "%s\n", gpuApi.last_error()); return false; }
	for (int y = 0; y < frameHeight(); y++) // the consumers of vfb (displayVFB*, takeScreenshot) stay as they are
		for (int x = 0; x < frameWidth(); x++) {
			const float* p = &rgb[3 * ((size_t) y * frameWidth() + x)];
			vfb[y][x] = Color(p[0], p[1], p[2]);
		}
	printf("fray_ref_gpu: %llu rays, %.3f ms on the device\n", (unsigned long long) st.rays, st.device_ms);
	return true;
}

} // namespace

int main(int argc, char** argv)
{
	if (argc < 4) {
		fprintf(stderr, "usage: fray_ref_gpu scene.fray --dump out.flat | --render out.f32 [--fp64] [--lib libfray_gpu.so]\n");
		return -1;
	}
	const char* dump = nullptr;
	const char* renderOut = nullptr;
	const char* lib = "libfray_gpu.so";
	int precision = FRAY_GPU_FP32;
	for (int i = 2; i < argc; i++) {
		if (!strcmp(argv[i], "--dump") && i + 1 < argc) dump = argv[++i];
		else if (!strcmp(argv[i], "--render") && i + 1 < argc) renderOut = argv[++i];
		else if (!strcmp(argv[i], "--lib") && i + 1 < argc) lib = argv[++i];
		else if (!strcmp(argv[i], "--fp64")) precision = FRAY_GPU_FP64;
	}
	initRandom(42);                                        // as main(), src/main.cpp:502-514
	if (!scene.parseScene(argv[1])) return -3;
	initGraphics(scene.settings.frameWidth, scene.settings.frameHeight, false);
	scene.beginRender();
	scene.beginFrame();
	FlatBuilder fb;
	if (!fb.build(scene)) { fprintf(stderr, "fray_ref_gpu: %s\n", fb.error.c_str()); return -4; }
	if (dump && frayDumpFlat(&fb.view, dump) != 0) { fprintf(stderr, "fray_ref_gpu: cannot write %s\n", dump); return -5; }
	if (renderOut) {
		if (!gpuApi.load(lib)) return -6;
		if (gpuApi.create(&fb.view, 0, precision, &gpu) != 0) { fprintf(stderr, "fray_ref_gpu: %s\n", gpuApi.last_error()); return -7; }
		if (!renderOnGpu()) return -8;
		FILE* f = fopen(renderOut, "wb");
		if (!f) return -9;
		int32_t wh[2] = { frameWidth(), frameHeight() };
		fwrite(wh, sizeof(wh), 1, f);
		for (int y = 0; y < frameHeight(); y++) fwrite(&vfb[y][0], sizeof(Color), frameWidth(), f);
		fclose(f);
		gpuApi.destroy(gpu);
	}
	return 0;
}
