// flatten.cpp -- turn the prepared Scene object graph into the POD tables of include/fray_gpu.h.
//
// This is the host half of the drop-in boundary: what the reference's RendMT worker reads through the global
// `scene` (/root/reference/src/scene.h:280-299) is written here as index-linked arrays that the GPU context
// uploads. Scene::beginRender() and Scene::beginFrame() must have run (KD trees built, bump maps differentiated,
// camera basis / light areas / glossy scalings derived), exactly the state render() sees at src/main.cpp:376.
#include <cstring>
#include <map>

#include "scene.h"

namespace fray {

static void put3(double* dst, const Vec3& v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }
static void put3(float* dst, const Color& c) { dst[0] = c.r; dst[1] = c.g; dst[2] = c.b; }
static void push3(std::vector<double>& v, const Vec3& p) { v.push_back(p.x); v.push_back(p.y); v.push_back(p.z); }

static void putTransform(FrayGpuTransform& out, const Transform& T)
{
	put3(out.offset, T.offset);
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) {
			out.m[3 * i + j] = T.m.m[i][j];
			out.inv[3 * i + j] = T.inv.m[i][j];
		}
}

void FlatScene::rebind()
{
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
}

void flattenCamera(const Scene& scene, FrayGpuCamera& c)
{
	const Camera& cam = *scene.camera;
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

namespace {
struct Flattener {
	Scene& scene;
	FlatScene& out;
	std::map<const Geometry*, int> geomIdx;
	std::map<const Shader*, int> shaderIdx;
	std::map<const Texture*, int> texIdx;

	int addBitmap(const Bitmap& bmp)
	{
		FrayGpuBitmap b;
		b.width = bmp.width;
		b.height = bmp.height;
		b.first_texel = (int64_t) out.texels.size() / 3;
		for (const Color& c: bmp.data) {
			out.texels.push_back(c.r);
			out.texels.push_back(c.g);
			out.texels.push_back(c.b);
		}
		out.bitmaps.push_back(b);
		return (int) out.bitmaps.size() - 1;
	}

	int texture(const Texture* t)
	{
		if (!t) return -1;
		auto it = texIdx.find(t);
		return it == texIdx.end() ? -1 : it->second;
	}

	int addMesh(const Mesh& m)
	{
		FrayGpuMesh fm;
		memset(&fm, 0, sizeof(fm));
		fm.flags = (m.faceted || m.normals.empty() ? FRAY_MESH_FACETED : 0) | (m.backfaceCulling ? FRAY_MESH_BACKFACE_CULL : 0) |
		           (!m.normals.empty() ? FRAY_MESH_HAS_NORMALS : 0) | (!m.uvs.empty() ? FRAY_MESH_HAS_UVS : 0);
		fm.first_vertex = (int32_t) (out.vertices.size() / 3);
		fm.num_vertices = (int32_t) m.vertices.size();
		for (const Vec3& v: m.vertices) push3(out.vertices, v);
		fm.first_normal = (int32_t) (out.normals.size() / 3);
		fm.num_normals = (int32_t) m.normals.size();
		for (const Vec3& v: m.normals) push3(out.normals, v);
		fm.first_uv = (int32_t) (out.uvs.size() / 3);
		fm.num_uvs = (int32_t) m.uvs.size();
		for (const Vec3& v: m.uvs) push3(out.uvs, v);
		fm.first_triangle = (int32_t) (out.tri_v.size() / 3);
		fm.num_triangles = (int32_t) m.triangles.size();
		for (const Triangle& t: m.triangles) {
			for (int k = 0; k < 3; k++) {
				out.tri_v.push_back(t.v[k]);
				out.tri_n.push_back(t.n[k]);
				out.tri_t.push_back(t.t[k]);
			}
			push3(out.tri_gnormal, t.gnormal);
			push3(out.tri_dndx, t.dNdx);
			push3(out.tri_dndy, t.dNdy);
			push3(out.tri_ab, t.AB);
			push3(out.tri_ac, t.AC);
			push3(out.tri_abxac, t.ABcrossAC);
		}
		fm.first_kd_node = (int32_t) out.kd_nodes.size();
		fm.num_kd_nodes = (int32_t) m.kdNodes.size();
		fm.kd_root = m.kdNodes.empty() ? -1 : 0;
		out.kd_nodes.insert(out.kd_nodes.end(), m.kdNodes.begin(), m.kdNodes.end());
		fm.first_leaf_ref = (int32_t) out.leaf_refs.size();
		fm.num_leaf_refs = (int32_t) m.leafRefs.size();
		out.leaf_refs.insert(out.leaf_refs.end(), m.leafRefs.begin(), m.leafRefs.end());
		put3(fm.bbox_min, m.bbox.vmin);
		put3(fm.bbox_max, m.bbox.vmax);
		out.meshes.push_back(fm);
		return (int) out.meshes.size() - 1;
	}

	int csgDepth(const Geometry* g)
	{
		const CsgOp* op = dynamic_cast<const CsgOp*>(g);
		if (!op) return 0;
		return 1 + std::max(csgDepth(op->left), csgDepth(op->right));
	}

	bool run()
	{
		out = FlatScene();
		memset(&out.view, 0, sizeof(out.view));
		FrayGpuScene& v = out.view;
		const GlobalSettings& st = scene.settings;
		v.settings.frame_width = st.frameWidth;
		v.settings.frame_height = st.frameHeight;
		v.settings.max_trace_depth = st.maxTraceDepth;
		v.settings.gi = st.gi;
		v.settings.num_paths = st.numPaths;
		v.settings.want_aa = st.wantAA;
		put3(v.settings.ambient, st.ambientLight);
		v.settings.saturation = st.saturation;
		flattenCamera(scene, v.camera);

		// textures first (shaders and nodes refer to them)
		for (Texture* t: scene.textures) {
			FrayGpuTexture ft;
			memset(&ft, 0, sizeof(ft));
			ft.type = t->texType();
			ft.bitmap = -1;
			ft.scaling = 1;
			if (auto* c = dynamic_cast<CheckerTexture*>(t)) {
				put3(ft.color1, c->color1);
				put3(ft.color2, c->color2);
				ft.scaling = c->scaling;
			} else if (auto* b = dynamic_cast<BitmapTexture*>(t)) {
				ft.scaling = b->scaling;
				ft.bitmap = addBitmap(b->bmp);
			} else if (auto* bm = dynamic_cast<BumpTexture*>(t)) {
				ft.scaling = bm->scaling;
				ft.bump_intensity = bm->bumpIntensity;
				ft.bitmap = addBitmap(bm->bumpTex);
			} else if (auto* f = dynamic_cast<FresnelTexture*>(t)) {
				ft.ior = f->ior;
			}
			texIdx[t] = (int) out.textures.size();
			out.textures.push_back(ft);
		}

		for (size_t i = 0; i < scene.shaders.size(); i++) shaderIdx[scene.shaders[i]] = (int) i;
		for (Shader* s: scene.shaders) {
			FrayGpuShader fs;
			memset(&fs, 0, sizeof(fs));
			fs.type = s->shaderType();
			fs.texture = texture(s->diffuseTex);
			if (auto* c = dynamic_cast<ConstantShader*>(s)) {
				put3(fs.color, c->color);
			} else if (auto* l = dynamic_cast<Lambert*>(s)) {
				put3(fs.color, l->color);
			} else if (auto* p = dynamic_cast<Phong*>(s)) {
				put3(fs.color, p->color);
				put3(fs.specular_color, p->specularColor);
				fs.exponent = p->exponent;
				fs.specular_multiplier = p->specularMultiplier;
			} else if (auto* r = dynamic_cast<Reflection*>(s)) {
				put3(fs.mult, r->mult);
				fs.num_samples = r->numSamples;
				fs.pure_reflection = r->pureReflection;
				fs.deflection_scaling = r->deflectionScaling;
			} else if (auto* rf = dynamic_cast<Refraction*>(s)) {
				put3(fs.mult, rf->mult);
				fs.ior = rf->ior;
			} else if (auto* ly = dynamic_cast<Layered*>(s)) {
				fs.first_layer = (int32_t) out.layers.size();
				fs.num_layers = (int32_t) ly->layers.size();
				for (const Layered::Layer& L: ly->layers) {
					FrayGpuLayer fl;
					memset(&fl, 0, sizeof(fl));
					fl.shader = shaderIdx.at(L.shader);
					fl.texture = texture(L.texture);
					put3(fl.opacity, L.opacity);
					out.layers.push_back(fl);
				}
			}
			out.shaders.push_back(fs);
		}

		for (size_t i = 0; i < scene.geometries.size(); i++) geomIdx[scene.geometries[i]] = (int) i;
		for (Geometry* g: scene.geometries) {
			FrayGpuGeometry fg;
			memset(&fg, 0, sizeof(fg));
			fg.type = g->geomType();
			fg.mesh = fg.left = fg.right = -1;
			if (auto* p = dynamic_cast<Plane*>(g)) {
				fg.p[0] = p->height;
				fg.p[1] = p->limit;
			} else if (auto* s = dynamic_cast<Sphere*>(g)) {
				fg.p[0] = s->O.x; fg.p[1] = s->O.y; fg.p[2] = s->O.z; fg.p[3] = s->R;
			} else if (auto* c = dynamic_cast<Cube*>(g)) {
				fg.p[0] = c->O.x; fg.p[1] = c->O.y; fg.p[2] = c->O.z; fg.p[3] = c->halfSide;
			} else if (auto* op = dynamic_cast<CsgOp*>(g)) {
				if (csgDepth(op) > FRAY_GPU_MAX_CSG_DEPTH) {
					scene.lastError = "CSG nesting deeper than the GPU back end supports: " + g->name;
					return false;
				}
				fg.left = geomIdx.at(op->left);
				fg.right = geomIdx.at(op->right);
			} else if (auto* m = dynamic_cast<Mesh*>(g)) {
				fg.mesh = addMesh(*m);
			}
			out.geometries.push_back(fg);
		}

		for (Node* n: scene.nodes) {
			FrayGpuNode fn;
			memset(&fn, 0, sizeof(fn));
			if (!n->geometry) {
				scene.lastError = "node without geometry: " + n->name;
				return false;
			}
			fn.geometry = geomIdx.at(n->geometry);
			fn.shader = shaderIdx.at(n->shader);
			// only textures that implement BumpMapperInterface change the normal (src/main.cpp:82-90)
			fn.bump = (n->bump && n->bump->getInterface(BumpMapperInterface::ID)) ? texture(n->bump) : -1;
			putTransform(fn.T, n->T);
			out.nodes.push_back(fn);
		}

		for (Light* l: scene.lights) {
			FrayGpuLight fl;
			memset(&fl, 0, sizeof(fl));
			fl.type = l->lightType();
			put3(fl.color, l->color);
			fl.power = l->power;
			fl.x_subd = fl.y_subd = 1;
			Transform identity;
			putTransform(fl.T, identity);
			if (auto* p = dynamic_cast<PointLight*>(l)) {
				put3(fl.pos, p->pos);
			} else if (auto* r = dynamic_cast<RectLight*>(l)) {
				fl.x_subd = r->xSubd;
				fl.y_subd = r->ySubd;
				putTransform(fl.T, r->T);
				put3(fl.center, r->center);
				fl.area = r->area;
			}
			out.lights.push_back(fl);
		}

		v.has_environment = 0;
		for (int i = 0; i < 6; i++) v.env_bitmaps[i] = -1;
		if (auto* env = dynamic_cast<CubemapEnvironment*>(scene.environment)) {
			v.has_environment = 1;
			for (int i = 0; i < 6; i++)
				if (env->maps[i] && env->maps[i]->isOK()) v.env_bitmaps[i] = addBitmap(*env->maps[i]);
		}
		out.rebind();
		return true;
	}
};
} // namespace

bool flatten(Scene& scene, FlatScene& out)
{
	if (!scene.camera) {
		scene.lastError = "scene has no camera";
		return false;
	}
	Flattener f{ scene, out };
	return f.run();
}

} // namespace fray
