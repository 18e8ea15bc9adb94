// scene_image.h -- builds the device-layout image of a scene from the C-ABI tables (include/fray_gpu.h).
//
// Host-side, plain C++. The image is ONE contiguous blob (16-byte aligned sections) plus a DScene<R> whose pointers
// are offsets into it; fray_gpu.cu copies the blob to HBM with a single cudaMemcpy and rebases the pointers.
// Conversions done here, once per scene:
//   * FP64 tables -> R (float for the fast path, double for the parity path)
//   * per-triangle vertex A gathered (the reference indexes vertices[T.v[0]] on every test, src/mesh.cpp:107)
//   * KD node boxes: the reference re-derives each child box by splitting the parent's while descending
//     (src/mesh.cpp:371-373, src/bbox.h:205-211); the same boxes are materialised per node here
//   * relative indices -> absolute, identity transforms flagged, "does anything on this node read (u,v)" flagged
#pragma once
#include <algorithm>
#include <array>
#include <functional>
#include <map>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "core.cuh"

namespace fray {

template <typename R> struct SceneImage {
	std::vector<unsigned char> blob;
	DScene<R> offsets; // pointer members hold byte offsets into blob
	int features = 0;  // FRAY_F_* needed by this scene

	template <typename T> size_t append(const std::vector<T>& v)
	{
		size_t off = (blob.size() + 15) & ~(size_t) 15;
		blob.resize(off + v.size() * sizeof(T) + 16, 0); // never empty: a null offset means "first table"
		if (!v.empty()) memcpy(blob.data() + off, v.data(), v.size() * sizeof(T));
		return off;
	}

	static void cvt3(R* dst, const double* src) { dst[0] = (R) src[0]; dst[1] = (R) src[1]; dst[2] = (R) src[2]; }
	static void cvtXform(DXform<R>& o, const FrayGpuTransform& t)
	{
		bool ident = true;
		for (int i = 0; i < 9; i++) {
			o.m[i] = (R) t.m[i];
			o.inv[i] = (R) t.inv[i];
			const double e = (i % 4 == 0) ? 1.0 : 0.0;
			if (t.m[i] != e || t.inv[i] != e) ident = false;
		}
		for (int i = 0; i < 3; i++) {
			o.off[i] = (R) t.offset[i];
			if (t.offset[i] != 0) ident = false;
		}
		o.identity = ident;
	}

	static bool textureReadsUV(const FrayGpuScene& s, int ti)
	{
		if (ti < 0) return false;
		int t = s.textures[ti].type;
		return t == FRAY_TEX_CHECKER || t == FRAY_TEX_BITMAP || t == FRAY_TEX_BUMP;
	}
	static bool shaderReadsUV(const FrayGpuScene& s, int si, int depth)
	{
		if (si < 0 || depth > 8) return false;
		const FrayGpuShader& sh = s.shaders[si];
		if (textureReadsUV(s, sh.texture)) return true;
		if (sh.type == FRAY_SHADER_LAYERED)
			for (int i = 0; i < sh.num_layers; i++) {
				const FrayGpuLayer& L = s.layers[sh.first_layer + i];
				if (textureReadsUV(s, L.texture) || shaderReadsUV(s, L.shader, depth + 1)) return true;
			}
		return false;
	}
	static bool shaderUsesTexture(const FrayGpuScene& s, int si, int depth)
	{
		if (si < 0 || depth > 8) return false;
		const FrayGpuShader& sh = s.shaders[si];
		if (sh.texture >= 0) return true;
		if (sh.type == FRAY_SHADER_LAYERED)
			for (int i = 0; i < sh.num_layers; i++) {
				const FrayGpuLayer& L = s.layers[sh.first_layer + i];
				if (L.texture >= 0 || shaderUsesTexture(s, L.shader, depth + 1)) return true;
			}
		return false;
	}
	static int layeredDepth(const FrayGpuScene& s, int si, int depth)
	{
		if (depth > 8) return 99;
		const FrayGpuShader& sh = s.shaders[si];
		if (sh.type != FRAY_SHADER_LAYERED) return 0;
		int d = 0;
		for (int i = 0; i < sh.num_layers; i++) d = std::max(d, layeredDepth(s, s.layers[sh.first_layer + i].shader, depth + 1));
		return 1 + d;
	}


	// ---- flat polygon table (flat.cuh), fast precision only --------------------------------------------------------------
	struct D3 { double x, y, z; };
	static D3 sub(D3 a, D3 b) { return D3{ a.x - b.x, a.y - b.y, a.z - b.z }; }
	static D3 crs(D3 a, D3 b) { return D3{ a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
	static double dt(D3 a, D3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
	static D3 scl(D3 a, double m) { return D3{ a.x * m, a.y * m, a.z * m }; }
	static D3 rowMul(D3 v, const double* m) // row vector times row-major 3x3 (src/matrix.h:53-60)
	{
		return D3{ v.x * m[0] + v.y * m[3] + v.z * m[6], v.x * m[1] + v.y * m[4] + v.z * m[7], v.x * m[2] + v.y * m[5] + v.z * m[8] };
	}
	static float4 plane4(D3 n, double c)
	{
		float4 r;
		r.x = (float) n.x; r.y = (float) n.y; r.z = (float) n.z; r.w = (float) c;
		return r;
	}
	// edge plane through a and b inside the polygon plane with normal n, scaled to 1 at `opposite` (inside = positive)
	static bool edgePlane(D3 a, D3 b, D3 n, D3 opposite, float4& out, double* atOthers = nullptr, const D3* others = nullptr, int numOthers = 0)
	{
		D3 m = crs(n, sub(b, a));
		double v = dt(m, sub(opposite, a));
		if (v == 0) return false;
		m = scl(m, 1 / v);
		out = plane4(m, -dt(m, a));
		for (int i = 0; i < numOthers; i++) atOthers[i] = dt(m, sub(others[i], a));
		return true;
	}

	std::vector<float4> flatPolys;
	std::vector<FlatInfo> flatInfo;

	std::vector<float4> flat2Polys; // two-sided records (flat.cuh): one record and one FlatInfo per polygon
	std::vector<FlatInfo> flat2Info;

	// Two-sided polygons go to the list of their own in the untextured kernel variants (smallpt: six Plane primitives, 12 -> 6
	// records per ray); the textured variants keep a front and a back copy in the main list -- their kernels are at the register
	// limit and lost more to the extra loops than the shorter table gave (zaphod 1.55 -> 1.75 ms)
	bool twoSidedList = false; // set by build() before buildFlat

	void pushFlat(const float4 rec[5], const FlatInfo& fi, bool twoSided)
	{
		if (twoSided && twoSidedList) {
			for (int k = 0; k < 5; k++) flat2Polys.push_back(rec[k]);
			flat2Info.push_back(fi);
			return;
		}
		for (int k = 0; k < 5; k++) flatPolys.push_back(rec[k]);
		flatInfo.push_back(fi);
		if (twoSided) { // the same polygon seen from behind: plane reversed, edges unchanged
			float4 back = rec[0];
			back.x = -back.x; back.y = -back.y; back.z = -back.z; back.w = -back.w;
			flatPolys.push_back(back);
			for (int k = 1; k < 5; k++) flatPolys.push_back(rec[k]);
			flatInfo.push_back(fi);
		}
	}


	// the points that span a light: the corners of a RectLight's unit square in world space, or the PointLight's position
	static std::vector<D3> lightPoints(const FrayGpuLight& l)
	{
		std::vector<D3> pts;
		if (l.type == FRAY_LIGHT_RECT) {
			const D3 off{ l.T.offset[0], l.T.offset[1], l.T.offset[2] };
			for (int k = 0; k < 4; k++) {
				const D3 q = rowMul(D3{ (k & 1) ? 0.5 : -0.5, 0, (k & 2) ? 0.5 : -0.5 }, l.T.m);
				pts.push_back(D3{ q.x + off.x, q.y + off.y, q.z + off.z });
			}
		} else {
			pts.push_back(D3{ l.pos[0], l.pos[1], l.pos[2] });
		}
		return pts;
	}
	// can the plane (Nu, d) have any point of the light behind it (i.e. can a front-face hit shadow that light at all)?
	static bool lightBehind(const D3& Nu, double d, const std::vector<D3>& pts)
	{
		double behind = 1e300, scale = fabs(d) + 1;
		for (const D3& q: pts) {
			behind = std::min(behind, dt(Nu, q) - d);
			scale = std::max(scale, std::max(fabs(q.x), std::max(fabs(q.y), fabs(q.z))));
		}
		return behind < 1e-4 * scale; // otherwise the whole light is clearly in front
	}

	// a flat record of a brute-force mesh before it is emitted, with what the hexahedron search needs
	struct Cand {
		float4 rec[5];
		FlatInfo fi;
		std::vector<D3> v; // polygon vertices
		D3 Nu;             // unit plane normal (front side)
		double d;          // Nu . p = d on the plane
		bool twoSided, attr;
		int hex;           // index of the hexahedron that swallowed the record, or -1
	};
	struct Hex {
		std::vector<int> faces;                       // candidate indices: the real faces
		std::vector<std::pair<D3, double>> caps;      // planes that close an open boundary loop: they clip, but cannot be hit
	};

	// Convex hexahedra (flat.cuh): connected groups of one-sided, faceted records that are ALL the faces of a convex polyhedron
	// seen from outside, with at most FRAY_HEX_PLANES planes including the caps of planar boundary loops.
	void findHexes(std::vector<Cand>& cands, std::vector<Hex>& hexes)
	{
		const int n = (int) cands.size();
		std::map<std::array<double, 3>, int> vid; // vertex ids by exact position
		std::vector<D3> vpos;
		std::map<std::pair<int, int>, std::vector<int>> edges; // undirected edge -> faces using it
		std::vector<std::vector<int>> cv(n);
		for (int c = 0; c < n; c++) {
			if (cands[c].twoSided || cands[c].attr) continue;
			for (const D3& p: cands[c].v) {
				const std::array<double, 3> key{ p.x, p.y, p.z };
				auto it = vid.find(key);
				if (it == vid.end()) {
					it = vid.emplace(key, (int) vpos.size()).first;
					vpos.push_back(p);
				}
				cv[c].push_back(it->second);
			}
			for (size_t k = 0; k < cv[c].size(); k++) {
				const int a = cv[c][k], b = cv[c][(k + 1) % cv[c].size()];
				edges[{ std::min(a, b), std::max(a, b) }].push_back(c);
			}
		}
		std::vector<int> parent(n);
		for (int i = 0; i < n; i++) parent[i] = i;
		std::function<int(int)> find = [&](int x) { return parent[x] == x ? x : parent[x] = find(parent[x]); };
		for (auto& e: edges)
			for (size_t k = 1; k < e.second.size(); k++) parent[find(e.second[k])] = find(e.second[0]);
		std::map<int, std::vector<int>> comps;
		for (int c = 0; c < n; c++)
			if (!cv[c].empty()) comps[find(c)].push_back(c);

		for (auto& kv: comps) {
			const std::vector<int>& members = kv.second;
			if (members.size() < 3 || members.size() > FRAY_HEX_PLANES || (int) hexes.size() >= FRAY_MAX_HEX) continue;
			std::vector<int> verts;
			double extent = 0;
			for (int c: members)
				for (int id: cv[c]) {
					verts.push_back(id);
					extent = std::max(extent, std::max(fabs(vpos[id].x), std::max(fabs(vpos[id].y), fabs(vpos[id].z))));
				}
			const double tol = 1e-9 * (extent + 1), ctol = 1e-7 * (extent + 1);
			// convex and seen from outside: every vertex on or behind every face plane; no two faces in one plane
			bool ok = true;
			for (int c: members)
				for (int id: verts)
					if (dt(cands[c].Nu, vpos[id]) - cands[c].d > tol) ok = false;
			for (size_t i = 0; i < members.size() && ok; i++)
				for (size_t j = i + 1; j < members.size(); j++)
					if (dt(cands[members[i]].Nu, cands[members[j]].Nu) > 1 - 1e-12) ok = false;
			if (!ok) continue;
			// every edge shared by exactly two faces, or on the boundary
			std::map<int, std::vector<int>> boundary; // vertex -> boundary neighbours
			for (int c: members)
				for (size_t k = 0; k < cv[c].size(); k++) {
					const int a = cv[c][k], b = cv[c][(k + 1) % cv[c].size()];
					const size_t users = edges[{ std::min(a, b), std::max(a, b) }].size();
					if (users > 2) ok = false;
					if (users == 1) { boundary[a].push_back(b); boundary[b].push_back(a); }
				}
			// boundary loops -> cap planes, oriented outwards like the faces
			std::vector<std::pair<D3, double>> caps;
			std::map<int, bool> seen;
			for (auto& bv: boundary) {
				if (!ok) break;
				if (seen[bv.first]) continue;
				std::vector<int> loop;
				int prev = -1, cur = bv.first;
				while (ok && !seen[cur]) {
					seen[cur] = true;
					loop.push_back(cur);
					const std::vector<int>& nb = boundary[cur];
					if (nb.size() != 2) { ok = false; break; }
					const int next = nb[0] != prev ? nb[0] : nb[1];
					prev = cur;
					cur = next;
				}
				if (!ok || loop.size() < 3) { ok = false; break; }
				D3 N{ 0, 0, 0 }; // Newell normal of the loop
				for (size_t k = 0; k < loop.size(); k++) {
					const D3 &a = vpos[loop[k]], &b = vpos[loop[(k + 1) % loop.size()]];
					N.x += (a.y - b.y) * (a.z + b.z);
					N.y += (a.z - b.z) * (a.x + b.x);
					N.z += (a.x - b.x) * (a.y + b.y);
				}
				const double l = sqrt(dt(N, N));
				if (!(l > 0)) { ok = false; break; }
				N = scl(N, 1 / l);
				double dcap = dt(N, vpos[loop[0]]);
				for (int id: loop)
					if (fabs(dt(N, vpos[id]) - dcap) > ctol) ok = false; // the loop is not planar
				double lo = 0, hi = 0;
				for (int id: verts) {
					lo = std::min(lo, dt(N, vpos[id]) - dcap);
					hi = std::max(hi, dt(N, vpos[id]) - dcap);
				}
				if (hi > ctol && lo < -ctol) ok = false; // the solid sticks out on both sides of the loop's plane
				else if (hi > ctol) { N = scl(N, -1); dcap = -dcap; }
				for (int c: members)
					if (dt(cands[c].Nu, N) > 1 - 1e-12) ok = false; // a cap in (or parallel behind) a face plane: leave the mesh alone
				caps.emplace_back(N, dcap);
			}
			if (!ok || members.size() + caps.size() > FRAY_HEX_PLANES || members.size() + caps.size() < 4 || caps.size() > FRAY_HEX_MAX_CAPS) continue;
			Hex hx;
			hx.faces = members;
			hx.caps = caps;
			for (int c: members) cands[c].hex = (int) hexes.size();
			hexes.push_back(hx);
		}
	}

	// Fills flatPolys / flatInfo and marks the nodes whose geometry moved into the table. Returns the feature bits used.
	int buildFlat(const FrayGpuScene& s, std::vector<DNode<R>>& nodes)
	{
		flatPolys.clear();
		flatInfo.clear();
		flat2Polys.clear();
		flat2Info.clear();
		int feat = 0;
		const float4 always = plane4(D3{ 0, 0, 0 }, 1.0);
		int numRect = 0;
		for (int i = 0; i < s.num_lights; i++) numRect += s.lights[i].type == FRAY_LIGHT_RECT;
		if (numRect > FRAY_MAX_FLAT / 2) return 0; // absurd: leave everything to the generic loops
		int room = FRAY_MAX_FLAT - numRect;
		std::vector<float4> spheres;
		std::vector<FlatInfo> sphereInfo;
		std::vector<Cand> cands;

		for (int ni = 0; ni < s.num_nodes; ni++) {
			const FrayGpuNode& n = s.nodes[ni];
			const FrayGpuGeometry& g = s.geometries[n.geometry];
			if (g.type == FRAY_GEOM_PLANE) {
				// Plane::intersect, src/geometry.cpp:30-50: the square |x|, |z| <= limit of the object-space plane y = height, hit
				// from either side, normal always +y. Under the node transform that is a world-space parallelogram: two records.
				if (room < (twoSidedList ? 1 : 2)) continue;
				const double height = g.p[0], limit = g.p[1];
				const double* I = n.T.inv;
				const D3 off{ n.T.offset[0], n.T.offset[1], n.T.offset[2] };
				const D3 cx{ I[0], I[3], I[6] }, cy{ I[1], I[4], I[7] }, cz{ I[2], I[5], I[8] }; // object x(p) = (p - off) . cx, ...
				const double ly = sqrt(dt(cy, cy));
				if (!(ly > 0) || !(limit > 0)) continue;
				float4 rec[5];
				const D3 Nu = scl(cy, 1 / ly);
				rec[0] = plane4(Nu, (height + dt(off, cy)) / ly);
				if (limit < 1e30) {
					rec[1] = plane4(scl(cx, -1 / limit), 1 + dt(off, cx) / limit);
					rec[2] = plane4(scl(cx, 1 / limit), 1 - dt(off, cx) / limit);
					rec[3] = plane4(scl(cz, -1 / limit), 1 + dt(off, cz) / limit);
					rec[4] = plane4(scl(cz, 1 / limit), 1 - dt(off, cz) / limit);
				} else {
					rec[1] = rec[2] = rec[3] = rec[4] = always;
				}
				FlatInfo fi;
				memset(&fi, 0, sizeof(fi));
				D3 w{ n.T.m[3], n.T.m[4], n.T.m[5] }; // (0, 1, 0) * m, normalised (src/geometry.cpp:204, src/matrix.cpp:153-156)
				const double lw = sqrt(dt(w, w));
				if (lw > 0) w = scl(w, 1 / lw);
				fi.nx = (float) w.x; fi.ny = (float) w.y; fi.nz = (float) w.z;
				const bool attr = nodes[ni].needsUV || n.bump >= 0;
				fi.node = ni; fi.tri0 = fi.tri1 = -1; fi.mesh = -1; fi.flags = attr ? (FRAY_FLAT_ATTR | FRAY_FLAT_PLANE) : 0;
				pushFlat(rec, fi, true);
				room -= twoSidedList ? 1 : 2;
				nodes[ni].inFlat = 1;
				feat |= FRAY_F_FLAT;
				if (attr) feat |= FRAY_F_ATTR;
				continue;
			}
			if (g.type == FRAY_GEOM_SPHERE) {
				// only under a pure translation: anything else turns the sphere into an ellipsoid (generic node loop)
				bool pure = true;
				for (int k = 0; k < 9; k++)
					if (n.T.m[k] != ((k % 4 == 0) ? 1.0 : 0.0)) pure = false;
				if (!pure || (int) spheres.size() >= 16) continue;
				float4 sp;
				sp.x = (float) (g.p[0] + n.T.offset[0]); sp.y = (float) (g.p[1] + n.T.offset[1]); sp.z = (float) (g.p[2] + n.T.offset[2]);
				sp.w = (float) (g.p[3] * g.p[3]);
				FlatInfo fi;
				memset(&fi, 0, sizeof(fi));
				const bool attr = nodes[ni].needsUV || n.bump >= 0;
				fi.node = ni; fi.tri0 = fi.tri1 = -1; fi.mesh = -1; fi.flags = FRAY_FLAT_SPHERE | (attr ? FRAY_FLAT_ATTR : 0);
				spheres.push_back(sp);
				sphereInfo.push_back(fi);
				nodes[ni].inFlat = 1;
				feat |= FRAY_F_FLAT | FRAY_F_SPHERES;
				if (attr) feat |= FRAY_F_ATTR;
				continue;
			}
			if (g.type != FRAY_GEOM_MESH) continue;
			const FrayGpuMesh& m = s.meshes[g.mesh];
			if (m.kd_root >= 0 || m.num_triangles <= 0) continue;
			const bool cull = (m.flags & FRAY_MESH_BACKFACE_CULL) != 0;
			const bool smooth = !(m.flags & FRAY_MESH_FACETED) && (m.flags & FRAY_MESH_HAS_NORMALS);
			const bool attr = smooth || ((m.flags & FRAY_MESH_HAS_UVS) && nodes[ni].needsUV) || n.bump >= 0;
			const double* M = n.T.m;
			const double det = M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
			if (det == 0) continue;
			const D3 off{ n.T.offset[0], n.T.offset[1], n.T.offset[2] };
			auto world = [&](int tri, int k) {
				const double* v = s.vertices + 3 * ((size_t) m.first_vertex + s.tri_v[3 * ((size_t) m.first_triangle + tri) + k]);
				const D3 p = rowMul(D3{ v[0], v[1], v[2] }, M);
				return D3{ p.x + off.x, p.y + off.y, p.z + off.z };
			};
			// records of this node, built first so that a node is flattened completely or not at all
			std::vector<Cand> mine;
			for (int t = 0; t < m.num_triangles; t++) {
				const size_t ti = (size_t) m.first_triangle + t;
				const D3 A = world(t, 0), B = world(t, 1), C = world(t, 2);
				D3 N = crs(sub(B, A), sub(C, A));
				const double nn = dt(N, N);
				if (!(nn > 0)) continue; // degenerate: |Dcr| < 1e-12 rejects it in the reference as well (src/triangle.cpp:72)
				// reference culls on dot(dir_object, gnormal_object) > 0, i.e. against the TRUE world normal times sign(det)
				const D3 Nf = det > 0 ? N : scl(N, -1);
				FlatInfo fi;
				memset(&fi, 0, sizeof(fi));
				{ // shading normal of a faceted hit: normalize(gnormal * m) -- by m, not its inverse transpose (src/matrix.cpp:153-156)
					const double* gn = s.tri_gnormal + 3 * ti;
					D3 w = rowMul(D3{ gn[0], gn[1], gn[2] }, M);
					const double l = sqrt(dt(w, w));
					if (l > 0) w = scl(w, 1 / l);
					fi.nx = (float) w.x; fi.ny = (float) w.y; fi.nz = (float) w.z;
				}
				fi.node = ni; fi.tri0 = (int) ti; fi.tri1 = -1; fi.mesh = g.mesh; fi.flags = attr ? FRAY_FLAT_ATTR : 0;
				float4 rec[5];
				const D3 Nu = scl(Nf, 1 / sqrt(nn)); // unit plane normal: N.d is the cosine, t is world distance
				rec[0] = plane4(Nu, dt(Nu, A));
				rec[4] = always;
				// try to merge with the next fan triangle (A, C, D) of the same face
				bool merged = false;
				if (!attr && t + 1 < m.num_triangles) {
					const D3 A2 = world(t + 1, 0), C2 = world(t + 1, 1), Dq = world(t + 1, 2);
					const double* g0 = s.tri_gnormal + 3 * ti;
					const double* g1 = s.tri_gnormal + 3 * (ti + 1);
					const bool sameA = A2.x == A.x && A2.y == A.y && A2.z == A.z && C2.x == C.x && C2.y == C.y && C2.z == C.z;
					const bool sameNormal = g0[0] == g1[0] && g0[1] == g1[1] && g0[2] == g1[2]; // one shading normal for both halves
					const double extent = sqrt(std::max(dt(sub(B, A), sub(B, A)), std::max(dt(sub(C, A), sub(C, A)), dt(sub(Dq, A), sub(Dq, A)))));
					const double offPlane = fabs(dt(N, sub(Dq, A))) / sqrt(nn);
					if (sameA && (sameNormal || offPlane <= 1e-9 * extent) && offPlane <= 1e-7 * extent) {
						// convex quadrilateral A, B, C, D: every vertex on the inner side of every edge
						const D3 P[4] = { A, B, C, Dq };
						float4 e[4];
						bool convex = true;
						for (int k = 0; k < 4 && convex; k++) {
							const D3 a = P[k], b = P[(k + 1) & 3];
							const D3 others[2] = { P[(k + 2) & 3], P[(k + 3) & 3] };
							double at[2];
							// scale by the farther of the two remaining vertices
							D3 mN = crs(N, sub(b, a));
							const double v0 = dt(mN, sub(others[0], a)), v1 = dt(mN, sub(others[1], a));
							if (!(v0 > 0 && v1 > 0)) { convex = false; break; }
							convex = edgePlane(a, b, N, v0 > v1 ? others[0] : others[1], e[k], at, others, 2);
						}
						if (convex) {
							for (int k = 0; k < 4; k++) rec[1 + k] = e[k];
							fi.tri1 = (int) ti + 1;
							fi.flags |= FRAY_FLAT_QUAD;
							float4 dg; // positive on B's side of the diagonal A-C
							if (edgePlane(A, C, N, B, dg)) {
								fi.diag = dg;
								merged = true;
							}
						}
					}
				}
				if (!merged) {
					fi.tri1 = -1;
					fi.flags &= ~FRAY_FLAT_QUAD;
					// barycentrics: lambda2 (0 on AC, 1 at B), lambda3 (0 on AB, 1 at C), lambda1 (0 on BC, 1 at A)
					if (!edgePlane(A, C, N, B, rec[1]) || !edgePlane(A, B, N, C, rec[2]) || !edgePlane(B, C, N, A, rec[3])) continue;
					rec[4] = always;
				}
				Cand cd;
				for (int k = 0; k < 5; k++) cd.rec[k] = rec[k];
				cd.fi = fi;
				cd.v = merged ? std::vector<D3>{ A, B, C, world(t + 1, 2) } : std::vector<D3>{ A, B, C };
				cd.Nu = Nu;
				cd.d = dt(Nu, A);
				cd.twoSided = !cull;
				cd.attr = attr;
				cd.hex = -1;
				mine.push_back(cd);
				if (merged) t++;
			}
			const int count = (int) mine.size() * ((cull || twoSidedList) ? 1 : 2);
			if (count > room) continue;
			room -= count;
			cands.insert(cands.end(), mine.begin(), mine.end());
			nodes[ni].inFlat = 1;
			feat |= FRAY_F_FLAT;
			if (attr) feat |= FRAY_F_ATTR;
		}
		std::vector<Hex> hexes;
		if (!getenv("FRAY_GPU_NO_HEX")) findHexes(cands, hexes);
		for (const Cand& cd: cands)
			if (cd.hex < 0) pushFlat(cd.rec, cd.fi, cd.twoSided);
		offsets.numFlatGeom = (int) flatInfo.size();
		// lights: the unit square of light space, seen from its -y side by rays travelling towards +y (src/lights.cpp:79-103)
		for (int li = 0; li < s.num_lights; li++) {
			const FrayGpuLight& l = s.lights[li];
			if (l.type != FRAY_LIGHT_RECT) continue;
			const double* I = l.T.inv;
			const D3 off{ l.T.offset[0], l.T.offset[1], l.T.offset[2] };
			const D3 cx{ I[0], I[3], I[6] }, cy{ I[1], I[4], I[7] }, cz{ I[2], I[5], I[8] }; // columns: x_l(p) = (p - off) . cx
			float4 rec[5];
			const D3 Nf = scl(cy, -1 / sqrt(dt(cy, cy)));
			rec[0] = plane4(Nf, dt(Nf, off));
			rec[1] = plane4(scl(cx, -1), 0.5 + dt(cx, off));
			rec[2] = plane4(cx, 0.5 - dt(cx, off));
			rec[3] = plane4(scl(cz, -1), 0.5 + dt(cz, off));
			rec[4] = plane4(cz, 0.5 - dt(cz, off));
			FlatInfo fi;
			memset(&fi, 0, sizeof(fi));
			fi.node = li; fi.tri0 = fi.tri1 = -1; fi.mesh = -1; fi.flags = FRAY_FLAT_LIGHT;
			pushFlat(rec, fi, false);
			feat |= FRAY_F_FLAT;
		}
		offsets.numFlatAll = (int) flatInfo.size();
		offsets.lightsInFlat = 1;
		// shadow sets (DScene::shadowFirst): a record can stop a ray towards light l only if some point of the light lies behind
		// its plane. For the light sources of a room nearly every wall fails that test.
		for (int l = 0; l < FRAY_SHADOW_LIGHTS; l++) { offsets.shadowFirst[l] = 0; offsets.shadowCount[l] = -1; }
		int shadowRoom = FRAY_MAX_SHADOW;
		for (int li = 0; li < s.num_lights && li < FRAY_SHADOW_LIGHTS; li++) {
			const FrayGpuLight& l = s.lights[li];
			const std::vector<D3> pts = lightPoints(l);
			std::vector<int> keep;
			for (int r = 0; r < offsets.numFlatGeom; r++) {
				const float4& pl = flatPolys[(size_t) FRAY_FLAT_POLY_VEC * r];
				if (lightBehind(D3{ pl.x, pl.y, pl.z }, pl.w, pts)) keep.push_back(r);
			}
			if ((int) keep.size() == offsets.numFlatGeom || (int) keep.size() > shadowRoom) continue; // nothing gained / no room
			shadowRoom -= (int) keep.size();
			offsets.shadowFirst[li] = (int) (flatPolys.size() / FRAY_FLAT_POLY_VEC);
			offsets.shadowCount[li] = (int) keep.size();
			for (int r: keep)
				for (int k = 0; k < FRAY_FLAT_POLY_VEC; k++) flatPolys.push_back(flatPolys[(size_t) FRAY_FLAT_POLY_VEC * r + k]);
		}
		offsets.numFlatTotal = (int) (flatPolys.size() / FRAY_FLAT_POLY_VEC);
		offsets.numFlatSpheres = (int) spheres.size();
		flatPolys.insert(flatPolys.end(), spheres.begin(), spheres.end());
		flatInfo.insert(flatInfo.end(), sphereInfo.begin(), sphereInfo.end());
		// convex hexahedra: header + six planes each (real faces first, then repeats of plane 0, caps in the last slots)
		offsets.numFlatHex = (int) hexes.size();
		for (int l = 0; l < FRAY_SHADOW_LIGHTS; l++) offsets.shadowHex[l] = 0xffffffffu;
		for (size_t k = 0; k < hexes.size(); k++) {
			const Hex& hx = hexes[k];
			int4 hd;
			hd.x = (int) flatInfo.size();
			hd.y = (int) hx.faces.size();
			hd.z = hx.caps.size() == 2 ? 3 : (hx.caps.size() == 1 ? 2 : 0); // which of slots 4, 5 hold caps
			hd.w = 0;
			float4 raw;
			memcpy(&raw, &hd, sizeof(raw));
			flatPolys.push_back(raw);
			std::vector<float4> planes;
			for (int c: hx.faces) {
				planes.push_back(cands[c].rec[0]);
				flatInfo.push_back(cands[c].fi);
			}
			while (planes.size() + hx.caps.size() < FRAY_HEX_PLANES) planes.push_back(planes[0]);
			for (const auto& cp: hx.caps) planes.push_back(plane4(cp.first, cp.second));
			flatPolys.insert(flatPolys.end(), planes.begin(), planes.end());
			for (int li = 0; li < s.num_lights && li < FRAY_SHADOW_LIGHTS; li++) {
				bool can = false;
				const std::vector<D3> pts = lightPoints(s.lights[li]);
				for (int c: hx.faces) can = can || lightBehind(cands[c].Nu, cands[c].d, pts);
				if (!can) offsets.shadowHex[li] &= ~(1u << k);
			}
			feat |= FRAY_F_HEX;
		}
		// two-sided records: after the hexahedra, their FlatInfo after the hexahedron faces'
		offsets.numFlat2 = (int) flat2Info.size();
		offsets.flat2InfoBase = (int) flatInfo.size();
		flatPolys.insert(flatPolys.end(), flat2Polys.begin(), flat2Polys.end());
		flatInfo.insert(flatInfo.end(), flat2Info.begin(), flat2Info.end());
		if (offsets.numFlat2 > 0) feat |= FRAY_F_TWOSIDED;
		offsets.numFlatInfo = (int) flatInfo.size();
		for (FlatInfo& fi: flatInfo) {
			if (fi.flags & FRAY_FLAT_LIGHT) continue;
			const FrayGpuShader& sh = s.shaders[s.nodes[fi.node].shader];
			const float* c = (sh.type == FRAY_SHADER_REFL || sh.type == FRAY_SHADER_REFR) ? sh.mult : sh.color;
			const int4 t{ sh.type, 0, 0, 0 };
			memcpy(&fi.shade, &t, sizeof(float4));
			fi.shade.y = c[0]; fi.shade.z = c[1]; fi.shade.w = c[2];
		}
		if (getenv("FRAY_GPU_VERBOSE")) {
			fprintf(stderr, "fray_gpu: flat table: %d geometry records, %d two-sided records, %d light records, %d spheres, %d convex hexahedra", offsets.numFlatGeom,
			        offsets.numFlat2, offsets.numFlatAll - offsets.numFlatGeom, offsets.numFlatSpheres, offsets.numFlatHex);
			for (int l = 0; l < s.num_lights && l < FRAY_SHADOW_LIGHTS; l++)
				if (offsets.shadowCount[l] >= 0) fprintf(stderr, ", shadow set of light %d: %d", l, offsets.shadowCount[l]);
			fprintf(stderr, "\n");
		}
		return feat;
	}


	// ---- alternative KD trees for the fast precision (OFF by default: FRAY_GPU_SAH_KD=1) -------------------------------------
	// The reference splits every node at the spatial median of its box, cycling x, y, z, until 20 triangles are left or depth
	// 64 is reached (src/mesh.cpp:315-355, src/constants.h:38-39). Those trees travel through the C ABI, the parity precision
	// walks them literally, and the fast precision walks them too. They look like a poor structure -- the teapot of boxed /
	// forest: depth 65, 22-30 triangle tests per walk -- so round 2 measured trees of our own over the same triangles: surface
	// area heuristic over 32 bins per axis, leaf size and traversal cost as parameters (FRAY_KD_LEAF, FRAY_KD_CT). On the CPU
	// model they cut the triangle tests 4x for 1.5x the inner steps; on the B200 (ms, reference trees / SAH leaf 4 / 8 / 16):
	// boxed 2.56 / 3.16 / 2.86 / 2.71, forest 4K 1.75 / 1.75 / 1.68 / 1.71, hw9/dragon 6.55 / 6.19 / 5.83 / 5.87. A leaf's
	// triangle loop is coherent and its loads are independent; inner steps are dependent loads and the irregular trees make
	// neighbouring rays part ways earlier. Two of the five BASELINE configurations against one: the reference's trees stay.
	struct SahTri { float lo[3], hi[3]; int tri; };

	static void buildSahNode(const std::vector<SahTri>& tris, const float* blo, const float* bhi, int depth, size_t nodeIdx, size_t firstNode,
	                         size_t firstRef, std::vector<DKdNode<R>>& kd, std::vector<int>& refs)
	{
		const int n = (int) tris.size();
		auto makeLeaf = [&]() {
			DKdNode<R> leaf;
			leaf.axis = 3;
			leaf.a = (int) refs.size(); // absolute index of the first reference
			leaf.b = n;
			leaf.split = 0;
			for (const SahTri& t: tris) refs.push_back(t.tri);
			kd[nodeIdx] = leaf;
		};
		static const int leafSize = getenv("FRAY_KD_LEAF") ? atoi(getenv("FRAY_KD_LEAF")) : 8;
		static const double ctEnv = getenv("FRAY_KD_CT") ? atof(getenv("FRAY_KD_CT")) : 2.0;
		if (n <= leafSize || depth >= 48) { makeLeaf(); return; }
		const float ext[3] = { bhi[0] - blo[0], bhi[1] - blo[1], bhi[2] - blo[2] };
		const double area = 2.0 * ((double) ext[0] * ext[1] + (double) ext[1] * ext[2] + (double) ext[2] * ext[0]);
		if (!(area > 0)) { makeLeaf(); return; }
		const int BINS = 32;
		const double costTraverse = ctEnv, costTri = 1.5;
		double bestCost = costTri * n; // the cost of not splitting
		int bestAxis = -1;
		float bestSplit = 0;
		for (int axis = 0; axis < 3; axis++) {
			if (!(ext[axis] > 0)) continue;
			const float inv = (float) BINS / ext[axis];
			int startCount[BINS] = { 0 }, endCount[BINS] = { 0 };
			for (int i = 0; i < n; i++) {
				const float lo = std::max(tris[i].lo[axis], blo[axis]), hi = std::min(tris[i].hi[axis], bhi[axis]);
				int b0 = (int) ((lo - blo[axis]) * inv), b1 = (int) ((hi - blo[axis]) * inv);
				b0 = std::min(std::max(b0, 0), BINS - 1);
				b1 = std::min(std::max(b1, b0), BINS - 1);
				startCount[b0]++;
				endCount[b1]++;
			}
			int nLeft = 0, nEnded = 0;
			for (int k = 1; k < BINS; k++) { // plane between bins k-1 and k
				nLeft += startCount[k - 1];   // triangles that start below the plane
				nEnded += endCount[k - 1];    // triangles that end below it
				const int nRight = n - nEnded;
				const float split = blo[axis] + ext[axis] * (float) k / (float) BINS;
				const int a1 = (axis + 1) % 3, a2 = (axis + 2) % 3;
				const double wl = split - blo[axis], wr = bhi[axis] - split;
				const double areaL = 2.0 * (wl * ext[a1] + (double) ext[a1] * ext[a2] + (double) ext[a2] * wl);
				const double areaR = 2.0 * (wr * ext[a1] + (double) ext[a1] * ext[a2] + (double) ext[a2] * wr);
				double cost = costTraverse + costTri * (areaL / area * nLeft + areaR / area * nRight);
				if (nLeft == 0 || nRight == 0) cost *= 0.8; // cutting off empty space is worth a node
				if (cost < bestCost) { bestCost = cost; bestAxis = axis; bestSplit = split; }
			}
		}
		if (bestAxis < 0) { makeLeaf(); return; }
		std::vector<SahTri> left, right;
		for (const SahTri& t: tris) {
			const float lo = t.lo[bestAxis], hi = t.hi[bestAxis];
			const bool inLeft = lo < bestSplit || hi <= bestSplit;
			const bool inRight = hi > bestSplit || (lo >= bestSplit && !inLeft);
			if (inLeft) left.push_back(t);
			if (inRight) right.push_back(t);
		}
		if ((int) left.size() == n && (int) right.size() == n) { makeLeaf(); return; } // nothing separated
		const size_t child = kd.size();
		kd.push_back(DKdNode<R>());
		kd.push_back(DKdNode<R>());
		DKdNode<R> inner;
		inner.axis = bestAxis;
		inner.a = (int) child; // absolute index of children[0]; children[1] follows
		inner.b = 0;
		inner.split = (R) bestSplit;
		kd[nodeIdx] = inner;
		float lhi[3] = { bhi[0], bhi[1], bhi[2] }, rlo[3] = { blo[0], blo[1], blo[2] };
		lhi[bestAxis] = bestSplit;
		rlo[bestAxis] = bestSplit;
		buildSahNode(left, blo, lhi, depth + 1, child, firstNode, firstRef, kd, refs);
		buildSahNode(right, rlo, bhi, depth + 1, child + 1, firstNode, firstRef, kd, refs);
	}

	// replaces the KD tree of mesh m (fast precision): nodes appended to `kd`, references (relative triangle indices) to `refs`
	static int buildSahTree(const FrayGpuScene& s, const FrayGpuMesh& m, std::vector<DKdNode<R>>& kd, std::vector<int>& refs, int& maxDepthOut)
	{
		std::vector<SahTri> tris;
		tris.reserve(m.num_triangles);
		for (int t = 0; t < m.num_triangles; t++) {
			const size_t ti = (size_t) m.first_triangle + t;
			const double nn = s.tri_abxac[3 * ti] * s.tri_abxac[3 * ti] + s.tri_abxac[3 * ti + 1] * s.tri_abxac[3 * ti + 1] + s.tri_abxac[3 * ti + 2] * s.tri_abxac[3 * ti + 2];
			if (!(nn > 0)) continue; // degenerate: can never be hit (its record is all zeros)
			SahTri st;
			st.tri = t;
			for (int k = 0; k < 3; k++) { st.lo[k] = FLT_MAX; st.hi[k] = -FLT_MAX; }
			for (int c = 0; c < 3; c++) {
				const double* v = s.vertices + 3 * ((size_t) m.first_vertex + s.tri_v[3 * ti + c]);
				for (int k = 0; k < 3; k++) {
					// outward rounding: the box must contain the triangle as the FP32 records see it
					const float f = (float) v[k];
					st.lo[k] = std::min(st.lo[k], std::nextafter(f, -FLT_MAX));
					st.hi[k] = std::max(st.hi[k], std::nextafter(f, FLT_MAX));
				}
			}
			tris.push_back(st);
		}
		float blo[3], bhi[3];
		for (int k = 0; k < 3; k++) {
			blo[k] = std::nextafter((float) m.bbox_min[k], -FLT_MAX);
			bhi[k] = std::nextafter((float) m.bbox_max[k], FLT_MAX);
		}
		const size_t root = kd.size();
		kd.push_back(DKdNode<R>());
		buildSahNode(tris, blo, bhi, 0, root, root, refs.size(), kd, refs);
		maxDepthOut = 48;
		return (int) root;
	}

	bool build(const FrayGpuScene& s, std::string& err)
	{
		blob.clear();
		memset(&offsets, 0, sizeof(offsets));
		features = 0;
		if (s.abi_version != FRAY_GPU_ABI_VERSION) { err = "FrayGpuScene.abi_version mismatch"; return false; }
		if (s.num_nodes < 0 || s.num_lights < 0 || s.settings.frame_width <= 0 || s.settings.frame_height <= 0) { err = "malformed scene header"; return false; }
		if (s.settings.frame_width > 65535 || s.settings.frame_height > 65535) { err = "frame larger than 65535 pixels a side"; return false; } // pixel coordinates travel as 16-bit pairs
		auto inRange = [](long long i, long long n) { return i >= 0 && i < n; };

		DScene<R>& d = offsets;
		// camera + settings
		cvt3(d.cam.pos, s.camera.pos);
		cvt3(d.cam.topLeft, s.camera.top_left);
		cvt3(d.cam.topRight, s.camera.top_right);
		cvt3(d.cam.bottomLeft, s.camera.bottom_left);
		cvt3(d.cam.front, s.camera.front);
		cvt3(d.cam.up, s.camera.up);
		cvt3(d.cam.right, s.camera.right);
		for (int i = 0; i < 3; i++) {
			d.cam.du[i] = (R) ((s.camera.top_right[i] - s.camera.top_left[i]) / s.camera.w);
			d.cam.dv[i] = (R) ((s.camera.bottom_left[i] - s.camera.top_left[i]) / s.camera.h);
		}
		d.cam.w = (R) s.camera.w;
		d.cam.h = (R) s.camera.h;
		d.cam.aperture = (R) s.camera.aperture_size;
		d.cam.focalDist = (R) s.camera.focal_plane_dist;
		d.cam.stereoSep = (R) s.camera.stereo_separation;
		for (int i = 0; i < 3; i++) {
			d.cam.leftMask[i] = s.camera.left_mask[i];
			d.cam.rightMask[i] = s.camera.right_mask[i];
			d.ambient[i] = s.settings.ambient[i];
		}
		d.cam.dof = s.camera.dof;
		d.maxTraceDepth = s.settings.max_trace_depth;
		d.gi = s.settings.gi;
		d.saturation = s.settings.saturation;
		d.numNodes = s.num_nodes;
		d.numLights = s.num_lights;
		d.hasEnv = s.has_environment;
		for (int i = 0; i < 6; i++) {
			d.env[i] = s.env_bitmaps[i];
			if (d.env[i] >= s.num_bitmaps) { err = "environment bitmap index out of range"; return false; }
		}

		std::vector<DGeom<R>> geoms(s.num_geometries);
		for (int i = 0; i < s.num_geometries; i++) {
			const FrayGpuGeometry& g = s.geometries[i];
			DGeom<R>& o = geoms[i];
			o.type = g.type; o.mesh = g.mesh; o.left = g.left; o.right = g.right;
			for (int k = 0; k < 4; k++) o.p[k] = (R) g.p[k];
			if (g.type == FRAY_GEOM_MESH && !inRange(g.mesh, s.num_meshes)) { err = "geometry.mesh out of range"; return false; }
			if (g.type >= FRAY_GEOM_CSG_PLUS && g.type <= FRAY_GEOM_CSG_MINUS) {
				if (!inRange(g.left, s.num_geometries) || !inRange(g.right, s.num_geometries)) { err = "CSG operand out of range"; return false; }
				features |= FRAY_F_CSG;
			} else if (g.type < 0 || g.type > FRAY_GEOM_MESH) { err = "unknown geometry type"; return false; }
		}

		std::vector<DMesh<R>> meshes(s.num_meshes);
		std::vector<DKdNode<R>> kd((size_t) s.num_kd_nodes);
		std::vector<R> kdBox((size_t) s.num_kd_nodes * 6);
		std::vector<int> leafRefs(s.leaf_refs, s.leaf_refs + s.num_leaf_refs);
		const size_t T = (size_t) s.num_triangles;
		std::vector<R> triA(3 * T), triAB(3 * T), triAC(3 * T), triN(3 * T), triG(3 * T), triDx(3 * T), triDy(3 * T);
		std::vector<int> triNi(s.tri_n, s.tri_n + 3 * T), triTi(s.tri_t, s.tri_t + 3 * T);
		for (size_t i = 0; i < 3 * T; i++) {
			triAB[i] = (R) s.tri_ab[i]; triAC[i] = (R) s.tri_ac[i]; triN[i] = (R) s.tri_abxac[i]; triG[i] = (R) s.tri_gnormal[i];
			triDx[i] = (R) s.tri_dndx[i]; triDy[i] = (R) s.tri_dndy[i];
		}
		for (int mi = 0; mi < s.num_meshes; mi++) {
			const FrayGpuMesh& m = s.meshes[mi];
			DMesh<R>& o = meshes[mi];
			if (m.first_triangle < 0 || (long long) m.first_triangle + m.num_triangles > s.num_triangles ||
			    m.first_vertex < 0 || (long long) m.first_vertex + m.num_vertices > s.num_vertices ||
			    m.first_kd_node < 0 || (long long) m.first_kd_node + m.num_kd_nodes > s.num_kd_nodes ||
			    m.first_leaf_ref < 0 || (long long) m.first_leaf_ref + m.num_leaf_refs > s.num_leaf_refs) { err = "mesh ranges out of bounds"; return false; }
			o.flags = m.flags;
			o.firstTri = m.first_triangle; o.numTris = m.num_triangles;
			o.firstNormal = m.first_normal; o.firstUV = m.first_uv;
			o.kdRoot = m.kd_root >= 0 ? m.first_kd_node + m.kd_root : -1;
			o.firstLeafRef = m.first_leaf_ref;
			o.pad = 0;
			cvt3(o.bmin, m.bbox_min);
			cvt3(o.bmax, m.bbox_max);
			for (int t = 0; t < m.num_triangles; t++) {
				const size_t ti = (size_t) m.first_triangle + t;
				for (int k = 0; k < 3; k++) {
					if (!inRange(s.tri_v[3 * ti + k], m.num_vertices)) { err = "triangle vertex index out of range"; return false; }
					if ((m.flags & FRAY_MESH_HAS_NORMALS) && !inRange(s.tri_n[3 * ti + k], m.num_normals)) { err = "triangle normal index out of range"; return false; }
					if ((m.flags & FRAY_MESH_HAS_UVS) && !inRange(s.tri_t[3 * ti + k], m.num_uvs)) { err = "triangle uv index out of range"; return false; }
				}
				cvt3(&triA[3 * ti], s.vertices + 3 * ((size_t) m.first_vertex + s.tri_v[3 * ti]));
			}
			// kd nodes: absolute indices + per-node boxes (in double, then rounded: the same values the reference's split() yields)
			if (m.kd_root >= 0) {
				struct Item { int node; int depth; double box[6]; };
				std::vector<Item> todo;
				Item root;
				root.node = m.kd_root;
				root.depth = 0;
				for (int k = 0; k < 3; k++) { root.box[k] = m.bbox_min[k]; root.box[3 + k] = m.bbox_max[k]; }
				todo.push_back(root);
				size_t visited = 0;
				while (!todo.empty()) {
					Item it = todo.back();
					todo.pop_back();
					if (!inRange(it.node, m.num_kd_nodes) || ++visited > (size_t) m.num_kd_nodes) { err = "malformed KD tree"; return false; }
					// the device walks keep one pending sibling per level on a fixed stack (core.cuh, FRAY_KD_STACK); the reference's
					// builder stops at MAX_DEPTH 64 (src/constants.h:39), a deeper tree handed in through the C ABI is refused here
					if (it.depth >= FRAY_KD_STACK - 2) { err = "KD tree deeper than the device traversal stack (FRAY_KD_STACK - 2 levels)"; return false; }
					const FrayGpuKdNode& n = s.kd_nodes[(size_t) m.first_kd_node + it.node];
					DKdNode<R>& dn = kd[(size_t) m.first_kd_node + it.node];
					for (int k = 0; k < 6; k++) kdBox[6 * ((size_t) m.first_kd_node + it.node) + k] = (R) it.box[k];
					dn.axis = n.axis;
					dn.split = (R) n.split;
					if (n.axis == 3) {
						if (n.a < 0 || n.b < 0 || n.a + n.b > m.num_leaf_refs) { err = "KD leaf range out of bounds"; return false; }
						dn.a = m.first_leaf_ref + n.a;
						dn.b = n.b;
					} else {
						if (n.axis < 0 || n.axis > 2 || !inRange(n.a, m.num_kd_nodes - 1)) { err = "malformed KD inner node"; return false; }
						dn.a = m.first_kd_node + n.a;
						dn.b = 0;
						Item l = it, r = it;
						l.node = n.a; r.node = n.a + 1;
						l.depth = r.depth = it.depth + 1;
						l.box[3 + n.axis] = n.split;
						r.box[n.axis] = n.split;
						todo.push_back(l);
						todo.push_back(r);
					}
				}
				for (int i = 0; i < m.num_leaf_refs; i++)
					if (!inRange(s.leaf_refs[(size_t) m.first_leaf_ref + i], m.num_triangles)) { err = "KD leaf reference out of range"; return false; }
			}
		}

		if (!Num<R>::kExact && getenv("FRAY_GPU_SAH_KD")) {
			// fast precision: our own trees over the same triangles (see buildSahTree); leaf references become absolute-in-mesh
			// offsets exactly as in the reference layout: leaf.a indexes leafRefs, entries are relative to mesh.firstTri
			kd.clear();
			leafRefs.clear();
			kdBox.clear();
			for (int mi = 0; mi < s.num_meshes; mi++) {
				const FrayGpuMesh& m = s.meshes[mi];
				if (m.kd_root < 0) continue;
				int depth = 0;
				meshes[mi].kdRoot = buildSahTree(s, m, kd, leafRefs, depth);
				meshes[mi].firstLeafRef = 0;
			}
			if (getenv("FRAY_GPU_VERBOSE")) fprintf(stderr, "fray_gpu: fast-precision KD trees: %zu nodes, %zu leaf references (reference trees: %lld nodes, %lld references)\n",
			                                        kd.size(), leafRefs.size(), (long long) s.num_kd_nodes, (long long) s.num_leaf_refs);
		}

		std::vector<DShader<R>> shaders(s.num_shaders);
		for (int i = 0; i < s.num_shaders; i++) {
			const FrayGpuShader& sh = s.shaders[i];
			DShader<R>& o = shaders[i];
			o.type = sh.type; o.texture = sh.texture; o.firstLayer = sh.first_layer; o.numLayers = sh.num_layers;
			o.numSamples = sh.num_samples; o.pureReflection = sh.pure_reflection;
			for (int k = 0; k < 3; k++) { o.color[k] = sh.color[k]; o.specularColor[k] = sh.specular_color[k]; o.mult[k] = sh.mult[k]; }
			o.exponent = (R) sh.exponent; o.specularMultiplier = (R) sh.specular_multiplier;
			o.deflectionScaling = (R) sh.deflection_scaling; o.ior = (R) sh.ior;
			if (sh.texture >= s.num_textures) { err = "shader texture out of range"; return false; }
			if (sh.type == FRAY_SHADER_LAYERED) {
				if (sh.first_layer < 0 || sh.num_layers < 0 || sh.first_layer + sh.num_layers > s.num_layers) { err = "layer range out of bounds"; return false; }
				for (int k = 0; k < sh.num_layers; k++) {
					const FrayGpuLayer& L = s.layers[sh.first_layer + k];
					if (!inRange(L.shader, s.num_shaders) || L.texture >= s.num_textures) { err = "layer references out of range"; return false; }
				}
			}
			if (sh.type == FRAY_SHADER_REFL && !sh.pure_reflection && sh.num_samples < 1) { err = "glossy numSamples must be >= 1"; return false; }
		}
		for (int i = 0; i < s.num_shaders; i++)
			if (layeredDepth(s, i, 0) > 2) { err = "Layered shaders nested deeper than 2 are not supported"; return false; }
		std::vector<DLayer> layers(s.num_layers);
		for (int i = 0; i < s.num_layers; i++) {
			layers[i].shader = s.layers[i].shader;
			layers[i].texture = s.layers[i].texture;
			for (int k = 0; k < 3; k++) layers[i].opacity[k] = s.layers[i].opacity[k];
		}
		std::vector<DTexture<R>> textures(s.num_textures);
		for (int i = 0; i < s.num_textures; i++) {
			const FrayGpuTexture& t = s.textures[i];
			DTexture<R>& o = textures[i];
			o.type = t.type; o.bitmap = t.bitmap;
			for (int k = 0; k < 3; k++) { o.color1[k] = t.color1[k]; o.color2[k] = t.color2[k]; }
			o.scaling = (R) t.scaling; o.ior = (R) t.ior; o.bumpIntensity = (R) t.bump_intensity;
			if ((t.type == FRAY_TEX_BITMAP || t.type == FRAY_TEX_BUMP) && !inRange(t.bitmap, s.num_bitmaps)) { err = "texture bitmap out of range"; return false; }
		}
		std::vector<DBitmap> bitmaps(s.num_bitmaps);
		for (int i = 0; i < s.num_bitmaps; i++) {
			const FrayGpuBitmap& b = s.bitmaps[i];
			bitmaps[i].width = b.width; bitmaps[i].height = b.height; bitmaps[i].firstTexel = b.first_texel;
			if (b.width <= 0 || b.height <= 0 || b.first_texel < 0 || b.first_texel + (long long) b.width * b.height > s.num_texels) { err = "bitmap texel range out of bounds"; return false; }
		}
		std::vector<DLight<R>> lights(s.num_lights);
		for (int i = 0; i < s.num_lights; i++) {
			const FrayGpuLight& l = s.lights[i];
			DLight<R>& o = lights[i];
			o.type = l.type; o.xSubd = l.x_subd; o.ySubd = l.y_subd;
			for (int k = 0; k < 3; k++) o.color[k] = l.color[k];
			o.power = l.power;
			cvt3(o.pos, l.pos);
			cvtXform(o.T, l.T);
			cvt3(o.center, l.center);
			o.area = (R) l.area;
			o.areaF = (float) l.area;
			if (l.type == FRAY_LIGHT_RECT && (l.x_subd < 1 || l.y_subd < 1)) { err = "RectLight subdivisions must be >= 1"; return false; }
		}
		d.lightSamples = d.lightDraws = 0;
		for (int i = 0; i < s.num_lights; i++) {
			const int ns = s.lights[i].type == FRAY_LIGHT_RECT ? s.lights[i].x_subd * s.lights[i].y_subd : 1;
			d.lightSamples += ns;
			if (s.lights[i].type == FRAY_LIGHT_RECT) d.lightDraws += 2 * ns;
		}
		// compact sampling records of the lights for the fast path tracer (core.cuh, LightRec)
		std::vector<float4> lightRecs((size_t) FRAY_LIGHT_REC_VEC * s.num_lights);
		for (int i = 0; i < s.num_lights; i++) {
			const FrayGpuLight& l = s.lights[i];
			float4* r = lightRecs.data() + (size_t) FRAY_LIGHT_REC_VEC * i;
			const int nx = std::max(1, l.x_subd), ny = std::max(1, l.y_subd);
			const int4 hd{ l.type, nx, ny, l.type == FRAY_LIGHT_RECT ? nx * ny : 1 };
			memcpy(&r[0], &hd, sizeof(float4));
			r[1] = float4{ (float) l.center[0], (float) l.center[1], (float) l.center[2], (float) l.area };
			// sample (column + r1, row + r2) of the unit square: T(((column + r1) / nx - 0.5, 0, (row + r2) / ny - 0.5)), src/lights.cpp:49-77
			const D3 corner = rowMul(D3{ -0.5, 0, -0.5 }, l.T.m);
			const D3 U = rowMul(D3{ 1.0 / nx, 0, 0 }, l.T.m), V = rowMul(D3{ 0, 0, 1.0 / ny }, l.T.m);
			if (l.type == FRAY_LIGHT_RECT) {
				r[2] = float4{ (float) (corner.x + l.T.offset[0]), (float) (corner.y + l.T.offset[1]), (float) (corner.z + l.T.offset[2]), 1.0f / (float) nx };
				r[3] = float4{ (float) U.x, (float) U.y, (float) U.z, 0.0f };
				r[4] = float4{ (float) V.x, (float) V.y, (float) V.z, 0.0f };
			} else {
				r[2] = float4{ (float) l.pos[0], (float) l.pos[1], (float) l.pos[2], 1.0f };
				r[3] = r[4] = float4{ 0, 0, 0, 0 };
			}
			r[5] = float4{ l.color[0] * l.power, l.color[1] * l.power, l.color[2] * l.power, 0.0f };
		}
		std::vector<DNode<R>> nodes(s.num_nodes);
		for (int i = 0; i < s.num_nodes; i++) {
			const FrayGpuNode& n = s.nodes[i];
			DNode<R>& o = nodes[i];
			if (!inRange(n.geometry, s.num_geometries) || !inRange(n.shader, s.num_shaders) || n.bump >= s.num_textures) { err = "node references out of range"; return false; }
			cvtXform(o.T, n.T);
			o.geom = n.geometry; o.shader = n.shader; o.bump = n.bump;
			o.needsUV = shaderReadsUV(s, n.shader, 0) || textureReadsUV(s, n.bump);
			o.inFlat = 0;
			o.pad[0] = o.pad[1] = o.pad[2] = 0;
		}
		// world-space boxes of the nodes (fast precision, wave.cuh): the eight corners of the object-space bounds through the node transform
		std::vector<float4> nodeBox;
		if (!Num<R>::kExact) {
			nodeBox.resize(2 * (size_t) s.num_nodes);
			for (int i = 0; i < s.num_nodes; i++) {
				const FrayGpuNode& n = s.nodes[i];
				const FrayGpuGeometry& g = s.geometries[n.geometry];
				double lo[3] = { -1e30, -1e30, -1e30 }, hi[3] = { 1e30, 1e30, 1e30 };
				bool bounded = true;
				if (g.type == FRAY_GEOM_MESH) {
					for (int k = 0; k < 3; k++) { lo[k] = s.meshes[g.mesh].bbox_min[k]; hi[k] = s.meshes[g.mesh].bbox_max[k]; }
				} else if (g.type == FRAY_GEOM_SPHERE || g.type == FRAY_GEOM_CUBE) {
					for (int k = 0; k < 3; k++) { lo[k] = g.p[k] - g.p[3]; hi[k] = g.p[k] + g.p[3]; }
				} else if (g.type == FRAY_GEOM_PLANE && g.p[1] < 1e15) {
					lo[0] = lo[2] = -g.p[1]; hi[0] = hi[2] = g.p[1]; lo[1] = hi[1] = g.p[0];
				} else {
					bounded = false; // CSG, unbounded planes: never culled
				}
				double wlo[3] = { 1e30, 1e30, 1e30 }, whi[3] = { -1e30, -1e30, -1e30 };
				if (bounded) {
					for (int c = 0; c < 8; c++) {
						const D3 p{ (c & 1) ? hi[0] : lo[0], (c & 2) ? hi[1] : lo[1], (c & 4) ? hi[2] : lo[2] };
						const D3 w = rowMul(p, n.T.m);
						const double q[3] = { w.x + n.T.offset[0], w.y + n.T.offset[1], w.z + n.T.offset[2] };
						for (int k = 0; k < 3; k++) { wlo[k] = std::min(wlo[k], q[k]); whi[k] = std::max(whi[k], q[k]); }
					}
					double ext = 0;
					for (int k = 0; k < 3; k++) ext = std::max(ext, std::max(std::fabs(wlo[k]), std::fabs(whi[k])));
					const double pad = 1e-4 * ext + 1e-4;
					for (int k = 0; k < 3; k++) { wlo[k] -= pad; whi[k] += pad; }
				} else {
					for (int k = 0; k < 3; k++) { wlo[k] = -3e38; whi[k] = 3e38; }
				}
				nodeBox[2 * i] = float4{ (float) std::max(wlo[0], -3e38), (float) std::max(wlo[1], -3e38), (float) std::max(wlo[2], -3e38), 0.0f };
				nodeBox[2 * i + 1] = float4{ (float) std::min(whi[0], 3e38), (float) std::min(whi[1], 3e38), (float) std::min(whi[2], 3e38), 0.0f };
			}
		}
		// feature bits this scene needs from the kernels
		// textures count only if something that is rendered refers to one (smallpt.fray defines a Fresnel texture and a Layered
		// shader that no node uses)
		if (s.has_environment) features |= FRAY_F_TEX;
		for (int i = 0; i < s.num_nodes; i++)
			if (shaderUsesTexture(s, s.nodes[i].shader, 0)) features |= FRAY_F_TEX;
		for (int i = 0; i < s.num_nodes; i++)
			if (s.nodes[i].bump >= 0) features |= FRAY_F_TEX;
		if (s.camera.dof || s.camera.stereo_separation > 0) features |= FRAY_F_LENS;
		if (!Num<R>::kExact) {
			// first with front and back copies of two-sided polygons; if the scene then fits the untextured variants, which have the
			// loops for the two-sided list, once more with that list
			auto flatFeatures = [&]() {
				int f = buildFlat(s, nodes);
				if (!offsets.lightsInFlat) f |= FRAY_F_NODES; // the generic light loop is compiled with the generic node loop
				for (int i = 0; i < s.num_nodes; i++)
					if (!nodes[i].inFlat) f |= FRAY_F_NODES;
				return f;
			};
			twoSidedList = false;
			int f = flatFeatures();
			const int untextured = Variants<float>::kLean | FRAY_F_SPHERES | FRAY_F_NODES;
			if (((features | f) & ~untextured) == 0 && !getenv("FRAY_GPU_NO_TWOSIDED_LIST")) {
				for (int i = 0; i < s.num_nodes; i++) nodes[i].inFlat = 0;
				twoSidedList = true;
				f = flatFeatures();
			}
			features |= f;
			// the node loop of the wavefront skips from one node outside the flat table to the next: .w of a node's upper box corner
			// holds the index of the next such node (num_nodes: none)
			int next = s.num_nodes;
			for (int i = s.num_nodes - 1; i >= 0; i--) {
				memcpy(&nodeBox[2 * i + 1].w, &next, sizeof(int));
				if (!nodes[i].inFlat) next = i;
			}
		}
		for (int i = 0; i < s.num_nodes; i++)
			if (!nodes[i].inFlat) features |= FRAY_F_NODES;
		if (Num<R>::kExact) features |= FRAY_F_GENERIC;
		std::vector<R> normals(3 * (size_t) s.num_normals), uvs(3 * (size_t) s.num_uvs);
		for (size_t i = 0; i < normals.size(); i++) normals[i] = (R) s.normals[i];
		for (size_t i = 0; i < uvs.size(); i++) uvs[i] = (R) s.uvs[i];
		std::vector<float> texels(s.texels, s.texels + 3 * (size_t) s.num_texels);
		// fast precision: plane + barycentric-plane records for the KD leaves (core.cuh, intersectMeshFast)
		std::vector<float4> kdTris;
		if (!Num<R>::kExact) {
			kdTris.resize(3 * T);
			for (int mi = 0; mi < s.num_meshes; mi++) {
				const FrayGpuMesh& m = s.meshes[mi];
				for (int t = 0; t < m.num_triangles; t++) {
					const size_t ti = (size_t) m.first_triangle + t;
					const double* a = s.vertices + 3 * ((size_t) m.first_vertex + s.tri_v[3 * ti]);
					const D3 A{ a[0], a[1], a[2] };
					const D3 AB{ s.tri_ab[3 * ti], s.tri_ab[3 * ti + 1], s.tri_ab[3 * ti + 2] }, AC{ s.tri_ac[3 * ti], s.tri_ac[3 * ti + 1], s.tri_ac[3 * ti + 2] };
					const D3 N{ s.tri_abxac[3 * ti], s.tri_abxac[3 * ti + 1], s.tri_abxac[3 * ti + 2] };
					const double nn = dt(N, N);
					float4 zero = plane4(D3{ 0, 0, 0 }, 0);
					if (!(nn > 0)) { kdTris[3 * ti] = kdTris[3 * ti + 1] = kdTris[3 * ti + 2] = zero; continue; }
					const D3 Nu = scl(N, 1 / sqrt(nn));
					const D3 m2 = scl(crs(AC, N), 1 / nn), m3 = scl(crs(N, AB), 1 / nn);
					kdTris[3 * ti] = plane4(Nu, dt(Nu, A));
					kdTris[3 * ti + 1] = plane4(m2, -dt(m2, A));
					kdTris[3 * ti + 2] = plane4(m3, -dt(m3, A));
				}
			}
		}

		std::vector<float4> kdLeafTris, kdLeafBox;
		if (!Num<R>::kExact) {
			kdLeafTris.resize(3 * leafRefs.size());
			for (int mi = 0; mi < s.num_meshes; mi++) {
				if (meshes[mi].kdRoot < 0) continue;
				const FrayGpuMesh& m = s.meshes[mi];
				// the references of mesh mi: every leaf of its tree (walk the nodes: a leaf owns [a, a + b))
				std::vector<int> todo(1, meshes[mi].kdRoot);
				while (!todo.empty()) {
					const int ni = todo.back();
					const DKdNode<R> n = kd[ni];
					todo.pop_back();
					if (n.axis == 3) {
						double lo[3] = { 1e30, 1e30, 1e30 }, hi[3] = { -1e30, -1e30, -1e30 };
						for (int i = 0; i < n.b; i++) {
							const size_t r = (size_t) n.a + i, t = (size_t) meshes[mi].firstTri + leafRefs[r];
							for (int k = 0; k < 3; k++) kdLeafTris[3 * r + k] = kdTris[3 * t + k];
							const size_t ti = (size_t) m.first_triangle + leafRefs[r];
							for (int c = 0; c < 3; c++) {
								const double* v = s.vertices + 3 * ((size_t) m.first_vertex + s.tri_v[3 * ti + c]);
								for (int k = 0; k < 3; k++) { lo[k] = std::min(lo[k], v[k]); hi[k] = std::max(hi[k], v[k]); }
							}
						}
						// the leaf's box (kdWalk tests it before the triangles), padded far beyond what FP32 can move a hit point
						double ext = 0;
						for (int k = 0; k < 3; k++) ext = std::max(ext, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
						const double pad = n.b > 0 ? 1e-4 * ext + 1e-6 : 0.0;
						const int box = (int) (kdLeafBox.size() / 2);
						kdLeafBox.push_back(float4{ (float) (lo[0] - pad), (float) (lo[1] - pad), (float) (lo[2] - pad), 0.0f });
						kdLeafBox.push_back(float4{ (float) (hi[0] + pad), (float) (hi[1] + pad), (float) (hi[2] + pad), 0.0f });
						memcpy(&kd[ni].split, &box, sizeof(int)); // (R is float here)
					} else {
						todo.push_back(n.a);
						todo.push_back(n.a + 1);
					}
				}
			}
		}

#define FRAY_PUT(member, vec) d.member = reinterpret_cast<decltype(d.member)>(append(vec))
		FRAY_PUT(nodes, nodes);
		FRAY_PUT(geoms, geoms);
		FRAY_PUT(meshes, meshes);
		FRAY_PUT(shaders, shaders);
		FRAY_PUT(layers, layers);
		FRAY_PUT(textures, textures);
		FRAY_PUT(bitmaps, bitmaps);
		FRAY_PUT(lights, lights);
		FRAY_PUT(triA, triA);
		FRAY_PUT(triAB, triAB);
		FRAY_PUT(triAC, triAC);
		FRAY_PUT(triN, triN);
		FRAY_PUT(triG, triG);
		FRAY_PUT(triDndx, triDx);
		FRAY_PUT(triDndy, triDy);
		FRAY_PUT(triNi, triNi);
		FRAY_PUT(triTi, triTi);
		FRAY_PUT(normals, normals);
		FRAY_PUT(uvs, uvs);
		FRAY_PUT(kd, kd);
		FRAY_PUT(kdBox, kdBox);
		FRAY_PUT(leafRefs, leafRefs);
		FRAY_PUT(texels, texels);
		FRAY_PUT(kdTris, kdTris);
		FRAY_PUT(kdLeafTris, kdLeafTris);
		FRAY_PUT(kdLeafBox, kdLeafBox);
		FRAY_PUT(lightRecs, lightRecs);
		FRAY_PUT(nodeBox, nodeBox);
		FRAY_PUT(flatPolys, flatPolys);
		FRAY_PUT(flatInfo, flatInfo);
#undef FRAY_PUT
		return true;
	}

	// the scene with pointers rebased onto `base` (host: blob.data(); device: the cudaMalloc'ed copy)
	DScene<R> bind(const void* base) const
	{
		DScene<R> d = offsets;
		const char* b = (const char*) base;
#define FRAY_REBASE(member) d.member = reinterpret_cast<decltype(d.member)>(b + (size_t) offsets.member)
		FRAY_REBASE(nodes); FRAY_REBASE(geoms); FRAY_REBASE(meshes); FRAY_REBASE(shaders); FRAY_REBASE(layers);
		FRAY_REBASE(textures); FRAY_REBASE(bitmaps); FRAY_REBASE(lights);
		FRAY_REBASE(triA); FRAY_REBASE(triAB); FRAY_REBASE(triAC); FRAY_REBASE(triN); FRAY_REBASE(triG);
		FRAY_REBASE(triDndx); FRAY_REBASE(triDndy); FRAY_REBASE(triNi); FRAY_REBASE(triTi);
		FRAY_REBASE(normals); FRAY_REBASE(uvs); FRAY_REBASE(kd); FRAY_REBASE(kdBox); FRAY_REBASE(leafRefs); FRAY_REBASE(texels);
		FRAY_REBASE(flatPolys); FRAY_REBASE(flatInfo); FRAY_REBASE(kdTris); FRAY_REBASE(kdLeafTris); FRAY_REBASE(kdLeafBox); FRAY_REBASE(lightRecs); FRAY_REBASE(nodeBox);
#undef FRAY_REBASE
		return d;
	}
};

// samples per pixel by the reference's rule, src/main.cpp:395-400
inline int samplesPerPixel(const FrayGpuScene& s)
{
	int spp = s.settings.want_aa ? 5 : 1;
	if (s.camera.dof) spp = std::max(spp, s.camera.num_dof_samples);
	if (s.settings.gi) spp = std::max(spp, s.settings.num_paths);
	return spp;
}

} // namespace fray
