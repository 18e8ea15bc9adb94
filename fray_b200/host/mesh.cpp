// mesh.cpp -- OBJ loading, per-triangle precompute and the parity-exact KD-tree builder (host side).
//
// The GPU only TRAVERSES the tree; which triangle wins a tie on a shared edge depends on the leaf contents
// and order, so the tree must be the reference's tree (SURVEY.md section 7 "KD tree shape"). This file
// therefore reproduces, in FP64 and in the same evaluation order:
//   OBJ reader            /root/reference/src/mesh.cpp:167-258  (1-based indices kept raw, dummy slot 0, fan triangulation)
//   prepareTriangles      src/mesh.cpp:260-313                  (AB, AC, AB^AC, unit gnormal, UV tangent frame)
//   bounding box          src/mesh.cpp:74-79                    (over ALL vertices, including the dummy (0,0,0))
//   buildKD               src/mesh.cpp:315-355                  (axis = depth % 3, midpoint split, <= 20 triangles or depth > 64 = leaf)
//   BBox::intersectTriangle / testIntersect / inside            src/bbox.h:79-134, 164-199
//   Triangle::intersect   src/triangle.cpp:32-64                (used by the box-edge-pierces-triangle case)
// The recursive pointer tree of the reference (src/mesh.h:35-53) is emitted directly in the flattened layout of
// include/fray_gpu.h (children adjacent, leaf triangle lists concatenated).
#include <cstdio>
#include <cstring>
#include <chrono>

#include "scene.h"

namespace fray {

static const int kMaxTrianglesPerLeaf = 20; // MAX_TRIANGLES_PER_LEAF, src/constants.h:38
static const int kMaxDepth = 64;            // MAX_DEPTH, src/constants.h:39

// ---- BBox ------------------------------------------------------------------------------------------

void BBox::add(const Vec3& p)
{
	vmin.x = std::min(vmin.x, p.x); vmax.x = std::max(vmax.x, p.x);
	vmin.y = std::min(vmin.y, p.y); vmax.y = std::max(vmax.y, p.y);
	vmin.z = std::min(vmin.z, p.z); vmax.z = std::max(vmax.z, p.z);
}

bool BBox::inside(const Vec3& v) const
{
	return vmin.x - 1e-6 <= v.x && v.x <= vmax.x + 1e-6 &&
	       vmin.y - 1e-6 <= v.y && v.y <= vmax.y + 1e-6 &&
	       vmin.z - 1e-6 <= v.z && v.z <= vmax.z + 1e-6;
}

// rdir as prepared by RRay::prepareForTracing (src/bbox.h:49-54)
static Vec3 reciprocalDir(const Vec3& d)
{
	return Vec3(std::fabs(d.x) > 1e-12 ? 1.0 / d.x : 1e12,
	            std::fabs(d.y) > 1e-12 ? 1.0 / d.y : 1e12,
	            std::fabs(d.z) > 1e-12 ? 1.0 / d.z : 1e12);
}

bool BBox::testIntersect(const Vec3& start, const Vec3& dir, const Vec3& rdir) const
{
	if (inside(start)) return true;
	for (int dim = 0; dim < 3; dim++) {
		if ((dir[dim] < 0 && start[dim] < vmin[dim]) || (dir[dim] > 0 && start[dim] > vmax[dim])) return false;
		if (std::fabs(dir[dim]) < 1e-9) continue;
		const double mul = rdir[dim];
		const int u = (dim == 0) ? 1 : 0;
		const int v = (dim == 2) ? 1 : 2;
		// near slab first; a slab plane behind the origin ends the work for this axis (src/bbox.h:97-113)
		const double planes[2] = { vmin[dim], vmax[dim] };
		for (double plane: planes) {
			double dist = (plane - start[dim]) * mul;
			if (dist < 0) break;
			double x = start[u] + dir[u] * dist;
			if (vmin[u] <= x && x <= vmax[u]) {
				double y = start[v] + dir[v] * dist;
				if (vmin[v] <= y && y <= vmax[v]) return true;
			}
		}
	}
	return false;
}

static inline double det3(const Vec3& a, const Vec3& b, const Vec3& c) { return dot(cross(a, b), c); }
static inline double signOf(double x) { return x > 0 ? +1 : -1; }

// general ray/triangle test with un-normalised direction, src/triangle.cpp:32-64
static bool triangleIntersect(const Vec3& start, const Vec3& dir, const Vec3& A, const Vec3& B, const Vec3& C, double maxGamma)
{
	Vec3 AB = B - A, AC = C - A, D = -dir;
	double Dcr = det3(AB, AC, D);
	if (std::fabs(Dcr) < 1e-12) return false;
	double rDcr = 1 / Dcr;
	Vec3 H = start - A;
	double gamma = det3(AB, AC, H) * rDcr;
	if (gamma < 0 || gamma > maxGamma) return false;
	double l2 = det3(H, AC, D) * rDcr;
	if (l2 < 0 || l2 > 1) return false;
	double l3 = det3(AB, H, D) * rDcr;
	if (l3 < 0 || l3 > 1) return false;
	return 1 - (l2 + l3) >= 0;
}

bool BBox::intersectTriangle(const Vec3& A, const Vec3& B, const Vec3& C) const
{
	if (inside(A) || inside(B) || inside(C)) return true;
	// an edge crosses the box (must be hit from both ends)
	const Vec3 t[3] = { A, B, C };
	for (int i = 0; i < 3; i++)
		for (int j = i + 1; j < 3; j++) {
			Vec3 d = t[j] - t[i];
			if (testIntersect(t[i], d, reciprocalDir(d))) {
				Vec3 e = t[i] - t[j];
				if (testIntersect(t[j], e, reciprocalDir(e))) return true;
			}
		}
	// a box edge pierces the triangle
	Vec3 AB = B - A, AC = C - A;
	Vec3 N = cross(AB, AC);
	double D = dot(A, N);
	for (int mask = 0; mask < 7; mask++)
		for (int j = 0; j < 3; j++) {
			if (mask & (1 << j)) continue;
			Vec3 s((mask & 1) ? vmax.x : vmin.x, (mask & 2) ? vmax.y : vmin.y, (mask & 4) ? vmax.z : vmin.z);
			Vec3 e = s;
			e[j] = vmax[j];
			if (signOf(dot(s, N) - D) != signOf(dot(e, N) - D)) {
				if (triangleIntersect(s, e - s, A, B, C, 1.0000001)) return true;
			}
		}
	return false;
}

void BBox::split(int axis, double where, BBox& left, BBox& right) const
{
	left = *this;
	right = *this;
	left.vmax[axis] = where;
	right.vmin[axis] = where;
}

// ---- OBJ -------------------------------------------------------------------------------------------

static std::vector<std::string> splitBlank(const char* s)
{
	std::vector<std::string> out;
	while (*s) {
		while (*s && isspace((unsigned char) *s)) s++;
		if (!*s) break;
		const char* e = s;
		while (*e && !isspace((unsigned char) *e)) e++;
		out.emplace_back(s, e);
		s = e;
	}
	return out;
}

static int toInt(const std::string& s)
{
	int x;
	return (!s.empty() && sscanf(s.c_str(), "%d", &x) == 1) ? x : 0;
}

static double toDouble(const std::string& s)
{
	double x;
	return (!s.empty() && sscanf(s.c_str(), "%lf", &x) == 1) ? x : 0;
}

// "v", "v/t", "v//n", "v/t/n"
static void parseCorner(const std::string& s, int& v, int& t, int& n)
{
	std::string part[3];
	int k = 0;
	for (char c: s) {
		if (c == '/') {
			if (++k > 2) break;
		} else {
			part[k] += c;
		}
	}
	v = toInt(part[0]);
	t = toInt(part[1]);
	n = toInt(part[2]);
}

bool Mesh::loadFromOBJ(const char* filename)
{
	FILE* f = fopen(filename, "rt");
	if (!f) return false;
	vertices.assign(1, Vec3());
	uvs.assign(1, Vec3());
	normals.assign(1, Vec3());
	triangles.clear();
	static thread_local char line[10000];
	while (fgets(line, sizeof(line), f)) {
		if (line[0] == '#') continue;
		std::vector<std::string> tok = splitBlank(line);
		if (tok.empty()) continue;
		auto num = [&](size_t i) { return i < tok.size() ? toDouble(tok[i]) : 0.0; };
		if (tok[0] == "v") vertices.push_back(Vec3(num(1), num(2), num(3)));
		else if (tok[0] == "vn") normals.push_back(Vec3(num(1), num(2), num(3)));
		else if (tok[0] == "vt") uvs.push_back(Vec3(num(1), num(2), 0));
		else if (tok[0] == "f") {
			for (int i = 0; i < (int) tok.size() - 3; i++) { // fan: (1, 2+i, 3+i)
				Triangle T;
				parseCorner(tok[1], T.v[0], T.t[0], T.n[0]);
				parseCorner(tok[2 + i], T.v[1], T.t[1], T.n[1]);
				parseCorner(tok[3 + i], T.v[2], T.t[2], T.n[2]);
				triangles.push_back(T);
			}
		}
	}
	fclose(f);
	if (normals.size() == 1) normals.clear();
	// indices outside the pools would read out of bounds in the reference; reject such files
	for (const Triangle& T: triangles)
		for (int k = 0; k < 3; k++) {
			if (T.v[k] < 0 || T.v[k] >= (int) vertices.size()) return false;
			if (!normals.empty() && (T.n[k] < 0 || T.n[k] >= (int) normals.size())) return false;
			if (T.t[k] < 0 || T.t[k] >= (int) uvs.size()) return false;
		}
	prepareTriangles();
	return true;
}

// x * A + y * B = C in the xy plane (Cramer), src/mesh.cpp:260-270
static void solve2D(const Vec3& A, const Vec3& B, const Vec3& C, double& x, double& y)
{
	double Dcr = A.x * B.y - A.y * B.x;
	x = (C.x * B.y - C.y * B.x) / Dcr;
	y = (A.x * C.y - A.y * C.x) / Dcr;
}

void Mesh::prepareTriangles()
{
	for (Triangle& t: triangles) {
		Vec3 A = vertices[t.v[0]], B = vertices[t.v[1]], C = vertices[t.v[2]];
		t.AB = B - A;
		t.AC = C - A;
		t.ABcrossAC = cross(t.AB, t.AC);
		t.gnormal = normalized(t.ABcrossAC);
		if (!uvs.empty() && !normals.empty()) {
			Vec3 tA = uvs[t.t[0]], tB = uvs[t.t[1]], tC = uvs[t.t[2]];
			Vec3 tAB = tB - tA, tAC = tC - tA;
			double px, qx, py, qy;
			solve2D(tAB, tAC, Vec3(1, 0, 0), px, qx);
			solve2D(tAB, tAC, Vec3(0, 1, 0), py, qy);
			t.dNdx = normalized(px * t.AB + qx * t.AC);
			t.dNdy = normalized(py * t.AB + qy * t.AC);
		} else {
			t.dNdx = Vec3();
			t.dNdy = Vec3();
		}
	}
	if (g_verbose) printf("Mesh loaded, %d triangles\n", (int) triangles.size());
}

// ---- KD build --------------------------------------------------------------------------------------

void Mesh::buildKD(int nodeIdx, const std::vector<int>& tris, const BBox& box, int depth)
{
	numNodes++;
	maxTreeDepth = std::max(maxTreeDepth, depth);
	nodeDepthSum += depth;
	if ((int) tris.size() <= kMaxTrianglesPerLeaf || depth > kMaxDepth) {
		FrayGpuKdNode& n = kdNodes[nodeIdx];
		n.axis = 3;
		n.a = (int32_t) leafRefs.size();
		n.b = (int32_t) tris.size();
		n.split = 0;
		leafRefs.insert(leafRefs.end(), tris.begin(), tris.end());
		return;
	}
	const int axis = depth % 3;
	const double split = (box.vmin[axis] + box.vmax[axis]) * 0.5; // findOptimalSplitPlane, src/mesh.cpp:315-318
	BBox lbox, rbox;
	box.split(axis, split, lbox, rbox);
	std::vector<int> ltris, rtris;
	for (int ti: tris) {
		const Triangle& T = triangles[ti];
		const Vec3& A = vertices[T.v[0]];
		const Vec3& B = vertices[T.v[1]];
		const Vec3& C = vertices[T.v[2]];
		if (lbox.intersectTriangle(A, B, C)) ltris.push_back(ti);
		if (rbox.intersectTriangle(A, B, C)) rtris.push_back(ti);
	}
	const int child = (int) kdNodes.size();
	kdNodes.push_back(FrayGpuKdNode{});
	kdNodes.push_back(FrayGpuKdNode{});
	{
		FrayGpuKdNode& n = kdNodes[nodeIdx];
		n.axis = axis;
		n.a = child;
		n.b = 0;
		n.split = split;
	}
	buildKD(child, ltris, lbox, depth + 1);
	buildKD(child + 1, rtris, rbox, depth + 1);
}

void Mesh::beginRender()
{
	bbox = BBox();
	for (const Vec3& v: vertices) bbox.add(v);
	kdNodes.clear();
	leafRefs.clear();
	numNodes = maxTreeDepth = 0;
	nodeDepthSum = 0;
	if (useKD && triangles.size() > (size_t) kMaxTrianglesPerLeaf) {
		auto t0 = std::chrono::steady_clock::now();
		std::vector<int> all(triangles.size());
		for (size_t i = 0; i < all.size(); i++) all[i] = (int) i;
		kdNodes.push_back(FrayGpuKdNode{});
		buildKD(0, all, bbox, 0);
		unsigned ms = (unsigned) std::chrono::duration_cast<std::chrono::milliseconds>(std::chrono::steady_clock::now() - t0).count();
		if (g_verbose) printf("KD Tree for %d triangles built in %u milliseconds (%d nodes, max depth = %d, avg depth = %.1f)\n",
		       (int) triangles.size(), ms, numNodes, maxTreeDepth, nodeDepthSum / float(numNodes));
	}
	if (normals.empty()) faceted = true;
}

} // namespace fray
