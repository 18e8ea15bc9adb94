// flat.cuh -- the "flat polygon table": fast-precision (FP32) intersection of everything the reference tests by brute force.
//
// The reference walks the node list for every ray (/root/reference/src/main.cpp:178-199, 250-271 and visible(), :64-80):
// per node it transforms the ray into object space (Node::intersect, src/geometry.cpp:196-208), rejects against the mesh
// box (Mesh::intersect, src/mesh.cpp:144-165), loops over the triangles of meshes with <= 20 triangles
// (src/mesh.cpp:85,154-158; Triangle::intersectFast, src/triangle.cpp:66-94) and afterwards tests the rectangular lights
// (RectLight::intersect, src/lights.cpp:79-103). On a GPU that control flow is the cost: 32 incoherent rays of a warp
// disagree at every box test and every early-out, and the small tables are fetched again by every ray.
//
// Here all of that geometry is brought into ONE table at upload time (scene_image.h, buildFlat):
//   * every triangle of every brute-force mesh, transformed to WORLD space through its node's transform, so no per-node
//     ray transform and no box test remain (a triangle hit implies the box hit);
//   * two fan triangles of one OBJ face that are coplanar and form a convex quadrilateral are merged into one record
//     (the union of the two acceptance regions is the quadrilateral), halving the work for quad meshes;
//   * every Plane primitive (src/geometry.cpp:30-50) as the world-space image of its bounded square, one record per side;
//   * every RectLight as the world-space image of its unit square, facing the way src/lights.cpp:85-86 demands.
// A record is a plane plus up to four inward edge planes (80 bytes, five float4):
//     plane = (N, dN)          t = (dN - N.o) / (N.d)   hit point p = o + t d   front side: N.d < 0 (back-face culling,
//                                                                               src/mesh.cpp:106; two-sided meshes get
//                                                                               a second record with N reversed)
//     edge_i = (m_i, c_i)      inside  <=>  m_i.p + c_i >= 0 for i = 0..3  (for a triangle these are the barycentrics
//                                                                               lambda2, lambda3, 1-lambda2-lambda3 and 1)
// The kernels stage the table in shared memory once per CTA; the loop below is branch-free, every lane of a warp walks the
// same records (128-bit broadcast reads), and a ray costs 36 issue slots per record whatever the other lanes do.
// Attributes of the winning record (node, triangle ids, world shading normal) are looked up afterwards in FlatInfo.
#pragma once
#include <stdint.h>

#if !defined(__CUDACC__)
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
#endif

namespace fray {

#define FRAY_FLAT_POLY_VEC 5   // float4 per record
#define FRAY_MAX_FLAT 96       // records staged per scene (7.5 KB of shared memory + 4.5 KB of FlatInfo)
#define FRAY_SHADOW_LIGHTS 8    // lights that get their own shadow set (core.cuh, DScene::shadowFirst)
#define FRAY_MAX_SHADOW 160    // records in all shadow sets together (12.5 KB of shared memory)

enum {
	FRAY_FLAT_LIGHT = 1, // node = light index
	FRAY_FLAT_ATTR = 2,  // the mesh interpolates normals and/or uvs: barycentrics are recomputed for the winner
	FRAY_FLAT_QUAD = 4,  // two triangles merged; `diag` tells them apart
	FRAY_FLAT_PLANE = 8, // a Plane primitive: with FRAY_FLAT_ATTR, (u, v) = object-space (x, z) of the hit
	FRAY_FLAT_SPHERE = 16 // a Sphere primitive under a pure translation: record = (centre, R^2) in the sphere list
};

struct FlatInfo {
	float nx, ny, nz;   // world shading normal of a faceted record: normalize(gnormal * m), src/geometry.cpp:204, src/matrix.cpp:153-156
	int node;           // node index (or light index when FRAY_FLAT_LIGHT)
	int tri0, tri1;     // absolute triangle indices (tri1: the second triangle of a merged quad, else -1)
	int mesh;
	int flags;
	float4 diag;        // FRAY_FLAT_QUAD: diag . (p, 1) < 0 <=> p lies in tri1
	float4 shade;       // the node's shader as the path tracer needs it: {type (int bits), r, g, b} with the colour of a Lambert
	                    // shader or the multiplier of a reflection / refraction (textured variants read the shader table instead)
	float4 pad;         // 80-byte stride: lanes that fetch different entries from shared memory spread over eight bank groups
};

// t = h / s. On the GPU: one MUFU.RCP and one FMUL (the records are scaled at upload so that neither over- nor underflows)
FRAY_HD float flatDivide(float h, float s)
{
#if defined(__CUDA_ARCH__)
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s));
	return h * r;
#else
	return h / s;
#endif
}

// One record against one ray: the ray parameter t of the plane hit and a margin that is >= 0 iff the hit counts: front
// side (-N.d > 0), in front of the origin (t >= 0), inside all four edge planes. All conditions are folded into one minimum
// (two FMNMX3 and one FMNMX) so that a record costs 20 FFMA + 2 FMUL + 1 MUFU + 3 min + 2 compares, and no branch.
// A ray parallel to the plane gives t = +-inf or NaN and fails the `t < tBest` test of the callers.
FRAY_HD float flatTest(const float4* __restrict__ rec, float ox, float oy, float oz, float dx, float dy, float dz, float& t)
{
	const float4 pl = rec[0], e0 = rec[1], e1 = rec[2], e2 = rec[3], e3 = rec[4];
	const float s = fmaf(pl.x, dx, fmaf(pl.y, dy, pl.z * dz));
	const float h = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
	t = flatDivide(h, s);
	const float px = fmaf(dx, t, ox), py = fmaf(dy, t, oy), pz = fmaf(dz, t, oz);
	const float a = fmaf(e0.x, px, fmaf(e0.y, py, fmaf(e0.z, pz, e0.w)));
	const float b = fmaf(e1.x, px, fmaf(e1.y, py, fmaf(e1.z, pz, e1.w)));
	const float c = fmaf(e2.x, px, fmaf(e2.y, py, fmaf(e2.z, pz, e2.w)));
	const float e = fmaf(e3.x, px, fmaf(e3.y, py, fmaf(e3.z, pz, e3.w)));
	return fminf(fminf(fminf(a, b), c), fminf(fminf(e, t), -s));
}

// closest record hit by the ray (o, d) with 0 <= t < tBest; `idx` is left alone when nothing is closer
FRAY_HD void flatClosest(const float4* __restrict__ P, int n, float ox, float oy, float oz, float dx, float dy, float dz, float& tBest, int& idx)
{
#if defined(__CUDACC__)
#pragma unroll 2
#endif
	for (int i = 0; i < n; i++) {
		float t;
		const bool ok = (flatTest(P + FRAY_FLAT_POLY_VEC * i, ox, oy, oz, dx, dy, dz, t) >= 0.0f) & (t < tBest);
		tBest = ok ? t : tBest;
		idx = ok ? i : idx;
	}
}

// is any record hit with 0 <= t < tMax ?
FRAY_HD bool flatAny(const float4* __restrict__ P, int n, float ox, float oy, float oz, float dx, float dy, float dz, float tMax)
{
	bool hit = false;
#if defined(__CUDACC__)
#pragma unroll 2
#endif
	for (int i = 0; i < n; i++) {
		float t;
		hit |= (flatTest(P + FRAY_FLAT_POLY_VEC * i, ox, oy, oz, dx, dy, dz, t) >= 0.0f) & (t < tMax);
	}
	return hit;
}

// Which of the first n <= 32 records can a ray that STARTS at o hit at all, whatever its direction? A record is hit from its
// front side only (s = N.d < 0) and at t = h / s >= 0, so h = dN - N.o must not be positive: an origin behind the plane sees
// the back. h is the very number flatTest computes (same operations, same operands), and with h > 1e-12 and s < 0 the
// quotient cannot round to zero, so dropping those records changes no result. Every shadow ray of a light loop leaves from
// the same point: the mask is computed once per point and light (wave.cuh), and in a closed room it is empty.
#define FRAY_FLAT_BEHIND 1e-12f
FRAY_HD unsigned flatOriginMask(const float4* __restrict__ P, int n, float ox, float oy, float oz)
{
	unsigned mask = 0;
	for (int i = 0; i < n; i++) {
		const float4 pl = P[FRAY_FLAT_POLY_VEC * i];
		const float h = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
		mask |= (h > FRAY_FLAT_BEHIND ? 0u : 1u) << i;
	}
	return mask;
}

// flatAny over the records whose bit is set
FRAY_HD bool flatAnyMasked(const float4* __restrict__ P, unsigned mask, float ox, float oy, float oz, float dx, float dy, float dz, float tMax)
{
	bool hit = false;
	while (mask != 0) {
#if defined(__CUDA_ARCH__)
		const int i = __ffs((int) mask) - 1;
#else
		const int i = __builtin_ctz(mask);
#endif
		mask &= mask - 1u;
		float t;
		hit |= (flatTest(P + FRAY_FLAT_POLY_VEC * i, ox, oy, oz, dx, dy, dz, t) >= 0.0f) & (t < tMax);
	}
	return hit;
}

// ---- two-sided records -----------------------------------------------------------------------------------------------
// Polygons that are hit from either side -- Plane primitives (src/geometry.cpp:30-50) and the triangles of meshes without
// back-face culling -- live in a list of their own: one record per polygon instead of a front and a back copy, tested without
// the front-side term (t = h / s has the right sign whichever way the ray crosses the plane).
FRAY_HD float flatTest2(const float4* __restrict__ rec, float ox, float oy, float oz, float dx, float dy, float dz, float& t)
{
	const float4 pl = rec[0], e0 = rec[1], e1 = rec[2], e2 = rec[3], e3 = rec[4];
	const float s = fmaf(pl.x, dx, fmaf(pl.y, dy, pl.z * dz));
	const float h = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
	t = flatDivide(h, s);
	const float px = fmaf(dx, t, ox), py = fmaf(dy, t, oy), pz = fmaf(dz, t, oz);
	const float a = fmaf(e0.x, px, fmaf(e0.y, py, fmaf(e0.z, pz, e0.w)));
	const float b = fmaf(e1.x, px, fmaf(e1.y, py, fmaf(e1.z, pz, e1.w)));
	const float c = fmaf(e2.x, px, fmaf(e2.y, py, fmaf(e2.z, pz, e2.w)));
	const float e = fmaf(e3.x, px, fmaf(e3.y, py, fmaf(e3.z, pz, e3.w)));
	return fminf(fminf(fminf(a, b), c), fminf(e, t));
}

FRAY_HD void flatClosest2(const float4* __restrict__ P, int n, float ox, float oy, float oz, float dx, float dy, float dz, float& tBest, int& idx, int idxBase)
{
#if defined(__CUDACC__)
#pragma unroll 2
#endif
	for (int i = 0; i < n; i++) {
		float t;
		const bool ok = (flatTest2(P + FRAY_FLAT_POLY_VEC * i, ox, oy, oz, dx, dy, dz, t) >= 0.0f) & (t < tBest);
		tBest = ok ? t : tBest;
		idx = ok ? idxBase + i : idx;
	}
}

FRAY_HD bool flatAny2(const float4* __restrict__ P, int n, float ox, float oy, float oz, float dx, float dy, float dz, float tMax)
{
	bool hit = false;
#if defined(__CUDACC__)
#pragma unroll 2
#endif
	for (int i = 0; i < n; i++) {
		float t;
		hit |= (flatTest2(P + FRAY_FLAT_POLY_VEC * i, ox, oy, oz, dx, dy, dz, t) >= 0.0f) & (t < tMax);
	}
	return hit;
}

// ---- spheres ---------------------------------------------------------------------------------------------------------
// Sphere::intersect (src/geometry.cpp:52-83) for sphere nodes whose transform is a pure translation, as world-space
// (centre, R^2) records walked like the polygon records. Nearest non-negative root, the far one from inside; evaluated in the
// cancellation-free form the fast path uses everywhere (core.cuh, intersectSphere): discriminant from the perpendicular
// offset, roots as q and c / q.
FRAY_HD bool flatSphereTest(const float4 sp, float ox, float oy, float oz, float dx, float dy, float dz, float& t)
{
	const float hx = ox - sp.x, hy = oy - sp.y, hz = oz - sp.z;
	const float b = -(dx * hx + dy * hy + dz * hz);
	const float qx = fmaf(dx, b, hx), qy = fmaf(dy, b, hy), qz = fmaf(dz, b, hz);
	const float disc = sp.w - (qx * qx + qy * qy + qz * qz);
	const float c = (hx * hx + hy * hy + hz * hz) - sp.w;
	const float sq = sqrtf(fmaxf(disc, 0.0f));
	const float q = b + (b < 0.0f ? -sq : sq);
	const float p1 = q, p2 = (q != 0.0f) ? c / q : 0.0f;
	const float smaller = fminf(p1, p2), larger = fmaxf(p1, p2);
	t = (smaller >= 0.0f) ? smaller : larger;
	return (disc >= 0.0f) & (larger >= 0.0f);
}

FRAY_HD void flatSpheresClosest(const float4* __restrict__ S, int n, float ox, float oy, float oz, float dx, float dy, float dz, float& tBest, int& idx, int idxBase)
{
	for (int i = 0; i < n; i++) {
		float t;
		const bool ok = flatSphereTest(S[i], ox, oy, oz, dx, dy, dz, t) & (t < tBest);
		tBest = ok ? t : tBest;
		idx = ok ? idxBase + i : idx;
	}
}

FRAY_HD bool flatSpheresAny(const float4* __restrict__ S, int n, float ox, float oy, float oz, float dx, float dy, float dz, float tMax)
{
	bool hit = false;
	for (int i = 0; i < n; i++) {
		float t;
		hit |= flatSphereTest(S[i], ox, oy, oz, dx, dy, dz, t) & (t < tMax);
	}
	return hit;
}

// ---- convex hexahedra ------------------------------------------------------------------------------------------------
// A group of one-sided records that is the COMPLETE set of faces of a convex polyhedron with at most six planes -- a box,
// a prism, a pyramid; possibly open along one planar loop that a CAP plane closes, like an open box standing on the floor --
// is found at upload (scene_image.h, findHexes) and stored as six unit planes (80 -> 16 bytes per face). With back-face
// culling "which face does the ray hit" is then a clipping problem: the ray is inside the solid for t in
// [max over front-facing planes, min over back-facing planes]; it hits the face that gives the maximum, provided the
// interval is not empty, starts at t >= 0, and that face is real (through a cap the ray enters an open solid, sees only
// back faces and leaves again: no hit, exactly what testing the faces one by one gives). 14 instructions per plane, no
// edge tests, no cracks along the edges. Equivalent to Triangle::intersectFast (src/triangle.cpp:66-94) with back-face
// culling (src/mesh.cpp:106) over the faces, except on rays that pass exactly through an edge.
#define FRAY_HEX_PLANES 6
#define FRAY_HEX_VEC 7 // header {first FlatInfo index, number of real faces, cap mask of slots 4 and 5, -} + six planes: real faces, then
                       // repeats of plane 0, caps (at most two) in the last slots
#define FRAY_HEX_MAX_CAPS 2
#define FRAY_MAX_HEX 16

FRAY_HD bool flatHexTest(const float4* __restrict__ planes, float ox, float oy, float oz, float dx, float dy, float dz, float& tHit, int& jHit)
{
	const float inf = 3.0e38f;
	tHit = -inf;
	jHit = 0;
	float tBound = inf;
#if defined(__CUDACC__)
#pragma unroll
#endif
	for (int j = 0; j < FRAY_HEX_PLANES; j++) {
		const float4 pl = planes[j];
		// seeded with +0 so that a ray parallel to the plane gives s = +0, never -0: then t = +inf (origin on the inner side:
		// no constraint), -inf (outer side: the bound kills the hit) or NaN (in the plane: ignored by fminf)
		const float s = fmaf(pl.x, dx, fmaf(pl.y, dy, fmaf(pl.z, dz, 0.0f)));
		const float h = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
		const float t = flatDivide(h, s);
		// written as guarded updates, not selects: they compile to predicated moves / min instead of compare + select pairs,
		// which halves the work of the ALU pipe, the busy one in this loop
		const bool front = s < 0.0f;
		const bool better = front & (t > tHit); // strict: a repeated plane never replaces the original
		if (better) { tHit = t; jHit = j; }
		if (!front) tBound = fminf(tBound, t);
	}
	return (tHit <= tBound) & (tHit >= 0.0f);
}

FRAY_HD void flatHexClosest(const float4* __restrict__ H, int n, float ox, float oy, float oz, float dx, float dy, float dz, float& tBest, int& idx)
{
	for (int i = 0; i < n; i++) {
		const int4 hd = *reinterpret_cast<const int4*>(H + FRAY_HEX_VEC * i);
		float t;
		int j;
		const bool ok = flatHexTest(H + FRAY_HEX_VEC * i + 1, ox, oy, oz, dx, dy, dz, t, j) & (j < hd.y) & (t < tBest);
		tBest = ok ? t : tBest;
		idx = ok ? hd.x + j : idx;
	}
}

// Any-hit form: no face index is needed, only "does the segment [0, tMax) enter the solid through a real face". The entry
// parameter is the maximum over the front-facing planes, the exit the minimum over the back-facing ones (a select per plane
// and FMNMX3 chains instead of the compare / select / select of the closest-hit form). Caps sit in the LAST slots (4 and 5, at
// most two, header .z = bit mask): a front-facing cap bounds the entry from below -- entering through a cap is no hit.
FRAY_HD bool flatHexAnyTest(const float4* __restrict__ planes, bool cap4, bool cap5, float ox, float oy, float oz, float dx, float dy, float dz, float tMax)
{
	const float inf = 3.0e38f;
	float tHit = -inf, tLow = -inf, tBound = inf;
#if defined(__CUDACC__)
#pragma unroll
#endif
	for (int j = 0; j < FRAY_HEX_PLANES; j++) {
		const float4 pl = planes[j];
		const float s = fmaf(pl.x, dx, fmaf(pl.y, dy, fmaf(pl.z, dz, 0.0f))); // +0 seed: see flatHexTest
		const float h = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
		const float t = flatDivide(h, s);
		const bool cap = (j == 4 && cap4) || (j == 5 && cap5);
		const bool front = s < 0.0f;
		if (front & !cap) tHit = fmaxf(tHit, t);
		if (front & cap) tLow = fmaxf(tLow, t);
		if (!front) tBound = fminf(tBound, t);
	}
	return (tHit <= tBound) & (tHit >= 0.0f) & (tHit >= tLow) & (tHit < tMax);
}

FRAY_HD bool flatHexAny(const float4* __restrict__ H, int n, unsigned mask, float ox, float oy, float oz, float dx, float dy, float dz, float tMax)
{
	bool hit = false;
	for (int i = 0; i < n; i++) {
		if (!((mask >> i) & 1u)) continue;
		const int4 hd = *reinterpret_cast<const int4*>(H + FRAY_HEX_VEC * i);
		hit |= flatHexAnyTest(H + FRAY_HEX_VEC * i + 1, (hd.z & 1) != 0, (hd.z & 2) != 0, ox, oy, oz, dx, dy, dz, tMax);
	}
	return hit;
}

} // namespace fray
