// core.cuh -- per-ray device code of the B200 renderer: intersection, shading and the two integrators.
//
// Everything is templated on the geometry scalar R:
//   R = double : "parity" precision. Same arithmetic types as the reference (FP64 geometry, FP32 colour) with the
//                reference's literal epsilons; compiled with -fmad=false (render_fp64.cu) so a*b+c is not fused.
//   R = float  : "fast" precision. Every absolute epsilon of the reference that assumes FP64 resolution
//                (1e-6 ray offsets, 1e-6 box slack) is replaced by one that scales with the magnitude of the
//                coordinates involved (Num<float>::eps), see DESIGN.md "FP32 epsilons".
// The control flow is iterative where the reference recurses: pathtrace's tail recursion
// (/root/reference/src/main.cpp:171-244) is a loop over path segments, Whitted recursion
// (src/main.cpp:246-285 via src/shading.cpp) is a per-thread stack of weighted ray tasks, the recursive KD descent
// (src/mesh.cpp:357-394) is a stack of node indices over per-node boxes precomputed at upload, and
// findAllIntersections / CsgOp::intersect (src/geometry.cpp:139-194) keep their hit lists in fixed arrays.
//
// The functions are __host__ __device__ so that tests/emul can compile the very same per-ray code with g++ and
// run it without a GPU while debugging; the product never does that (fray_gpu.cu refuses to run without CUDA).
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string.h>

#include "rng.cuh"
#include "flat.cuh"
#include "../../include/fray_gpu.h"

namespace fray {

#if defined(FRAY_DEBUG_TRACE) && !defined(__CUDA_ARCH__)
extern thread_local int g_frayTrace; // tests/emul: print the rays of one pixel
#endif

// ---------------------------------------------------------------------------------------------------
// numeric traits
// ---------------------------------------------------------------------------------------------------
template <typename R> struct Num;

template <> struct Num<double> {
	static constexpr bool kExact = true;
	// the reference's literals: 1e-6 ray-origin offsets (src/main.cpp:144,221; src/shading.cpp:74,93,164,252; src/geometry.cpp:152)
	// and 1e-6 box slack (src/bbox.h:81-83; src/geometry.cpp:96-98)
	FRAY_HD static double offsetEps(double /*magnitude*/) { return 1e-6; }
	FRAY_HD static double reshootEps(double /*magnitude*/) { return 1e-6; }
	FRAY_HD static double slackEps(double /*magnitude*/) { return 1e-6; }
	FRAY_HD static double selfEps(double /*magnitude*/) { return 0; } // parity precision keeps the reference's rules literally
	FRAY_HD static double rcpLen(double s) { return 1.0 / sqrt(s); }
	FRAY_HD static double sqrtR(double s) { return sqrt(s); }
	FRAY_HD static double powR(double a, double b) { return pow(a, b); }
	FRAY_HD static void sincos2pi(double u, double& s, double& c)
	{
		const double a = u * 2 * 3.141592653589793238; // `randdouble() * 2 * PI`, src/random_generator.cpp:76
		s = sin(a);
		c = cos(a);
	}
	template <typename RNG> FRAY_HD static double draw(RNG& rng) { return rng.randdouble(); }
	FRAY_HD static float overPi(float c) { return (float) (c / 3.141592653589793238); } // `cosTerm / PI` is a double division
	FRAY_HD static double big() { return 1e99; }                                        // INF, src/constants.h:32
};

template <> struct Num<float> {
	static constexpr bool kExact = false;
	// FP32 cannot resolve 1e-6 next to coordinates of a few hundred (ulp(512) = 6e-5): the epsilons scale with the
	// magnitude of the coordinates involved and never drop below the reference's 1e-6 (DESIGN.md "FP32 epsilons")
#ifndef FRAY_F32_OFFSET_SCALE
#define FRAY_F32_OFFSET_SCALE 1e-6f
#endif
#ifndef FRAY_F32_RESHOOT_SCALE
#define FRAY_F32_RESHOOT_SCALE 1e-6f
#endif
#ifndef FRAY_F32_SLACK_SCALE
#define FRAY_F32_SLACK_SCALE 1e-6f
#endif
	FRAY_HD static float offsetEps(float magnitude) { return fmaxf(1e-6f, magnitude * FRAY_F32_OFFSET_SCALE); }
	FRAY_HD static float reshootEps(float magnitude) { return fmaxf(1e-6f, magnitude * FRAY_F32_RESHOOT_SCALE); }
	FRAY_HD static float slackEps(float magnitude) { return fmaxf(1e-6f, magnitude * FRAY_F32_SLACK_SCALE); }
	// A ray that starts ON a closed surface (a secondary ray: its origin is a hit point moved by offsetEps) meets that same
	// surface again at a parameter of the order of the rounding error of the hit point, with either sign. Crossings of the
	// ORIGIN'S OWN node below this bound are not hits (see intersectSphere / intersectCube / csgCrossings); the bound is a few
	// times the offset, far below any chord the scene can show at FP32 resolution.
	FRAY_HD static float selfEps(float magnitude) { return fmaxf(1e-5f, magnitude * 3e-5f); }
	FRAY_HD static float rcpLen(float s)
	{
#if defined(__CUDA_ARCH__)
		return rsqrtf(s);
#else
		return 1.0f / sqrtf(s);
#endif
	}
	FRAY_HD static float sqrtR(float s) { return sqrtf(s); }
	FRAY_HD static float powR(float a, float b) { return powf(a, b); }
	FRAY_HD static void sincos2pi(float u, float& s, float& c)
	{
#if defined(__CUDA_ARCH__)
		sincospif(2.0f * u, &s, &c);
#else
		s = sinf(6.28318530717958647692f * u);
		c = cosf(6.28318530717958647692f * u);
#endif
	}
	template <typename RNG> FRAY_HD static float draw(RNG& rng) { return rng.randdoubleAsFloat(); }
	FRAY_HD static float overPi(float c) { return c * 0.318309886183790672f; }
	FRAY_HD static float big() { return FLT_MAX; }
};

#define FRAY_PI 3.141592653589793238

template <typename R> struct V3 {
	R x, y, z;
	FRAY_HD V3() {}
	FRAY_HD V3(R x_, R y_, R z_): x(x_), y(y_), z(z_) {}
	FRAY_HD R get(int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
template <typename R> FRAY_HD V3<R> operator+(const V3<R>& a, const V3<R>& b) { return V3<R>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename R> FRAY_HD V3<R> operator-(const V3<R>& a, const V3<R>& b) { return V3<R>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename R> FRAY_HD V3<R> operator-(const V3<R>& a) { return V3<R>(-a.x, -a.y, -a.z); }
template <typename R> FRAY_HD V3<R> operator*(const V3<R>& a, R m) { return V3<R>(a.x * m, a.y * m, a.z * m); }
template <typename R> FRAY_HD R dot(const V3<R>& a, const V3<R>& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename R> FRAY_HD V3<R> cross(const V3<R>& a, const V3<R>& b)
{
	return V3<R>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename R> FRAY_HD R lengthSqr(const V3<R>& a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
template <typename R> FRAY_HD R length(const V3<R>& a) { return Num<R>::sqrtR(lengthSqr(a)); }
template <typename R> FRAY_HD V3<R> normalized(const V3<R>& a) { return a * Num<R>::rcpLen(lengthSqr(a)); } // src/vector.h:84-88
template <typename R> FRAY_HD R dist3(const V3<R>& a, const V3<R>& b) { return length(a - b); }
template <typename R> FRAY_HD R maxAbs(const V3<R>& a) { return fmax(fabs(a.x), fmax(fabs(a.y), fabs(a.z))); }
template <typename R> FRAY_HD V3<R> faceforward(const V3<R>& d, const V3<R>& n) { return dot(d, n) < 0 ? n : -n; } // src/vector.h:169-175
template <typename R> FRAY_HD V3<R> reflect(const V3<R>& i, const V3<R>& n) { return i + n * (2 * dot(-i, n)); } // src/vector.h:178-181
template <typename R> FRAY_HD V3<R> load3(const R* p) { return V3<R>(p[0], p[1], p[2]); }

struct Col {
	float r, g, b;
	FRAY_HD Col() {}
	FRAY_HD Col(float r_, float g_, float b_): r(r_), g(g_), b(b_) {}
	FRAY_HD float intensity() const { return (r + g + b) / 3; } // src/color.h:81-84
};
FRAY_HD Col operator+(const Col& a, const Col& b) { return Col(a.r + b.r, a.g + b.g, a.b + b.b); }
FRAY_HD Col operator-(const Col& a, const Col& b) { return Col(a.r - b.r, a.g - b.g, a.b - b.b); }
FRAY_HD Col operator*(const Col& a, const Col& b) { return Col(a.r * b.r, a.g * b.g, a.b * b.b); }
FRAY_HD Col operator*(const Col& a, float m) { return Col(a.r * m, a.g * m, a.b * m); }
FRAY_HD Col operator/(const Col& a, float d) { return Col(a.r / d, a.g / d, a.b / d); }
FRAY_HD Col loadCol(const float* p) { return Col(p[0], p[1], p[2]); }

// ---------------------------------------------------------------------------------------------------
// device scene (built by upload.cuh from the FrayGpuScene tables)
// ---------------------------------------------------------------------------------------------------
template <typename R> struct DXform { // struct Transform, row vectors: p' = p * m + offset
	R m[9], inv[9], off[3];
	int identity; // m == inv == I and offset == 0 (fast precision skips the arithmetic)
};

template <typename R> struct DNode {
	DXform<R> T;
	int geom, shader, bump;
	int needsUV; // some texture on this node reads (u, v); otherwise sphere uv (atan2/asin) is skipped in fast precision
	int inFlat;  // fast precision: the node's triangles live in the flat polygon table (flat.cuh) and the node loop skips it
	int pad[3];
};

template <typename R> struct DGeom {
	int type, mesh, left, right;
	R p[4];
};

template <typename R> struct DMesh {
	int flags;
	int firstTri, numTris;  // absolute ranges
	int firstNormal, firstUV;
	int kdRoot;             // absolute kd node index or -1
	int firstLeafRef;
	int pad;
	R bmin[3], bmax[3];
};

template <typename R> struct DKdNode {
	int axis; // 0..2 inner, 3 leaf
	int a;    // inner: absolute index of children[0]; leaf: absolute index of first leaf ref
	int b;    // leaf: count
	R split;
};

template <typename R> struct DShader {
	int type, texture, firstLayer, numLayers, numSamples, pureReflection;
	float color[3], specularColor[3], mult[3];
	R exponent, specularMultiplier, deflectionScaling, ior;
};

struct DLayer {
	int shader, texture;
	float opacity[3];
};

template <typename R> struct DTexture {
	int type, bitmap;
	float color1[3], color2[3];
	R scaling, ior, bumpIntensity;
};

struct DBitmap {
	int width, height;
	long long firstTexel; // RGB float triplets
};

template <typename R> struct DLight {
	int type, xSubd, ySubd;
	float color[3], power;
	R pos[3];
	DXform<R> T;
	R center[3];
	R area;
	float areaF; // (float) area, as used by RectLight::getNthSample
};

#define FRAY_LIGHT_REC_VEC 6

FRAY_HD int floatBits(float f) // the int stored in a float4 lane
{
#if defined(__CUDA_ARCH__)
	return __float_as_int(f);
#else
	int i;
	memcpy(&i, &f, sizeof(i));
	return i;
#endif
}

FRAY_HD float intBitsToFloat(int i)
{
#if defined(__CUDA_ARCH__)
	return __int_as_float(i);
#else
	float f;
	memcpy(&f, &i, sizeof(f));
	return f;
#endif
}

template <typename R> struct DCamera {
	R pos[3], topLeft[3], topRight[3], bottomLeft[3], front[3], up[3], right[3];
	R w, h, aperture, focalDist, stereoSep;
	float leftMask[3], rightMask[3];
	int dof;
	R du[3], dv[3]; // (topRight - topLeft) / w and (bottomLeft - topLeft) / h: the per-pixel steps of the screen plane (fast precision)
};

template <typename R> struct DScene {
	DCamera<R> cam;
	int maxTraceDepth, gi, numNodes, numLights, hasEnv;
	int env[6];
	float ambient[3], saturation;
	const DNode<R>* nodes;
	const DGeom<R>* geoms;
	const DMesh<R>* meshes;
	const DShader<R>* shaders;
	const DLayer* layers;
	const DTexture<R>* textures;
	const DBitmap* bitmaps;
	const DLight<R>* lights;
	// triangle planes (3 R per entry unless noted)
	const R* triA;
	const R* triAB;
	const R* triAC;
	const R* triN;      // AB ^ AC
	const R* triG;      // unit geometric normal
	const R* triDndx;
	const R* triDndy;
	const int* triNi;   // 3 normal indices per triangle (relative to mesh.firstNormal)
	const int* triTi;   // 3 uv indices per triangle (relative to mesh.firstUV)
	const R* normals;
	const R* uvs;
	const DKdNode<R>* kd;
	const R* kdBox;     // 6 R per kd node: vmin xyz, vmax xyz
	const int* leafRefs; // triangle indices relative to mesh.firstTri
	const float* texels;
	// fast precision only: one 48-byte record per triangle for the KD leaves (intersectMeshFast): plane (N, N.A) with
	// N = AB ^ AC normalised, and the two barycentric planes lambda2(p) = e2.(p,1), lambda3(p) = e3.(p,1)
	const float4* kdTris;
	// the same records once more in LEAF order: record r belongs to leaf reference leafRefs[r] (wave.cuh, kdWalk: a leaf's
	// triangles are then consecutive in memory and their loads do not wait for the reference to arrive)
	const float4* kdLeafTris;
	// fast precision only: the bounding box of the triangles of every KD leaf, two vectors {min, -} {max, -} per leaf, padded;
	// a leaf node carries the index of its box in the `split` word. The reference's trees split cells at the spatial median
	// down to 20 triangles: most rays that cross a leaf cell pass its triangles by, and the box says so before they are tested.
	const float4* kdLeafBox;
	// fast precision only: FRAY_LIGHT_REC_VEC float4 per light, what explicitLightSample needs in six 128-bit loads:
	// {type, xSubd, ySubd, samples} {centre, area} {sample-grid corner, 1 / xSubd} {column step} {row step} {colour * power}
	const float4* lightRecs;
	// fast precision only: world-space bounding box of every node's geometry, two vectors per node {min, -} {max, -}, padded by
	// more than the object-space box slack: a ray that misses it cannot hit the node (wave.cuh skips the node's ray transform)
	const float4* nodeBox;
	// fast precision only (flat.cuh): world-space convex polygons of the brute-force meshes and the rectangular lights
	const float4* flatPolys;   // FRAY_FLAT_POLY_VEC float4 per polygon
	const FlatInfo* flatInfo;  // one per polygon
	int numFlatGeom;           // polygons [0, numFlatGeom) are node geometry (occluders)
	int numFlatAll;            // polygons [numFlatGeom, numFlatAll) are lights
	int lightsInFlat;          // the rectangular lights are in the table: the light loop of closestHit is skipped
	// Shadow sets: the records that can occlude a ray towards light l at all, i.e. whose plane has some point of the light
	// behind it (a ray hits a record only from its front side, so its end point must lie behind the plane). They are compact
	// copies appended to flatPolys at [shadowFirst[l], shadowFirst[l] + shadowCount[l]); shadowCount[l] < 0 = no set, use all.
	int shadowFirst[FRAY_SHADOW_LIGHTS], shadowCount[FRAY_SHADOW_LIGHTS];
	int numFlatTotal;          // records in flatPolys including the shadow sets (what the kernels stage)
	int numFlatSpheres;        // (centre, R^2) vectors that follow the records in flatPolys; their FlatInfo follow the lights'
	int numFlatHex;            // convex hexahedra (FRAY_HEX_VEC vectors each) after the spheres; FlatInfo of their faces after the spheres'
	int numFlat2;              // two-sided records (FRAY_FLAT_POLY_VEC vectors each) after the hexahedra; one FlatInfo each, after the hexahedron faces'
	int flat2InfoBase;         // FlatInfo index of the first two-sided record
	int numFlatInfo;           // all FlatInfo entries
	unsigned shadowHex[FRAY_SHADOW_LIGHTS]; // bit k: hexahedron k can occlude a ray towards light l
	// the light loop of Lambert / Phong as the wavefront integrator sees it (wave.cuh): samples of all lights together, and the
	// random draws they consume (two per RectLight sample, src/lights.cpp:49-77)
	int lightSamples, lightDraws;
};

template <typename R> struct Ray {
	V3<R> start, dir;
};

// closest-hit record. For mesh hits the shading attributes are derived later from (tri, l2, l3).
template <typename R> struct Hit {
	R dist;
	V3<R> ip, norm;
	R u, v;
	R l2, l3;
	int tri;  // absolute triangle index or -1
	int mesh; // mesh index (for attribute fetch) or -1
	int geom; // geometry that produced the hit (CSG side bookkeeping)
	int flat; // fast precision: FlatInfo index of a hit in the flat table
};

// per-thread counts: 32 bits in the kernels (a persistent lane traces a few thousand rays per call; the warp sums are 64-bit)
#if defined(__CUDA_ARCH__)
typedef unsigned int RayCount;
#else
typedef unsigned long long RayCount;
#endif
struct RayCounters {
	RayCount rays, primary, shadow;
};

// ---------------------------------------------------------------------------------------------------
// transforms (src/matrix.cpp:142-161)
// ---------------------------------------------------------------------------------------------------
template <typename R> FRAY_HD V3<R> mulRow(const V3<R>& v, const R* m)
{
	return V3<R>(v.x * m[0] + v.y * m[3] + v.z * m[6], v.x * m[1] + v.y * m[4] + v.z * m[7], v.x * m[2] + v.y * m[5] + v.z * m[8]);
}
template <typename R> FRAY_HD V3<R> xfPoint(const DXform<R>& T, const V3<R>& p)
{
	if (!Num<R>::kExact && T.identity) return p;
	return mulRow(p, T.m) + load3(T.off);
}
template <typename R> FRAY_HD V3<R> xfUnpoint(const DXform<R>& T, const V3<R>& p)
{
	if (!Num<R>::kExact && T.identity) return p;
	return mulRow(p - load3(T.off), T.inv);
}
template <typename R> FRAY_HD V3<R> xfDir(const DXform<R>& T, const V3<R>& d)
{
	if (!Num<R>::kExact && T.identity) return d;
	return normalized(mulRow(d, T.m));
}
template <typename R> FRAY_HD V3<R> xfUndir(const DXform<R>& T, const V3<R>& d)
{
	if (!Num<R>::kExact && T.identity) return d;
	return normalized(mulRow(d, T.inv));
}

// ---------------------------------------------------------------------------------------------------
// analytic primitives, object space
// ---------------------------------------------------------------------------------------------------

// Plane::intersect, src/geometry.cpp:30-50
template <typename R> FRAY_HD bool intersectPlane(const DGeom<R>& g, const Ray<R>& ray, Hit<R>& h)
{
	const R height = g.p[0], limit = g.p[1];
	if (ray.start.y > height && ray.dir.y >= 0) return false;
	if (ray.start.y < height && ray.dir.y <= 0) return false;
	const R scaling = fabs(ray.start.y - height) / fabs(ray.dir.y);
	const V3<R> ip = ray.start + ray.dir * scaling;
	if (fabs(ip.x) > limit) return false;
	if (fabs(ip.z) > limit) return false;
	h.ip = ip;
	h.dist = dist3(ray.start, ip);
	h.norm = V3<R>(0, 1, 0);
	h.u = ip.x;
	h.v = ip.z;
	h.tri = -1;
	h.mesh = -1;
	return true;
}

// Sphere::intersect, src/geometry.cpp:52-83
// `self` (fast precision): the ray starts on this very sphere -- only the far crossing can be a hit
template <typename R> FRAY_HD bool intersectSphere(const DGeom<R>& g, const Ray<R>& ray, Hit<R>& h, bool needUV, bool self = false)
{
	const V3<R> O(g.p[0], g.p[1], g.p[2]);
	const R Rad = g.p[3];
	const V3<R> H = ray.start - O;
	R d;
	if (Num<R>::kExact) {
		const R B = 2 * dot(ray.dir, H);
		const R C = lengthSqr(H) - Rad * Rad;
		const R Disc = B * B - 4 * C;
		if (Disc < 0) return false;
		const R sq = Num<R>::sqrtR(Disc);
		const R p1 = (-B + sq) / 2, p2 = (-B - sq) / 2;
		const R smaller = fmin(p1, p2), larger = fmax(p1, p2);
		if (larger < 0) return false;
		d = (smaller >= 0) ? smaller : larger;
	} else {
		// same roots, evaluated without the cancellation of B*B - 4*C (which costs ~1e-3 of absolute accuracy in FP32 at
		// the distances of data/smallpt.fray): discriminant from the perpendicular offset vector, roots as q and c/q.
		const R b = -dot(ray.dir, H);
		const V3<R> perp = H + ray.dir * b;
		const R disc = Rad * Rad - lengthSqr(perp);
		if (disc < 0) return false;
		const R c = lengthSqr(H) - Rad * Rad;
		const R q = b + (b < 0 ? -Num<R>::sqrtR(disc) : Num<R>::sqrtR(disc));
		const R p1 = q, p2 = (q != 0) ? c / q : 0;
		const R smaller = fmin(p1, p2), larger = fmax(p1, p2);
		if (self) {
			if (!(larger > Num<R>::selfEps(fmax(maxAbs(ray.start), Rad)))) return false;
			d = larger;
		} else {
			if (larger < 0) return false;
			d = (smaller >= 0) ? smaller : larger;
		}
	}
	h.ip = ray.start + ray.dir * d;
	h.dist = dist3(ray.start, h.ip);
	h.norm = normalized(h.ip - O);
	if (Num<R>::kExact || needUV) {
		h.u = (R) ((atan2(h.norm.z, h.norm.x) / (R) FRAY_PI * 180 + 180) / 360);
		h.v = (R) (1 - (asin(h.norm.y) / (R) FRAY_PI * 180 + 90) / 180);
	} else {
		h.u = h.v = 0;
	}
	h.tri = -1;
	h.mesh = -1;
	return true;
}

// Cube::intersect + intersectCubeSide, src/geometry.cpp:85-137
template <typename R> FRAY_HD bool intersectCube(const DGeom<R>& g, const Ray<R>& ray, Hit<R>& h, bool self = false)
{
	if constexpr (!Num<R>::kExact) {
		// Fast precision: the cube as three slabs -- entry = latest near plane, exit = earliest far plane; from outside the hit is
		// the entry, from inside the exit (what the closest of the six side tests gives), for a ray that starts on this cube only
		// the exit (its entry is its own origin). Unlike six bounded side tests, whose 1e-6 slack is below the FP32 error of a hit
		// point at distance 20, the slab form has no cracks along the edges.
		const V3<R> O(g.p[0], g.p[1], g.p[2]);
		const R hs = g.p[3];
		R tIn = -FLT_MAX, tOut = FLT_MAX;
		int aIn = 0, aOut = 0;
		R sIn = 0, sOut = 0;
#if defined(__CUDACC__)
#pragma unroll
#endif
		for (int axis = 0; axis < 3; axis++) {
			const R s = ray.start.get(axis), d = ray.dir.get(axis), c = O.get(axis);
			if (fabs(d) < (R) 1e-9) { // parallel to the slab (src/geometry.cpp:110): inside it or never
				if (s < c - hs || s > c + hs) return false;
				continue;
			}
			const R r = 1 / d;
			const R tLo = (c - hs - s) * r, tHi = (c + hs - s) * r;
			const R tn = fmin(tLo, tHi), tf = fmax(tLo, tHi);
			if (tn > tIn) { tIn = tn; aIn = axis; sIn = d > 0 ? (R) -1 : (R) 1; }
			if (tf < tOut) { tOut = tf; aOut = axis; sOut = d > 0 ? (R) 1 : (R) -1; }
		}
		if (!(tIn <= tOut)) return false;
		const R tMin = self ? Num<R>::selfEps(maxAbs(O) + hs) : (R) 0;
		const bool entry = !self && tIn >= 0;
		if (!entry && !(tOut >= tMin)) return false;
		const R t = entry ? tIn : tOut;
		const int axis = entry ? aIn : aOut;
		const R sgn = entry ? sIn : sOut;
		h.ip = ray.start + ray.dir * t;
		h.dist = t; // |dir| = 1
		h.norm = V3<R>(axis == 0 ? sgn : 0, axis == 1 ? sgn : 0, axis == 2 ? sgn : 0);
		h.u = axis == 0 ? h.ip.y : h.ip.x;
		h.v = axis == 2 ? h.ip.y : h.ip.z;
		h.tri = -1;
		h.mesh = -1;
		return true;
	}
	const V3<R> O(g.p[0], g.p[1], g.p[2]);
	const R hs = g.p[3];
	const R slack = Num<R>::slackEps(maxAbs(O) + hs);
	const R tMin = self ? Num<R>::selfEps(maxAbs(O) + hs) : (R) 0; // a ray that starts on this cube cannot hit the face it leaves
	R best = (R) 1e30;
	bool found = false;
#if defined(__CUDACC__)
#pragma unroll
#endif
	for (int side = 0; side < 6; side++) {
		const int axis = side >> 1;
		const R sgn = (side & 1) ? (R) 1 : (R) -1;
		const R start = ray.start.get(axis), dir = ray.dir.get(axis);
		const R target = O.get(axis) + sgn * hs;
		if (fabs(dir) < (R) 1e-9) continue;
		const R mult = (target - start) / dir;
		if (mult < tMin) continue;
		const V3<R> ip = ray.start + ray.dir * mult;
		if (ip.x < O.x - hs - slack || ip.x > O.x + hs + slack) continue;
		if (ip.y < O.y - hs - slack || ip.y > O.y + hs + slack) continue;
		if (ip.z < O.z - hs - slack || ip.z > O.z + hs + slack) continue;
		const R d = dist3(ray.start, ip);
		if (d < best) {
			best = d;
			found = true;
			h.dist = d;
			h.ip = ip;
			h.norm = V3<R>(axis == 0 ? sgn : 0, axis == 1 ? sgn : 0, axis == 2 ? sgn : 0);
			if (axis == 0) { h.u = ip.y; h.v = ip.z; }
			else if (axis == 1) { h.u = ip.x; h.v = ip.z; }
			else { h.u = ip.x; h.v = ip.y; }
		}
	}
	h.tri = -1;
	h.mesh = -1;
	return found;
}

// ---------------------------------------------------------------------------------------------------
// meshes
// ---------------------------------------------------------------------------------------------------

// BBox::inside, src/bbox.h:79-84
template <typename R> FRAY_HD bool boxInside(const R* b, const V3<R>& v, R slack)
{
	return b[0] - slack <= v.x && v.x <= b[3] + slack && b[1] - slack <= v.y && v.y <= b[4] + slack && b[2] - slack <= v.z && v.z <= b[5] + slack;
}

// BBox::testIntersect, src/bbox.h:87-134 (with the "near slab behind the origin => skip this axis" shortcut)
template <typename R> FRAY_HD bool boxTest(const R* b, const Ray<R>& ray, const V3<R>& rdir, R slack)
{
	if (boxInside(b, ray.start, slack)) return true;
#if defined(__CUDACC__)
#pragma unroll
#endif
	for (int dim = 0; dim < 3; dim++) {
		const R d = ray.dir.get(dim), s = ray.start.get(dim);
		const R lo = b[dim], hi = b[3 + dim];
		if ((d < 0 && s < lo) || (d > 0 && s > hi)) return false;
		if (fabs(d) < (R) 1e-9) continue;
		const R mul = rdir.get(dim);
		const int u = (dim == 0) ? 1 : 0;
		const int v = (dim == 2) ? 1 : 2;
		R t = (lo - s) * mul;
		if (t < 0) continue;
		R x = ray.start.get(u) + ray.dir.get(u) * t;
		if (b[u] <= x && x <= b[3 + u]) {
			R y = ray.start.get(v) + ray.dir.get(v) * t;
			if (b[v] <= y && y <= b[3 + v]) return true;
		}
		t = (hi - s) * mul;
		if (t < 0) continue;
		x = ray.start.get(u) + ray.dir.get(u) * t;
		if (b[u] <= x && x <= b[3 + u]) {
			R y = ray.start.get(v) + ray.dir.get(v) * t;
			if (b[v] <= y && y <= b[3 + v]) return true;
		}
	}
	return false;
}

// Mesh::intersectTriangle's test part = Triangle::intersectFast, src/triangle.cpp:66-94, src/mesh.cpp:102-109.
// `maxT` plays info.dist: a hit farther than it is rejected, an equal one is accepted again (later triangle wins ties).
template <typename R>
FRAY_HD bool triangleTest(const DScene<R>& sc, int tri, bool cull, const Ray<R>& ray, R maxT, R& gamma, R& l2, R& l3)
{
	const size_t o = 3 * (size_t) tri;
	if (cull && dot(ray.dir, load3(sc.triG + o)) > 0) return false;
	const V3<R> N = load3(sc.triN + o);
	const V3<R> D = -ray.dir;
	const R Dcr = dot(N, D);
	if (fabs(Dcr) < (R) 1e-12) return false;
	const R rDcr = 1 / Dcr;
	const V3<R> H = ray.start - load3(sc.triA + o);
	gamma = dot(N, H) * rDcr;
	if (gamma < 0 || gamma > maxT) return false;
	const V3<R> AC = load3(sc.triAC + o);
	l2 = dot(cross(H, AC), D) * rDcr;
	if (l2 < 0 || l2 > 1) return false;
	const V3<R> AB = load3(sc.triAB + o);
	l3 = dot(cross(AB, H), D) * rDcr;
	if (l3 < 0 || l3 > 1) return false;
	return 1 - (l2 + l3) >= 0;
}

// the attribute part of Mesh::intersectTriangle, src/mesh.cpp:110-138 (object space)
template <typename R> FRAY_HD void triangleAttributes(const DScene<R>& sc, int meshIdx, int tri, R l2, R l3, V3<R>& norm, R& u, R& v)
{
	const DMesh<R>& m = sc.meshes[meshIdx];
	const size_t o = 3 * (size_t) tri;
	if ((m.flags & FRAY_MESH_FACETED) || !(m.flags & FRAY_MESH_HAS_NORMALS)) {
		norm = load3(sc.triG + o);
	} else {
		const V3<R> nA = load3(sc.normals + 3 * (size_t) (m.firstNormal + sc.triNi[o]));
		const V3<R> nB = load3(sc.normals + 3 * (size_t) (m.firstNormal + sc.triNi[o + 1]));
		const V3<R> nC = load3(sc.normals + 3 * (size_t) (m.firstNormal + sc.triNi[o + 2]));
		norm = normalized(nA + (nB - nA) * l2 + (nC - nA) * l3);
	}
	if (!(m.flags & FRAY_MESH_HAS_UVS)) {
		u = v = 0;
	} else {
		const R* tA = sc.uvs + 3 * (size_t) (m.firstUV + sc.triTi[o]);
		const R* tB = sc.uvs + 3 * (size_t) (m.firstUV + sc.triTi[o + 1]);
		const R* tC = sc.uvs + 3 * (size_t) (m.firstUV + sc.triTi[o + 2]);
		u = tA[0] + (tB[0] - tA[0]) * l2 + (tC[0] - tA[0]) * l3;
		v = tA[1] + (tB[1] - tA[1]) * l2 + (tC[1] - tA[1]) * l3;
	}
}

#define FRAY_KD_STACK 72 // MAX_DEPTH 64 (src/constants.h:39) + slack: one pending sibling per level

// Mesh::intersect, src/mesh.cpp:144-165 + Mesh::intersectKD, src/mesh.cpp:357-394. Object space.
// ANYHIT: stop at the first triangle hit with gamma <= maxT (used by visible(): the reference runs the full
// closest-hit query and compares the distance afterwards, which gives the same boolean).
template <typename R, bool ANYHIT>
FRAY_HD_HOT bool intersectMesh(const DScene<R>& sc, int meshIdx, const Ray<R>& ray, R maxT, Hit<R>& h)
{
	const DMesh<R>& m = sc.meshes[meshIdx];
	const V3<R> rdir(fabs(ray.dir.x) > (R) 1e-12 ? 1 / ray.dir.x : (R) 1e12, fabs(ray.dir.y) > (R) 1e-12 ? 1 / ray.dir.y : (R) 1e12,
	                 fabs(ray.dir.z) > (R) 1e-12 ? 1 / ray.dir.z : (R) 1e12); // RRay::prepareForTracing, src/bbox.h:49-54
	const R slack = Num<R>::slackEps(fmax(maxAbs(load3(m.bmin)), maxAbs(load3(m.bmax))));
	if (!boxTest(m.bmin, ray, rdir, slack)) return false;
	const bool cull = (m.flags & FRAY_MESH_BACKFACE_CULL) != 0;
	R best = maxT;
	bool found = false;
	R gamma, l2, l3;
	if (m.kdRoot < 0) {
		for (int t = m.firstTri; t < m.firstTri + m.numTris; t++)
			if (triangleTest(sc, t, cull, ray, best, gamma, l2, l3)) {
				best = gamma;
				found = true;
				h.tri = t;
				h.l2 = l2;
				h.l3 = l3;
				if (ANYHIT) break;
			}
	} else {
		int stack[FRAY_KD_STACK];
		int sp = 0;
		stack[sp++] = m.kdRoot;
		bool pending = false; // a hit was recorded but lies outside the leaf that produced it
		while (sp > 0) {
			const int ni = stack[--sp];
			const DKdNode<R> n = sc.kd[ni];
			if (n.axis == 3) {
				bool hitHere = false;
				for (int i = 0; i < n.b; i++) {
					const int t = m.firstTri + sc.leafRefs[n.a + i];
					if (triangleTest(sc, t, cull, ray, best, gamma, l2, l3)) {
						best = gamma;
						hitHere = true;
						h.tri = t;
						h.l2 = l2;
						h.l3 = l3;
						if (ANYHIT) break;
					}
				}
				if (hitHere) {
					if (ANYHIT) { found = true; break; }
					// `return found && bbox.inside(info.ip)`: a valid hit ends the whole descent
					if (boxInside(sc.kdBox + 6 * (size_t) ni, ray.start + ray.dir * best, slack)) { found = true; pending = false; break; }
					pending = true; // keeps tightening `best`, exactly like info.dist does in the reference
				}
			} else {
				const int near = ray.start.get(n.axis) < n.split ? 0 : 1;
				const int cNear = n.a + near, cFar = n.a + (1 - near);
				if (boxTest(sc.kdBox + 6 * (size_t) cFar, ray, rdir, slack) && sp < FRAY_KD_STACK) stack[sp++] = cFar;
				if (boxTest(sc.kdBox + 6 * (size_t) cNear, ray, rdir, slack) && sp < FRAY_KD_STACK) stack[sp++] = cNear;
			}
		}
		(void) pending; // a hit that no leaf box confirmed is dropped: the reference returns false in that case too
	}
	if (!found) return false;
	h.dist = best;
	h.ip = ray.start + ray.dir * best;
	h.mesh = meshIdx;
	return true;
}


// ---- fast-precision mesh traversal ------------------------------------------------------------------------------------
// Same hits as Mesh::intersect / intersectKD (src/mesh.cpp:144-165, 357-394) up to ties, organised for the GPU:
//   * the classic front-to-back KD walk over ray-parameter intervals: one 16-byte node fetch (LDG.128), one multiply and
//     two compares per inner node, instead of the reference's two full box tests per inner node; the far child goes on a
//     stack with its interval. The near child is the one on the origin's side of the split (src/mesh.cpp:368), a leaf hit
//     ends the walk if it lies within the leaf's interval (the reference: within the leaf's box, src/mesh.cpp:384-391),
//     otherwise the walk goes on with the tightened distance, exactly like info.dist does there;
//   * triangles as 48-byte plane + barycentric-plane records (three LDG.128 per test, no division until the plane hit).
template <bool ANYHIT>
FRAY_HD_HOT bool intersectMeshFast(const DScene<float>& sc, int meshIdx, const Ray<float>& ray, float maxT, Hit<float>& h)
{
	const DMesh<float>& m = sc.meshes[meshIdx];
	const float ox = ray.start.x, oy = ray.start.y, oz = ray.start.z, dx = ray.dir.x, dy = ray.dir.y, dz = ray.dir.z;
	// RRay::prepareForTracing, src/bbox.h:49-54
	const float rx = fabsf(dx) > 1e-12f ? 1.0f / dx : 1e12f, ry = fabsf(dy) > 1e-12f ? 1.0f / dy : 1e12f, rz = fabsf(dz) > 1e-12f ? 1.0f / dz : 1e12f;
	// root box as three slabs (BBox::testIntersect, src/bbox.h:87-134, with its +-1e-6 slack scaled to FP32)
	const float slack = Num<float>::slackEps(fmaxf(maxAbs(load3(m.bmin)), maxAbs(load3(m.bmax))));
	float tmin, tmax;
	{
		const float ax0 = (m.bmin[0] - slack - ox) * rx, ax1 = (m.bmax[0] + slack - ox) * rx;
		const float ay0 = (m.bmin[1] - slack - oy) * ry, ay1 = (m.bmax[1] + slack - oy) * ry;
		const float az0 = (m.bmin[2] - slack - oz) * rz, az1 = (m.bmax[2] + slack - oz) * rz;
		tmin = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
		tmax = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), maxT));
		if (!(tmin <= tmax)) return false;
	}
	const bool cull = (m.flags & FRAY_MESH_BACKFACE_CULL) != 0;
	float best = maxT;
	int bestTri = -1;
	float bestL2 = 0, bestL3 = 0;

	// one triangle: plane hit, then the two barycentric planes (Triangle::intersectFast, src/triangle.cpp:66-94)
	auto testTriangle = [&](int t) {
		const float4 pl = sc.kdTris[3 * (size_t) t], e2 = sc.kdTris[3 * (size_t) t + 1], e3 = sc.kdTris[3 * (size_t) t + 2];
		const float s = fmaf(pl.x, dx, fmaf(pl.y, dy, pl.z * dz));
		if (cull && s > 0.0f) return; // dot(dir, gnormal) > 0, src/mesh.cpp:106
		const float hh = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
		const float tt = flatDivide(hh, s);
		if (!(tt >= 0.0f && tt <= best)) return; // also rejects the parallel case (inf / NaN); `<=`: a later triangle wins ties
		const float px = fmaf(dx, tt, ox), py = fmaf(dy, tt, oy), pz = fmaf(dz, tt, oz);
		const float l2 = fmaf(e2.x, px, fmaf(e2.y, py, fmaf(e2.z, pz, e2.w)));
		const float l3 = fmaf(e3.x, px, fmaf(e3.y, py, fmaf(e3.z, pz, e3.w)));
		if (l2 < 0.0f || l3 < 0.0f || l2 + l3 > 1.0f) return;
		best = tt;
		bestTri = t;
		bestL2 = l2;
		bestL3 = l3;
	};

	bool found = false;
	if (m.kdRoot < 0) {
		for (int t = m.firstTri; t < m.firstTri + m.numTris; t++) {
			testTriangle(t);
			if (ANYHIT && bestTri >= 0) break;
		}
		found = bestTri >= 0;
	} else {
		int stackNode[FRAY_KD_STACK];
		float stackTmin[FRAY_KD_STACK], stackTmax[FRAY_KD_STACK];
		int sp = 0;
		int ni = m.kdRoot;
		const int4* nodes = reinterpret_cast<const int4*>(sc.kd);
		for (;;) {
			int4 n = nodes[ni];
			while (n.x != 3) { // inner node: axis n.x, children n.y / n.y + 1, split position in n.w
#if defined(__CUDA_ARCH__)
				const float split = __int_as_float(n.w);
#else
				float split;
				memcpy(&split, &n.w, sizeof(split));
#endif
				const float o = n.x == 0 ? ox : (n.x == 1 ? oy : oz), r = n.x == 0 ? rx : (n.x == 1 ? ry : rz);
				const float ts = (split - o) * r;
				// near child = the origin's side (src/mesh.cpp:368); an origin exactly on the split belongs to the side it moves into
				const bool lowFirst = o < split || (o == split && r <= 0.0f);
				const int nearChild = n.y + (lowFirst ? 0 : 1), farChild = n.y + (lowFirst ? 1 : 0);
				if (ts > tmax || ts <= 0.0f) {
					ni = nearChild; // the far side lies beyond the interval, or behind the origin
				} else if (ts < tmin) {
					ni = farChild;
				} else {
					if (sp < FRAY_KD_STACK) { stackNode[sp] = farChild; stackTmin[sp] = ts; stackTmax[sp] = tmax; sp++; }
					ni = nearChild;
					tmax = ts;
				}
				n = nodes[ni];
			}
			// leaf: n.y = first leaf reference, n.z = count
			for (int i = 0; i < n.z; i++) testTriangle(m.firstTri + sc.leafRefs[n.y + i]);
			if (bestTri >= 0) {
				if (ANYHIT) { found = true; break; }
				// a hit inside this leaf's interval is the closest one: everything still on the stack starts farther away
				if (best <= tmax + slack + 1e-6f * tmax) { found = true; break; }
			}
			// next pending subtree that can still contain something closer
			bool more = false;
			while (sp > 0) {
				sp--;
				if (stackTmin[sp] <= best) { ni = stackNode[sp]; tmin = stackTmin[sp]; tmax = stackTmax[sp]; more = true; break; }
			}
			if (!more) break;
		}
		// a hit that no leaf interval confirmed is still a hit of a triangle of this mesh: unlike the reference's leaf-box
		// rule there is nothing beyond the last interval, so it stands
		if (!found && bestTri >= 0) found = true;
	}
	if (!found) return false;
	h.dist = best;
	h.ip = ray.start + ray.dir * best;
	h.mesh = meshIdx;
	h.tri = bestTri;
	h.l2 = bestL2;
	h.l3 = bestL3;
	return true;
}

// geometry dispatch without CSG (the leaves of a CSG tree and plain nodes)
template <typename R, bool ANYHIT>
FRAY_HD bool intersectLeafGeom(const DScene<R>& sc, int gi, const Ray<R>& ray, R maxT, Hit<R>& h, bool needUV, bool needAttr, bool self = false)
{
	const DGeom<R>& g = sc.geoms[gi];
	bool ok;
	switch (g.type) {
		case FRAY_GEOM_PLANE: ok = intersectPlane(g, ray, h); break;
		case FRAY_GEOM_SPHERE: ok = intersectSphere(g, ray, h, needUV, self); break;
		case FRAY_GEOM_CUBE: ok = intersectCube(g, ray, h, self); break;
		case FRAY_GEOM_MESH:
			if constexpr (Num<R>::kExact) ok = intersectMesh<R, ANYHIT>(sc, g.mesh, ray, maxT, h);
			else ok = intersectMeshFast<ANYHIT>(sc, g.mesh, ray, maxT, h);
			if (ok && needAttr) triangleAttributes(sc, g.mesh, h.tri, h.l2, h.l3, h.norm, h.u, h.v);
			break;
		default: ok = false; break;
	}
	if (ok) h.geom = gi;
	return ok;
}

// the analytic primitives only (wave.cuh: meshes go through kdWalk there, and this keeps the mesh walk out of its kernels)
template <typename R>
FRAY_HD bool intersectAnalytic(const DScene<R>& sc, int gi, const Ray<R>& ray, Hit<R>& h, bool needUV, bool self)
{
	const DGeom<R>& g = sc.geoms[gi];
	bool ok;
	switch (g.type) {
		case FRAY_GEOM_PLANE: ok = intersectPlane(g, ray, h); break;
		case FRAY_GEOM_SPHERE: ok = intersectSphere(g, ray, h, needUV, self); break;
		case FRAY_GEOM_CUBE: ok = intersectCube(g, ray, h, self); break;
		default: ok = false; break;
	}
	if (ok) h.geom = gi;
	return ok;
}

#define FRAY_CSG_MAX_HITS 30 // `counter = 30`, src/geometry.cpp:144

template <typename R> struct CsgHit {
	R dist;
	V3<R> ip, norm;
	R u, v;
	int tri, mesh, geom;
};

template <typename R, int LEVEL> struct CsgEval;

template <typename R, int LEVEL>
FRAY_HD bool intersectGeomCsg(const DScene<R>& sc, int gi, const Ray<R>& ray, Hit<R>& h)
{
	const int type = sc.geoms[gi].type;
	if (type >= FRAY_GEOM_CSG_PLUS && type <= FRAY_GEOM_CSG_MINUS) return CsgEval<R, LEVEL>::run(sc, gi, ray, h, false);
	return intersectLeafGeom<R, false>(sc, gi, ray, Num<R>::big(), h, true, true);
}

// findAllIntersections, src/geometry.cpp:139-157
template <typename R, int LEVEL>
FRAY_HD_COLD int csgAllHits(const DScene<R>& sc, int gi, const Ray<R>& rayIn, CsgHit<R>* out)
{
	Ray<R> ray = rayIn;
	int n = 0;
	int counter = FRAY_CSG_MAX_HITS;
	Hit<R> h;
	while (intersectGeomCsg<R, LEVEL>(sc, gi, ray, h) && counter-- > 0) {
		CsgHit<R>& c = out[n];
		c.dist = n == 0 ? h.dist : dist3(h.ip, rayIn.start);
		c.ip = h.ip; c.norm = h.norm; c.u = h.u; c.v = h.v; c.tri = h.tri; c.mesh = h.mesh; c.geom = gi;
		n++;
		ray.start = h.ip + ray.dir * Num<R>::reshootEps(maxAbs(h.ip));
	}
	return n;
}

// ---- fast precision: every boundary crossing of a geometry along the ray, ascending ---------------------------------------
// findAllIntersections (src/geometry.cpp:139-157) restarts the ray 1e-6 behind every hit. In FP32 a restarted ray sits within
// rounding error of the surface it just left and meets it again (a sphere answers 17.3754501 and then 17.3754520), the parity of
// the hit count flips and CsgOp::intersect (:159-194) reports the wrong side. So the crossings are enumerated from the ORIGINAL
// ray instead: both roots of a sphere, entry and exit of a cube, the one crossing of a plane, and for an operand that is itself
// a CSG every change of its inside state. Same lists as the reference's in exact arithmetic. `tMin`: crossings below it are
// dropped -- zero for an ordinary ray, selfEps for a ray that starts on this very node, whose own surface passes through its
// origin (the inside state then follows from the crossings AHEAD, i.e. it is that of the side the ray travels into, which is the
// side the reference's 1e-6 offset puts it on).
template <int LEVEL> struct CsgCross {
	static FRAY_HD_COLD int run(const DScene<float>& sc, int gi, const Ray<float>& ray, float tMin, CsgHit<float>* out);
};
template <> struct CsgCross<FRAY_GPU_MAX_CSG_DEPTH + 1> {
	static FRAY_HD int run(const DScene<float>&, int, const Ray<float>&, float, CsgHit<float>*) { return 0; }
};

FRAY_HD bool csgOp(int type, bool a, bool b)
{
	return type == FRAY_GEOM_CSG_PLUS ? (a || b) : (type == FRAY_GEOM_CSG_AND ? (a && b) : (a && !b));
}

template <int LEVEL>
FRAY_HD_COLD int CsgCross<LEVEL>::run(const DScene<float>& sc, int gi, const Ray<float>& ray, float tMin, CsgHit<float>* out)
{
	const DGeom<float>& g = sc.geoms[gi];
	int n = 0;
	auto put = [&](float t, const V3<float>& ip, const V3<float>& norm, float u, float v) {
		CsgHit<float>& c = out[n++];
		c.dist = t; c.ip = ip; c.norm = norm; c.u = u; c.v = v; c.tri = -1; c.mesh = -1; c.geom = gi;
	};
	switch (g.type) {
		case FRAY_GEOM_SPHERE: { // Sphere::intersect, src/geometry.cpp:52-83: both roots, cancellation-free (see intersectSphere)
			const V3<float> O(g.p[0], g.p[1], g.p[2]);
			const float Rad = g.p[3];
			const V3<float> H = ray.start - O;
			const float b = -dot(ray.dir, H);
			const V3<float> perp = H + ray.dir * b;
			const float disc = Rad * Rad - lengthSqr(perp);
			if (disc < 0) return 0;
			const float c = lengthSqr(H) - Rad * Rad;
			const float sq = sqrtf(disc);
			const float q = b + (b < 0 ? -sq : sq);
			const float p1 = q, p2 = (q != 0) ? c / q : 0;
			const float roots[2] = { fminf(p1, p2), fmaxf(p1, p2) };
			for (int k = 0; k < 2; k++) {
				if (!(roots[k] >= tMin)) continue;
				const V3<float> ip = ray.start + ray.dir * roots[k];
				const V3<float> nn = normalized(ip - O);
				put(roots[k], ip, nn, (float) ((atan2f(nn.z, nn.x) / (float) FRAY_PI * 180 + 180) / 360), (float) (1 - (asinf(nn.y) / (float) FRAY_PI * 180 + 90) / 180));
			}
			return n;
		}
		case FRAY_GEOM_CUBE: { // Cube::intersect, src/geometry.cpp:85-137, as three slabs: entry = latest near plane, exit = earliest far plane
			const V3<float> O(g.p[0], g.p[1], g.p[2]);
			const float hs = g.p[3];
			float tIn = -FLT_MAX, tOut = FLT_MAX;
			int aIn = 0, aOut = 0;
			float sIn = 0, sOut = 0;
			for (int axis = 0; axis < 3; axis++) {
				const float s = ray.start.get(axis), d = ray.dir.get(axis), c = O.get(axis);
				if (fabsf(d) < 1e-9f) { // parallel to the slab (src/geometry.cpp:110): inside it or never
					if (s < c - hs || s > c + hs) return 0;
					continue;
				}
				const float r = 1.0f / d;
				const float tLo = (c - hs - s) * r, tHi = (c + hs - s) * r;
				const float tn = fminf(tLo, tHi), tf = fmaxf(tLo, tHi);
				if (tn > tIn) { tIn = tn; aIn = axis; sIn = d > 0 ? -1.0f : 1.0f; }
				if (tf < tOut) { tOut = tf; aOut = axis; sOut = d > 0 ? 1.0f : -1.0f; }
			}
			if (!(tIn <= tOut)) return 0;
			for (int k = 0; k < 2; k++) {
				const float t = k == 0 ? tIn : tOut;
				const int axis = k == 0 ? aIn : aOut;
				const float sgn = k == 0 ? sIn : sOut;
				if (!(t >= tMin)) continue;
				const V3<float> ip = ray.start + ray.dir * t;
				const V3<float> nn(axis == 0 ? sgn : 0, axis == 1 ? sgn : 0, axis == 2 ? sgn : 0);
				put(t, ip, nn, axis == 0 ? ip.y : ip.x, axis == 2 ? ip.y : ip.z);
			}
			return n;
		}
		case FRAY_GEOM_PLANE: {
			Hit<float> h;
			if (intersectPlane(g, ray, h) && h.dist >= tMin) put(h.dist, h.ip, h.norm, h.u, h.v);
			return n;
		}
		case FRAY_GEOM_MESH: { // restart behind every hit like the reference; a re-hit within selfEps of the previous one is the same crossing
			Ray<float> r = ray;
			Hit<float> h;
			float last = -FLT_MAX;
			int counter = FRAY_CSG_MAX_HITS;
			while (n < FRAY_CSG_MAX_HITS && counter-- > 0 && intersectLeafGeom<float, false>(sc, gi, r, Num<float>::big(), h, true, true)) {
				const float t = dist3(h.ip, ray.start);
				const float sep = Num<float>::selfEps(maxAbs(h.ip));
				if (t >= tMin && t > last + sep) {
					CsgHit<float>& c = out[n++];
					c.dist = t; c.ip = h.ip; c.norm = h.norm; c.u = h.u; c.v = h.v; c.tri = h.tri; c.mesh = h.mesh; c.geom = gi;
					last = t;
				}
				r.start = h.ip + r.dir * Num<float>::reshootEps(maxAbs(h.ip));
			}
			return n;
		}
		default: { // CsgOp::intersect, src/geometry.cpp:159-194, carried on past the first change of state
			if (g.type < FRAY_GEOM_CSG_PLUS || g.type > FRAY_GEOM_CSG_MINUS) return 0;
			CsgHit<float> L[FRAY_CSG_MAX_HITS], Rr[FRAY_CSG_MAX_HITS];
			const int nL = CsgCross<LEVEL + 1>::run(sc, g.left, ray, tMin, L);
			const int nR = CsgCross<LEVEL + 1>::run(sc, g.right, ray, tMin, Rr);
			bool inL = (nL & 1) == 1, inR = (nR & 1) == 1;
#if defined(FRAY_DEBUG_TRACE) && !defined(__CUDA_ARCH__)
			if (g_frayTrace) {
				printf("    csg geom %d level %d: left %d crossings:", gi, LEVEL, nL);
				for (int k = 0; k < nL; k++) printf(" %.7f", (double) L[k].dist);
				printf(" | right %d crossings:", nR);
				for (int k = 0; k < nR; k++) printf(" %.7f", (double) Rr[k].dist);
				printf("\n");
			}
#endif
			bool state = csgOp(g.type, inL, inR);
			int i = 0, j = 0;
			while ((i < nL || j < nR) && n < FRAY_CSG_MAX_HITS) {
				const bool takeL = (j >= nR) || (i < nL && !(Rr[j].dist < L[i].dist));
				const CsgHit<float>& c = takeL ? L[i] : Rr[j];
				if (takeL || g.left == g.right) inL = !inL; else inR = !inR;
				if (takeL) i++; else j++;
				const bool now = csgOp(g.type, inL, inR);
				if (now != state) {
					out[n] = c;
					out[n].geom = gi;
					n++;
					state = now;
				}
			}
			return n;
		}
	}
}

template <typename R, int LEVEL> struct CsgEval {
	// CsgOp::intersect, src/geometry.cpp:159-194 (the two hit lists are already sorted along the ray; they are merged
	// with the left operand first on equal distance). `self`: the ray starts on the node this geometry belongs to.
	FRAY_HD_COLD static bool run(const DScene<R>& sc, int gi, const Ray<R>& ray, Hit<R>& h, bool self)
	{
		const DGeom<R>& g = sc.geoms[gi];
		if constexpr (!Num<R>::kExact) {
			if (LEVEL == 0) { // the first change of state along the ray
				CsgHit<float> X[FRAY_CSG_MAX_HITS];
				const float tMin = self ? Num<float>::selfEps(maxAbs(ray.start)) : 0.0f;
				if (CsgCross<0>::run(sc, gi, ray, tMin, X) == 0) return false;
				const CsgHit<float>& c = X[0];
				h.dist = c.dist; h.ip = c.ip; h.norm = c.norm; h.u = c.u; h.v = c.v; h.tri = c.tri; h.mesh = c.mesh;
				h.geom = gi;
				return true;
			}
		}
		CsgHit<R> L[FRAY_CSG_MAX_HITS], Rr[FRAY_CSG_MAX_HITS];
		const int nL = csgAllHits<R, LEVEL + 1>(sc, g.left, ray, L);
		const int nR = csgAllHits<R, LEVEL + 1>(sc, g.right, ray, Rr);
		bool inL = (nL & 1) == 1, inR = (nR & 1) == 1;
#if defined(FRAY_DEBUG_TRACE) && !defined(__CUDA_ARCH__)
		if (g_frayTrace) {
			printf("    csg geom %d level %d: left %d hits:", gi, LEVEL, nL);
			for (int k = 0; k < nL; k++) printf(" %.7f", (double) L[k].dist);
			printf(" | right %d hits:", nR);
			for (int k = 0; k < nR; k++) printf(" %.7f", (double) Rr[k].dist);
			printf("\n");
		}
#endif
		const int type = g.type;
		const bool start = csgOp(type, inL, inR);
		int i = 0, j = 0;
		while (i < nL || j < nR) {
			const bool takeL = (j >= nR) || (i < nL && !(Rr[j].dist < L[i].dist));
			const CsgHit<R>& c = takeL ? L[i] : Rr[j];
			// `if (ip.geom == left) inLeft = !inLeft; else inRight = !inRight;`
			if (takeL || g.left == g.right) inL = !inL; else inR = !inR;
			if (takeL) i++; else j++;
			if (csgOp(type, inL, inR) != start) {
				h.dist = c.dist; h.ip = c.ip; h.norm = c.norm; h.u = c.u; h.v = c.v; h.tri = c.tri; h.mesh = c.mesh;
				h.geom = gi;
				return true;
			}
		}
		return false;
	}
};
template <typename R> struct CsgEval<R, FRAY_GPU_MAX_CSG_DEPTH> {
	FRAY_HD static bool run(const DScene<R>&, int, const Ray<R>&, Hit<R>&, bool) { return false; } // rejected at flatten time
};

// feature bits (template parameter F of the kernels): code paths that cost registers, local memory and instruction-cache
// footprint are compiled in only for scenes that need them. Parity precision always runs with FRAY_F_NODES | FRAY_F_TEX.
#define FRAY_F_CSG 1    // CSG geometry exists
#define FRAY_F_NODES 2  // some node is traced through the generic node loop (analytic primitives, KD meshes)
#define FRAY_F_TEX 4    // textures, bump maps or an environment exist
#define FRAY_F_FLAT 8   // fast precision: the flat polygon table (flat.cuh) holds the brute-force meshes and the lights
#define FRAY_F_ATTR 16  // some flat record interpolates normals / uvs
#define FRAY_F_SPHERES 32 // the flat table has a sphere list
#define FRAY_F_HEX 64     // the flat table has convex hexahedra
#define FRAY_F_LENS 128   // the camera uses depth of field or stereo
#define FRAY_F_TWOSIDED 256 // the flat table has a list of two-sided records (the untextured variants beyond the lean one)
#define FRAY_F_GENERIC (FRAY_F_NODES | FRAY_F_TEX | FRAY_F_LENS)

// The kernel variants compiled per precision, smallest first; a scene runs on the first one that covers its feature bits.
template <typename R> struct Variants;
template <> struct Variants<float> {
	static constexpr int count = 6;
	static constexpr int kLean = FRAY_F_FLAT | FRAY_F_HEX;
	static constexpr int kAll = kLean | FRAY_F_SPHERES | FRAY_F_ATTR | FRAY_F_NODES | FRAY_F_TEX | FRAY_F_LENS;
	static constexpr int mask(int i)
	{
		return i == 0 ? kLean                                                      // brute-force meshes, planes, lights (cornell_box)
		     : i == 1 ? (kLean | FRAY_F_SPHERES | FRAY_F_TWOSIDED)                   // + translated spheres, two-sided polygons (smallpt)
		     : i == 2 ? (FRAY_F_FLAT | FRAY_F_ATTR | FRAY_F_TEX | FRAY_F_LENS)       // textured brute-force meshes and planes, any camera (zaphod)
		     : i == 3 ? (kLean | FRAY_F_SPHERES | FRAY_F_NODES | FRAY_F_TWOSIDED)    // + KD meshes / other primitives, untextured
		     : i == 4 ? kAll                                                         // everything but CSG
		              : (kAll | FRAY_F_CSG | FRAY_F_TWOSIDED);                       // the fallback for every combination (e.g. a lens added later to a scene with a two-sided list)
	}
};
template <> struct Variants<double> {
	static constexpr int count = 2;
	static constexpr int mask(int i) { return i == 0 ? FRAY_F_GENERIC : (FRAY_F_GENERIC | FRAY_F_CSG); }
};

// where the kernels staged the flat table (shared memory on the GPU, the scene blob in the host emulator)
struct FlatTab {
	const float4* polys;
	const FlatInfo* info;
	const float4* spheres;
	const float4* hexes;
	const float4* polys2; // two-sided records
};

// Node::intersect, src/geometry.cpp:196-208. On success h is in WORLD space (ip, norm, dist).
// `maxDist` bounds the search in world units: hits farther away may be dropped (closest-hit pruning against the best
// node so far, fast precision only; the reference compares afterwards, src/main.cpp:256) and, for ANYHIT, the mesh
// descent stops at the first triangle within it (h then only carries dist).
// `origin`: the node the ray starts on (a secondary or shadow ray), or -1. Fast precision uses it to tell the ray's own
// surface from a hit (Num<float>::selfEps); parity precision ignores it.
template <typename R, bool ANYHIT, int F>
FRAY_HD bool intersectNode(const DScene<R>& sc, int ni, const Ray<R>& ray, R maxDist, Hit<R>& h, int origin = -1)
{
	const DNode<R>& n = sc.nodes[ni];
	const bool self = !Num<R>::kExact && ni == origin;
	Ray<R> local;
	R scale = 1; // |dir * inv|: object-space ray parameter = scale * world distance
	if (!Num<R>::kExact && n.T.identity) {
		local = ray;
	} else {
		local.start = mulRow(ray.start - load3(n.T.off), n.T.inv);
		const V3<R> d = mulRow(ray.dir, n.T.inv);
		const R l2 = lengthSqr(d);
		const R rl = Num<R>::rcpLen(l2);
		local.dir = d * rl;
		scale = l2 * rl;
	}
	const int type = sc.geoms[n.geom].type;
	bool ok;
	if (type >= FRAY_GEOM_CSG_PLUS && type <= FRAY_GEOM_CSG_MINUS) {
		if (F & FRAY_F_CSG) ok = CsgEval<R, 0>::run(sc, n.geom, local, h, self);
		else ok = false;
	} else {
		R maxT = Num<R>::big();
		if (ANYHIT || !Num<R>::kExact) maxT = maxDist < Num<R>::big() / 4 ? maxDist * scale * (ANYHIT ? (R) 1 : (R) 1.0001) : Num<R>::big();
		if (ANYHIT) ok = intersectLeafGeom<R, true>(sc, n.geom, local, maxT, h, false, false, self);
		else ok = intersectLeafGeom<R, false>(sc, n.geom, local, maxT, h, n.needsUV != 0, true, self);
	}
	if (!ok) return false;
	h.ip = xfPoint(n.T, h.ip);
	h.dist = dist3(ray.start, h.ip);
	if (!ANYHIT) h.norm = xfDir(n.T, h.norm);
	return true;
}

// visible(), src/main.cpp:64-80
// `light`: index of the light the end point b was sampled on (selects the light's shadow set), or -1
// `origin`: the node the point a lies on (see intersectNode)
template <typename R, int F> FRAY_HD_HOT bool visible(const DScene<R>& sc, const FlatTab& ft, const V3<R>& a, const V3<R>& b, int light, RayCounters& cnt, int origin = -1)
{
	cnt.rays++;
	cnt.shadow++;
	Ray<R> ray;
	ray.dir = b - a;
	ray.start = a;
	const R maxDist = length(ray.dir);
	ray.dir = normalized(ray.dir);
	if constexpr ((F & FRAY_F_FLAT) != 0 && !Num<R>::kExact) {
		int first = 0, count = sc.numFlatGeom;
		if (light >= 0 && light < FRAY_SHADOW_LIGHTS && sc.shadowCount[light] >= 0) {
			first = sc.shadowFirst[light];
			count = sc.shadowCount[light];
		}
		if (flatAny(ft.polys + FRAY_FLAT_POLY_VEC * first, count, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		if (F & FRAY_F_HEX) {
			const unsigned hexMask = (light >= 0 && light < FRAY_SHADOW_LIGHTS) ? sc.shadowHex[light] : 0xffffffffu;
			if (flatHexAny(ft.hexes, sc.numFlatHex, hexMask, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		}
		if ((F & FRAY_F_SPHERES) && flatSpheresAny(ft.spheres, sc.numFlatSpheres, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		if ((F & FRAY_F_TWOSIDED) && flatAny2(ft.polys2, sc.numFlat2, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
	}
	if (F & FRAY_F_NODES) {
		for (int n = 0; n < sc.numNodes; n++) {
			if ((F & FRAY_F_FLAT) && sc.nodes[n].inFlat) continue;
			Hit<R> h;
			if (intersectNode<R, true, F>(sc, n, ray, maxDist, h, origin) && h.dist < maxDist) return false;
		}
	}
	return true;
}

// RectLight::intersect, src/lights.cpp:79-103
template <typename R> FRAY_HD bool intersectLight(const DLight<R>& l, const Ray<R>& ray, R& dist)
{
	if (l.type != FRAY_LIGHT_RECT) return false;
	const V3<R> s = xfUnpoint(l.T, ray.start);
	const V3<R> d = xfUndir(l.T, ray.dir);
	if (s.y >= 0) return false;
	if (d.y <= 0) return false;
	const R scaling = fabs(s.y) / fabs(d.y);
	const V3<R> ip = s + d * scaling;
	if (fabs(ip.x) > (R) 0.5 || fabs(ip.z) > (R) 0.5) return false;
	dist = dist3(ray.start, xfPoint(l.T, ip));
	return true;
}

// barycentrics of world point `p` in triangle `tri` of node `nd` (flat records that interpolate attributes)
template <typename R> FRAY_HD void flatBarycentrics(const DScene<R>& sc, const DNode<R>& nd, int tri, const V3<R>& p, R& l2, R& l3)
{
	const size_t o = 3 * (size_t) tri;
	const V3<R> H = xfUnpoint(nd.T, p) - load3(sc.triA + o);
	const V3<R> N = load3(sc.triN + o), AB = load3(sc.triAB + o), AC = load3(sc.triAC + o);
	const R rNN = 1 / dot(N, N);
	l2 = dot(cross(H, AC), N) * rNN;
	l3 = dot(cross(AB, H), N) * rNN;
}

// the winner `idx` of the flat-table loops (a FlatInfo index) as a hit: node or light, hit point, shading normal, (u, v)
template <typename R, int F>
FRAY_HD void flatResolve(const DScene<R>& sc, const FlatTab& ft, int idx, const Ray<R>& ray, int& node, int& light, Hit<R>& best)
{
	const FlatInfo& fi = ft.info[idx];
	if (fi.flags & FRAY_FLAT_LIGHT) {
		light = fi.node;
	} else {
		node = fi.node;
		best.flat = idx;
		best.ip = ray.start + ray.dir * best.dist;
		best.norm = V3<R>(fi.nx, fi.ny, fi.nz);
		best.u = best.v = 0;
		best.mesh = fi.mesh;
		best.tri = fi.tri0;
		if ((fi.flags & FRAY_FLAT_QUAD) && fi.diag.x * best.ip.x + fi.diag.y * best.ip.y + fi.diag.z * best.ip.z + fi.diag.w < 0) best.tri = fi.tri1;
		if ((F & FRAY_F_SPHERES) && (fi.flags & FRAY_FLAT_SPHERE)) { // info.norm = ip - O, normalised (src/geometry.cpp:71-72)
			const float4 sp = ft.spheres[idx - sc.numFlatAll];
			best.norm = normalized(best.ip - V3<R>(sp.x, sp.y, sp.z));
		}
		if ((F & FRAY_F_ATTR) && (fi.flags & FRAY_FLAT_ATTR)) {
			const DNode<R>& nd = sc.nodes[node];
			if (fi.flags & FRAY_FLAT_SPHERE) { // spherical coordinates of the normal, src/geometry.cpp:73-80
				best.u = (R) ((atan2(best.norm.z, best.norm.x) / (R) FRAY_PI * 180 + 180) / 360);
				best.v = (R) (1 - (asin(best.norm.y) / (R) FRAY_PI * 180 + 90) / 180);
			} else if (fi.flags & FRAY_FLAT_PLANE) { // info.u = ip.x, info.v = ip.z in object space, src/geometry.cpp:45-46
				const V3<R> q = xfUnpoint(nd.T, best.ip);
				best.u = q.x;
				best.v = q.z;
			} else {
				flatBarycentrics(sc, nd, best.tri, best.ip, best.l2, best.l3);
				triangleAttributes(sc, fi.mesh, best.tri, best.l2, best.l3, best.norm, best.u, best.v);
				best.norm = xfDir(nd.T, best.norm);
			}
		}
	}
}

// the two closest-hit loops of raytrace()/pathtrace(), src/main.cpp:178-199, 250-271
template <typename R, int F>
FRAY_HD_HOT void closestHit(const DScene<R>& sc, const FlatTab& ft, const Ray<R>& ray, int& node, int& light, Hit<R>& best, int origin = -1)
{
	node = -1;
	light = -1;
	best.dist = Num<R>::big();
	best.tri = -1;
	best.mesh = -1;
	if constexpr ((F & FRAY_F_FLAT) != 0 && !Num<R>::kExact) {
		int idx = -1;
		flatClosest(ft.polys, sc.numFlatAll, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, best.dist, idx);
		if (F & FRAY_F_HEX) flatHexClosest(ft.hexes, sc.numFlatHex, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, best.dist, idx);
		if (F & FRAY_F_SPHERES) flatSpheresClosest(ft.spheres, sc.numFlatSpheres, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, best.dist, idx, sc.numFlatAll);
		if (F & FRAY_F_TWOSIDED) flatClosest2(ft.polys2, sc.numFlat2, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, best.dist, idx, sc.flat2InfoBase);
		if (idx >= 0) flatResolve<R, F>(sc, ft, idx, ray, node, light, best);
	}
	if (F & FRAY_F_NODES) {
		for (int n = 0; n < sc.numNodes; n++) {
			if ((F & FRAY_F_FLAT) && sc.nodes[n].inFlat) continue;
			Hit<R> h;
			if (intersectNode<R, false, F>(sc, n, ray, best.dist, h, origin) && h.dist < best.dist) {
				best = h;
				node = n;
				light = -1;
			}
		}
	}
	if (!(F & FRAY_F_FLAT) || ((F & FRAY_F_NODES) && !sc.lightsInFlat)) { // (lights outside the table force FRAY_F_NODES, scene_image.h)
		for (int l = 0; l < sc.numLights; l++) {
			R d;
			if (intersectLight(sc.lights[l], ray, d) && d < best.dist) {
				best.dist = d;
				light = l;
			}
		}
	}
}

// ---------------------------------------------------------------------------------------------------
// textures, bump, environment
// ---------------------------------------------------------------------------------------------------
template <typename R> FRAY_HD Col bitmapPixel(const DScene<R>& sc, int bi, int x, int y) // Bitmap::getPixel, src/bitmap.cpp:67-71
{
	if (bi < 0) return Col(0, 0, 0);
	const DBitmap& b = sc.bitmaps[bi];
	if (x < 0 || x >= b.width || y < 0 || y >= b.height) return Col(0, 0, 0);
	return loadCol(sc.texels + 3 * ((size_t) b.firstTexel + x + (size_t) y * b.width));
}

template <typename R> FRAY_HD void wrappedTexel(const DBitmap& b, R u, R v, R scaling, int& ix, int& iy) // src/shading.cpp:149-155
{
	ix = (int) floor(u * scaling * b.width);
	iy = (int) floor(v * scaling * b.height);
	ix %= b.width;
	iy %= b.height;
	if (ix < 0) ix += b.width;
	if (iy < 0) iy += b.height;
}

FRAY_HD float fresnelSchlick(float NdotI, float ior) // fresnel(), src/shading.cpp:230-236
{
	const float q = (1.0f - ior) / (1.0f + ior);
	const float f = q * q;
	return f + (1.0f - f) * powf(1.0f - NdotI, 5.0f);
}

template <typename R> FRAY_HD Col sampleTexture(const DScene<R>& sc, int ti, const V3<R>& rayDir, const V3<R>& norm, R u, R v)
{
	const DTexture<R>& t = sc.textures[ti];
	switch (t.type) {
		case FRAY_TEX_CHECKER: { // src/shading.cpp:40-46: floor, divide by 5, truncate toward zero
			const int ix = (int) (floor(u * t.scaling) / 5);
			const int iy = (int) (floor(v * t.scaling) / 5);
			return ((ix + iy) % 2 == 0) ? loadCol(t.color1) : loadCol(t.color2);
		}
		case FRAY_TEX_BITMAP: { // src/shading.cpp:147-158
			int ix, iy;
			wrappedTexel(sc.bitmaps[t.bitmap], u, v, t.scaling, ix, iy);
			return bitmapPixel(sc, t.bitmap, ix, iy);
		}
		case FRAY_TEX_FRESNEL: { // src/shading.cpp:369-385
			const R d = dot(rayDir, norm);
			// n = +-norm so that dot(n, dir) <= 0; NdotI = -dot(n, dir) = |d|
			const float ior = d < 0 ? (float) t.ior : (float) (1 / t.ior);
			const float f = fresnelSchlick((float) fabs(d), ior);
			return Col(f, f, f);
		}
		default: return Col(0, 0, 0); // BumpTexture::sample, src/shading.cpp:392-395
	}
}

// applyBumpMapping, src/main.cpp:82-90 + BumpTexture::getDeflection/modifyNormal, src/shading.cpp:397-418.
// dNdx / dNdy stay in object space (src/geometry.cpp:204-206) and are zero for non-mesh hits.
template <typename R> FRAY_HD void applyBump(const DScene<R>& sc, const DNode<R>& node, Hit<R>& h)
{
	if (node.bump < 0) return;
	const DTexture<R>& t = sc.textures[node.bump];
	if (t.type != FRAY_TEX_BUMP) return;
	int ix, iy;
	wrappedTexel(sc.bitmaps[t.bitmap], h.u, h.v, t.scaling, ix, iy);
	const Col tex = bitmapPixel(sc, t.bitmap, ix, iy);
	const float dx = (float) (tex.r * t.bumpIntensity);
	const float dy = (float) (tex.g * t.bumpIntensity);
	V3<R> dNdx(0, 0, 0), dNdy(0, 0, 0);
	if (h.tri >= 0) {
		dNdx = load3(sc.triDndx + 3 * (size_t) h.tri);
		dNdy = load3(sc.triDndy + 3 * (size_t) h.tri);
	}
	h.norm = normalized(h.norm + (dNdx * (R) dx + dNdy * (R) dy) * t.bumpIntensity);
}

// CubemapEnvironment::getEnvironment / getSide, src/environment.cpp:64-98
template <typename R> FRAY_HD Col environmentLookup(const DScene<R>& sc, const V3<R>& dir)
{
	const R ax = fabs(dir.x), ay = fabs(dir.y), az = fabs(dir.z);
	int dim = 0;
	R maxVal = ax;
	if (ay > maxVal) { dim = 1; maxVal = ay; }
	if (az > maxVal) dim = 2;
	const R dv = dir.get(dim);
	const R m = 1 / fabs(dv);
	const V3<R> s = dir * m;
	const int face = (dv > 0 ? 3 : 0) + dim;
	R x, y;
	switch (face) {
		case 0: x = s.z; y = -s.y; break;
		case 3: x = -s.z; y = -s.y; break;
		case 1: x = s.x; y = -s.z; break;
		case 4: x = s.x; y = s.z; break;
		case 2: x = s.x; y = s.y; break;
		default: x = s.x; y = -s.y; break;
	}
	const int bi = sc.env[face];
	if (bi < 0) return Col(0, 0, 0);
	const DBitmap& b = sc.bitmaps[bi];
	const int ix = (int) (((x + 1) / 2) * b.width);
	const int iy = (int) (((y + 1) / 2) * b.height);
	return bitmapPixel(sc, bi, ix, iy);
}

// ---------------------------------------------------------------------------------------------------
// lights
// ---------------------------------------------------------------------------------------------------
template <typename R> FRAY_HD int lightNumSamples(const DLight<R>& l) { return l.type == FRAY_LIGHT_RECT ? l.xSubd * l.ySubd : 1; }
template <typename R> FRAY_HD Col lightEmission(const DLight<R>& l) { return loadCol(l.color) * l.power; } // Light::getColor

// PointLight::getNthSample src/lights.cpp:31-35; RectLight::getNthSample src/lights.cpp:49-77
// the `color` result of getNthSample: it depends on the light and the shaded point only, not on the sample (the wavefront's
// light loop evaluates it once per light)
template <typename R>
FRAY_HD Col lightColorAt(const DLight<R>& l, const V3<R>& shadePos)
{
	if (l.type == FRAY_LIGHT_POINT) return loadCol(l.color) * l.power;
	const V3<R> q = xfUnpoint(l.T, shadePos);
	if (q.y > 0) return Col(0, 0, 0);
	const float cosWeight = (float) (-q.y / length(q));
	return loadCol(l.color) * l.power * l.areaF * cosWeight;
}

template <typename R, typename RNG>
FRAY_HD void lightSample(const DLight<R>& l, RNG& rng, int sampleIdx, const V3<R>& shadePos, V3<R>& samplePos, Col& color, bool wantColor)
{
	if (l.type == FRAY_LIGHT_POINT) {
		samplePos = load3(l.pos);
		if (wantColor) color = loadCol(l.color) * l.power;
		return;
	}
	const R sx = (R) 1 / l.xSubd, sy = (R) 1 / l.ySubd;
	int row;
	if (Num<R>::kExact) row = sampleIdx / l.xSubd;
	else row = (int) (((float) sampleIdx + 0.5f) * (float) sx); // exact for the small integers involved, no integer division
	const int column = sampleIdx - row * l.xSubd;
	const R px = column * sx + sx * (R) rng.randfloat();
	const R py = row * sy + sy * (R) rng.randfloat();
	if (wantColor) color = lightColorAt(l, shadePos);
	samplePos = xfPoint(l.T, V3<R>(px - (R) 0.5, 0, py - (R) 0.5));
}

// ---------------------------------------------------------------------------------------------------
// Whitted shading: local terms (Lambert / Phong), src/shading.cpp:48-80, 101-144
// ---------------------------------------------------------------------------------------------------
template <typename R, int F, typename RNG>
FRAY_HD Col shadeDirectBody(const DScene<R>& sc, const FlatTab& ft, const DShader<R>& s, const V3<R>& rayDir, const Hit<R>& h, int node, RNG& rng, RayCounters& cnt)
{
	Col diffuse = loadCol(s.color);
	if ((F & FRAY_F_TEX) && s.texture >= 0) diffuse = diffuse * sampleTexture(sc, s.texture, rayDir, h.norm, h.u, h.v);
	Col result = diffuse * loadCol(sc.ambient);
	const V3<R> n = faceforward(rayDir, h.norm);
	const V3<R> shadowStart = h.ip + n * Num<R>::offsetEps(maxAbs(h.ip));
	for (int li = 0; li < sc.numLights; li++) {
		const DLight<R>& light = sc.lights[li];
		Col sum(0, 0, 0);
		const int ns = lightNumSamples(light);
		for (int si = 0; si < ns; si++) {
			Col lightCol;
			V3<R> lightPos;
			lightSample(light, rng, si, h.ip, lightPos, lightCol, true);
			const V3<R> toL = lightPos - h.ip;
			const R distSqr = lengthSqr(toL);
			const V3<R> toLight = normalized(toL);
			const float cosAngle = (float) dot(toLight, n);
			float lambertTerm = (float) (cosAngle / distSqr);
			lambertTerm = fmaxf(0.0f, lambertTerm);
			if (visible<R, F>(sc, ft, shadowStart, lightPos, li, cnt, node)) {
				Col c = diffuse * lightCol * lambertTerm;
				if (s.type == FRAY_SHADER_PHONG) {
					const V3<R> r = reflect(-toLight, n);
					const R cosRefl = dot(-rayDir, r);
					if (cosRefl > 0)
						c = c + lightCol / (float) distSqr * loadCol(s.specularColor) * (float) Num<R>::powR(cosRefl, s.exponent) *
						            (float) s.specularMultiplier;
				}
				sum = sum + c;
			}
		}
		result = result + sum / (float) ns;
	}
	return result;
}

// One out-of-line copy for the nested levels of Layered shaders (and for the whole parity unit, where compile time matters
// more than speed). The node's own shader (level 0) gets the body inlined in the fast-precision kernels: behind a call the
// scene and the shared-memory tables are reached through generic pointers (LD + R2UR instead of LDS / LDC), and with 32
// shadow rays per hit in data/boxed.fray this function is the Whitted hot loop.
template <typename R, int F, typename RNG>
FRAY_HD_COLD Col shadeDirect(const DScene<R> sc, const FlatTab ft, const DShader<R>& s, const V3<R>& rayDir, const Hit<R>& h, int node, RNG& rng, RayCounters& cnt)
{
	// `sc` BY VALUE: a reference would force the caller to keep the kernel's parameter block addressable, i.e. in local
	// memory, and every table pointer of the whole kernel would then be loaded from there and dereferenced generically
	return shadeDirectBody<R, F>(sc, ft, s, rayDir, h, node, rng, cnt);
}

// refract(), src/vector.h:184-191; returns false on total internal reflection
template <typename R> FRAY_HD bool refractDir(const V3<R>& i, const V3<R>& n, R ior, V3<R>& out)
{
	const R NdotI = dot(i, n);
	const R k = 1 - (ior * ior) * (1 - NdotI * NdotI);
	if (k < 0) return false;
	out = normalized(i * ior - n * (ior * NdotI + Num<R>::sqrtR(k)));
	return !(out.x == 0 && out.y == 0 && out.z == 0);
}

// orthonormalSystem, src/vector.h:197-213 (c is not normalised)
template <typename R> FRAY_HD void orthonormalSystem(const V3<R>& a, V3<R>& b, V3<R>& c)
{
	V3<R> test(1, 0, 0);
	if (fabs(a.x) > (R) 0.9) test = V3<R>(0, 1, 0);
	b = normalized(cross(a, test));
	c = cross(a, b);
}

// One pending piece of the Whitted ray tree. FRAY_TASK_RAY: radiance(ray) * weight is added to the pixel.
// FRAY_TASK_GLOSSY: the samples k..ns-1 of a glossy reflection (Reflection::shade, src/shading.cpp:176-200) that have not
// been traced yet -- a generator: popping it produces sample k and puts the rest back, so that a hit with 25 glossy samples
// occupies one stack entry instead of 25 (the per-thread stacks live in local memory; what they touch has to fit in L2).
enum { FRAY_TASK_RAY = 0, FRAY_TASK_GLOSSY = 1 };

template <typename R> struct RayTask {
	V3<R> start, dir; // GLOSSY: dir = direction of the INCOMING ray
	Col weight;       // GLOSSY: weight of one sample
	int depth;        // GLOSSY: depth of the reflecting invocation (samples start at depth + 1)
	uint32_t branch;  // RNG stream of the raytrace() invocation this ray starts; GLOSSY: stream of the reflecting invocation
	uint32_t count;   // draws already consumed from that stream; GLOSSY: draws consumed when the reflection was spawned
	int kind;
	int origin;       // node the ray starts on (-1: the camera), see intersectNode
	// GLOSSY only
	V3<R> n;          // face-forwarded normal at the hit
	int shader, k, ns;
	uint32_t k0;      // child number of sample 0 within the reflecting invocation
};

#define FRAY_TASK_STACK 32

template <typename R> struct WhittedState {
	RayTask<R> stack[FRAY_TASK_STACK]; // local memory on the GPU
	int sp;
	int overflow;
	// the root ray of a sample waits here (registers) instead of making a round trip through the stack: most samples of the
	// Whitted scenes are a primary ray and its shadow rays, and a stack entry is 21 words of local memory per lane
	bool rootPending;
	V3<R> rootStart, rootDir;

	FRAY_HD void setRoot(const V3<R>& start, const V3<R>& dir) { rootStart = start; rootDir = dir; rootPending = true; sp = 0; }
	FRAY_HD bool done() const { return sp == 0 && !rootPending; }
};

// Reflection::shade (src/shading.cpp:160-205), Refraction::shade (:238-263), Layered::shade (:357-367), expressed as
// "add weight * local shading now, push weight' * raytrace(child) for later". `spawn` numbers the children of this
// raytrace() invocation in the order the reference would create them (RNG contract).
template <typename R, int LEVEL, int F, typename RNG>
FRAY_HD void shadeWhitted(const DScene<R>& sc, const FlatTab& ft, int shaderIdx, int node, const V3<R>& rayDir, int depth, const Hit<R>& h, const Col& weight,
                          RNG& rng, uint32_t& spawn, WhittedState<R>& ws, Col& accum, RayCounters& cnt)
{
	const DShader<R>& s = sc.shaders[shaderIdx];
	switch (s.type) {
		case FRAY_SHADER_CONST: accum = accum + weight * loadCol(s.color); return; // src/shading.cpp:35-38
		case FRAY_SHADER_LAMBERT:
		case FRAY_SHADER_PHONG:
			if (LEVEL == 0 && !Num<R>::kExact) accum = accum + weight * shadeDirectBody<R, F>(sc, ft, s, rayDir, h, node, rng, cnt);
			else accum = accum + weight * shadeDirect<R, F>(sc, ft, s, rayDir, h, node, rng, cnt);
			return;
		case FRAY_SHADER_REFL: {
			const V3<R> n = faceforward(rayDir, h.norm);
			const V3<R> start = h.ip + n * Num<R>::offsetEps(maxAbs(h.ip));
			const uint32_t drawsAtSpawn = rng.count;
			if (s.pureReflection) {
				RayTask<R> t;
				t.kind = FRAY_TASK_RAY;
				t.origin = node;
				t.start = start;
				t.dir = reflect(rayDir, n);
				t.weight = weight * loadCol(s.mult);
				t.depth = depth + 1;
				t.branch = rngChildBranch(rng.branch, drawsAtSpawn, spawn++);
				t.count = 0;
				if (ws.sp < FRAY_TASK_STACK) ws.stack[ws.sp++] = t; else ws.overflow = 1;
				return;
			}
			const int ns = depth == 0 ? s.numSamples : 3; // LOW_GLOSSY_SAMPLES, src/constants.h:36
			RayTask<R> t;
			t.kind = FRAY_TASK_GLOSSY;
			t.origin = node;
			t.start = start;
			t.dir = rayDir;
			t.weight = weight * loadCol(s.mult) / (float) ns;
			t.depth = depth;
			t.branch = rng.branch;
			t.count = drawsAtSpawn;
			t.n = n;
			t.shader = shaderIdx;
			t.k = 0;
			t.ns = ns;
			t.k0 = spawn;
			spawn += (uint32_t) ns;
			if (ns > 0) { if (ws.sp < FRAY_TASK_STACK) ws.stack[ws.sp++] = t; else ws.overflow = 1; }
			return;
		}
		case FRAY_SHADER_REFR: {
			const V3<R> n = faceforward(rayDir, h.norm);
			const R ior = dot(n, h.norm) > 0 ? 1 / s.ior : s.ior;
			V3<R> refracted;
			if (!refractDir(rayDir, n, ior, refracted)) return; // total internal reflection: black
			RayTask<R> t;
			t.kind = FRAY_TASK_RAY;
			t.origin = node;
			t.start = h.ip - n * Num<R>::offsetEps(maxAbs(h.ip));
			t.dir = refracted;
			t.weight = weight * loadCol(s.mult);
			t.depth = depth + 1;
			t.branch = rngChildBranch(rng.branch, rng.count, spawn++);
			t.count = 0;
			if (ws.sp < FRAY_TASK_STACK) ws.stack[ws.sp++] = t; else ws.overflow = 1;
			return;
		}
		default: { // LAYERED: res = L_i*op_i + (1-op_i)*res bottom-up  ==>  sum_i L_i * op_i * prod_{j>i} (1-op_j)
			if (LEVEL >= 2) return; // Layered inside Layered inside Layered is not supported (rejected at upload)
			const int nl = s.numLayers;
			for (int i = 0; i < nl; i++) {
				Col w = weight;
				for (int j = nl - 1; j >= i; j--) {
					const DLayer& L = sc.layers[s.firstLayer + j];
					const Col op = ((F & FRAY_F_TEX) && L.texture >= 0) ? sampleTexture(sc, L.texture, rayDir, h.norm, h.u, h.v) : loadCol(L.opacity);
					w = w * (j == i ? op : (Col(1, 1, 1) - op));
				}
				shadeWhitted<R, (LEVEL < 2 ? LEVEL + 1 : 2), F>(sc, ft, sc.layers[s.firstLayer + i].shader, node, rayDir, depth, h, w, rng, spawn, ws, accum, cnt);
			}
			return;
		}
	}
}

// raytrace(), src/main.cpp:246-285: one ray of the Whitted tree. Adds weight * (what this invocation returns minus
// what its secondary rays return) to accum and pushes the secondary rays.
template <typename R, int F, typename RNG>
FRAY_HD void whittedStep(const DScene<R>& sc, const FlatTab& ft, const RayTask<R>& task, RNG& rng, WhittedState<R>& ws, Col& accum, RayCounters& cnt)
{
	if (task.depth > sc.maxTraceDepth) return;
	cnt.rays++;
	Ray<R> ray;
	ray.start = task.start;
	ray.dir = task.dir;
	int node, light;
	Hit<R> h;
	closestHit<R, F>(sc, ft, ray, node, light, h, task.origin);
#if defined(FRAY_DEBUG_TRACE) && !defined(__CUDA_ARCH__)
	if (g_frayTrace)
		printf("  ray depth %d start (%.7f %.7f %.7f) dir (%.6f %.6f %.6f) -> node %d light %d dist %.7f ip (%.6f %.6f %.6f) n (%.4f %.4f %.4f) w %.4f\n", task.depth, (double) ray.start.x,
		       (double) ray.start.y, (double) ray.start.z, (double) ray.dir.x, (double) ray.dir.y, (double) ray.dir.z, node, light, (double) h.dist, (double) h.ip.x, (double) h.ip.y, (double) h.ip.z,
		       node >= 0 ? (double) h.norm.x : 0.0, node >= 0 ? (double) h.norm.y : 0.0, node >= 0 ? (double) h.norm.z : 0.0, task.weight.intensity());
#endif
	if (light >= 0) { accum = accum + task.weight * lightEmission(sc.lights[light]); return; }
	if (node < 0) {
		if ((F & FRAY_F_TEX) && sc.hasEnv) accum = accum + task.weight * environmentLookup(sc, ray.dir);
		return;
	}
	const DNode<R>& nd = sc.nodes[node];
	if (F & FRAY_F_TEX) applyBump(sc, nd, h);
	uint32_t spawn = 0;
	shadeWhitted<R, 0, F>(sc, ft, nd.shader, node, ray.dir, task.depth, h, task.weight, rng, spawn, ws, accum, cnt);
}

// Takes the top entry off the ray-task stack and traces it. `primary` is the stream of the pixel sample (branch 0), which
// the primary invocation draws from directly; every other invocation owns the derived stream recorded in its task.
template <typename R, int F, typename RNG>
FRAY_HD void whittedPop(const DScene<R>& sc, const FlatTab& ft, RNG& primary, WhittedState<R>& ws, Col& accum, RayCounters& cnt)
{
	RayTask<R> t;
	if (ws.rootPending) {
		ws.rootPending = false;
		t.start = ws.rootStart; t.dir = ws.rootDir; t.weight = Col(1, 1, 1);
		t.depth = 0; t.branch = 0; t.count = 0; t.kind = FRAY_TASK_RAY; t.origin = -1;
		t.n = V3<R>(0, 0, 0); t.shader = 0; t.k = 0; t.ns = 0; t.k0 = 0;
	} else {
		t = ws.stack[--ws.sp];
	}
	if (t.kind == FRAY_TASK_GLOSSY) {
		// sample t.k of a glossy reflection: Reflection::shade, src/shading.cpp:176-200
		const DShader<R>& s = sc.shaders[t.shader];
		RNG child;
		child.initBranch(primary, rngChildBranch(t.branch, t.count, t.k0 + (uint32_t) t.k));
		V3<R> b, c;
		orthonormalSystem(t.n, b, c);
		V3<R> reflected;
		for (;;) {
			// Random::unitDiscSample, src/random_generator.cpp:71-80
			R sn, cs;
			Num<R>::sincos2pi(Num<R>::draw(child), sn, cs);
			const R rad = Num<R>::sqrtR(Num<R>::draw(child));
			const R x = sn * rad * s.deflectionScaling, y = cs * rad * s.deflectionScaling;
			const V3<R> nn = normalized(t.n + b * x + c * y);
			reflected = reflect(t.dir, nn);
			if (dot(reflected, t.n) > 0) break;
		}
		if (t.k + 1 < t.ns) { // the remaining samples go back under whatever this one spawns
			ws.stack[ws.sp] = t;
			ws.stack[ws.sp].k = t.k + 1;
			ws.sp++;
		}
		RayTask<R> ray;
		ray.kind = FRAY_TASK_RAY;
		ray.origin = t.origin;
		ray.start = t.start;
		ray.dir = reflected;
		ray.weight = t.weight;
		ray.depth = t.depth + 1;
		ray.branch = child.branch;
		ray.count = child.count;
		whittedStep<R, F>(sc, ft, ray, child, ws, accum, cnt);
		return;
	}
	if (t.branch == 0) {
		whittedStep<R, F>(sc, ft, t, primary, ws, accum, cnt);
	} else {
		RNG child;
		child.initBranch(primary, t.branch);
		child.skip(t.count);
		whittedStep<R, F>(sc, ft, t, child, ws, accum, cnt);
	}
}

// ---------------------------------------------------------------------------------------------------
// path tracing, src/main.cpp:92-244
// ---------------------------------------------------------------------------------------------------
#define FRAY_RF_DIFFUSE 2u // src/vector.h:215-219

template <typename R> struct PathState {
	V3<R> start, dir;
	Col mult;   // pathMultiplier
	int depth;
	unsigned flags;
	int origin; // node the segment starts on (-1: the camera), see intersectNode
};

// hemisphereSample(), src/main.cpp:92-116. cos(phi) = 2v-1 and sin(phi) = sqrt(1 - cos^2) replace acos/sin/cos.
template <typename R, typename RNG> FRAY_HD V3<R> hemisphereSample(RNG& rng, const V3<R>& norm)
{
	const R u = Num<R>::draw(rng);
	const R v = Num<R>::draw(rng);
	R st, ct;
	Num<R>::sincos2pi(u, st, ct);
	const R cphi = 2 * v - 1;
	const R sphi = Num<R>::sqrtR(fmax((R) 0, 1 - cphi * cphi));
	const V3<R> d(sphi * ct, cphi, sphi * st);
	return dot(d, norm) > 0 ? d : -d;
}

// the cut-off at the top of pathtrace(), src/main.cpp:173-176
template <typename R> FRAY_HD bool pathAlive(const DScene<R>& sc, const PathState<R>& ps)
{
	return !(ps.depth > sc.maxTraceDepth || ps.mult.intensity() <= 0.01f /* float < 0.01 (double) */);
}

// One iteration of pathtrace(): returns false when the path ended. `accum` receives the terms the reference adds up
// (contribLight of every level and the terminal term); FP32 summation order differs from the recursion's unwinding.
// The cut-off of the NEXT level is evaluated at the end of this one, so that a lane whose path is over learns it in
// the same iteration and never spends a whole trip round the warp loop just to find out.
template <typename R, int F, typename RNG> FRAY_HD bool pathSegment(const DScene<R>& sc, const FlatTab& ft, PathState<R>& ps, RNG& rng, Col& accum, RayCounters& cnt)
{
	if (!pathAlive(sc, ps)) return false; // only the first level can fail here (maxTraceDepth < 0)
	cnt.rays++;
	Ray<R> ray;
	ray.start = ps.start;
	ray.dir = ps.dir;
	int node, light;
	Hit<R> h;
	closestHit<R, F>(sc, ft, ray, node, light, h, (F & FRAY_F_NODES) ? ps.origin : -1);
#if defined(FRAY_DEBUG_TRACE) && !defined(__CUDA_ARCH__)
	printf("  seg depth %d start (%.6f %.6f %.6f) dir (%.6f %.6f %.6f) -> node %d light %d dist %.6f ip (%.5f %.5f %.5f) mult %.5f draws %u\n", ps.depth, (double) ray.start.x, (double) ray.start.y,
	       (double) ray.start.z, (double) ray.dir.x, (double) ray.dir.y, (double) ray.dir.z, node, light, (double) h.dist, (double) h.ip.x, (double) h.ip.y, (double) h.ip.z, ps.mult.intensity(), rng.count);
#endif
	if (light >= 0) {
		if (!(ps.flags & FRAY_RF_DIFFUSE)) {
			if constexpr (!Num<R>::kExact) {
				const float4 em = sc.lightRecs[FRAY_LIGHT_REC_VEC * light + 5];
				accum = accum + Col(em.x, em.y, em.z) * ps.mult;
			} else {
				accum = accum + lightEmission(sc.lights[light]) * ps.mult;
			}
		}
		return false;
	}
	if (node < 0) {
		if ((F & FRAY_F_TEX) && sc.hasEnv) accum = accum + environmentLookup(sc, ray.dir) * ps.mult;
		return false;
	}
	const DNode<R>& nd = sc.nodes[node];
	const DShader<R>& s = sc.shaders[nd.shader];
	if (F & FRAY_F_TEX) applyBump(sc, nd, h);
	const R eps = Num<R>::offsetEps(maxAbs(h.ip));
	// shader type and colour (Lambert: colour, Refl / Refr: multiplier): from the staged FlatInfo when every hit of this kernel
	// variant comes out of the flat table, else from the shader table
	constexpr bool kFlatShade = !Num<R>::kExact && (F & FRAY_F_FLAT) != 0 && (F & (FRAY_F_NODES | FRAY_F_TEX)) == 0;
	int sType;
	if constexpr (kFlatShade) sType = floatBits(ft.info[h.flat].shade.x);
	else sType = s.type;
	auto shaderCol = [&]() {
		if constexpr (kFlatShade) {
			const float4 sh = ft.info[h.flat].shade;
			return Col(sh.y, sh.z, sh.w);
		} else {
			return loadCol((sType == FRAY_SHADER_REFL || sType == FRAY_SHADER_REFR) ? s.mult : s.color);
		}
	};
	const bool lambert = sType == FRAY_SHADER_LAMBERT;

	// all draws of the segment: 2 skipped + 4 for the light sample + 2 for the new direction (Lambert): two Philox blocks; 4 otherwise
	rng.ensure(lambert ? 8 : 4);
	// the first, discarded spawnRay (src/main.cpp:219-224) only advances the stream (its two randdouble() draws for Lambert)
	if (lambert) rng.skip(2);

	// explicitLightSample(), src/main.cpp:118-169
	if constexpr (!Num<R>::kExact) {
		// fast precision: the same steps on the compact light records (scene_image.h) -- identical draws, the sample point as
		// corner + (column + r1) * columnStep + (row + r2) * rowStep instead of through the light's transform
		if (sc.numLights > 0) {
			const int li = rng.randint(0, sc.numLights - 1);
			const float4* rec = sc.lightRecs + FRAY_LIGHT_REC_VEC * li;
			const float4 hdRaw = rec[0];
			const int lType = floatBits(hdRaw.x), xSubd = floatBits(hdRaw.y), nSamples = floatBits(hdRaw.w);
			if (lType == FRAY_LIGHT_RECT) {
				const float4 ca = rec[1];
				const float solidAngle = ca.w / fmaxf(1.0f, lengthSqr(h.ip - V3<R>(ca.x, ca.y, ca.z)));
				if (solidAngle != 0) {
					const int si = rng.randint(0, nSamples - 1);
					const float4 c0 = rec[2], cu = rec[3], cv = rec[4];
					const int row = (int) (((float) si + 0.5f) * c0.w); // exact for the small integers involved, no integer division
					const int column = si - row * xSubd;
					const float a = (float) column + rng.randfloat();
					const float b = (float) row + rng.randfloat();
					const V3<R> onLight(fmaf(b, cv.x, fmaf(a, cu.x, c0.x)), fmaf(b, cv.y, fmaf(a, cu.y, c0.y)), fmaf(b, cv.z, fmaf(a, cu.z, c0.z)));
					Col brdf;
					bool brdfZero;
					if (lambert) {
						const V3<R> wOut = normalized(onLight - h.ip);
						const float cosTerm = fmaxf(0.0f, dot(h.norm, wOut));
						brdf = shaderCol() * Num<R>::overPi(cosTerm);
						brdfZero = brdf.intensity() == 0;
					} else if (sType == FRAY_SHADER_REFL || sType == FRAY_SHADER_REFR) {
						brdfZero = true;
					} else {
						brdf = Col(1, 0, 0);
						brdfZero = false;
					}
					// a zero BRDF makes the shadow ray pointless (the reference tests visibility first; same result)
					if (!brdfZero && visible<R, F>(sc, ft, h.ip + h.norm * eps, onLight, li, cnt, node)) {
						const float4 em = rec[5];
						const float probHit = 1.0f / solidAngle;
						const float probPick = 1.0f / (float) sc.numLights;
						accum = accum + Col(em.x, em.y, em.z) * ps.mult * brdf / (probHit * probPick);
					}
				}
			}
		}
	} else if (sc.numLights > 0) {
		const int li = rng.randint(0, sc.numLights - 1);
		const DLight<R>& L = sc.lights[li];
		if (L.type == FRAY_LIGHT_RECT) { // solidAngle() == 0 for point lights
			const R solidAngle = L.area / fmax((R) 1, lengthSqr(h.ip - load3(L.center)));
			if (solidAngle != 0) {
				const int si = rng.randint(0, lightNumSamples(L) - 1);
				V3<R> onLight;
				Col unused;
				lightSample(L, rng, si, h.ip, onLight, unused, false);
				// brdf first: a zero BRDF makes the shadow ray pointless (the reference tests visibility first; same result)
				Col brdf;
				bool brdfZero;
				const V3<R> wOut = normalized(onLight - h.ip);
				if (lambert) {
					const float cosTerm = (float) fmax((R) 0, dot(h.norm, wOut));
					brdf = shaderCol() * Num<R>::overPi(cosTerm);
					brdfZero = brdf.intensity() == 0;
				} else if (sType == FRAY_SHADER_REFL || sType == FRAY_SHADER_REFR) {
					brdfZero = true;
				} else {
					brdf = Col(1, 0, 0);
					brdfZero = false;
				}
				if (Num<R>::kExact || !brdfZero) {
					const bool vis = visible<R, F>(sc, ft, h.ip + h.norm * eps, onLight, li, cnt);
					if (vis && !brdfZero) {
						const float probHit = (float) (1.0f / solidAngle);
						const float probPick = 1.0f / (float) sc.numLights;
						accum = accum + lightEmission(L) * ps.mult * brdf / (probHit * probPick);
					}
				}
			}
		}
	}

	// the second spawnRay (src/main.cpp:232-236): Lambert src/shading.cpp:88-99, Refl :215-227, Refr :270-299, other src/shading.h:128-134
	Col brdf;
	float pdf;
	if (lambert) {
		ps.start = h.ip + h.norm * eps;
		ps.dir = hemisphereSample(rng, h.norm);
		ps.flags |= FRAY_RF_DIFFUSE;
		const float cosTerm = (float) fmax((R) 0, dot(h.norm, ps.dir));
		brdf = shaderCol() * Num<R>::overPi(cosTerm);
		pdf = 0.15915494309189535f; // (float) (1 / (2 * PI))
	} else if (sType == FRAY_SHADER_REFL) {
		const V3<R> n = faceforward(ray.dir, h.norm);
		ps.start = h.ip + n * eps;
		ps.dir = reflect(ray.dir, h.norm);
		ps.flags &= ~FRAY_RF_DIFFUSE;
		brdf = shaderCol() * 1e9f;
		pdf = 1e9f;
	} else if (sType == FRAY_SHADER_REFR) {
		const V3<R> n = faceforward(ray.dir, h.norm);
		const R ior = dot(n, h.norm) > 0 ? 1 / s.ior : s.ior;
		V3<R> refracted;
		if (refractDir(ray.dir, n, ior, refracted)) {
			ps.start = h.ip - n * eps;
			ps.dir = refracted;
			ps.flags &= ~FRAY_RF_DIFFUSE;
			brdf = shaderCol() * 1e9f;
			pdf = 1e9f;
		} else {
			brdf = Col(0, 0, 0); // total internal reflection: the path goes on with zero throughput and is cut next round
			pdf = 1.0f;
		}
	} else {
		brdf = Col(1, 0, 0); // shaders without a BRDF (Phong, Layered, Const) render red in GI mode
		pdf = 1.0f;
	}
	ps.depth++;
	if (F & FRAY_F_NODES) ps.origin = node;
	ps.mult = ps.mult * brdf / pdf;
	return pathAlive(sc, ps);
}

// ---------------------------------------------------------------------------------------------------
// camera, src/camera.cpp:59-92
// ---------------------------------------------------------------------------------------------------
template <typename R> FRAY_HD Ray<R> screenRay(const DCamera<R>& c, R x, R y, int which)
{
	const V3<R> tl = load3(c.topLeft), tr = load3(c.topRight), bl = load3(c.bottomLeft);
	Ray<R> r;
	if (Num<R>::kExact) r.dir = normalized(tl + (tr - tl) * (x / c.w) + (bl - tl) * (y / c.h)); // src/camera.cpp:62-64, operation for operation
	else r.dir = normalized(tl + load3(c.du) * x + load3(c.dv) * y);
	r.start = load3(c.pos);
	if (which == 1) r.start = r.start + load3(c.right) * -c.stereoSep;
	else if (which == 2) r.start = r.start + load3(c.right) * c.stereoSep;
	return r;
}

template <typename R, typename RNG> FRAY_HD Ray<R> dofRay(const DCamera<R>& c, RNG& rng, R x, R y, int which)
{
	Ray<R> ray = screenRay(c, x, y, which);
	const R M = c.focalDist / dot(load3(c.front), ray.dir);
	const V3<R> T = load3(c.pos) + ray.dir * M;
	R sn, cs;
	Num<R>::sincos2pi(Num<R>::draw(rng), sn, cs);
	const R rad = Num<R>::sqrtR(Num<R>::draw(rng));
	const R u = sn * rad * c.aperture, v = cs * rad * c.aperture;
	ray.start = ray.start + (load3(c.right) * u + load3(c.up) * v);
	ray.dir = normalized(T - ray.start);
	return ray;
}

// getRay, src/main.cpp:296-302. `lens`: the kernel variant was compiled with FRAY_F_LENS (depth of field / stereo possible)
template <typename R, typename RNG> FRAY_HD Ray<R> cameraRay(const DCamera<R>& c, RNG& rng, R x, R y, int which, bool lens)
{
	return (lens && c.dof) ? dofRay(c, rng, x, y, which) : screenRay(c, x, y, lens ? which : 0);
}

FRAY_HD Col adjustSaturation(const Col& c, float amount) // src/color.h:128-134
{
	const float mid = (c.r + c.g + c.b) / 3.0f;
	return Col(mid + (c.r - mid) * amount, mid + (c.g - mid) * amount, mid + (c.b - mid) * amount);
}

// pixel-sample offsets, src/main.cpp:55-61 and :351-357
template <typename RNG> FRAY_HD void sampleOffset(bool randomOffsets, int sampleIdx, RNG& rng, float& ox, float& oy)
{
	if (randomOffsets) {
		ox = rng.randfloat();
		oy = rng.randfloat();
	} else {
		const int i = sampleIdx < 5 ? sampleIdx : 0;
		ox = (i == 1 || i == 4) ? 0.6f : (i == 2 ? 0.3f : 0.0f);
		oy = (i == 3 || i == 4) ? 0.6f : (i == 2 ? 0.3f : 0.0f);
	}
}

// One complete pixel sample, the body of the `for i < samplesPerPixel` loop in RendMT::entry (src/main.cpp:348-359)
// through raytraceSinglePixel (src/main.cpp:304-321). Used by the AOV pass, the test emulator and as the
// straight-line reference for the warp-scheduled kernels in render_kernels.cuh.
template <typename R, int F>
FRAY_HD Col renderSample(const DScene<R>& sc, const FlatTab& ft, uint32_t seed, int px, int py, int width, int sampleIdx, WhittedState<R>* ws, RayCounters& cnt, bool centred = false)
{
	Rng rng;
	rng.init(seed, (uint32_t) (py * width + px), (uint32_t) sampleIdx, 0);
	float ox = 0, oy = 0;
	if (!centred) sampleOffset(sc.cam.dof || sc.gi, sampleIdx, rng, ox, oy); // centred: the prepass shoots through (x, y) itself, src/main.cpp:386
	// x + offsetX is evaluated in float in the reference (int + float), then widened to double
	const R fx = (R) ((float) px + ox), fy = (R) ((float) py + oy);
	const bool stereo = (F & FRAY_F_LENS) && sc.cam.stereoSep > 0;
	const int eyes = stereo ? 2 : 1;
	Ray<R> rays[2];
	for (int e = 0; e < eyes; e++) rays[e] = cameraRay(sc.cam, rng, fx, fy, stereo ? 1 + e : 0, (F & FRAY_F_LENS) != 0);
	Col total(0, 0, 0);
	for (int e = 0; e < eyes; e++) {
		cnt.primary++;
		Col c(0, 0, 0);
		if (sc.gi) {
			PathState<R> ps;
			ps.start = rays[e].start;
			ps.dir = rays[e].dir;
			ps.mult = Col(1, 1, 1);
			ps.depth = 0;
			ps.flags = 0;
			ps.origin = -1;
			while (pathSegment<R, F>(sc, ft, ps, rng, c, cnt)) {}
		} else {
			ws->setRoot(rays[e].start, rays[e].dir);
			while (!ws->done()) whittedPop<R, F>(sc, ft, rng, *ws, c, cnt);
		}
		if (stereo) {
			if (sc.saturation != 1) c = adjustSaturation(c, sc.saturation);
			c = c * (e == 0 ? loadCol(sc.cam.leftMask) : loadCol(sc.cam.rightMask));
		}
		total = total + c;
	}
	return total;
}

} // namespace fray
