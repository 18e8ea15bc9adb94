// wave.cuh -- the Whitted integrator of the fast precision as a WAVEFRONT: the bodies of its three stages.
//
// The reference recurses: raytrace() finds the closest hit, the shader loops over the lights shooting shadow rays through
// visible() and calls raytrace() again for reflections and refractions (/root/reference/src/main.cpp:64-80, 246-285,
// src/shading.cpp:48-144, 160-205, 238-263, 357-367). Round 1 ran that recursion as one megakernel with a per-thread stack of
// ray tasks; on the KD scenes it was bound by instruction fetch and local-memory traffic. Here every level of the
// recursion ("wave") is three lean passes over queues in HBM:
//
//   TRACE   one thread per ray of the wave: closest hit (flat table, analytic nodes, KD meshes) -> a 24-byte hit record.
//           Nothing else: no shading state, a KD short stack in shared memory, ~48 registers.
//   SHADE   one thread per ray: hit attributes, textures, bump, the shader tree. Emits (a) radiance it knows already (ambient,
//           constant, environment, emission), (b) one LIT record per Lambert / Phong evaluation -- the point, the normal, the
//           colours: everything the light loop needs except visibility, (c) the secondary rays, appended to the next wave's
//           queue through warp-aggregated atomics.
//   SHADOW  one thread per lit record: the light loop. The sample points come from the record's counter-based stream (the
//           record carries the stream position of its first draw), every sample is one any-hit traversal, and the Lambert /
//           Phong term is evaluated only if the light is visible. There is NO shadow-ray queue.
//
// Radiance is accumulated per pixel in 64-bit fixed point (2^-32): integer addition is associative, so the frame is
// independent of the order in which threads, warps or GPUs deliver their terms -- bit-identical run to run, and tile or
// sample shards add up exactly.
//
// Every secondary ray owns a derived random stream (rngChildBranch, as in round 1), so no stage depends on the order of the
// queues. The functions are __host__ __device__: tests/emul drives the very same stages with std::vector queues on the CPU.
#pragma once
#include "core.cuh"

namespace fray {

// ---- KD short stack ------------------------------------------------------------------------------------------------------
// The pending far children of the KD walk, newest FRAY_KD_SHORT (8) of them: (node, end of its interval). Its interval starts
// where the leaf that is finished when it is popped ends, so two words per entry suffice. When more are pending than fit, the
// oldest is dropped and remembered as lost; once the stack runs empty the walk restarts at the root with the ray interval cut
// to what has not been visited (kd-restart), and finds the dropped subtrees again. STORE says where the entries live: a
// per-thread column of shared memory on the GPU (no local memory, nothing to spill), a plain array on the host.
// 8 entries: 8 KB of shared memory per CTA instead of 16, i.e. 64 KB more L1 per SM for the KD nodes and triangle records at 8
// resident CTAs, which pays for the few more restarts (forest 4K -1.7 %, dragon -2 %; 4 entries: boxed +9 %)
#ifndef FRAY_KD_SHORT
#define FRAY_KD_SHORT 8
#endif

template <int D> struct KdStoreArray {
	int node[D];
	float tmax[D];
	FRAY_HD void put(unsigned i, int n, float t) { node[i] = n; tmax[i] = t; }
	FRAY_HD void get(unsigned i, int& n, float& t) const { n = node[i]; t = tmax[i]; }
};

template <typename STORE, int D> struct KdShortStack {
	STORE st;
	unsigned sp, base; // entries [base, sp) are valid; base > 0: older ones were dropped
	FRAY_HD void reset() { sp = base = 0; }
	FRAY_HD void push(int node, float tmax)
	{
		st.put(sp & (unsigned) (D - 1), node, tmax);
		sp++;
		if (sp - base > (unsigned) D) base = sp - (unsigned) D;
	}
	FRAY_HD bool pop(int& node, float& tmax)
	{
		if (sp == base) return false;
		sp--;
		st.get(sp & (unsigned) (D - 1), node, tmax);
		return true;
	}
	FRAY_HD bool lost() const { return base > 0; }
};

// Mesh::intersect + Mesh::intersectKD (src/mesh.cpp:144-165, 357-394) in the fast precision: the interval walk of
// intersectMeshFast (core.cuh) -- same node and triangle records, same near-child and leaf-acceptance rules, so the same hit
// ids -- over a short stack. Object space; returns the ray parameter, the absolute triangle index and the barycentrics.
template <bool ANYHIT, typename STK>
FRAY_HD_HOT bool kdWalk(const DScene<float>& sc, const DMesh<float>& m, const Ray<float>& ray, float maxT, STK& stk, float& tHit, int& triHit, float& l2Hit, float& l3Hit)
{
	const float ox = ray.start.x, oy = ray.start.y, oz = ray.start.z, dx = ray.dir.x, dy = ray.dir.y, dz = ray.dir.z;
	// RRay::prepareForTracing, src/bbox.h:49-54
	const float rx = fabsf(dx) > 1e-12f ? 1.0f / dx : 1e12f, ry = fabsf(dy) > 1e-12f ? 1.0f / dy : 1e12f, rz = fabsf(dz) > 1e-12f ? 1.0f / dz : 1e12f;
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
	const float rootTmax = tmax;
	const bool cull = (m.flags & FRAY_MESH_BACKFACE_CULL) != 0;
	float best = maxT;
	int bestTri = -1;
	float bestL2 = 0, bestL3 = 0;

	// Triangle::intersectFast, src/triangle.cpp:66-94, on a 48-byte plane record; `t` is what the caller wants back on a hit
	// All three vectors of the record are fetched before anything is tested (one load latency per triangle instead of three in
	// a chain) and the tests are folded into one minimum, as flatTest does: the hit counts iff
	//     min(l2, l3, 1 - (l2 + l3), t, [culling: -s]) >= 0   and   t <= best.
	// Same operations on the same operands as the early-out form, and the folded tests decide exactly as the separate ones:
	// 1 - x is exact for x in [1/2, 2] and keeps the sign outside, fminf passes -0 as "not negative" just as `< 0` does, and a
	// NaN or infinite t (a ray in the triangle's plane) fails `t <= best`.
	const float cullSign = cull ? -1.0f : 0.0f;
	auto testTriangle = [&](const float4* rec, int t) {
		const float4 pl = rec[0], e2 = rec[1], e3 = rec[2];
		const float s = fmaf(pl.x, dx, fmaf(pl.y, dy, pl.z * dz));
		const float hh = fmaf(-pl.x, ox, fmaf(-pl.y, oy, fmaf(-pl.z, oz, pl.w)));
		const float tt = flatDivide(hh, s);
		if constexpr (ANYHIT) {
			// most triangles an occlusion walk meets are missed by every lane: the plane test alone decides that (the two edge
			// vectors are in flight already)
			if (!((fminf(tt, cullSign * s) >= 0.0f) & (tt <= best))) return;
		}
		const float px = fmaf(dx, tt, ox), py = fmaf(dy, tt, oy), pz = fmaf(dz, tt, oz);
		const float l2 = fmaf(e2.x, px, fmaf(e2.y, py, fmaf(e2.z, pz, e2.w)));
		const float l3 = fmaf(e3.x, px, fmaf(e3.y, py, fmaf(e3.z, pz, e3.w)));
		const float inside = fminf(fminf(l2, l3), 1.0f - (l2 + l3));
		const float margin = fminf(fminf(inside, tt), cullSign * s);
		const bool hit = (margin >= 0.0f) & (tt <= best);
		best = hit ? tt : best;
		bestTri = hit ? t : bestTri;
		bestL2 = hit ? l2 : bestL2;
		bestL3 = hit ? l3 : bestL3;
	};

	bool found = false;
	if (m.kdRoot < 0) {
		for (int t = m.firstTri; t < m.firstTri + m.numTris; t++) {
			testTriangle(sc.kdTris + 3 * (size_t) t, t);
			if (ANYHIT && bestTri >= 0) break;
		}
		found = bestTri >= 0;
	} else {
		stk.reset();
		int ni = m.kdRoot;
		bool strict = false; // the descent that follows a restart: a split AT the interval's start belongs to what was visited
		const int4* nodes = reinterpret_cast<const int4*>(sc.kd);
		for (;;) {
			int4 n = nodes[ni];
			while (n.x != 3) { // inner node: axis n.x, children n.y / n.y + 1, split position in n.w
				const float split = intBitsToFloat(n.w);
				const float o = n.x == 0 ? ox : (n.x == 1 ? oy : oz), r = n.x == 0 ? rx : (n.x == 1 ? ry : rz);
				const float ts = (split - o) * r;
				const bool lowFirst = o < split || (o == split && r <= 0.0f); // near child = the origin's side (src/mesh.cpp:368)
				const int nearChild = n.y + (lowFirst ? 0 : 1), farChild = n.y + (lowFirst ? 1 : 0);
				if (ts > tmax || ts <= 0.0f) {
					ni = nearChild;
				} else if (ts < tmin || (strict && ts <= tmin)) {
					ni = farChild;
				} else {
					stk.push(farChild, tmax);
					ni = nearChild;
					tmax = ts;
				}
				n = nodes[ni];
			}
			strict = false;
			// leaf: its records are consecutive in kdLeafTris; a hit remembers the reference, the triangle id is looked up at the end
			// ... if the ray touches the box of the leaf's triangles within [0, best] at all (DScene::kdLeafBox, padded: "no" is safe)
			bool touches = n.z > 0;
#if !defined(FRAY_KD_NO_LEAFBOX)
			if (touches) {
				const float4 lo = sc.kdLeafBox[2 * (size_t) n.w], hi = sc.kdLeafBox[2 * (size_t) n.w + 1];
				const float ax0 = (lo.x - ox) * rx, ax1 = (hi.x - ox) * rx;
				const float ay0 = (lo.y - oy) * ry, ay1 = (hi.y - oy) * ry;
				const float az0 = (lo.z - oz) * rz, az1 = (hi.z - oz) * rz;
				const float tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
				const float tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), best));
				touches = tn <= tf;
			}
#endif
			if (touches)
				for (int i = 0; i < n.z; i++) testTriangle(sc.kdLeafTris + 3 * (size_t) (n.y + i), n.y + i);
			if (bestTri >= 0) {
				if (ANYHIT) { found = true; break; }
				// a hit inside this leaf's interval is the closest one: everything still pending starts farther away
				if (best <= tmax + slack + 1e-6f * tmax) { found = true; break; }
			}
			// what is pending starts where this leaf ended
			tmin = tmax;
			if (tmin > best) break;
			if (stk.pop(ni, tmax)) continue;
			if (!stk.lost() || tmin >= rootTmax) break;
			stk.reset(); // kd-restart
			ni = m.kdRoot;
			tmax = rootTmax;
			strict = true;
		}
		if (!found && bestTri >= 0) found = true; // see intersectMeshFast
	}
	if (!found) return false;
	tHit = best;
	triHit = m.kdRoot < 0 ? bestTri : m.firstTri + sc.leafRefs[bestTri]; // KD walk: bestTri is the index of the leaf reference
	l2Hit = bestL2;
	l3Hit = bestL3;
	return true;
}

// ---- records ---------------------------------------------------------------------------------------------------------------
// What TRACE leaves for SHADE. node >= 0: node index; -1: nothing hit; <= -2: light -2 - node.
struct WaveHit {
	float t;      // world distance
	float l2, l3; // barycentrics (KD / brute-force mesh hits of the node loop)
	int node;
	int tri;      // absolute triangle index or -1
	int flat;     // FlatInfo index of a flat-table hit or -1
};

// A ray of the Whitted tree as SHADE sees it.
struct WaveRay {
	V3<float> start, dir;
	Col weight;       // product of the reflection / refraction / layer factors down to this ray
	int pixel;        // y * width + x: the accumulator it feeds and the `pixel` word of its random streams
	int sample;
	int depth;        // Ray::depth, src/vector.h:222-230
	int origin;       // node the ray starts on, -1 for camera rays (core.cuh, intersectNode)
	int eye;          // 0: mono; 1 / 2: left / right eye of a stereo pair (saturation + mask at accumulation, src/main.cpp:304-321)
	uint32_t branch;  // random stream of the raytrace() invocation this ray starts
	uint32_t count;   // draws already consumed from it
};

// One Lambert / Phong evaluation waiting for its light loop (src/shading.cpp:48-80, 101-144).
struct WaveLit {
	V3<float> ip, n;    // hit point, face-forwarded shading normal
	V3<float> rayDir;   // Phong only
	Col diffuse;        // weight * shader colour (* texture)
	Col specular;       // Phong: weight * specularColor * specularMultiplier
	float exponent;     // Phong
	int phong;
	int pixel, sample, origin, eye;
	uint32_t branch, count; // stream and position of the first light-sample draw
};

#define FRAY_FIX_ONE 4294967296.0f // 2^32: accumulator units per unit of radiance

FRAY_HD long long waveFixed(float c)
{
	// |c| beyond 2^30 cannot come out of a renderable scene; the clamp only keeps the conversion defined
	const float v = fminf(fmaxf(c, -1.0e9f), 1.0e9f) * FRAY_FIX_ONE;
#if defined(__CUDA_ARCH__)
	return __float2ll_rn(v);
#else
	return (long long) llrintf(v);
#endif
}

// the per-eye colour operator of raytraceSinglePixel (src/main.cpp:309-316): linear, so it is applied term by term
FRAY_HD Col waveEyeColor(const DScene<float>& sc, int eye, Col c)
{
	if (eye == 0) return c;
	if (sc.saturation != 1) c = adjustSaturation(c, sc.saturation);
	return c * (eye == 1 ? loadCol(sc.cam.leftMask) : loadCol(sc.cam.rightMask));
}

// ---- TRACE -----------------------------------------------------------------------------------------------------------------
// object-space copy of a world ray for node nd; scale = object-space ray parameter per unit of world distance
FRAY_HD void waveLocalRay(const DNode<float>& nd, const Ray<float>& ray, Ray<float>& local, float& scale)
{
	if (nd.T.identity) {
		local = ray;
		scale = 1;
		return;
	}
	local.start = mulRow(ray.start - load3(nd.T.off), nd.T.inv);
	const V3<float> d = mulRow(ray.dir, nd.T.inv);
	const float l2 = lengthSqr(d);
	const float rl = Num<float>::rcpLen(l2);
	local.dir = d * rl;
	scale = l2 * rl;
}

// does the world ray touch the box within [0, tMax] ? (DScene::nodeBox: padded, so "no" is safe)
FRAY_HD bool waveBoxHit(const float4 lo, const float4 hi, const Ray<float>& ray, const V3<float>& inv, float tMax)
{
	const float ax0 = (lo.x - ray.start.x) * inv.x, ax1 = (hi.x - ray.start.x) * inv.x;
	const float ay0 = (lo.y - ray.start.y) * inv.y, ay1 = (hi.y - ray.start.y) * inv.y;
	const float az0 = (lo.z - ray.start.z) * inv.z, az1 = (hi.z - ray.start.z) * inv.z;
	const float tn = fmaxf(fmaxf(fminf(ax0, ax1), fminf(ay0, ay1)), fmaxf(fminf(az0, az1), 0.0f));
	const float tf = fminf(fminf(fmaxf(ax0, ax1), fmaxf(ay0, ay1)), fminf(fmaxf(az0, az1), tMax));
	return tn <= tf;
}

// The node loop of raytrace() / visible() (src/main.cpp:64-80, 250-258) with Node::intersect (src/geometry.cpp:196-208): nodes
// whose padded world box (DScene::nodeBox) the ray misses are skipped before their ray transform; meshes go through kdWalk.
// CLOSEST: wh holds the best hit so far (from the flat table) and is improved. ANY: true as soon as something lies within
// maxDist (wh is not touched).
// (Measured alternative: the node loop folded INTO the walk as one while-while state machine, so that lanes walking different
// nodes share instructions. With the 3-7 nodes of the bundled scenes it was slower -- boxed shadow pass 2.99 -> 3.20 ms: every
// trip then pays for the node set-up block with one or two lanes active, while here all lanes set a node up together.)
template <bool ANYHIT, int F, typename STK>
FRAY_HD_HOT bool waveNodes(const DScene<float>& sc, const Ray<float>& ray, float maxDist, int origin, STK& stk, WaveHit& wh)
{
	// RRay::prepareForTracing, src/bbox.h:49-54, for the world ray
	const V3<float> inv(fabsf(ray.dir.x) > 1e-12f ? 1.0f / ray.dir.x : 1e12f, fabsf(ray.dir.y) > 1e-12f ? 1.0f / ray.dir.y : 1e12f,
	                    fabsf(ray.dir.z) > 1e-12f ? 1.0f / ray.dir.z : 1e12f);
	// nodes that live in the flat table are stepped over: the upper corner of a node's box carries the index of the next node
	// that does not (scene_image.h)
	for (int n = 0, nextNode; n < sc.numNodes; n = nextNode) {
		const DNode<float>& nd = sc.nodes[n];
		const float4 boxLo = sc.nodeBox[2 * n], boxHi = sc.nodeBox[2 * n + 1];
		nextNode = (F & FRAY_F_FLAT) ? floatBits(boxHi.w) : n + 1;
		if ((F & FRAY_F_FLAT) && n == 0 && nd.inFlat) continue;
		if (!waveBoxHit(boxLo, boxHi, ray, inv, ANYHIT ? maxDist : wh.t)) continue;
		Ray<float> local;
		float scale;
		waveLocalRay(nd, ray, local, scale);
		const DGeom<float>& g = sc.geoms[nd.geom];
		// hits farther than the best so far may be dropped (the reference compares afterwards, src/main.cpp:256)
		const float maxT = ANYHIT ? maxDist * scale : (wh.t < Num<float>::big() / 4 ? wh.t * scale * 1.0001f : Num<float>::big());
		float tObj, l2 = 0, l3 = 0;
		int tri = -1;
		V3<float> ipObj;
		if (g.type == FRAY_GEOM_MESH) {
			if (!kdWalk<ANYHIT>(sc, sc.meshes[g.mesh], local, maxT, stk, tObj, tri, l2, l3)) continue;
			ipObj = local.start + local.dir * tObj;
		} else {
			Hit<float> h;
			if (!intersectAnalytic(sc, nd.geom, local, h, false, n == origin)) continue;
			ipObj = h.ip;
		}
		const float tw = dist3(ray.start, xfPoint(nd.T, ipObj)); // info.dist of Node::intersect, src/geometry.cpp:203
		if (ANYHIT) {
			if (tw < maxDist) return true;
		} else if (tw < wh.t) {
			wh.t = tw;
			wh.node = n;
			wh.tri = tri;
			wh.l2 = l2;
			wh.l3 = l3;
			wh.flat = -1;
		}
	}
	return false;
}

// the closest-hit loops of raytrace(), src/main.cpp:250-271: ids and distance only, attributes are SHADE's business
template <int F, typename STK>
FRAY_HD_HOT void waveClosest(const DScene<float>& sc, const FlatTab& ft, const Ray<float>& ray, int origin, STK& stk, WaveHit& wh)
{
	wh.t = Num<float>::big();
	wh.l2 = wh.l3 = 0;
	wh.node = -1;
	wh.tri = -1;
	wh.flat = -1;
	if constexpr ((F & FRAY_F_FLAT) != 0) {
		int idx = -1;
		flatClosest(ft.polys, sc.numFlatAll, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, wh.t, idx);
		if (F & FRAY_F_HEX) flatHexClosest(ft.hexes, sc.numFlatHex, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, wh.t, idx);
		if (F & FRAY_F_SPHERES) flatSpheresClosest(ft.spheres, sc.numFlatSpheres, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, wh.t, idx, sc.numFlatAll);
		if (F & FRAY_F_TWOSIDED) flatClosest2(ft.polys2, sc.numFlat2, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, wh.t, idx, sc.flat2InfoBase);
		if (idx >= 0) {
			const FlatInfo& fi = ft.info[idx];
			wh.flat = idx;
			wh.node = (fi.flags & FRAY_FLAT_LIGHT) ? -2 - fi.node : fi.node;
		}
	}
	waveNodes<false, F>(sc, ray, 0.0f, origin, stk, wh);
	if (!(F & FRAY_F_FLAT) || !sc.lightsInFlat) {
		for (int l = 0; l < sc.numLights; l++) {
			float d;
			if (intersectLight(sc.lights[l], ray, d) && d < wh.t) {
				wh.t = d;
				wh.node = -2 - l;
				wh.tri = -1;
				wh.flat = -1;
			}
		}
	}
}

// visible(), src/main.cpp:64-80 (cf. visible() in core.cuh)
// `flatMask`: the records of the light's shadow set the start point can hit at all (waveShadowSet), or FRAY_FLAT_UNMASKED
#define FRAY_FLAT_UNMASKED 0xffffffffu // (sets of up to 31 records are masked)
template <int F, typename STK>
FRAY_HD_HOT bool waveVisible(const DScene<float>& sc, const FlatTab& ft, const V3<float>& a, const V3<float>& b, int light, int origin, STK& stk, unsigned flatMask = FRAY_FLAT_UNMASKED)
{
	Ray<float> ray;
	ray.dir = b - a;
	ray.start = a;
	const float maxDist = length(ray.dir);
	ray.dir = normalized(ray.dir);
	if constexpr ((F & FRAY_F_FLAT) != 0) {
		int first = 0, count = sc.numFlatGeom;
		if (light >= 0 && light < FRAY_SHADOW_LIGHTS && sc.shadowCount[light] >= 0) {
			first = sc.shadowFirst[light];
			count = sc.shadowCount[light];
		}
		if (flatMask != FRAY_FLAT_UNMASKED) {
			if (flatAnyMasked(ft.polys + FRAY_FLAT_POLY_VEC * first, flatMask, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		} else if (flatAny(ft.polys + FRAY_FLAT_POLY_VEC * first, count, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		if (F & FRAY_F_HEX) {
			const unsigned hexMask = (light >= 0 && light < FRAY_SHADOW_LIGHTS) ? sc.shadowHex[light] : 0xffffffffu;
			if (flatHexAny(ft.hexes, sc.numFlatHex, hexMask, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		}
		if ((F & FRAY_F_SPHERES) && flatSpheresAny(ft.spheres, sc.numFlatSpheres, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
		if ((F & FRAY_F_TWOSIDED) && flatAny2(ft.polys2, sc.numFlat2, ray.start.x, ray.start.y, ray.start.z, ray.dir.x, ray.dir.y, ray.dir.z, maxDist)) return false;
	}
	WaveHit unused;
	return !waveNodes<true, F>(sc, ray, maxDist, origin, stk, unused);
}

// ---- SHADE -----------------------------------------------------------------------------------------------------------------
// The hit of a ray as the shaders want it (IntersectionInfo, src/geometry.h:33-39) from the ids TRACE left.
template <int F>
FRAY_HD void waveHitAttributes(const DScene<float>& sc, const FlatTab& ft, const WaveRay& r, const WaveHit& wh, Hit<float>& h)
{
	Ray<float> ray;
	ray.start = r.start;
	ray.dir = r.dir;
	h.dist = wh.t;
	h.tri = -1;
	h.mesh = -1;
	h.flat = -1;
	h.u = h.v = 0;
	h.l2 = h.l3 = 0;
	if (wh.flat >= 0) {
		if constexpr ((F & FRAY_F_FLAT) != 0) {
			int node = -1, light = -1;
			flatResolve<float, F>(sc, ft, wh.flat, ray, node, light, h);
		}
		return;
	}
	const DNode<float>& nd = sc.nodes[wh.node];
	const DGeom<float>& g = sc.geoms[nd.geom];
	if (g.type == FRAY_GEOM_MESH) {
		h.ip = ray.start + ray.dir * wh.t;
		h.tri = wh.tri;
		h.mesh = g.mesh;
		h.l2 = wh.l2;
		h.l3 = wh.l3;
		triangleAttributes(sc, g.mesh, wh.tri, wh.l2, wh.l3, h.norm, h.u, h.v);
		h.norm = xfDir(nd.T, h.norm);
		return;
	}
	// analytic primitives: the intersection once more, now with normal and (u, v) (Node::intersect, src/geometry.cpp:196-208)
	Ray<float> local;
	float scale;
	waveLocalRay(nd, ray, local, scale);
	if (!intersectAnalytic(sc, nd.geom, local, h, nd.needsUV != 0, wh.node == r.origin)) {
		h.ip = local.start + local.dir * (wh.t * scale); // cannot happen (TRACE just found it); keep the numbers finite
		h.norm = V3<float>(0, 1, 0);
	}
	h.ip = xfPoint(nd.T, h.ip);
	h.norm = xfDir(nd.T, h.norm);
	h.dist = wh.t;
}

// Reflection::shade (src/shading.cpp:160-205), Refraction::shade (:238-263), Layered::shade (:357-367), Lambert / Phong
// (:48-80, 101-144) and ConstantShader (:35-38) for one hit: what is known now goes to sink.add(), the light loops to
// sink.lit(), the secondary rays to sink.ray(). `count` / `spawn`: draws consumed from / children spawned by this raytrace()
// invocation so far, in the reference's program order (RNG contract).
template <int LEVEL, int F, typename SINK>
FRAY_HD void waveShadeLevel(const DScene<float>& sc, int shaderIdx, int node, const WaveRay& r, const Hit<float>& h, const Col& weight, const uint32_t* keys, uint32_t seed,
                            uint32_t& count, uint32_t& spawn, SINK& sink)
{
	const DShader<float>& s = sc.shaders[shaderIdx];
	switch (s.type) {
		case FRAY_SHADER_CONST: sink.add(r, weight * loadCol(s.color)); return;
		case FRAY_SHADER_LAMBERT:
		case FRAY_SHADER_PHONG: {
			Col diffuse = loadCol(s.color);
			if ((F & FRAY_F_TEX) && s.texture >= 0) diffuse = diffuse * sampleTexture(sc, s.texture, r.dir, h.norm, h.u, h.v);
			sink.add(r, weight * (diffuse * loadCol(sc.ambient)));
			if (sc.numLights > 0) {
				WaveLit L;
				L.ip = h.ip;
				L.n = faceforward(r.dir, h.norm);
				L.rayDir = r.dir;
				L.diffuse = weight * diffuse;
				L.phong = s.type == FRAY_SHADER_PHONG;
				L.specular = weight * (loadCol(s.specularColor) * s.specularMultiplier);
				L.exponent = s.exponent;
				L.pixel = r.pixel; L.sample = r.sample; L.origin = node; L.eye = r.eye;
				L.branch = r.branch; L.count = count;
				sink.lit(L);
				count += (uint32_t) sc.lightDraws; // two draws per RectLight sample, none for a PointLight (src/lights.cpp:31-77)
			}
			return;
		}
		case FRAY_SHADER_REFL: {
			if (r.depth + 1 > sc.maxTraceDepth) { // the child would return black at once (src/main.cpp:248); its stream ids are still consumed
				spawn += s.pureReflection ? 1u : (uint32_t) (r.depth == 0 ? s.numSamples : 3);
				return;
			}
			const V3<float> n = faceforward(r.dir, h.norm);
			WaveRay c;
			c.start = h.ip + n * Num<float>::offsetEps(maxAbs(h.ip));
			c.pixel = r.pixel; c.sample = r.sample; c.depth = r.depth + 1; c.origin = node; c.eye = r.eye;
			if (s.pureReflection) {
				c.dir = reflect(r.dir, n);
				c.weight = weight * loadCol(s.mult);
				c.branch = rngChildBranch(r.branch, count, spawn++);
				c.count = 0;
				sink.ray(c);
				return;
			}
			const int ns = r.depth == 0 ? s.numSamples : 3; // LOW_GLOSSY_SAMPLES, src/constants.h:36
			c.weight = weight * loadCol(s.mult) / (float) ns;
			V3<float> b, cc;
			orthonormalSystem(n, b, cc);
			for (int k = 0; k < ns; k++) { // Reflection::shade, src/shading.cpp:176-200
				RngT<FRAY_RNG_KEYED> child;
				child.keys = keys;
				child.init(seed, (uint32_t) r.pixel, (uint32_t) r.sample, rngChildBranch(r.branch, count, spawn++));
				for (;;) { // Random::unitDiscSample, src/random_generator.cpp:71-80
					float sn, cs;
					Num<float>::sincos2pi(Num<float>::draw(child), sn, cs);
					const float rad = sqrtf(Num<float>::draw(child));
					const float x = sn * rad * s.deflectionScaling, y = cs * rad * s.deflectionScaling;
					const V3<float> nn = normalized(n + b * x + cc * y);
					c.dir = reflect(r.dir, nn);
					if (dot(c.dir, n) > 0) break;
				}
				c.branch = child.branch;
				c.count = child.count;
				sink.ray(c);
			}
			return;
		}
		case FRAY_SHADER_REFR: {
			const V3<float> n = faceforward(r.dir, h.norm);
			const float ior = dot(n, h.norm) > 0 ? 1 / s.ior : s.ior;
			WaveRay c;
			if (!refractDir(r.dir, n, ior, c.dir)) return; // total internal reflection: black
			const uint32_t id = spawn++;
			if (r.depth + 1 > sc.maxTraceDepth) return;
			c.start = h.ip - n * Num<float>::offsetEps(maxAbs(h.ip));
			c.weight = weight * loadCol(s.mult);
			c.pixel = r.pixel; c.sample = r.sample; c.depth = r.depth + 1; c.origin = node; c.eye = r.eye;
			c.branch = rngChildBranch(r.branch, count, id);
			c.count = 0;
			sink.ray(c);
			return;
		}
		default: { // LAYERED: res = L_i*op_i + (1-op_i)*res bottom-up  ==>  sum_i L_i * op_i * prod_{j>i} (1-op_j)
			if (LEVEL >= 2) return;
			const int nl = s.numLayers;
			for (int i = 0; i < nl; i++) {
				Col w = weight;
				for (int j = nl - 1; j >= i; j--) {
					const DLayer& L = sc.layers[s.firstLayer + j];
					const Col op = ((F & FRAY_F_TEX) && L.texture >= 0) ? sampleTexture(sc, L.texture, r.dir, h.norm, h.u, h.v) : loadCol(L.opacity);
					w = w * (j == i ? op : (Col(1, 1, 1) - op));
				}
				waveShadeLevel<(LEVEL < 2 ? LEVEL + 1 : 2), F>(sc, sc.layers[s.firstLayer + i].shader, node, r, h, w, keys, seed, count, spawn, sink);
			}
			return;
		}
	}
}

// raytrace() after its closest-hit loops, src/main.cpp:272-284
template <int F, typename SINK>
FRAY_HD void waveShade(const DScene<float>& sc, const FlatTab& ft, const WaveRay& r, const WaveHit& wh, const uint32_t* keys, uint32_t seed, uint32_t& count, SINK& sink)
{
	if (wh.node <= -2) {
		sink.add(r, r.weight * lightEmission(sc.lights[-2 - wh.node]));
		return;
	}
	if (wh.node < 0) {
		if ((F & FRAY_F_TEX) && sc.hasEnv) sink.add(r, r.weight * environmentLookup(sc, r.dir));
		return;
	}
	Hit<float> h;
	waveHitAttributes<F>(sc, ft, r, wh, h);
	const DNode<float>& nd = sc.nodes[wh.node];
	if (F & FRAY_F_TEX) applyBump(sc, nd, h);
	uint32_t spawn = 0;
	waveShadeLevel<0, F>(sc, nd.shader, wh.node, r, h, r.weight, keys, seed, count, spawn, sink);
}

// ---- SHADOW ----------------------------------------------------------------------------------------------------------------
// The light loop of Lambert::shade / Phong::shade (src/shading.cpp:55-78, 108-141) for one lit record: every light, every
// sample of it in order -- the sample point from the record's random stream (RectLight::getNthSample draws two floats per
// sample, src/lights.cpp:49-77; the stream position of the record's first draw is L.count), the shadow ray, and the Lambert /
// Phong term when the light is visible. One thread owns a record, so the lanes of a warp -- 32 neighbouring hit points -- shoot
// at the SAME light sample at the same time and walk the same parts of the scene (32 samples of one point on 32 lanes, which
// was the first version, kept ~10 lanes busy in the KD walks). Returns the sum the shader adds to its ambient term.
template <int F, typename STK>
FRAY_HD_HOT Col waveLightLoop(const DScene<float>& sc, const FlatTab& ft, const WaveLit& L, const uint32_t* keys, uint32_t seed, STK& stk, unsigned& shadowRays)
{
	RngT<FRAY_RNG_KEYED> rng;
	rng.keys = keys;
	rng.init(seed, (uint32_t) L.pixel, (uint32_t) L.sample, L.branch);
	rng.skip(L.count);
	const V3<float> shadowStart = L.ip + L.n * Num<float>::offsetEps(maxAbs(L.ip));
	Col result(0, 0, 0);
	for (int li = 0; li < sc.numLights; li++) {
		const DLight<float>& light = sc.lights[li];
		const int ns = lightNumSamples(light);
		const Col lightCol = lightColorAt(light, L.ip); // the same for every sample of the light
		// every shadow ray of this loop starts at shadowStart: the records of the light's shadow set it lies behind are out
		unsigned flatMask = FRAY_FLAT_UNMASKED;
		if constexpr ((F & FRAY_F_FLAT) != 0) {
			int first = 0, count = sc.numFlatGeom;
			if (li < FRAY_SHADOW_LIGHTS && sc.shadowCount[li] >= 0) {
				first = sc.shadowFirst[li];
				count = sc.shadowCount[li];
			}
			if (count <= 31) flatMask = flatOriginMask(ft.polys + FRAY_FLAT_POLY_VEC * first, count, shadowStart.x, shadowStart.y, shadowStart.z);
		}
		Col sum(0, 0, 0);
		for (int si = 0; si < ns; si++) {
			Col unused;
			V3<float> lightPos;
			lightSample(light, rng, si, L.ip, lightPos, unused, false);
			shadowRays++;
			if (!waveVisible<F>(sc, ft, shadowStart, lightPos, li, L.origin, stk, flatMask)) continue;
			const V3<float> toL = lightPos - L.ip;
			const float distSqr = lengthSqr(toL);
			const V3<float> toLight = normalized(toL);
			const float cosAngle = dot(toLight, L.n);
			const float lambertTerm = fmaxf(0.0f, cosAngle / distSqr);
			Col c = L.diffuse * lightCol * lambertTerm;
			if (L.phong) {
				const V3<float> rr = reflect(-toLight, L.n);
				const float cosRefl = dot(-L.rayDir, rr);
				if (cosRefl > 0) c = c + lightCol / distSqr * L.specular * Num<float>::powR(cosRefl, L.exponent);
			}
			sum = sum + c;
		}
		result = result + sum / (float) ns;
	}
	return result;
}

// ---- the camera ray of primary sample (pixel, s) and the state its raytrace() starts in (src/main.cpp:296-321, 348-359) ---
template <int F>
FRAY_HD void wavePrimary(const DScene<float>& sc, const uint32_t* keys, uint32_t seed, int px, int py, int width, int s, bool randomOffsets, WaveRay& left, WaveRay& right, bool& stereo)
{
	RngT<FRAY_RNG_KEYED> rng;
	rng.keys = keys;
	rng.init(seed, (uint32_t) (py * width + px), (uint32_t) s, 0);
	float ox, oy;
	sampleOffset(randomOffsets, s, rng, ox, oy);
	const float fx = (float) px + ox, fy = (float) py + oy;
	stereo = (F & FRAY_F_LENS) && sc.cam.stereoSep > 0;
	const Ray<float> first = cameraRay(sc.cam, rng, fx, fy, stereo ? 1 : 0, (F & FRAY_F_LENS) != 0);
	left.start = first.start; left.dir = first.dir;
	left.weight = Col(1, 1, 1);
	left.pixel = py * width + px; left.sample = s; left.depth = 0; left.origin = -1; left.eye = stereo ? 1 : 0;
	left.branch = 0;
	if (stereo) {
		const Ray<float> second = cameraRay(sc.cam, rng, fx, fy, 2, true);
		right = left;
		right.start = second.start; right.dir = second.dir;
		right.eye = 2;
	}
	left.count = rng.count; // the right eye continues where the left eye's light loops stop: SHADE of the left eye queues it
}

} // namespace fray
