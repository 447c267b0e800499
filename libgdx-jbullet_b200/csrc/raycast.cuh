// raycast.cuh — batched closest-hit ray tests against the bodies of the world.
//
// Replaces, per ray: CollisionWorld.rayTest with a ClosestRayResultCallback (disp/CollisionWorld.java:553-590, 697-729):
//   objects in world order, needsCollision filter (:664-670), AabbUtil2.rayAabb (lm/AabbUtil2.java:40-108) with the running
//   closestHitFraction as the exit bound, rayTestSingle's convex branch (:260-300) = SubsimplexConvexCast of a zero-radius,
//   zero-margin sphere (np/SubsimplexConvexCast.java:63-190) on the Voronoi simplex solver of gjk.cuh.
// One block per ray.  The reference loop is sequential (every accepted hit lowers the bound the next rayAabb sees), but
// an object's cast result does not depend on that bound, so: (1) all threads test the ray against every body's AABB with
// the widest bound and collect the candidates, (2) the candidates are cast in parallel, (3) one thread replays the
// reference's loop over the candidates in body order with the stored cast results — the same accept/reject decisions, in
// the same order.
// rayTestSingle's other branches (disp/CollisionWorld.java:301-356) follow the same scheme, because what an object reports
// under the widest bound decides what it reports under any bound b: a triangle mesh / static plane ends on the FIRST
// triangle (in traversal order) that attains its smallest hit distance m (np/TriangleRaycastCallback.java:46-117 accepts
// only distance < hitFraction, and the BVH ray walk, sh/OptimizedBvh.java:817-931, never looks at the bound), a compound on
// the first child attaining its smallest cast fraction (:333-352) — and under bound b it reports exactly that if m < b and
// nothing otherwise.  So step (2) stores one (valid, fraction, world normal) per candidate whatever its shape type.
#pragma once
#include "broadphase.cuh"
#include "epa.cuh"
#include "gjk.cuh"
#include "narrowphase.cuh"

namespace b2c {

#ifndef B2C_RAY_THREADS
#define B2C_RAY_THREADS 128
#endif
constexpr int RAY_THREADS = B2C_RAY_THREADS;
#ifndef B2C_RAY_MAX_CAND
#define B2C_RAY_MAX_CAND 1024
#endif
constexpr int RAY_MAX_CAND = B2C_RAY_MAX_CAND;   // candidates one round keeps in shared memory

__device__ __forceinline__ int rayOutcode(f3 p, f3 h) {  // lm/AabbUtil2.java:40-43
    return (p.x < -h.x ? 0x01 : 0) | (p.x > h.x ? 0x08 : 0) | (p.y < -h.y ? 0x02 : 0) | (p.y > h.y ? 0x10 : 0) |
           (p.z < -h.z ? 0x04 : 0) | (p.z > h.z ? 0x20 : 0);
}
__device__ __forceinline__ float f3get(f3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// lm/AabbUtil2.java:45-108 (the hit normal of the box is not used by the ray test and is not produced)
__device__ __forceinline__ bool rayAabb(f3 rayFrom, f3 rayTo, f3 aabbMin, f3 aabbMax, float param) {
    const f3 he = scl3(sub3(aabbMax, aabbMin), 0.5f);
    const f3 ce = scl3(add3(aabbMax, aabbMin), 0.5f);
    const f3 source = sub3(rayFrom, ce), target = sub3(rayTo, ce);
    const int so = rayOutcode(source, he), to = rayOutcode(target, he);
    if ((so & to) != 0) return false;
    float lambdaEnter = 0.f, lambdaExit = param;
    const f3 r = sub3(target, source);
    float normSign = 1.f;
    int bit = 1;
    for (int j = 0; j < 2; j++) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (so & bit) {
                float lambda = (-f3get(source, i) - f3get(he, i) * normSign) / f3get(r, i);
                if (lambdaEnter <= lambda) lambdaEnter = lambda;
            } else if (to & bit) {
                float lambda = (-f3get(source, i) - f3get(he, i) * normSign) / f3get(r, i);
                lambdaExit = jminf(lambdaExit, lambda);
            }
            bit <<= 1;
        }
        normSign = -1.f;
    }
    return lambdaEnter <= lambdaExit;
}

// np/SubsimplexConvexCast.java:63-190 with convexA = zero sphere (support (0,0,0)), A's basis = identity, B fixed.
// Returns true when a time of impact is reported; fraction / normal as CastResult.
__device__ __forceinline__ bool rayConvexCast(f3 rayFrom, f3 rayTo, const AnyS& B, const Xf& xfB, float& fractionOut, f3& normalOut) {
    Simplex S;
    S.reset();
    const f3 linVelA = sub3(rayTo, rayFrom);
    const f3 linVelB = sub3(xfB.o, xfB.o);
    float lambda = 0.f;
    f3 originA = rayFrom;
    Xf interpB = xfB;
    const f3 r = sub3(linVelA, linVelB);
    const f3 zero = mk3(0.f, 0.f, 0.f);
    // I * (0,0,0) + origin, evaluated as Transform.transform does
    f3 supA = add3(mk3(zero.x * 1.f + zero.y * 0.f + zero.z * 0.f, zero.x * 0.f + zero.y * 1.f + zero.z * 0.f,
                       zero.x * 0.f + zero.y * 0.f + zero.z * 1.f), rayFrom);
    f3 supB = xfPoint(xfB, B.supportMargin(mulMtV(xfB.m, r)));
    f3 v = sub3(supA, supB);
    int maxIter = 32;
    f3 n = zero;
    float dist2 = len2_3(v);
    const float epsilon = 0.0001f;
    while ((dist2 > epsilon) && (maxIter--) != 0) {
        supA = add3(mk3(zero.x * 1.f + zero.y * 0.f + zero.z * 0.f, zero.x * 0.f + zero.y * 1.f + zero.z * 0.f,
                        zero.x * 0.f + zero.y * 0.f + zero.z * 1.f), originA);
        supB = xfPoint(interpB, B.supportMargin(mulMtV(interpB.m, v)));
        f3 w = sub3(supA, supB);
        const float VdotW = dot3(v, w);
        if (lambda > 1.f) return false;
        if (VdotW > 0.f) {
            const float VdotR = dot3(v, r);
            if (VdotR >= -(B2C_FLT_EPSILON * B2C_FLT_EPSILON)) return false;
            lambda = lambda - VdotW / VdotR;
            const float s = 1.f - lambda;  // lm/VectorUtil.java:137-141 setInterpolate3
            originA = mk3(s * rayFrom.x + lambda * rayTo.x, s * rayFrom.y + lambda * rayTo.y, s * rayFrom.z + lambda * rayTo.z);
            interpB.o = mk3(s * xfB.o.x + lambda * xfB.o.x, s * xfB.o.y + lambda * xfB.o.y, s * xfB.o.z + lambda * xfB.o.z);
            w = sub3(supA, supB);
            n = v;
        }
        S.addVertex(w, supA, supB);
        const bool ok = S.update();
        v = S.cachedV;
        dist2 = ok ? len2_3(v) : 0.f;
    }
    fractionOut = lambda;
    f3 nn = zero;
    if (len2_3(n) >= B2C_FLT_EPSILON * B2C_FLT_EPSILON) nn = nor3(n);
    normalOut = nn;
    if (dot3(nn, r) >= -0.f) return false;  // allowedPenetration = 0
    return true;
}

// np/TriangleRaycastCallback.java:46-117 processTriangle with BridgeTriangleRaycastCallback.reportHit ->
// ClosestRayResultCallback.addSingleResult folded in: hitFraction / normal of the last reported triangle
struct TriRay {
    f3 from, to;        // in the concave object's local space
    float hitFraction;
    bool hit;
    f3 normalLocal;     // not normalised, as in the reference
    __device__ __forceinline__ void processTriangle(f3 vert0, f3 vert1, f3 vert2) {
        const f3 v10 = sub3(vert1, vert0), v20 = sub3(vert2, vert0);
        const f3 triangleNormal = crs3(v10, v20);
        const float dist = dot3(vert0, triangleNormal);
        float dist_a = dot3(triangleNormal, from);
        dist_a -= dist;
        float dist_b = dot3(triangleNormal, to);
        dist_b -= dist;
        if (dist_a * dist_b >= 0.f) return;  // same sign
        const float proj_length = dist_a - dist_b;
        const float distance = dist_a / proj_length;
        if (distance < hitFraction) {
            float edge_tolerance = len2_3(triangleNormal);
            edge_tolerance *= -0.0001f;
            const float s = 1.f - distance;  // lm/VectorUtil.java:137-141 setInterpolate3
            const f3 point = mk3(s * from.x + distance * to.x, s * from.y + distance * to.y, s * from.z + distance * to.z);
            const f3 v0p = sub3(vert0, point), v1p = sub3(vert1, point);
            const f3 cp0 = crs3(v0p, v1p);
            if (dot3(cp0, triangleNormal) >= edge_tolerance) {
                const f3 v2p = sub3(vert2, point);
                const f3 cp1 = crs3(v1p, v2p);
                if (dot3(cp1, triangleNormal) >= edge_tolerance) {
                    const f3 cp2 = crs3(v2p, v0p);
                    if (dot3(cp2, triangleNormal) >= edge_tolerance) {
                        hit = true;
                        normalLocal = dist_a > 0.f ? triangleNormal : neg3(triangleNormal);
                        hitFraction = distance;
                    }
                }
            }
        }
    }
};

// sh/BvhTriangleMeshShape.java:135-142 performRaycast -> sh/OptimizedBvh.java:817-931 walkStacklessQuantizedTreeAgainstRay
// with zero box-cast extents: quantised box prune, then rayAabb (exit bound 1) on the unquantised node box
__device__ __forceinline__ void rayMeshWalk(const MeshDev& md, TriRay& tr) {
    f3 rmin = mk3(jminf(tr.from.x, tr.to.x), jminf(tr.from.y, tr.to.y), jminf(tr.from.z, tr.to.z));
    f3 rmax = mk3(jmaxf(tr.from.x, tr.to.x), jmaxf(tr.from.y, tr.to.y), jmaxf(tr.from.z, tr.to.z));
    const f3 zero = mk3(0.f, 0.f, 0.f);
    rmin = add3(rmin, zero);
    rmax = add3(rmax, zero);
    uint32_t qmin[3], qmax[3];
    quantizeClamp(md, rmin, qmin);
    quantizeClamp(md, rmax, qmax);
    int cur = 0;
    const int end = md.numNodes;
    while (cur < end) {
        const int4 nd = __ldg(md.nodes + cur);
        const uint32_t nminx = (uint32_t)nd.x & 0xFFFFu, nminy = ((uint32_t)nd.x >> 16) & 0xFFFFu, nminz = (uint32_t)nd.y & 0xFFFFu;
        const uint32_t nmaxx = ((uint32_t)nd.y >> 16) & 0xFFFFu, nmaxy = (uint32_t)nd.z & 0xFFFFu, nmaxz = ((uint32_t)nd.z >> 16) & 0xFFFFu;
        const bool boxBox = !(qmin[0] > nmaxx || qmax[0] < nminx) && !(qmin[2] > nmaxz || qmax[2] < nminz) &&
                            !(qmin[1] > nmaxy || qmax[1] < nminy);
        const bool leaf = nd.w >= 0;
        bool rayBox = false;
        if (boxBox) {  // unQuantize (:1058-1068) + the zero extents
            f3 b0 = add3(mk3(__uint2float_rn(nminx) / md.quant[0], __uint2float_rn(nminy) / md.quant[1], __uint2float_rn(nminz) / md.quant[2]),
                         mk3(md.qmin[0], md.qmin[1], md.qmin[2]));
            f3 b1 = add3(mk3(__uint2float_rn(nmaxx) / md.quant[0], __uint2float_rn(nmaxy) / md.quant[1], __uint2float_rn(nmaxz) / md.quant[2]),
                         mk3(md.qmin[0], md.qmin[1], md.qmin[2]));
            b0 = add3(b0, zero);
            b1 = add3(b1, zero);
            rayBox = rayAabb(tr.from, tr.to, b0, b1, 1.f);
        }
        if (leaf && rayBox) {
            const TriS t = loadTri(md, nd.w, 0.f);
            tr.processTriangle(t.a, t.b, t.c);
        }
        if (rayBox || leaf) cur++;
        else cur += -nd.w;
    }
}

// sh/StaticPlaneShape.java:60-122 processAllTriangles over the ray's local AABB (including the `set(aabbMax).set(aabbMin)`
// half-extent line, :67) with lm/TransformUtil.java:45-61 planeSpace1
__device__ __forceinline__ void rayPlaneTriangles(f3 planeNormal, float planeConstant, TriRay& tr) {
    const f3 aabbMin = mk3(jminf(tr.from.x, tr.to.x), jminf(tr.from.y, tr.to.y), jminf(tr.from.z, tr.to.z));
    const f3 aabbMax = mk3(jmaxf(tr.from.x, tr.to.x), jmaxf(tr.from.y, tr.to.y), jmaxf(tr.from.z, tr.to.z));
    const f3 halfExtents = scl3(aabbMin, 0.5f);  // (sic)
    const float radius = len3(halfExtents);
    const f3 center = scl3(add3(aabbMax, aabbMin), 0.5f);
    f3 t0, t1;
    const f3 n = planeNormal;
    if (fabsf(n.z) > 0.7071067811865475244008443621048490f) {
        const float a = n.y * n.y + n.z * n.z;
        const float k = 1.f / jsqrtf(a);
        t0 = mk3(0.f, -n.z * k, n.y * k);
        t1 = mk3(a * k, -n.x * t0.z, n.x * t0.y);
    } else {
        const float a = n.x * n.x + n.y * n.y;
        const float k = 1.f / jsqrtf(a);
        t0 = mk3(-n.y * k, n.x * k, 0.f);
        t1 = mk3(-n.z * t0.y, n.z * t0.x, a * k);
    }
    const f3 projectedCenter = sub3(center, scl3(planeNormal, dot3(planeNormal, center) - planeConstant));
    const f3 tmp1 = scl3(t0, radius), tmp2 = scl3(t1, radius);
    const f3 both = mk3(projectedCenter.x + tmp1.x + tmp2.x, projectedCenter.y + tmp1.y + tmp2.y, projectedCenter.z + tmp1.z + tmp2.z);
    const f3 diff = sub3(tmp1, tmp2);
    tr.processTriangle(both, add3(projectedCenter, diff), sub3(projectedCenter, diff));
    tr.processTriangle(sub3(projectedCenter, diff), sub3(projectedCenter, add3(tmp1, tmp2)), both);
}

// tight shape AABBs (shape.getAabb(worldTransform), no contact threshold) of the bodies a ray can hit; others get an empty box
__global__ void __launch_bounds__(256)
k_ray_aabbs(BodyArrays B, const ShapeDev* __restrict__ shapes, int n, float4* __restrict__ rmin, float4* __restrict__ rmax) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 mn = make_float4(1.f, 1.f, 1.f, 0.f), mx = make_float4(-1.f, -1.f, -1.f, 0.f);  // w = 0: not a candidate
    const uint8_t fl = B.flags[i];
    if (fl & BF_ALIVE) {
        const ShapeDev s = shapes[B.shape[i]];
        {
            Xf t = loadXf(B.xf4, i);
            f3 a, b;
            shapeAabb(s, t, a, b);
            mn = make_float4(a.x, a.y, a.z, 1.f);
            mx = make_float4(b.x, b.y, b.z, 0.f);
        }
    }
    rmin[i] = mn;
    rmax[i] = mx;
}

// Culling of the body loop.  Bodies are taken in the order the last broadphase sorted them (grid row, then x: consecutive
// positions are neighbours in space; bodies created since follow in index order) and grouped into chunks of RAY_CHUNK; a chunk
// box is the union of its members' boxes, inflated a little so that a segment touching a member box can never miss the union
// through rounding.  A ray tests the chunk boxes first and only the members of the chunks it meets: the candidate set stays a
// superset of "bodies whose box the ray meets", so the results are unchanged.
constexpr int RAY_CHUNK = 64;
__device__ __forceinline__ int rayBodyAt(const float4* __restrict__ sortedMin, int nSorted, int pos) {
    return pos < nSorted ? (int)__float_as_uint(__ldg(&sortedMin[pos].w)) : pos;
}
__global__ void __launch_bounds__(128)
k_ray_chunks(const float4* __restrict__ rmin, const float4* __restrict__ rmax, const float4* __restrict__ sortedMin, int nSorted, int n,
             float4* __restrict__ cmin, float4* __restrict__ cmax) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int nChunks = (n + RAY_CHUNK - 1) / RAY_CHUNK;
    if (c >= nChunks) return;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    int members = 0;
    for (int k = 0; k < RAY_CHUNK; k++) {
        const int pos = c * RAY_CHUNK + k;
        if (pos >= n) break;
        const int i = rayBodyAt(sortedMin, nSorted, pos);
        if (i < 0 || i >= n) continue;
        const float4 a = __ldg(rmin + i);
        if (a.w == 0.f) continue;
        const float4 b = __ldg(rmax + i);
        lo[0] = fminf(lo[0], a.x); lo[1] = fminf(lo[1], a.y); lo[2] = fminf(lo[2], a.z);
        hi[0] = fmaxf(hi[0], b.x); hi[1] = fmaxf(hi[1], b.y); hi[2] = fmaxf(hi[2], b.z);
        members++;
    }
    for (int d = 0; d < 3; d++) {  // conservative inflation: absolute + relative
        lo[d] = lo[d] - (1e-3f + 1e-5f * fabsf(lo[d]));
        hi[d] = hi[d] + (1e-3f + 1e-5f * fabsf(hi[d]));
    }
    cmin[c] = make_float4(lo[0], lo[1], lo[2], members ? 1.f : 0.f);
    cmax[c] = make_float4(hi[0], hi[1], hi[2], 0.f);
}

struct RayOut {
    int uid;          // 0 = no hit
    float fraction;   // closestHitFraction (1 = no hit)
    float normal[3];  // hitNormalWorld
    float point[3];   // hitPointWorld
};

__global__ void __launch_bounds__(RAY_THREADS)
k_ray_test(BodyArrays B, const ShapeDev* __restrict__ shapes, const float4* __restrict__ hullPts, const MeshDev* __restrict__ meshes,
           const CompoundChildDev* __restrict__ children, const float4* __restrict__ sortedMin, int nSorted,
           const float4* __restrict__ cmin, const float4* __restrict__ cmax, int n,
           const float4* __restrict__ rmin, const float4* __restrict__ rmax, const float* __restrict__ rayFrom,
           const float* __restrict__ rayTo, int numRays, uint32_t cbFilter /* group | mask << 16 */, RayOut* __restrict__ out,
           uint32_t* __restrict__ overflow) {
    __shared__ uint32_t sCount;
    __shared__ uint32_t sCand[RAY_MAX_CAND];
    __shared__ float sFrac[RAY_MAX_CAND];
    __shared__ float sNrm[RAY_MAX_CAND][3];
    __shared__ uint8_t sValid[RAY_MAX_CAND];
    for (int ray = blockIdx.x; ray < numRays; ray += gridDim.x) {
        const f3 from = mk3(rayFrom[3 * ray], rayFrom[3 * ray + 1], rayFrom[3 * ray + 2]);
        const f3 to = mk3(rayTo[3 * ray], rayTo[3 * ray + 1], rayTo[3 * ray + 2]);
        __syncthreads();
        if (threadIdx.x == 0) sCount = 0;
        __syncthreads();
        // (1) every body whose AABB the ray meets under the widest bound (closestHitFraction = 1)
        const int nChunks = (n + RAY_CHUNK - 1) / RAY_CHUNK;
        for (int c = threadIdx.x; c < nChunks; c += RAY_THREADS) {
            const float4 cn = __ldg(cmin + c);
            if (cn.w == 0.f) continue;
            const float4 cx = __ldg(cmax + c);
            if (!rayAabb(from, to, mk3(cn.x, cn.y, cn.z), mk3(cx.x, cx.y, cx.z), 1.f)) continue;
            for (int m = 0; m < RAY_CHUNK; m++) {
                const int pos = c * RAY_CHUNK + m;
                if (pos >= n) break;
                const int i = rayBodyAt(sortedMin, nSorted, pos);
                if (i < 0 || i >= n) continue;
                const float4 mn = __ldg(rmin + i);
                if (mn.w == 0.f) continue;
                if (!filterPass(cbFilter, B.filt[i])) continue;  // RayResultCallback.needsCollision (disp/CollisionWorld.java:664-670)
                const float4 mx = __ldg(rmax + i);
                if (rayAabb(from, to, mk3(mn.x, mn.y, mn.z), mk3(mx.x, mx.y, mx.z), 1.f)) {
                    uint32_t k = atomicAdd(&sCount, 1u);
                    if (k < RAY_MAX_CAND) sCand[k] = (uint32_t)i;
                }
            }
        }
        // (2) the casts, one candidate per thread
        auto castAll = [&](uint32_t cnt) {
        for (uint32_t k = threadIdx.x; k < cnt; k += RAY_THREADS) {
            const int i = (int)sCand[k];
            const ShapeDev s = shapes[B.shape[i]];
            const Xf t = loadXf(B.xf4, i);
            float fr = 1.f;
            f3 nn = mk3(0.f, 0.f, 0.f);
            bool valid = false;
            auto castConvex = [&](const ShapeDev& cs, const Xf& cx, float bound) {
                AnyS shp;
                shp.type = cs.type;
                shp.h = mk3(cs.dims[0], cs.dims[1], cs.dims[2]);
                shp.ta = shp.tb = shp.tc = mk3(0.f, 0.f, 0.f);
                shp.pts = hullPts + cs.pointOffset;
                shp.n = cs.numPoints;
                shp.margin = cs.margin;
                float f = 1.f;
                f3 c = mk3(0.f, 0.f, 0.f);
                const bool hit = rayConvexCast(from, to, shp, cx, f, c);
                if (hit && len2_3(c) > 0.0001f && f < bound) {
                    // castResult.normal.mul(rayFromTrans.basis) with the identity basis, then nor()
                    nn = nor3(mk3(c.x * 1.f + c.y * 0.f + c.z * 0.f, c.x * 0.f + c.y * 1.f + c.z * 0.f, c.x * 0.f + c.y * 0.f + c.z * 1.f));
                    fr = f;
                    valid = true;
                }
            };
            if (s.type == SH_BOX || s.type == SH_SPHERE || s.type == SH_HULL) {
                castConvex(s, t, 1.f);
                if (!valid) fr = 1.f;
            } else if (s.type == SH_COMPOUND) {  // :333-352: the children in order, each against the bound its predecessors left
                for (int ch = 0; ch < s.numPoints; ch++) {
                    const CompoundChildDev& cd = children[s.pointOffset + ch];
                    castConvex(shapes[cd.shape], compoundChildWorld(t, children, cd), fr);
                }
            } else {  // :301-331: static plane / triangle mesh in the object's local space
                Xf inv;  // Transform.inverse (lm/Transform.java:101-105)
                for (int r = 0; r < 3; r++)
                    for (int c = 0; c < 3; c++) inv.m[r][c] = t.m[c][r];
                inv.o = mulMV(inv.m, neg3(t.o));
                TriRay tr;
                tr.from = xfPoint(inv, from);
                tr.to = xfPoint(inv, to);
                tr.hitFraction = 1.f;
                tr.hit = false;
                tr.normalLocal = mk3(0.f, 0.f, 0.f);
                if (s.type == SH_MESH) rayMeshWalk(meshes[s.mesh], tr);
                else rayPlaneTriangles(mk3(s.plane[0], s.plane[1], s.plane[2]), s.plane[3], tr);
                if (tr.hit) {
                    valid = true;
                    fr = tr.hitFraction;
                    nn = mulMV(t.m, tr.normalLocal);  // hitNormalWorld.mul(collisionObject.getWorldTransform().basis), not normalised
                }
            }
            sValid[k] = valid ? 1 : 0;
            sFrac[k] = fr;
            sNrm[k][0] = nn.x; sNrm[k][1] = nn.y; sNrm[k][2] = nn.z;
        }
        };
        // (3) the reference's loop over the listed candidates in ascending body index, continuing from (closest, hitBody, hn)
        float closest = 1.f;
        int hitBody = -1;
        f3 hn = mk3(0.f, 0.f, 0.f);
        auto replay = [&](uint32_t cnt) {
            uint32_t done = 0;
            int lastIdx = -1;
            while (done < cnt) {
                // next candidate in ascending body index (selection: the candidate lists are short)
                int best = 0x7fffffff;
                uint32_t bk = 0;
                for (uint32_t k = 0; k < cnt; k++) {
                    const int i = (int)sCand[k];
                    if (i > lastIdx && i < best) { best = i; bk = k; }
                }
                lastIdx = best;
                done++;
                if (closest == 0.f) break;
                const float4 mn = __ldg(rmin + best), mx = __ldg(rmax + best);
                if (!rayAabb(from, to, mk3(mn.x, mn.y, mn.z), mk3(mx.x, mx.y, mx.z), closest)) continue;
                if (sValid[bk] && sFrac[bk] < closest) {
                    closest = sFrac[bk];
                    hitBody = best;
                    hn = mk3(sNrm[bk][0], sNrm[bk][1], sNrm[bk][2]);
                }
            }
        };
        __syncthreads();
        const uint32_t total = sCount;
        if (total <= RAY_MAX_CAND) {
            castAll(total);
            __syncthreads();
            if (threadIdx.x == 0) replay(total);
        } else {
            // A ray that meets more boxes than one round holds (a very long ray through a dense pile) takes the bodies in
            // INDEX order instead, RAY_MAX_CAND consecutive indices at a time: collect, cast, replay — the replay state carries
            // over from tile to tile, so the decisions are still the reference's, in its order; only the culling is lost.
            if (threadIdx.x == 0) atomicMax(overflow, total);   // informational: how many boxes the longest such ray met
            for (int t0 = 0; t0 < n; t0 += RAY_MAX_CAND) {
                __syncthreads();
                if (threadIdx.x == 0) sCount = 0;
                __syncthreads();
                const int t1 = min(n, t0 + RAY_MAX_CAND);
                for (int i = t0 + threadIdx.x; i < t1; i += RAY_THREADS) {
                    const float4 mn = __ldg(rmin + i);
                    if (mn.w == 0.f) continue;
                    if (!filterPass(cbFilter, B.filt[i])) continue;
                    const float4 mx = __ldg(rmax + i);
                    if (rayAabb(from, to, mk3(mn.x, mn.y, mn.z), mk3(mx.x, mx.y, mx.z), 1.f)) sCand[atomicAdd(&sCount, 1u)] = (uint32_t)i;
                }
                __syncthreads();
                const uint32_t c2 = sCount;
                castAll(c2);
                __syncthreads();
                if (threadIdx.x == 0) replay(c2);
            }
        }
        if (threadIdx.x == 0) {
            RayOut o;
            o.uid = hitBody + 1;
            o.fraction = closest;
            o.normal[0] = hn.x; o.normal[1] = hn.y; o.normal[2] = hn.z;
            const float s = 1.f - closest;
            f3 pt = hitBody >= 0 ? mk3(s * from.x + closest * to.x, s * from.y + closest * to.y, s * from.z + closest * to.z) : mk3(0.f, 0.f, 0.f);
            o.point[0] = pt.x; o.point[1] = pt.y; o.point[2] = pt.z;
            out[ray] = o;
        }
    }
}

}  // namespace b2c
