// convexcast.cuh — batched closest-hit convex sweeps against the bodies of the world (the CCD query).
//
// Replaces, per sweep: CollisionWorld.convexSweepTest with a ClosestConvexResultCallback (disp/CollisionWorld.java:596-651,
// 765-800) for a TRANSLATIONAL sweep (the cast shape keeps one basis from start to end):
//   * cast-shape box: castShape.calculateTemporalAabb(R, linVel, angVel, 1) (sh/CollisionShape.java:74-122).  For equal
//     bases the angular velocity TransformUtil.calculateVelocity derives (lm/TransformUtil.java:105-153, through a
//     quaternion) is rounding noise; it is taken as exactly zero here, and R = the sweep's basis (the reference round-trips
//     it through a quaternion, :617-618).  This box only culls.
//   * objects in world order, needsCollision (:752-756), the object's box expanded by the cast box
//     (lm/AabbUtil2.java:35-38), rayAabb with the constant exit bound 1 (:634), then objectQuerySingle (:392-551):
//       convex        np/GjkConvexCast.java:66-196 — conservative advancement over GjkPairDetector WITHOUT a penetration
//                     solver, one PointCollector for the whole cast (np/PointCollector.java:44-52);
//       triangle mesh sh/BvhTriangleMeshShape.java:144-152 -> sh/OptimizedBvh.java:1017-1036 box-cast walk with the cast
//                     shape's box in mesh space, np/TriangleConvexcastCallback.java:53-88: np/SubsimplexConvexCast.java
//                     :63-190 against each triangle (margin = the mesh's), BridgeTriangleConvexcastCallback.reportHit
//                     (:362-388);
//       compound      every child in order with colObjWorld * childTrans (:528-545);
//       static plane  the reference calls calcTimeOfImpact on a null caster (:470-477, "Start buggy") and throws; a sweep
//                     that reaches that branch reports uid -1.
// Same scheme as raycast.cuh: one block per sweep; (1) all threads collect the bodies whose expanded box the segment
// meets, (2) the candidates are cast in parallel, (3) the callback's decisions are replayed.  Because the exit bound of the
// box test is constant and every accept is a strict "fraction < closest" between objects, (3) is: the smallest fraction
// wins, the lowest body index among equals — an order-independent reduction, so (1) and (2) alternate in rounds over a
// fixed-size candidate list and the number of candidates a sweep may have is unbounded.  Inside one object: a compound keeps the first child attaining its minimum;
// a mesh keeps the LAST triangle (in traversal order) attaining its minimum, since reportHit accepts "<=" — and either
// reports that minimum m under a bound b exactly when m < b, so one record per candidate is enough.
#pragma once
#include "raycast.cuh"

namespace b2c {

constexpr int SWEEP_THREADS = RAY_THREADS;
#ifndef B2C_SWEEP_MAX_CAND
#define B2C_SWEEP_MAX_CAND 4096
#define B2C_SWEEP_HIT_CAP 2048
#endif
constexpr int SWEEP_MAX_CAND = B2C_SWEEP_MAX_CAND;  // candidate list of one round (>= RAY_CHUNK)
constexpr int SWEEP_HIT_CAP = B2C_SWEEP_HIT_CAP;    // chunk boxes examined per round = capacity of the hit-chunk list

__device__ __forceinline__ AnyS anyShapeOf(const ShapeDev& cs, const float4* __restrict__ hullPts) {
    AnyS shp;
    shp.type = cs.type;
    shp.h = mk3(cs.dims[0], cs.dims[1], cs.dims[2]);
    shp.ta = shp.tb = shp.tc = mk3(0.f, 0.f, 0.f);
    shp.pts = hullPts + cs.pointOffset;
    shp.n = cs.numPoints;
    shp.margin = cs.margin;
    return shp;
}

// np/GjkPairDetector.java:73-303 with penetrationDepthSolver == null and maximumDistanceSquared = Float.MAX_VALUE, feeding a
// PointCollector: (distance, normalOnBInWorld, pointInWorld) replace the collector's only when the distance is smaller
struct PointCollectorDev {
    bool hasResult;
    float distance;
    f3 normalOnBInWorld, pointInWorld;
};
__device__ __noinline__ void gjkCollect(const AnyS& A, const Xf& ta, const AnyS& B, const Xf& tb, PointCollectorDev& pc) {
    GjkLane L;
    L.begin(ta, tb, A.margin, B.margin, B2C_SIMD_INFINITY);
    for (;;) {
        const f3 pW = add3(mulMV(ta.m, A.support(L.dirA(ta))), L.laO);
        const f3 qW = add3(mulMV(tb.m, B.support(L.dirB(tb))), L.lbO);
        if (L.iterate(pW, qW)) break;
    }
    GjkResult r;
    L.finish(r);
    if (r.isValid && r.distance < pc.distance) {
        pc.hasResult = true;
        pc.distance = r.distance;
        pc.normalOnBInWorld = r.normalInB;
        pc.pointInWorld = add3(r.pointOnB, r.positionOffset);
    }
}

// np/GjkConvexCast.java:66-196 calcTimeOfImpact, B at rest
__device__ __forceinline__ bool gjkConvexCast(const AnyS& A, const Xf& fromA, f3 toOrigin, const AnyS& B, const Xf& xfB,
                                              float allowedPenetration, float& fractionOut, f3& normalOut, f3& pointOut) {
    const f3 linVelA = sub3(toOrigin, fromA.o);
    const f3 linVelB = sub3(xfB.o, xfB.o);
    const float radius = 0.001f;
    float lambda = 0.f;
    const f3 r = sub3(linVelA, linVelB);
    float lastLambda = lambda;
    int numIter = 0;
    PointCollectorDev pc;
    pc.hasResult = false;
    pc.distance = 1e30f;
    pc.normalOnBInWorld = pc.pointInWorld = mk3(0.f, 0.f, 0.f);
    Xf inA = fromA, inB = xfB;
    gjkCollect(A, inA, B, inB, pc);
    if (!pc.hasResult) return false;
    f3 c = pc.pointInWorld;
    float dist = pc.distance;
    f3 n = pc.normalOnBInWorld;
    while (dist > radius) {
        numIter++;
        if (numIter > 32) return false;
        const float projectedLinearVelocity = dot3(r, n);
        const float dLambda = dist / projectedLinearVelocity;
        lambda = lambda - dLambda;
        if (lambda > 1.f) return false;
        if (lambda < 0.f) return false;
        if (lambda <= lastLambda) return false;
        lastLambda = lambda;
        const float s = 1.f - lambda;  // lm/VectorUtil.java:137-141 setInterpolate3
        inA.o = mk3(s * fromA.o.x + lambda * toOrigin.x, s * fromA.o.y + lambda * toOrigin.y, s * fromA.o.z + lambda * toOrigin.z);
        inB.o = mk3(s * xfB.o.x + lambda * xfB.o.x, s * xfB.o.y + lambda * xfB.o.y, s * xfB.o.z + lambda * xfB.o.z);
        gjkCollect(A, inA, B, inB, pc);
        if (pc.distance < 0.f) {
            fractionOut = lastLambda;
            normalOut = pc.normalOnBInWorld;
            pointOut = pc.pointInWorld;
            return true;
        }
        c = pc.pointInWorld;
        n = pc.normalOnBInWorld;
        dist = pc.distance;
    }
    if (dot3(n, r) >= -allowedPenetration) return false;
    fractionOut = lambda;
    normalOut = n;
    pointOut = c;
    return true;
}

// np/SubsimplexConvexCast.java:63-190 for a moving A against B at rest (fromB == toB); hitPoint = the simplex's point on B
__device__ __forceinline__ bool subsimplexConvexCast(const AnyS& A, const Xf& fromA, f3 toOrigin, const AnyS& B, const Xf& xfB,
                                                     float& fractionOut, f3& normalOut, f3& pointOut) {
    Simplex S;
    S.reset();
    const f3 linVelA = sub3(toOrigin, fromA.o);
    const f3 linVelB = sub3(xfB.o, xfB.o);
    float lambda = 0.f;
    Xf interpA = fromA, interpB = xfB;
    const f3 r = sub3(linVelA, linVelB);
    f3 supA = xfPoint(fromA, A.supportMargin(mulMtV(fromA.m, neg3(r))));
    f3 supB = xfPoint(xfB, B.supportMargin(mulMtV(xfB.m, r)));
    f3 v = sub3(supA, supB);
    int maxIter = 32;
    const f3 zero = mk3(0.f, 0.f, 0.f);
    f3 n = zero;
    float dist2 = len2_3(v);
    const float epsilon = 0.0001f;
    while ((dist2 > epsilon) && (maxIter--) != 0) {
        supA = xfPoint(interpA, A.supportMargin(mulMtV(interpA.m, neg3(v))));
        supB = xfPoint(interpB, B.supportMargin(mulMtV(interpB.m, v)));
        f3 w = sub3(supA, supB);
        const float VdotW = dot3(v, w);
        if (lambda > 1.f) return false;
        if (VdotW > 0.f) {
            const float VdotR = dot3(v, r);
            if (VdotR >= -(B2C_FLT_EPSILON * B2C_FLT_EPSILON)) return false;
            lambda = lambda - VdotW / VdotR;
            const float s = 1.f - lambda;
            interpA.o = mk3(s * fromA.o.x + lambda * toOrigin.x, s * fromA.o.y + lambda * toOrigin.y, s * fromA.o.z + lambda * toOrigin.z);
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
    if (dot3(nn, r) >= -0.f) return false;  // CastResult.allowedPenetration = 0 for the triangle casts
    pointOut = S.cachedP2;                  // compute_points: the cache is current
    return true;
}

struct SweepRec {  // what one candidate reports under the widest bound
    float fraction;
    float normal[3];
    float point[3];
};

// sh/OptimizedBvh.java:817-931 walkStacklessQuantizedTreeAgainstRay with box-cast extents: the cast of `A` against every
// triangle the walk reports, folded as BridgeTriangleConvexcastCallback does under the entry bound 1
__device__ __forceinline__ void sweepMeshWalk(const MeshDev& md, const AnyS& A, const Xf& fromA, f3 toOrigin, const Xf& meshXf,
                                              float meshMargin, f3 fromLocal, f3 toLocal, f3 boxMin, f3 boxMax, bool& valid,
                                              SweepRec& rec, bool notMe, f3 rel) {
    f3 rmin = mk3(jminf(fromLocal.x, toLocal.x), jminf(fromLocal.y, toLocal.y), jminf(fromLocal.z, toLocal.z));
    f3 rmax = mk3(jmaxf(fromLocal.x, toLocal.x), jmaxf(fromLocal.y, toLocal.y), jmaxf(fromLocal.z, toLocal.z));
    rmin = add3(rmin, boxMin);
    rmax = add3(rmax, boxMax);
    uint32_t qmin[3], qmax[3];
    quantizeClamp(md, rmin, qmin);
    quantizeClamp(md, rmax, qmax);
    int cur = 0;
    const int end = md.numNodes;
    float closest = 1.f;
    while (cur < end) {
        const int4 nd = __ldg(md.nodes + cur);
        const uint32_t nminx = (uint32_t)nd.x & 0xFFFFu, nminy = ((uint32_t)nd.x >> 16) & 0xFFFFu, nminz = (uint32_t)nd.y & 0xFFFFu;
        const uint32_t nmaxx = ((uint32_t)nd.y >> 16) & 0xFFFFu, nmaxy = (uint32_t)nd.z & 0xFFFFu, nmaxz = ((uint32_t)nd.z >> 16) & 0xFFFFu;
        const bool boxBox = !(qmin[0] > nmaxx || qmax[0] < nminx) && !(qmin[2] > nmaxz || qmax[2] < nminz) &&
                            !(qmin[1] > nmaxy || qmax[1] < nminy);
        const bool leaf = nd.w >= 0;
        bool rayBox = false;
        if (boxBox) {
            f3 b0 = add3(mk3(__uint2float_rn(nminx) / md.quant[0], __uint2float_rn(nminy) / md.quant[1], __uint2float_rn(nminz) / md.quant[2]),
                         mk3(md.qmin[0], md.qmin[1], md.qmin[2]));
            f3 b1 = add3(mk3(__uint2float_rn(nmaxx) / md.quant[0], __uint2float_rn(nmaxy) / md.quant[1], __uint2float_rn(nmaxz) / md.quant[2]),
                         mk3(md.qmin[0], md.qmin[1], md.qmin[2]));
            b0 = add3(b0, boxMin);
            b1 = add3(b1, boxMax);
            rayBox = rayAabb(fromLocal, toLocal, b0, b1, 1.f);
        }
        if (leaf && rayBox) {
            const TriS t = loadTri(md, nd.w, meshMargin);
            AnyS T;
            T.type = SH_TRIANGLE;
            T.h = mk3(0.f, 0.f, 0.f);
            T.ta = t.a; T.tb = t.b; T.tc = t.c;
            T.pts = nullptr;
            T.n = 0;
            T.margin = meshMargin;
            float f = 1.f;
            f3 nn, pt;
            if (subsimplexConvexCast(A, fromA, toOrigin, T, meshXf, f, nn, pt)) {
                if (len2_3(nn) > 0.0001f && f < 1.f) {
                    nn = nor3(nn);
                    if (f <= closest && !(notMe && dot3(nn, rel) >= -0.f)) {   // ClosestNotMe...addSingleResult (:1157)
                        closest = f;
                        valid = true;
                        rec.fraction = f;
                        rec.normal[0] = nn.x; rec.normal[1] = nn.y; rec.normal[2] = nn.z;
                        rec.point[0] = pt.x; rec.point[1] = pt.y; rec.point[2] = pt.z;
                    }
                }
            }
        }
        if (rayBox || leaf) cur++;
        else cur += -nd.w;
    }
}

// DiscreteDynamicsWorld's ClosestNotMeConvexResultCallback (dyn/DiscreteDynamicsWorld.java:1129-1199) for the CCD motion
// clamping sweeps of integrateTransforms (:700-729): per sweep the swept body `me` (its filter group / mask are the
// callback's), a sphere of its ccdSweptSphereRadius as the cast shape, and the pair cache with its manifolds: an object the
// body already has contact points with is skipped (needsCollision, :1181-1196), so is a result whose normal does not
// oppose the motion (addSingleResult, :1157).
struct SweepNotMe {
    const int* me;               // [numSweeps] 0-based body index, or null: plain ClosestConvexResultCallback sweeps
    const float* radius;         // [numSweeps] radius of the swept sphere
    const uint64_t* keys;        // sorted pair keys of the last pair calculation
    const uint32_t* numPairs;    // null: no pair cache yet
    const uint32_t* first;       // first pair of every uid0
    int uidBits;
    const ManifoldHdr* mhdr;     // manifold header of every pair
    const ManifoldHdr* compH;    // child manifolds of compound pairs (or null)
};
__device__ __forceinline__ bool sweepAlreadyTouching(const SweepNotMe& nm, const BodyArrays& B, const ShapeDev* __restrict__ shapes, int me,
                                                     int other) {
    if (!nm.numPairs) return false;
    const uint32_t n = *nm.numPairs;
    const uint32_t u0 = (uint32_t)(me < other ? me : other) + 1u, u1 = (uint32_t)(me < other ? other : me) + 1u;
    const int p = findPairIndex(nm.keys, n, nm.first, ((uint64_t)u0 << nm.uidBits) | u1, nm.uidBits);
    if (p < 0) return false;
    const ManifoldHdr h = nm.mhdr[p];
    if (h.algorithm == 5) {   // compound pair: any child manifold (getAllContactManifolds of the compound algorithm)
        if (!nm.compH) return false;
        const ShapeDev& S0 = shapes[B.shape[u0 - 1]];
        const ShapeDev& S1 = shapes[B.shape[u1 - 1]];
        const uint32_t cnt = (S0.type == SH_COMPOUND ? (uint32_t)S0.numPoints : 1u) * (S1.type == SH_COMPOUND ? (uint32_t)S1.numPoints : 1u);
        for (uint32_t k = 0; k < cnt; k++)
            if (nm.compH[(uint32_t)h.pad1 + k].num_contacts > 0) return true;
        return false;
    }
    return h.algorithm != 0 && h.num_contacts > 0;
}

// One candidate's cast, folded into the thread's running best (smallest fraction, lowest body index among equals).
struct SweepBest {
    float fraction;
    int body;      // -1 = none
    SweepRec rec;
    __device__ __forceinline__ void offer(bool valid, int i, const SweepRec& r) {
        if (!valid) return;
        if (r.fraction < fraction || (body >= 0 && r.fraction == fraction && i < body)) { fraction = r.fraction; body = i; rec = r; }
    }
};

__global__ void __launch_bounds__(SWEEP_THREADS)
k_convex_sweep(BodyArrays B, const ShapeDev* __restrict__ shapes, const float4* __restrict__ hullPts, const MeshDev* __restrict__ meshes,
               const CompoundChildDev* __restrict__ children, const float4* __restrict__ sortedMin, int nSorted,
               const float4* __restrict__ cmin, const float4* __restrict__ cmax, int n, const float4* __restrict__ rmin,
               const float4* __restrict__ rmax, const int* __restrict__ castShapes, const float* __restrict__ basis9,
               const float* __restrict__ sweepFrom, const float* __restrict__ sweepTo, int numSweeps, uint32_t cbFilterIn,
               float allowedPenetration, RayOut* __restrict__ out, uint32_t* __restrict__ maxCandidates, SweepNotMe nm) {
    // The candidate set of a sweep has no useful bound (the reference expands every body's box by the cast shape's box
    // INCLUDING its whole linear motion), so it is produced and consumed in rounds: chunk boxes of a range of SWEEP_HIT_CAP
    // chunks -> hit-chunk list; members of as many hit chunks as the candidate list has room for (64 each, worst case) ->
    // candidate list; when it is full, or at the end, all threads cast its entries into their running best.
    __shared__ uint32_t sHitCount, sCount, sTotal;
    __shared__ int sUnsupported;
    __shared__ uint32_t sHit[SWEEP_HIT_CAP];
    __shared__ uint32_t sCand[SWEEP_MAX_CAND];
    __shared__ float sBestFrac[SWEEP_THREADS];
    __shared__ int sBestBody[SWEEP_THREADS];
    __shared__ SweepRec sBestRec[SWEEP_THREADS];
    for (int sw = blockIdx.x; sw < numSweeps; sw += gridDim.x) {
        uint32_t cbFilter = cbFilterIn;
        Xf fromT;
        for (int r = 0; r < 3; r++)
            for (int c = 0; c < 3; c++) fromT.m[r][c] = nm.me ? (r == c ? 1.f : 0.f) : basis9[9 * (size_t)sw + 3 * r + c];
        fromT.o = nm.me ? mk3(0.f, 0.f, 0.f) : mk3(sweepFrom[3 * sw], sweepFrom[3 * sw + 1], sweepFrom[3 * sw + 2]);
        const f3 to = mk3(sweepTo[3 * sw], sweepTo[3 * sw + 1], sweepTo[3 * sw + 2]);
        const int me = nm.me ? nm.me[sw] : -1;
        ShapeDev castS;
        if (nm.me) {   // SphereShape(body.getCcdSweptSphereRadius())
            castS = ShapeDev{};
            castS.type = SH_SPHERE;
            castS.dims[0] = nm.radius[sw];
            castS.margin = nm.radius[sw];
        } else {
            castS = shapes[castShapes[sw]];
        }
        const AnyS A = anyShapeOf(castS, hullPts);
        if (nm.me) {   // the body's own transform is the start of the sweep; its filter is the callback's
            const Xf mt = loadXf(B.xf4, me);
            for (int r = 0; r < 3; r++)
                for (int c = 0; c < 3; c++) fromT.m[r][c] = mt.m[r][c];
            fromT.o = mt.o;
            cbFilter = B.filt[me];
        }
        const f3 from = fromT.o;
        const f3 rel = sub3(sub3(to, from), mk3(0.f, 0.f, 0.f));   // linVelA - linVelB
        // calculateTemporalAabb(R, linVel, 0, 1)
        f3 castMin, castMax;
        {
            Xf R = fromT;
            R.o = mk3(0.f, 0.f, 0.f);
            shapeAabb(castS, R, castMin, castMax);
            f3 lin = sub3(to, from);
            lin = scl3(lin, 1.f / 1.f);
            lin = scl3(lin, 1.f);
            if (lin.x > 0.f) castMax.x += lin.x; else castMin.x += lin.x;
            if (lin.y > 0.f) castMax.y += lin.y; else castMin.y += lin.y;
            if (lin.z > 0.f) castMax.z += lin.z; else castMin.z += lin.z;
            const f3 am = mk3(0.f, 0.f, 0.f);
            castMin = sub3(castMin, am);
            castMax = add3(castMax, am);
        }
        SweepBest best;
        best.fraction = 1.f;
        best.body = -1;
        best.rec.fraction = 1.f;
        for (int d = 0; d < 3; d++) best.rec.normal[d] = best.rec.point[d] = 0.f;
        __syncthreads();
        if (threadIdx.x == 0) { sCount = 0; sTotal = 0; sUnsupported = 0; }

        // (2) the casts of the listed candidates, strided over the threads
        auto flush = [&]() {
            __syncthreads();
            const uint32_t cnt = sCount;
            for (uint32_t k = threadIdx.x; k < cnt; k += SWEEP_THREADS) {
                const int i = (int)sCand[k];
                const ShapeDev s = shapes[B.shape[i]];
                const Xf t = loadXf(B.xf4, i);
                bool valid = false;
                SweepRec rec;
                rec.fraction = 1.f;
                rec.normal[0] = rec.normal[1] = rec.normal[2] = 0.f;
                rec.point[0] = rec.point[1] = rec.point[2] = 0.f;
                auto castConvex = [&](const ShapeDev& cs, const Xf& cx) {
                    const AnyS shp = anyShapeOf(cs, hullPts);
                    float f = 1.f;
                    f3 nn, pt;
                    if (gjkConvexCast(A, fromT, to, shp, cx, allowedPenetration, f, nn, pt)) {
                        if (len2_3(nn) > 0.0001f && f < rec.fraction) {
                            nn = nor3(nn);
                            if (nm.me && dot3(nn, rel) >= -0.f) return;   // ClosestNotMe...addSingleResult (:1157)
                            valid = true;
                            rec.fraction = f;
                            rec.normal[0] = nn.x; rec.normal[1] = nn.y; rec.normal[2] = nn.z;
                            rec.point[0] = pt.x; rec.point[1] = pt.y; rec.point[2] = pt.z;
                        }
                    }
                };
                if (s.type == SH_BOX || s.type == SH_SPHERE || s.type == SH_HULL) {
                    castConvex(s, t);
                } else if (s.type == SH_COMPOUND) {
                    for (int ch = 0; ch < s.numPoints; ch++) {
                        const CompoundChildDev& cd = children[s.pointOffset + ch];
                        castConvex(shapes[cd.shape], compoundChildWorld(t, children, cd));
                    }
                } else if (s.type == SH_MESH) {
                    Xf inv;  // Transform.inverse (lm/Transform.java:101-105)
                    for (int r = 0; r < 3; r++)
                        for (int c = 0; c < 3; c++) inv.m[r][c] = t.m[c][r];
                    inv.o = mulMV(inv.m, neg3(t.o));
                    Xf rot;  // MeshRotation^-1 * ConvexToRotation, zero origin
                    mulMM(inv.m, fromT.m, rot.m);
                    rot.o = mk3(0.f, 0.f, 0.f);
                    f3 boxMin, boxMax;
                    shapeAabb(castS, rot, boxMin, boxMax);
                    sweepMeshWalk(meshes[s.mesh], A, fromT, to, t, s.margin, xfPoint(inv, from), xfPoint(inv, to), boxMin, boxMax, valid, rec, nm.me != nullptr,
                                  rel);
                } else if (s.type == SH_PLANE) {
                    sUnsupported = 1;
                }
                best.offer(valid, i, rec);
            }
            __syncthreads();
            if (threadIdx.x == 0) sCount = 0;
            __syncthreads();
        };

        // (1) candidates, in rounds
        const int nChunks = (n + RAY_CHUNK - 1) / RAY_CHUNK;
        for (int c0 = 0; c0 < nChunks; c0 += SWEEP_HIT_CAP) {
            __syncthreads();
            if (threadIdx.x == 0) sHitCount = 0;
            __syncthreads();
            const int c1 = min(nChunks, c0 + SWEEP_HIT_CAP);
            for (int c = c0 + threadIdx.x; c < c1; c += SWEEP_THREADS) {
                const float4 cn = __ldg(cmin + c);
                if (cn.w == 0.f) continue;
                const float4 cx = __ldg(cmax + c);
                if (rayAabb(from, to, add3(mk3(cn.x, cn.y, cn.z), castMin), add3(mk3(cx.x, cx.y, cx.z), castMax), 1.f))
                    sHit[atomicAdd(&sHitCount, 1u)] = (uint32_t)c;
            }
            __syncthreads();
            const uint32_t nHit = sHitCount;
            uint32_t h = 0;
            while (h < nHit) {
                const uint32_t room = (SWEEP_MAX_CAND - sCount) / RAY_CHUNK;   // hit chunks whose members surely fit
                if (room == 0) { flush(); continue; }
                const uint32_t g = min(nHit - h, room);
                __syncthreads();   // everyone has read sCount
                for (uint32_t w = threadIdx.x; w < g * RAY_CHUNK; w += SWEEP_THREADS) {
                    const int pos = (int)sHit[h + w / RAY_CHUNK] * RAY_CHUNK + (int)(w % RAY_CHUNK);
                    if (pos >= n) continue;
                    const int i = rayBodyAt(sortedMin, nSorted, pos);
                    if (i < 0 || i >= n) continue;
                    const float4 mn = __ldg(rmin + i);
                    if (mn.w == 0.f) continue;
                    if (i == me) continue;                           // ClosestNotMe...needsCollision (:1169)
                    if (!filterPass(cbFilter, B.filt[i])) continue;  // ConvexResultCallback.needsCollision (:752-756)
                    if (nm.me && sweepAlreadyTouching(nm, B, shapes, me, i)) continue;
                    const float4 mx = __ldg(rmax + i);
                    if (rayAabb(from, to, add3(mk3(mn.x, mn.y, mn.z), castMin), add3(mk3(mx.x, mx.y, mx.z), castMax), 1.f)) {
                        sCand[atomicAdd(&sCount, 1u)] = (uint32_t)i;
                        atomicAdd(&sTotal, 1u);
                    }
                }
                __syncthreads();
                h += g;
            }
        }
        flush();

        // (3) smallest fraction, lowest body index among equals
        sBestFrac[threadIdx.x] = best.fraction;
        sBestBody[threadIdx.x] = best.body;
        sBestRec[threadIdx.x] = best.rec;
        __syncthreads();
        if (threadIdx.x == 0) {
            float closest = 1.f;
            int hitBody = -1;
            int bk = 0;
            for (int k = 0; k < SWEEP_THREADS; k++) {
                const int i = sBestBody[k];
                if (i < 0) continue;
                const float f = sBestFrac[k];
                if (f < closest || (hitBody >= 0 && f == closest && i < hitBody)) { closest = f; hitBody = i; bk = k; }
            }
            RayOut o;
            o.uid = sUnsupported ? -1 : hitBody + 1;
            o.fraction = closest;
            for (int d = 0; d < 3; d++) {
                o.normal[d] = hitBody >= 0 ? sBestRec[bk].normal[d] : 0.f;
                o.point[d] = hitBody >= 0 ? sBestRec[bk].point[d] : 0.f;
            }
            out[sw] = o;
            atomicMax(maxCandidates, sTotal);
        }
    }
}

}  // namespace b2c
