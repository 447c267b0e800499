// ORACLE (test infrastructure only; never linked into the product) — "parity unpinned": the reference ships no tests.
//
// convexcast.h — CPU restatement of CollisionWorld.convexSweepTest with a ClosestConvexResultCallback,
// for TRANSLATIONAL sweeps (convexFromWorld.basis == convexToWorld.basis):
//   disp/CollisionWorld.java:596-651 convexSweepTest: cast-shape temporal AABB, objects in world order, needsCollision
//       filter (:752-756), object AABB expanded by the cast AABB (lm/AabbUtil2.java:35-38), rayAabb with exit bound 1
//       (:634 "could use closestHitFraction, but needs testing"), then objectQuerySingle;
//   sh/CollisionShape.java:74-122 calculateTemporalAabb.  TransformUtil.calculateVelocity derives the angular velocity from
//       the libgdx quaternion of toBasis * fromBasis^-1 (lm/TransformUtil.java:105-153), whose float sequence is not
//       available here (libgdx is not vendored); for equal bases it is rounding noise times the shape's angular motion
//       disc.  This restatement takes it as exactly zero — the culling box differs from the reference's by that noise,
//       the casts themselves do not depend on it;
//   disp/CollisionWorld.java:392-428 objectQuerySingle, convex branch: np/GjkConvexCast.java:66-196 (conservative
//       advancement on GjkPairDetector without a penetration solver, radius 0.001, <= 32 iterations, ONE PointCollector
//       for the whole cast, np/PointCollector.java:44-52), accepted when normal.len2 > 1e-4 and fraction < closest;
//   :429-456 triangle mesh: the cast box in mesh space, sh/BvhTriangleMeshShape.java:144-152 performConvexcast ->
//       sh/OptimizedBvh.java:1017-1036,817-931 box-cast walk, np/TriangleConvexcastCallback.java:53-88 per triangle:
//       np/SubsimplexConvexCast.java:63-190 (cast shape vs TriangleShape with the mesh margin), hits with fraction < the
//       bound at entry reported through BridgeTriangleConvexcastCallback.reportHit (:362-388, fraction <= running closest);
//   :457-497 static plane: the reference dereferences a null caster (FIXME at :470) — throws; reported as unsupported;
//   :528-545 compound: every child in order with colObjWorldTransform * childTrans;
//   :765-800 ClosestConvexResultCallback.addSingleResult.
#pragma once
#include "gjk.h"
#include "raycast.h"

namespace orc {

struct ConvexCastResult {
    V3 normal, hitPoint;
    float fraction = 1e30f;
    float allowedPenetration = 0.f;
};

// np/GjkConvexCast.java:66-196 calcTimeOfImpact; B does not move (fromB == toB)
static inline bool gjkConvexCast(const Shape& convexA, const Xf& fromA, const Xf& toA, const Shape& convexB, const Xf& xfB,
                                 ConvexCastResult& result) {
    V3 linVelA; linVelA.set(toA.origin).sub(fromA.origin);
    V3 linVelB; linVelB.set(xfB.origin).sub(xfB.origin);
    const float radius = 0.001f;
    float lambda = 0.f;
    const int maxIter = 32;
    V3 n(0, 0, 0), c;
    V3 r; r.set(linVelA).sub(linVelB);
    float lastLambda = lambda;
    int numIter = 0;
    // the PointCollector lives for the whole cast: a later query only replaces it with a SMALLER distance (:44-52)
    bool hasResult = false;
    V3 pcNormal, pcPoint;
    float pcDistance = 1e30f;
    auto query = [&](const Xf& ta, const Xf& tb) {
        GjkOut o;
        gjkGetClosestPoints(&convexA, &convexB, ta, tb, SIMD_INFINITY_, o, false);  // input.init(): maximumDistanceSquared = Float.MAX_VALUE
        if (o.hasContact && o.depth < pcDistance) {
            hasResult = true;
            pcNormal.set(o.normalOnBInWorld);
            pcPoint.set(o.pointInWorld);
            pcDistance = o.depth;
        }
    };
    Xf inA; inA.set(fromA);
    Xf inB; inB.set(xfB);
    query(inA, inB);
    c.set(pcPoint);
    if (!hasResult) return false;
    float dist = pcDistance;
    n.set(pcNormal);
    while (dist > radius) {
        numIter++;
        if (numIter > maxIter) return false;
        const float projectedLinearVelocity = r.dot(n);
        const float dLambda = dist / projectedLinearVelocity;
        lambda = lambda - dLambda;
        if (lambda > 1.f) return false;
        if (lambda < 0.f) return false;
        if (lambda <= lastLambda) return false;
        lastLambda = lambda;
        {   // VectorUtil.setInterpolate3 (lm/VectorUtil.java:137-141)
            const float s = 1.f - lambda;
            inA.origin.set(s * fromA.origin.x + lambda * toA.origin.x, s * fromA.origin.y + lambda * toA.origin.y,
                           s * fromA.origin.z + lambda * toA.origin.z);
            inB.origin.set(s * xfB.origin.x + lambda * xfB.origin.x, s * xfB.origin.y + lambda * xfB.origin.y,
                           s * xfB.origin.z + lambda * xfB.origin.z);
        }
        query(inA, inB);
        // pointCollector.hasResult stays true once set
        if (pcDistance < 0.f) {
            result.fraction = lastLambda;
            n.set(pcNormal);
            result.normal.set(n);
            result.hitPoint.set(pcPoint);
            return true;
        }
        c.set(pcPoint);
        n.set(pcNormal);
        dist = pcDistance;
    }
    if (n.dot(r) >= -result.allowedPenetration) return false;
    result.fraction = lambda;
    result.normal.set(n);
    result.hitPoint.set(c);
    return true;
}

// np/SubsimplexConvexCast.java:63-190 calcTimeOfImpact in full generality (the ray test's copy in raycast.h has convexA = a
// zero sphere folded in)
static inline bool subsimplexConvexCast(const Shape& convexA, const Xf& fromA, const Xf& toA, const Shape& convexB, const Xf& fromB,
                                        const Xf& toB, ConvexCastResult& result) {
    VoronoiSimplexSolver simplex;
    simplex.reset();
    V3 tmp;
    V3 linVelA; linVelA.set(toA.origin).sub(fromA.origin);
    V3 linVelB; linVelB.set(toB.origin).sub(fromB.origin);
    float lambda = 0.f;
    Xf interpA; interpA.set(fromA);
    Xf interpB; interpB.set(fromB);
    V3 r; r.set(linVelA).sub(linVelB);
    V3 v;
    tmp.set(r).scl(-1.f);
    transposeTransform(tmp, tmp, fromA.basis);
    V3 supA; localGetSupportingVertex(convexA, tmp, supA);
    fromA.transform(supA);
    transposeTransform(tmp, r, fromB.basis);
    V3 supB; localGetSupportingVertex(convexB, tmp, supB);
    fromB.transform(supB);
    v.set(supA).sub(supB);
    int maxIter = 32;
    V3 n(0, 0, 0);
    float dist2 = v.len2();
    const float epsilon = 0.0001f;
    V3 w;
    while ((dist2 > epsilon) && (maxIter--) != 0) {
        tmp.set(v).scl(-1.f);
        transposeTransform(tmp, tmp, interpA.basis);
        localGetSupportingVertex(convexA, tmp, supA);
        interpA.transform(supA);
        transposeTransform(tmp, v, interpB.basis);
        localGetSupportingVertex(convexB, tmp, supB);
        interpB.transform(supB);
        w.set(supA).sub(supB);
        const float VdotW = v.dot(w);
        if (lambda > 1.f) return false;
        if (VdotW > 0.f) {
            const float VdotR = v.dot(r);
            if (VdotR >= -(FLT_EPSILON_ * FLT_EPSILON_)) return false;
            lambda = lambda - VdotW / VdotR;
            const float s = 1.f - lambda;
            interpA.origin.set(s * fromA.origin.x + lambda * toA.origin.x, s * fromA.origin.y + lambda * toA.origin.y,
                               s * fromA.origin.z + lambda * toA.origin.z);
            interpB.origin.set(s * fromB.origin.x + lambda * toB.origin.x, s * fromB.origin.y + lambda * toB.origin.y,
                               s * fromB.origin.z + lambda * toB.origin.z);
            w.set(supA).sub(supB);
            n.set(v);
        }
        simplex.addVertex(w, supA, supB);
        if (simplex.closest(v)) dist2 = v.len2();
        else dist2 = 0.f;
    }
    result.fraction = lambda;
    if (n.len2() >= FLT_EPSILON_ * FLT_EPSILON_) { result.normal.set(n); result.normal.nor(); }
    else result.normal.set(0, 0, 0);
    if (result.normal.dot(r) >= -result.allowedPenetration) return false;
    V3 hitA, hitB;
    simplex.compute_points(hitA, hitB);
    result.hitPoint.set(hitB);
    return true;
}

struct ConvexSweepHit {
    int uid = 0;              // 0 = no hit (hitCollisionObject == null)
    float fraction = 1.f;     // closestHitFraction
    V3 normal, point;         // hitNormalWorld, hitPointWorld
    bool unsupported = false; // the sweep met a static plane: the reference throws there (disp/CollisionWorld.java:470-495)
};

}  // namespace orc
