// ORACLE (test infrastructure only; never linked into the product) — "parity unpinned": the reference ships no tests.
//
// raycast.h — CPU restatement of CollisionWorld.rayTest with a ClosestRayResultCallback for convex shapes
// (SURVEY §8f rank 4):
//   disp/CollisionWorld.java:553-590 rayTest: objects in world order, needsCollision filter on the callback's group/mask
//       (disp/CollisionWorld.java:664-670), shape AABB under the object's transform, AabbUtil2.rayAabb with the running
//       closestHitFraction as the exit bound, then rayTestSingle;
//   lm/AabbUtil2.java:40-108 outcode / rayAabb;
//   disp/CollisionWorld.java:260-300 rayTestSingle, convex branch: SubsimplexConvexCast of a zero-radius, zero-margin
//       sphere against the shape, accepted when normal.len2 > 1e-4 and fraction < closestHitFraction;
//   np/SubsimplexConvexCast.java:63-190 calcTimeOfImpact (<= 32 iterations, epsilon 1e-4) on the Voronoi simplex solver;
//   disp/CollisionWorld.java:697-729 ClosestRayResultCallback.addSingleResult.
// Concave shapes (planes, meshes) are not cast here: the ray passes through them (stated in include/b2c.h).
#pragma once
#include "jmath.h"
#include "shapes.h"
#include "voronoi.h"

namespace orc {

static inline int rayOutcode(const V3& p, const V3& h) {  // lm/AabbUtil2.java:40-43
    return (p.x < -h.x ? 0x01 : 0) | (p.x > h.x ? 0x08 : 0) | (p.y < -h.y ? 0x02 : 0) | (p.y > h.y ? 0x10 : 0) |
           (p.z < -h.z ? 0x04 : 0) | (p.z > h.z ? 0x20 : 0);
}

// lm/AabbUtil2.java:45-108; param is in/out (the entry parameter on success)
static inline bool rayAabb(const V3& rayFrom, const V3& rayTo, const V3& aabbMin, const V3& aabbMax, float& param, V3& normal) {
    V3 he; he.set(aabbMax).sub(aabbMin); he.scl(0.5f);
    V3 ce; ce.set(aabbMax).add(aabbMin); ce.scl(0.5f);
    V3 source; source.set(rayFrom).sub(ce);
    V3 target; target.set(rayTo).sub(ce);
    const int so = rayOutcode(source, he), to = rayOutcode(target, he);
    if ((so & to) == 0) {
        float lambdaEnter = 0.f, lambdaExit = param;
        V3 r; r.set(target).sub(source);
        float normSign = 1.f;
        V3 hitNormal(0, 0, 0);
        int bit = 1;
        for (int j = 0; j < 2; j++) {
            for (int i = 0; i != 3; ++i) {
                if (so & bit) {
                    float lambda = (-source.get(i) - he.get(i) * normSign) / r.get(i);
                    if (lambdaEnter <= lambda) {
                        lambdaEnter = lambda;
                        hitNormal.set(0, 0, 0);
                        hitNormal.setc(i, normSign);
                    }
                } else if (to & bit) {
                    float lambda = (-source.get(i) - he.get(i) * normSign) / r.get(i);
                    lambdaExit = jminf(lambdaExit, lambda);
                }
                bit <<= 1;
            }
            normSign = -1.f;
        }
        if (lambdaEnter <= lambdaExit) {
            param = lambdaEnter;
            normal.set(hitNormal);
            return true;
        }
    }
    return false;
}

struct CastResult {
    V3 normal;
    float fraction = 1e30f;
    float allowedPenetration = 0.f;
    int iterations = 0;
};

// np/SubsimplexConvexCast.java:63-190 with convexA = SphereShape(0) with margin 0 (support (0,0,0)), fromA/toA = identity
// basis at the ray end points, fromB = toB = the object's transform
static inline bool rayConvexCast(const V3& rayFrom, const V3& rayTo, const Shape& shapeB, const Xf& xfB, CastResult& result) {
    VoronoiSimplexSolver simplex;
    simplex.reset();
    V3 linVelA; linVelA.set(rayTo).sub(rayFrom);
    V3 linVelB; linVelB.set(xfB.origin).sub(xfB.origin);
    float lambda = 0.f;
    V3 originA = rayFrom;            // interpolatedTransA.origin (basis stays identity)
    Xf interpB; interpB.set(xfB);
    V3 r; r.set(linVelA).sub(linVelB);
    V3 v, tmp;
    // supVertexA = fromA.transform(localSupport) = I*(0,0,0) + origin
    V3 supA(0, 0, 0);
    { V3 z(0, 0, 0); supA.set(z.x * 1.f + z.y * 0.f + z.z * 0.f, z.x * 0.f + z.y * 1.f + z.z * 0.f, z.x * 0.f + z.y * 0.f + z.z * 1.f); supA.add(rayFrom); }
    V3 supB;
    transposeTransform(tmp, r, xfB.basis);
    localGetSupportingVertex(shapeB, tmp, supB);
    xfB.transform(supB);
    v.set(supA).sub(supB);
    int maxIter = 32;
    V3 n(0, 0, 0);
    float dist2 = v.len2();
    const float epsilon = 0.0001f;
    V3 w;
    result.iterations = 0;
    while ((dist2 > epsilon) && (maxIter--) != 0) {
        result.iterations++;
        // convexA's support is (0,0,0) whatever the direction; transformed by the interpolated A (identity basis)
        { V3 z(0, 0, 0); supA.set(z.x * 1.f + z.y * 0.f + z.z * 0.f, z.x * 0.f + z.y * 1.f + z.z * 0.f, z.x * 0.f + z.y * 0.f + z.z * 1.f); supA.add(originA); }
        transposeTransform(tmp, v, interpB.basis);
        localGetSupportingVertex(shapeB, tmp, supB);
        interpB.transform(supB);
        w.set(supA).sub(supB);
        float VdotW = v.dot(w);
        if (lambda > 1.f) return false;
        if (VdotW > 0.f) {
            float VdotR = v.dot(r);
            if (VdotR >= -(FLT_EPSILON_ * FLT_EPSILON_)) return false;
            lambda = lambda - VdotW / VdotR;
            // setInterpolate3 (lm/VectorUtil.java:137-141)
            {
                float s = 1.f - lambda;
                originA.set(s * rayFrom.x + lambda * rayTo.x, s * rayFrom.y + lambda * rayTo.y, s * rayFrom.z + lambda * rayTo.z);
                interpB.origin.set(s * xfB.origin.x + lambda * xfB.origin.x, s * xfB.origin.y + lambda * xfB.origin.y,
                                   s * xfB.origin.z + lambda * xfB.origin.z);
            }
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
    return true;
}

struct RayHit {
    int uid = 0;  // 0 = no hit
    float fraction = 1.f;
    V3 normal, point;
};

}  // namespace orc
