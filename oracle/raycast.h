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
// Concave and compound shapes (rayTestSingle's other branches, disp/CollisionWorld.java:301-356):
//   np/TriangleRaycastCallback.java:46-117 processTriangle + BridgeTriangleRaycastCallback (:807-830);
//   sh/BvhTriangleMeshShape.java:135-142 performRaycast -> sh/OptimizedBvh.java:999-1015 reportRayOverlappingNodex
//       -> :817-931 walkStacklessQuantizedTreeAgainstRay (quantised box prune, then rayAabb on the unquantised node box);
//   sh/StaticPlaneShape.java:60-122 processAllTriangles (two triangles sized by the ray's local AABB, including the
//       reference's `set(aabbMax).set(aabbMin)` half-extent quirk at :67) with lm/TransformUtil.java:45-61 planeSpace1;
//   compound: every child in order with colObjWorldTransform * childTrans (:333-352).
#pragma once
#include "bvh.h"
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

// np/TriangleRaycastCallback.java:46-117 with reportHit = BridgeTriangleRaycastCallback.reportHit ->
// ClosestRayResultCallback.addSingleResult (normalInWorldSpace = false): returns the new hitFraction.
struct TriangleRaycast {
    V3 from, to;           // in the concave object's local space
    float hitFraction = 1.f;
    bool hit = false;      // at least one triangle was reported
    V3 hitNormalLocal;     // of the last reported triangle (not normalised, as in the reference)
    int triangleIndex = -1;
    void processTriangle(const V3 tri[3], int partId, int triIndex) {
        (void)partId;
        const V3& vert0 = tri[0]; const V3& vert1 = tri[1]; const V3& vert2 = tri[2];
        V3 v10; v10.set(vert1).sub(vert0);
        V3 v20; v20.set(vert2).sub(vert0);
        V3 triangleNormal; triangleNormal.set(v10).crs(v20);
        const float dist = vert0.dot(triangleNormal);
        float dist_a = triangleNormal.dot(from);
        dist_a -= dist;
        float dist_b = triangleNormal.dot(to);
        dist_b -= dist;
        if (dist_a * dist_b >= 0.f) return;  // same sign
        const float proj_length = dist_a - dist_b;
        const float distance = (dist_a) / (proj_length);
        if (distance < hitFraction) {
            float edge_tolerance = triangleNormal.len2();
            edge_tolerance *= -0.0001f;
            V3 point;
            { float s = 1.f - distance; point.set(s * from.x + distance * to.x, s * from.y + distance * to.y, s * from.z + distance * to.z); }
            V3 v0p; v0p.set(vert0).sub(point);
            V3 v1p; v1p.set(vert1).sub(point);
            V3 cp0; cp0.set(v0p).crs(v1p);
            if (cp0.dot(triangleNormal) >= edge_tolerance) {
                V3 v2p; v2p.set(vert2).sub(point);
                V3 cp1; cp1.set(v1p).crs(v2p);
                if (cp1.dot(triangleNormal) >= edge_tolerance) {
                    V3 cp2; cp2.set(v2p).crs(v0p);
                    if (cp2.dot(triangleNormal) >= edge_tolerance) {
                        hit = true;
                        triangleIndex = triIndex;
                        if (dist_a > 0.f) hitNormalLocal.set(triangleNormal);
                        else { hitNormalLocal.set(triangleNormal); hitNormalLocal.scl(-1.f); }
                        hitFraction = distance;  // reportHit returns rayResult.hitFraction
                    }
                }
            }
        }
    }
};

// sh/OptimizedBvh.java:817-931 walkStacklessQuantizedTreeAgainstRay with zero box-cast extents (:999-1005)
// and with the box-cast extents of sh/OptimizedBvh.java:1017-1036 reportBoxCastOverlappingNodex
template <class F>
static inline void bvhReportBoxCastOverlappingNodex(const Bvh& bvh, const V3& raySource, const V3& rayTarget, const V3& aabbMin,
                                                    const V3& aabbMax, F cb);
template <class F>
static inline void bvhReportRayOverlappingNodex(const Bvh& bvh, const V3& raySource, const V3& rayTarget, F cb) {
    V3 zero(0, 0, 0);
    bvhReportBoxCastOverlappingNodex(bvh, raySource, rayTarget, zero, zero, cb);
}
template <class F>
static inline void bvhReportBoxCastOverlappingNodex(const Bvh& bvh, const V3& raySource, const V3& rayTarget, const V3& aabbMin,
                                                    const V3& aabbMax, F cb) {
    V3 rayAabbMin(jminf(raySource.x, rayTarget.x), jminf(raySource.y, rayTarget.y), jminf(raySource.z, rayTarget.z));
    V3 rayAabbMax(jmaxf(raySource.x, rayTarget.x), jmaxf(raySource.y, rayTarget.y), jmaxf(raySource.z, rayTarget.z));
    rayAabbMin.add(aabbMin);  // :856-858 add box cast extents to bounding box
    rayAabbMax.add(aabbMax);
    uint16_t qmin[3], qmax[3];
    bvh.quantizeWithClamp(rayAabbMin, qmin);
    bvh.quantizeWithClamp(rayAabbMax, qmax);
    int curIndex = 0;
    const int endNodeIndex = bvh.curNodeIndex;
    while (curIndex < endNodeIndex) {
        const QNode& n = bvh.nodes[curIndex];
        float param = 1.f;
        bool rayBoxOverlap = false;
        bool boxBoxOverlap = true;  // :563-585
        boxBoxOverlap = (qmin[0] > n.mx[0] || qmax[0] < n.mn[0]) ? false : boxBoxOverlap;
        boxBoxOverlap = (qmin[2] > n.mx[2] || qmax[2] < n.mn[2]) ? false : boxBoxOverlap;
        boxBoxOverlap = (qmin[1] > n.mx[1] || qmax[1] < n.mn[1]) ? false : boxBoxOverlap;
        const bool isLeaf = n.isLeaf();
        if (boxBoxOverlap) {
            V3 b0 = bvh.unQuantize(n.mn), b1 = bvh.unQuantize(n.mx);
            b0.add(aabbMin);  // :901-903 add box cast extents
            b1.add(aabbMax);
            V3 normal;
            rayBoxOverlap = rayAabb(raySource, rayTarget, b0, b1, param, normal);
        }
        if (isLeaf && rayBoxOverlap) {
            int v = n.escapeIndexOrTriangleIndex;
            int tri = v & ~((~0) << (31 - Bvh::MAX_NUM_PARTS_IN_BITS));
            int part = (int)((uint32_t)v >> (31 - Bvh::MAX_NUM_PARTS_IN_BITS));
            cb(part, tri);
        }
        if (rayBoxOverlap || isLeaf) curIndex++;
        else curIndex += -n.escapeIndexOrTriangleIndex;
    }
}

// lm/TransformUtil.java:45-61
static inline void planeSpace1(const V3& n, V3& p, V3& q) {
    const float SIMDSQRT12 = 0.7071067811865475244008443621048490f;
    if (jabsf(n.z) > SIMDSQRT12) {
        float a = n.y * n.y + n.z * n.z;
        float k = 1.f / jsqrt(a);
        p.set(0, -n.z * k, n.y * k);
        q.set(a * k, -n.x * p.z, n.x * p.y);
    } else {
        float a = n.x * n.x + n.y * n.y;
        float k = 1.f / jsqrt(a);
        p.set(-n.y * k, n.x * k, 0);
        q.set(-n.z * p.y, n.z * p.x, a * k);
    }
}

// sh/StaticPlaneShape.java:60-122 processAllTriangles(callback, aabbMin, aabbMax)
template <class F>
static inline void planeProcessAllTriangles(const V3& planeNormal, float planeConstant, const V3& aabbMin, const V3& aabbMax, F cb) {
    V3 halfExtents; halfExtents.set(aabbMax); halfExtents.set(aabbMin);  // :67 (sic)
    halfExtents.scl(0.5f);
    const float radius = halfExtents.len();
    V3 center; center.set(aabbMax).add(aabbMin);
    center.scl(0.5f);
    V3 t0, t1;
    planeSpace1(planeNormal, t0, t1);
    V3 tmp; tmp.set(planeNormal).scl(planeNormal.dot(center) - planeConstant);
    V3 projectedCenter; projectedCenter.set(center).sub(tmp);
    V3 tmp1, tmp2;
    tmp1.set(t0).scl(radius);
    tmp2.set(t1).scl(radius);
    V3 tri[3];
    tri[0].set(projectedCenter.x + tmp1.x + tmp2.x, projectedCenter.y + tmp1.y + tmp2.y, projectedCenter.z + tmp1.z + tmp2.z);
    tmp.set(tmp1).sub(tmp2);
    tri[1].set(projectedCenter.x + tmp.x, projectedCenter.y + tmp.y, projectedCenter.z + tmp.z);
    tri[2].set(projectedCenter).sub(tmp);
    cb(tri, 0, 0);
    tmp.set(tmp1).sub(tmp2);
    tri[0].set(projectedCenter).sub(tmp);
    tmp.set(tmp1).add(tmp2);
    tri[1].set(projectedCenter).sub(tmp);
    tri[2].set(projectedCenter.x + tmp1.x + tmp2.x, projectedCenter.y + tmp1.y + tmp2.y, projectedCenter.z + tmp1.z + tmp2.z);
    cb(tri, 0, 1);
}

}  // namespace orc
