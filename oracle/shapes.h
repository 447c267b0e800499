// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// shapes.h — restatement of the collision shapes on the hot path: support mappings and AABBs.
// Reference files are under /root/reference/src/com/bulletphysics/collision/shapes/ (sh/)
// and linearmath/ (lm/).
#pragma once
#include <vector>
#include "jmath.h"

namespace orc {

enum ShapeType { SH_BOX = 0, SH_SPHERE = 1, SH_HULL = 2, SH_TRIANGLE = 3, SH_PLANE = 4, SH_MESH = 5, SH_COMPOUND = 6 };

struct Bvh;  // bvh.h

struct CompoundChild {  // sh/CompoundShapeChild.java:34-39
    Xf transform;
    int shape = -1;      // index of the child shape in World::shapes (childShape)
};

struct Shape {
    int type = SH_BOX;
    V3 implicitDims;               // sh/ConvexInternalShape.java:42 implicitShapeDimensions
    V3 localScaling{1, 1, 1};      // sh/ConvexInternalShape.java:41
    float collisionMargin = CONVEX_DISTANCE_MARGIN;  // sh/ConvexInternalShape.java:43
    std::vector<V3> points;        // hull points (sh/ConvexHullShape.java:38)
    V3 localAabbMin{1, 1, 1}, localAabbMax{-1, -1, -1};  // sh/PolyhedralConvexShape.java:57-58
    V3 tri[3];                     // sh/TriangleShape.java vertices1
    V3 planeNormal; float planeConstant = 0;  // sh/StaticPlaneShape.java:41-42
    Bvh* bvh = nullptr;            // sh/BvhTriangleMeshShape.java:50
    std::vector<CompoundChild> children;  // sh/CompoundShape.java:43 (localAabbMin/Max above double as :44-45)

    bool isConvex() const { return type == SH_BOX || type == SH_SPHERE || type == SH_HULL || type == SH_TRIANGLE; }
    bool isConcave() const { return type == SH_MESH || type == SH_PLANE; }  // bp/BroadphaseNativeType.java:95-98
    bool isCompound() const { return type == SH_COMPOUND; }                  // bp/BroadphaseNativeType.java:104-106

    // sh/SphereShape.java:83-97 ; sh/ConvexInternalShape.java:111 ; sh/ConcaveShape.java:39
    float getMargin() const {
        if (type == SH_SPHERE) return implicitDims.x * localScaling.x;
        return collisionMargin;
    }
};

// lm/ScalarUtil.java:35-37
static inline float fsel(float a, float b, float c) { return a >= 0 ? b : c; }

// lm/VectorUtil.java:41-58 maxAxis (first strict max, init -1e30)
static inline int maxAxis(const V3& v) {
    int maxIndex = -1;
    float maxVal = -1e30f;
    if (v.x > maxVal) { maxIndex = 0; maxVal = v.x; }
    if (v.y > maxVal) { maxIndex = 1; maxVal = v.y; }
    if (v.z > maxVal) { maxIndex = 2; maxVal = v.z; }
    return maxIndex;
}

// sh/BoxShape.java:46-50 constructor
static inline void initBox(Shape& s, const V3& boxHalfExtents) {
    s.type = SH_BOX;
    V3 margin(s.getMargin(), s.getMargin(), s.getMargin());
    s.implicitDims.set(boxHalfExtents.x * s.localScaling.x, boxHalfExtents.y * s.localScaling.y,
                       boxHalfExtents.z * s.localScaling.z);
    s.implicitDims.sub(margin);
}
// sh/SphereShape.java:38-41
static inline void initSphere(Shape& s, float radius) {
    s.type = SH_SPHERE;
    s.implicitDims.x = radius;
    s.collisionMargin = radius;
}

static void localGetSupportingVertexWithoutMargin(const Shape& s, const V3& vec0, V3& out);

// sh/PolyhedralConvexShape.java:177-201 recalcLocalAabb via batched support on the 6 axis directions
// (sh/ConvexHullShape.java:105-139: per direction the first strict max of dir . (p*scaling)).
static inline void recalcLocalAabbHull(Shape& s) {
    static const V3 dirs[6] = {V3(1, 0, 0), V3(0, 1, 0), V3(0, 0, 1), V3(-1, 0, 0), V3(0, -1, 0), V3(0, 0, -1)};
    V3 sup[6];
    float w[6];
    for (int j = 0; j < 6; j++) w[j] = -1e30f;
    for (size_t i = 0; i < s.points.size(); i++) {
        V3 vtx(s.points[i].x * s.localScaling.x, s.points[i].y * s.localScaling.y, s.points[i].z * s.localScaling.z);
        for (int j = 0; j < 6; j++) {
            float newDot = dirs[j].dot(vtx);
            if (newDot > w[j]) { sup[j].set(vtx); w[j] = newDot; }
        }
    }
    for (int i = 0; i < 3; i++) {
        s.localAabbMax.setc(i, sup[i].get(i) + s.collisionMargin);
        s.localAabbMin.setc(i, sup[i + 3].get(i) - s.collisionMargin);
    }
}
// sh/ConvexHullShape.java:45-53
static inline void initHull(Shape& s, const float* pts, int n) {
    s.type = SH_HULL;
    s.points.clear();
    for (int i = 0; i < n; i++) s.points.push_back(V3(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]));
    recalcLocalAabbHull(s);
}
// sh/StaticPlaneShape.java:45-48
static inline void initPlane(Shape& s, const V3& n, float c) {
    s.type = SH_PLANE;
    s.collisionMargin = 0.0f;  // sh/ConcaveShape.java:35
    s.planeNormal.set(n).nor();
    s.planeConstant = c;
}

// Support mapping without margin.
static void localGetSupportingVertexWithoutMargin(const Shape& s, const V3& vec0, V3& out) {
    switch (s.type) {
    case SH_BOX: {  // sh/BoxShape.java:89-97
        const V3& h = s.implicitDims;
        out.set(fsel(vec0.x, h.x, -h.x), fsel(vec0.y, h.y, -h.y), fsel(vec0.z, h.z, -h.z));
        return;
    }
    case SH_SPHERE:  // sh/SphereShape.java:44-47
        out.set(0, 0, 0);
        return;
    case SH_HULL: {  // sh/ConvexHullShape.java:75-102
        out.set(0, 0, 0);
        float newDot, maxDot = -1e30f;
        V3 vec = vec0;
        float lenSqr = vec.len2();
        if (lenSqr < 0.0001f) {
            vec.set(1, 0, 0);
        } else {
            float rlen = 1.0f / jsqrt(lenSqr);
            vec.scl(rlen);
        }
        for (size_t i = 0; i < s.points.size(); i++) {
            V3 vtx(s.points[i].x * s.localScaling.x, s.points[i].y * s.localScaling.y, s.points[i].z * s.localScaling.z);
            newDot = vec.dot(vtx);
            if (newDot > maxDot) { maxDot = newDot; out.set(vtx); }
        }
        return;
    }
    case SH_TRIANGLE: {  // sh/TriangleShape.java:93-100
        V3 dots(vec0.dot(s.tri[0]), vec0.dot(s.tri[1]), vec0.dot(s.tri[2]));
        int k = maxAxis(dots);
        if (k < 0) k = 0;  // Java would throw on NaN input; never hit on finite data
        out.set(s.tri[k]);
        return;
    }
    default:
        out.set(0, 0, 0);
    }
}

// Support mapping with margin.
static inline void localGetSupportingVertex(const Shape& s, const V3& vec, V3& out) {
    if (s.type == SH_BOX) {  // sh/BoxShape.java:73-86
        V3 h = s.implicitDims;
        float margin = s.getMargin();
        h.x += margin; h.y += margin; h.z += margin;
        out.set(fsel(vec.x, h.x, -h.x), fsel(vec.y, h.y, -h.y), fsel(vec.z, h.z, -h.z));
        return;
    }
    // sh/ConvexInternalShape.java:85-100 and sh/ConvexHullShape.java:142-157 (identical bodies)
    localGetSupportingVertexWithoutMargin(s, vec, out);
    if (s.getMargin() != 0.0f) {
        V3 vecnorm = vec;
        if (vecnorm.len2() < (FLT_EPSILON_ * FLT_EPSILON_)) vecnorm.set(-1, -1, -1);
        vecnorm.nor();
        vecnorm.scl(s.getMargin());
        out.add(vecnorm);
    }
}

// lm/AabbUtil2.java:133-163 transformAabb(halfExtents, margin, t, ...)
static inline void transformAabbHE(const V3& halfExtents, float margin, const Xf& t, V3& mn, V3& mx) {
    V3 he(halfExtents.x + margin, halfExtents.y + margin, halfExtents.z + margin);
    V3 center = t.origin;
    V3 extent, tmp;
    for (int r = 0; r < 3; r++) {
        tmp.set(jabsf(t.basis.m[r][0]), jabsf(t.basis.m[r][1]), jabsf(t.basis.m[r][2]));
        extent.setc(r, tmp.dot(he));
    }
    mn.set(center).sub(extent);
    mx.set(center).add(extent);
}
// lm/AabbUtil2.java:165-209 transformAabb(localMin, localMax, margin, trans, ...)
static inline void transformAabbMM(const V3& lmin, const V3& lmax, float margin, const Xf& t, V3& mn, V3& mx) {
    V3 he; he.set(lmax).sub(lmin); he.scl(0.5f);
    he.x += margin; he.y += margin; he.z += margin;
    V3 lc; lc.set(lmax).add(lmin); lc.scl(0.5f);
    V3 center = lc;
    t.transform(center);
    V3 extent, tmp;
    for (int r = 0; r < 3; r++) {
        tmp.set(jabsf(t.basis.m[r][0]), jabsf(t.basis.m[r][1]), jabsf(t.basis.m[r][2]));
        extent.setc(r, tmp.dot(he));
    }
    mn.set(center).sub(extent);
    mx.set(center).add(extent);
}

// CollisionShape.getAabb(t, min, max) for each type on the path.
static inline void shapeGetAabb(const Shape& s, const Xf& t, V3& mn, V3& mx) {
    switch (s.type) {
    case SH_BOX:  // sh/BoxShape.java:147-151
        transformAabbHE(s.implicitDims, s.getMargin(), t, mn, mx);
        return;
    case SH_SPHERE: {  // sh/SphereShape.java:57-65
        V3 extent(s.getMargin(), s.getMargin(), s.getMargin());
        mn.set(t.origin).sub(extent);
        mx.set(t.origin).add(extent);
        return;
    }
    case SH_HULL:  // sh/PolyhedralConvexShape.java:169-171 (margin applied a second time: SURVEY Q8)
        transformAabbMM(s.localAabbMin, s.localAabbMax, s.getMargin(), t, mn, mx);
        return;
    case SH_COMPOUND:  // sh/CompoundShape.java:124-160: the same float sequence as lm/AabbUtil2.java:165-209 (margin 0 unless set)
        transformAabbMM(s.localAabbMin, s.localAabbMax, s.getMargin(), t, mn, mx);
        return;
    case SH_PLANE:  // sh/StaticPlaneShape.java:125-128
        mn.set(-1e30f, -1e30f, -1e30f);
        mx.set(1e30f, 1e30f, 1e30f);
        return;
    case SH_MESH: {  // sh/TriangleMeshShape.java:95-128
        V3 he; he.set(s.localAabbMax).sub(s.localAabbMin); he.scl(0.5f);
        V3 lc; lc.set(s.localAabbMax).add(s.localAabbMin); lc.scl(0.5f);
        V3 center = lc;
        t.transform(center);
        V3 extent, tmp;
        for (int r = 0; r < 3; r++) {
            tmp.set(jabsf(t.basis.m[r][0]), jabsf(t.basis.m[r][1]), jabsf(t.basis.m[r][2]));
            extent.setc(r, tmp.dot(he));
        }
        V3 margin(s.getMargin(), s.getMargin(), s.getMargin());
        extent.add(margin);
        mn.set(center).sub(extent);
        mx.set(center).add(extent);
        return;
    }
    case SH_TRIANGLE: {  // sh/TriangleShape.java:84-87 -> sh/ConvexInternalShape.java:54-82 getAabbSlow
        float margin = s.getMargin();
        for (int i = 0; i < 3; i++) {
            V3 vec(0, 0, 0), tmp1, tmp2;
            vec.setc(i, 1.0f);
            transposeTransform(tmp1, vec, t.basis);
            localGetSupportingVertex(s, tmp1, tmp2);
            t.transform(tmp2);
            mx.setc(i, tmp2.get(i) + margin);
            vec.setc(i, -1.0f);
            transposeTransform(tmp1, vec, t.basis);
            localGetSupportingVertex(s, tmp1, tmp2);
            t.transform(tmp2);
            mn.setc(i, tmp2.get(i) - margin);
        }
        return;
    }
    }
}

}  // namespace orc
