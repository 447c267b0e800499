// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// manifold.h — restatement of np/PersistentManifold.java, np/ManifoldPoint.java and
// disp/ManifoldResult.java (disp/ = collision/dispatch/).
#pragma once
#include "jmath.h"

namespace orc {

struct ManifoldPoint {  // np/ManifoldPoint.java:36-62
    V3 localPointA, localPointB, positionWorldOnB, positionWorldOnA, normalWorldOnB;
    float distance1 = 0, combinedFriction = 0, combinedRestitution = 0;
    int partId0 = 0, partId1 = 0, index0 = 0, index1 = 0;
    float appliedImpulse = 0;
    bool lateralFrictionInitialized = false;
    float appliedImpulseLateral1 = 0, appliedImpulseLateral2 = 0;
    int lifeTime = 0;
    // not in the reference: which slot of the manifold at the START of this step the point continues
    // (-1 = created this step); lets a host shim keep solver warm-start state attached.
    int srcSlot = -1;

    void init(const V3& pointA, const V3& pointB, const V3& normal, float distance) {  // :76-89
        localPointA.set(pointA);
        localPointB.set(pointB);
        normalWorldOnB.set(normal);
        distance1 = distance;
        combinedFriction = 0; combinedRestitution = 0;
        appliedImpulse = 0; lateralFrictionInitialized = false;
        appliedImpulseLateral1 = 0; appliedImpulseLateral2 = 0;
        lifeTime = 0;
        srcSlot = -1;
    }
};

struct PersistentManifold {
    static constexpr int MANIFOLD_CACHE_SIZE = 4;  // np/PersistentManifold.java:52
    ManifoldPoint pointCache[4];
    int body0 = -1, body1 = -1;  // body indices
    int cachedPoints = 0;
    float breakingThreshold = 0.02f;  // BulletGlobals.java:63

    // lm/VectorUtil.java:60-90 closestAxis4 / maxAxis4
    static int closestAxis4(float x, float y, float z, float w) {
        x = jabsf(x); y = jabsf(y); z = jabsf(z); w = jabsf(w);
        int maxIndex = -1;
        float maxVal = -1e30f;
        if (x > maxVal) { maxIndex = 0; maxVal = x; }
        if (y > maxVal) { maxIndex = 1; maxVal = y; }
        if (z > maxVal) { maxIndex = 2; maxVal = z; }
        if (w > maxVal) { maxIndex = 3; maxVal = w; }
        return maxIndex;
    }

    int sortCachedPoints(const ManifoldPoint& pt) const {  // :83-156
        int maxPenetrationIndex = -1;
        float maxPenetration = pt.distance1;
        for (int i = 0; i < 4; i++) {
            if (pointCache[i].distance1 < maxPenetration) {
                maxPenetrationIndex = i;
                maxPenetration = pointCache[i].distance1;
            }
        }
        float res0 = 0, res1 = 0, res2 = 0, res3 = 0;
        if (maxPenetrationIndex != 0) {
            V3 a0 = pt.localPointA; a0.sub(pointCache[1].localPointA);
            V3 b0 = pointCache[3].localPointA; b0.sub(pointCache[2].localPointA);
            V3 cross; cross.set(a0).crs(b0);
            res0 = cross.len2();
        }
        if (maxPenetrationIndex != 1) {
            V3 a1 = pt.localPointA; a1.sub(pointCache[0].localPointA);
            V3 b1 = pointCache[3].localPointA; b1.sub(pointCache[2].localPointA);
            V3 cross; cross.set(a1).crs(b1);
            res1 = cross.len2();
        }
        if (maxPenetrationIndex != 2) {
            V3 a2 = pt.localPointA; a2.sub(pointCache[0].localPointA);
            V3 b2 = pointCache[3].localPointA; b2.sub(pointCache[1].localPointA);
            V3 cross; cross.set(a2).crs(b2);
            res2 = cross.len2();
        }
        if (maxPenetrationIndex != 3) {
            V3 a3 = pt.localPointA; a3.sub(pointCache[0].localPointA);
            V3 b3 = pointCache[2].localPointA; b3.sub(pointCache[1].localPointA);
            V3 cross; cross.set(a3).crs(b3);
            res3 = cross.len2();
        }
        return closestAxis4(res0, res1, res2, res3);
    }

    int getCacheEntry(const ManifoldPoint& newPoint) const {  // :214-233
        float shortestDist = breakingThreshold * breakingThreshold;
        int nearestPoint = -1;
        for (int i = 0; i < cachedPoints; i++) {
            V3 diffA; diffA.set(pointCache[i].localPointA).sub(newPoint.localPointA);
            float d = diffA.dot(diffA);
            if (d < shortestDist) { shortestDist = d; nearestPoint = i; }
        }
        return nearestPoint;
    }
    int addManifoldPoint(const ManifoldPoint& newPoint) {  // :235-257
        int insertIndex = cachedPoints;
        if (insertIndex == MANIFOLD_CACHE_SIZE) {
            insertIndex = sortCachedPoints(newPoint);
        } else {
            cachedPoints++;
        }
        pointCache[insertIndex] = newPoint;
        return insertIndex;
    }
    void removeContactPoint(int index) {  // :259-278
        int last = cachedPoints - 1;
        if (index != last) {
            pointCache[index] = pointCache[last];
            pointCache[last].appliedImpulse = 0;
            pointCache[last].lateralFrictionInitialized = false;
            pointCache[last].appliedImpulseLateral1 = 0;
            pointCache[last].appliedImpulseLateral2 = 0;
            pointCache[last].lifeTime = 0;
            pointCache[last].srcSlot = -1;
        }
        cachedPoints--;
    }
    void replaceContactPoint(const ManifoldPoint& newPoint, int insertIndex) {  // :280-305
        int lifeTime = pointCache[insertIndex].lifeTime;
        float ai = pointCache[insertIndex].appliedImpulse;
        float l1 = pointCache[insertIndex].appliedImpulseLateral1;
        float l2 = pointCache[insertIndex].appliedImpulseLateral2;
        int src = pointCache[insertIndex].srcSlot;
        pointCache[insertIndex] = newPoint;
        pointCache[insertIndex].appliedImpulse = ai;
        pointCache[insertIndex].appliedImpulseLateral1 = l1;
        pointCache[insertIndex].appliedImpulseLateral2 = l2;
        pointCache[insertIndex].lifeTime = lifeTime;
        pointCache[insertIndex].srcSlot = src;
    }
    bool validContactDistance(const ManifoldPoint& pt) const { return pt.distance1 <= breakingThreshold; }  // :307-309

    void refreshContactPoints(const Xf& trA, const Xf& trB) {  // :312-372
        V3 tmp;
        for (int i = cachedPoints - 1; i >= 0; i--) {
            ManifoldPoint& mp = pointCache[i];
            mp.positionWorldOnA.set(mp.localPointA);
            trA.transform(mp.positionWorldOnA);
            mp.positionWorldOnB.set(mp.localPointB);
            trB.transform(mp.positionWorldOnB);
            tmp.set(mp.positionWorldOnA);
            tmp.sub(mp.positionWorldOnB);
            mp.distance1 = tmp.dot(mp.normalWorldOnB);
            mp.lifeTime++;
        }
        float distance2d;
        V3 projectedDifference, projectedPoint;
        for (int i = cachedPoints - 1; i >= 0; i--) {
            ManifoldPoint& mp = pointCache[i];
            if (!validContactDistance(mp)) {
                removeContactPoint(i);
            } else {
                tmp.set(mp.normalWorldOnB).scl(mp.distance1);
                projectedPoint.set(mp.positionWorldOnA).sub(tmp);
                projectedDifference.set(mp.positionWorldOnB).sub(projectedPoint);
                distance2d = projectedDifference.dot(projectedDifference);
                if (distance2d > breakingThreshold * breakingThreshold) removeContactPoint(i);
            }
        }
    }
};

// disp/ManifoldResult.java
struct ManifoldResult {
    PersistentManifold* manifoldPtr = nullptr;
    Xf rootTransA, rootTransB;
    int body0 = -1, body1 = -1;
    float friction0 = 0.5f, friction1 = 0.5f, restitution0 = 0, restitution1 = 0;
    int partId0 = 0, partId1 = 0, index0 = 0, index1 = 0;
    int addedContacts = 0;  // accepted addContactPoint calls (metric: contacts/s)

    void addContactPoint(const V3& normalOnBInWorld, const V3& pointInWorld, float depth) {  // :92-157
        if (depth > manifoldPtr->breakingThreshold) return;
        addedContacts++;
        bool isSwapped = manifoldPtr->body0 != body0;
        V3 pointA; pointA.set(normalOnBInWorld).scl(depth).add(pointInWorld);
        V3 localA, localB;
        if (isSwapped) {
            rootTransB.invXform(pointA, localA);
            rootTransA.invXform(pointInWorld, localB);
        } else {
            rootTransA.invXform(pointA, localA);
            rootTransB.invXform(pointInWorld, localB);
        }
        ManifoldPoint newPt;
        newPt.init(localA, localB, normalOnBInWorld, depth);
        newPt.positionWorldOnA.set(pointA);
        newPt.positionWorldOnB.set(pointInWorld);
        int insertIndex = manifoldPtr->getCacheEntry(newPt);
        // :160-175
        float friction = friction0 * friction1;
        const float MAX_FRICTION = 10.0f;
        if (friction < -MAX_FRICTION) friction = -MAX_FRICTION;
        if (friction > MAX_FRICTION) friction = MAX_FRICTION;
        newPt.combinedFriction = friction;
        newPt.combinedRestitution = restitution0 * restitution1;
        newPt.partId0 = partId0; newPt.partId1 = partId1;
        newPt.index0 = index0; newPt.index1 = index1;
        if (insertIndex >= 0) manifoldPtr->replaceContactPoint(newPt, insertIndex);
        else manifoldPtr->addManifoldPoint(newPt);
    }
    void refreshContactPoints() {  // :177-191
        if (manifoldPtr->cachedPoints == 0) return;
        bool isSwapped = manifoldPtr->body0 != body0;
        if (isSwapped) manifoldPtr->refreshContactPoints(rootTransB, rootTransA);
        else manifoldPtr->refreshContactPoints(rootTransA, rootTransB);
    }
};

}  // namespace orc
