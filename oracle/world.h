// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// world.h — restatement of the per-step collision path
//   disp/CollisionWorld.java:123-151 performDiscreteCollisionDetection
//     :231-245 updateAabbs / :195-229 updateSingleAabb
//     bp/DbvtBroadphase.java:89-228 (as the per-proxy effective-AABB state machine)
//     bp/SimpleBroadphase.java:81-110 ("tight" mode)
//     disp/CollisionDispatcher.java:198-257 + disp/DefaultNearCallback.java:39-66
//     disp/DefaultCollisionConfiguration.java:149-213 (algorithm table)
//     disp/SphereSphereCollisionAlgorithm.java:73-134, disp/ConvexPlaneCollisionAlgorithm.java:75-136,
//     disp/ConvexConvexAlgorithm.java:90-139, disp/ConvexConcaveCollisionAlgorithm.java:65-93,
//     disp/ConvexTriangleCallback.java:83-172
#pragma once
#include <algorithm>
#include <cstdint>
#include <map>
#include <memory>
#include <set>
#include <vector>
#include "bvh.h"
#include "dbvt_literal.h"
#include "sap_literal.h"
#include "gjk.h"
#include "convexcast.h"
#include "raycast.h"
#include "jmath.h"
#include "manifold.h"
#include "shapes.h"

namespace orc {

enum BroadphaseMode { BP_TIGHT = 0, BP_DBVT = 1, BP_DBVT_LITERAL = 2,
                      BP_SAP16 = 3, BP_SAP32 = 4,                     // AxisSweep3 / AxisSweep3_32 as a stateless predicate on quantised AABBs
                      BP_SAP16_LITERAL = 5, BP_SAP32_LITERAL = 6 };  // 2: the reference's tree, literally (dbvt_literal.h)

struct Body {
    int shape = -1;
    Xf xf;
    int16_t group = 1, mask = -1;
    int uid = 0;
    bool isStatic = false;
    bool active = true;
    bool alive = true;
    float friction = 0.5f, restitution = 0.0f;  // disp/CollisionObject.java:95,71
    // broadphase proxy state
    V3 effMin, effMax;    // DbvtProxy.aabb (bp/DbvtProxy.java:35) / SimpleBroadphaseProxy min,max
    V3 leafMin, leafMax;  // proxy.leaf.volume
    bool inFixed = false; // stage == STAGECOUNT
    int lastSetStep = -1; // step index of the last setAabb (createProxy counts as one)
    bool aabbOverflow = false;
    LProxy* lit = nullptr; // BP_DBVT_LITERAL proxy
    int world = 0;        // batched independent worlds: each is its own CollisionWorld in the reference
    unsigned qmin[3] = {0, 0, 0}, qmax[3] = {1, 1, 1};  // SAP modes: quantised bounds (bp/AxisSweep3Internal.java:201-216)
    int sapHandle = 0;    // literal SAP handle (== uid while no handle is ever reused)
};

struct RawContact {  // one detector result, before ManifoldResult
    int uid0, uid1;  // the broadphase pair
    int tri;         // triangle index for mesh pairs, else -1
    int hasContact;
    float normal[3], point[3], depth;
    int method;      // GjkPairDetector.lastUsedMethod, or 10 sphere-sphere, 11 convex-plane
    int iters;
};

struct PairState {
    PersistentManifold manifold;
    bool hasManifold = false;
    // CompoundCollisionAlgorithm.childCollisionAlgorithms (disp/CompoundCollisionAlgorithm.java:46): one child algorithm, and
    // so one manifold, per child (per child x child when both shapes are compounds), in the order processCollision visits them
    std::vector<PairState> kids;
    std::vector<std::pair<int, int>> kidChild;  // (child index in manifold.body0's shape, in body1's shape), -1 = not a compound
    bool isCompound = false;
};

struct MeshShapeData {
    MeshData mesh;
    Bvh bvh;
};

struct World {
    int mode = BP_TIGHT;
    float breakingThreshold = 0.02f;  // BulletGlobals.java:63
    float dbvtMargin = 0.05f;         // bp/DbvtBroadphase.java:35
    float predictedFrames = 2.0f;     // bp/DbvtBroadphase.java:73
    bool bruteForcePairs = false;     // O(N^2) pair enumeration instead of sort-and-sweep
    std::vector<Shape> shapes;
    std::vector<std::unique_ptr<MeshShapeData>> meshes;  // indexed by shape id (null for non-mesh)
    std::vector<Body> bodies;  // index = uid-1
    int step = 0;
    std::vector<std::pair<int, int>> pairs;  // sorted (uid0<uid1)
    std::vector<std::pair<int, int>> prevPairs;  // the list before the last calculateOverlappingPairs (pair add/remove deltas)
    std::map<std::pair<int, int>, PairState> pairState;
    std::vector<RawContact> raw;
    std::set<std::pair<int, int>> noCollide;  // (uid0 < uid1) pairs whose bodies are linked by a collision-disabling constraint
    LDbvtBroadphase literal;
    // AxisSweep3 modes
    V3 worldAabbMin = V3(-1000.f, -1000.f, -1000.f), worldAabbMax = V3(1000.f, 1000.f, 1000.f);
    SapQuantizer sapQ;
    LSap sapLit;
    bool sapReady = false;
    std::vector<int> sapHandleUid;
    bool isSap() const { return mode >= BP_SAP16 && mode <= BP_SAP32_LITERAL; }
    bool isSapLiteral() const { return mode == BP_SAP16_LITERAL || mode == BP_SAP32_LITERAL; }
    void sapInit() {
        if (sapReady) return;
        const bool wide = (mode == BP_SAP32 || mode == BP_SAP32_LITERAL);
        sapQ.init(worldAabbMin, worldAabbMax, wide);
        if (isSapLiteral()) sapLit.init(worldAabbMin, worldAabbMax, wide, 1 << 16);
        sapReady = true;
    }
    // quantise, and keep the monotone float image of the quantised box as the "effective" AABB
    void sapSet(Body& b, const V3& mn, const V3& mx) {
        sapQ.quantize(b.qmin, mn, 0);
        sapQ.quantize(b.qmax, mx, 1);
        b.effMin.set(sapQ.dequant(b.qmin[0], 0), sapQ.dequant(b.qmin[1], 1), sapQ.dequant(b.qmin[2], 2));
        b.effMax.set(sapQ.dequant(b.qmax[0], 0), sapQ.dequant(b.qmax[1], 1), sapQ.dequant(b.qmax[2], 2));
    }
    static bool sapOverlap(const Body& a, const Body& b) {
        for (int k = 0; k < 3; k++)
            if (a.qmax[k] < b.qmin[k] || b.qmax[k] < a.qmin[k]) return false;
        return true;
    }
    // counters
    long gjkChecks = 0, deepPenetrationChecks = 0, addedContacts = 0, bvhNodesVisited = 0, trianglesTested = 0;

    int addShape(const Shape& s) {
        shapes.push_back(s);
        meshes.emplace_back(nullptr);
        return (int)shapes.size() - 1;
    }
    // sh/CompoundShape.java:50-82 addChildShape for every child: the local AABB is the running Math.min / Math.max of the
    // children's AABBs under their local transforms.  Children are convex (box, sphere, hull) or compounds themselves (their
    // box is then CompoundShape.getAabb of the nested shape under the child transform).
    int addCompound(int n, const int* childShapes, const float* childXf12) {
        Shape s;
        s.type = SH_COMPOUND;
        s.collisionMargin = 0.0f;  // sh/CompoundShape.java:49
        s.localAabbMin.set(1e30f, 1e30f, 1e30f);     // :44
        s.localAabbMax.set(-1e30f, -1e30f, -1e30f);  // :45
        for (int i = 0; i < n; i++) {
            CompoundChild c;
            const float* t = childXf12 + 12 * i;
            for (int r = 0; r < 3; r++)
                for (int k = 0; k < 3; k++) c.transform.basis.m[r][k] = t[r * 3 + k];
            c.transform.origin.set(t[9], t[10], t[11]);
            c.shape = childShapes[i];
            V3 mn, mx;
            shapeGetAabb(shapes[c.shape], c.transform, mn, mx);
            s.localAabbMin.set(jminf(s.localAabbMin.x, mn.x), jminf(s.localAabbMin.y, mn.y), jminf(s.localAabbMin.z, mn.z));  // lm/VectorUtil.java:176-180
            s.localAabbMax.set(jmaxf(s.localAabbMax.x, mx.x), jmaxf(s.localAabbMax.y, mx.y), jmaxf(s.localAabbMax.z, mx.z));  // :182-186
            s.children.push_back(c);
        }
        return addShape(s);
    }
    int addMesh(const float* verts, int nv, const int* idx, int ntri) {
        return addMeshParts(1, verts, &nv, idx, &ntri);
    }
    // sh/TriangleIndexVertexArray.java:72-100: nparts IndexedMesh entries; verts / idx are the parts' arrays back to back
    // (indices local to their part)
    int addMeshParts(int nparts, const float* verts, const int* nv, const int* idx, const int* ntri) {
        std::unique_ptr<MeshShapeData> md(new MeshShapeData());
        md->mesh.parts.resize(nparts);
        for (int p = 0; p < nparts; p++) {
            md->mesh.parts[p].verts.assign(verts, verts + 3 * nv[p]);
            md->mesh.parts[p].idx.assign(idx, idx + 3 * ntri[p]);
            verts += 3 * nv[p];
            idx += 3 * ntri[p];
        }
        // sh/StridingMeshInterface.java calculateAabbBruteForce
        V3 mn(1e30f, 1e30f, 1e30f), mx(-1e30f, -1e30f, -1e30f);
        for (int p = 0; p < nparts; p++)
        for (int t = 0; t < ntri[p]; t++) {
            V3 tri[3];
            md->mesh.getTriangle(p, t, tri);
            for (int k = 0; k < 3; k++) {
                mn.x = jminf(mn.x, tri[k].x); mn.y = jminf(mn.y, tri[k].y); mn.z = jminf(mn.z, tri[k].z);
                mx.x = jmaxf(mx.x, tri[k].x); mx.y = jmaxf(mx.y, tri[k].y); mx.z = jmaxf(mx.z, tri[k].z);
            }
        }
        md->bvh.build(md->mesh, mn, mx);
        Shape s;
        s.type = SH_MESH;
        s.collisionMargin = 0.0f;  // sh/ConcaveShape.java:35
        // sh/TriangleMeshShape.java:79-93 recalcLocalAabb: extreme vertex coordinate per axis +- margin
        for (int i = 0; i < 3; i++) {
            s.localAabbMax.setc(i, mx.get(i) + s.collisionMargin);
            s.localAabbMin.setc(i, mn.get(i) - s.collisionMargin);
        }
        s.bvh = &md->bvh;
        shapes.push_back(s);
        meshes.push_back(std::move(md));
        return (int)shapes.size() - 1;
    }

    // disp/CollisionWorld.java:102-121 addCollisionObject + bp/DbvtBroadphase.java:173-182 createProxy
    int addBody(int shape, const Xf& xf, int group, int mask, bool isStatic, int world = 0) {
        Body b;
        b.shape = shape;
        b.xf.set(xf);
        b.group = (int16_t)group;
        b.mask = (int16_t)mask;
        b.isStatic = isStatic;
        b.world = world;
        b.uid = (int)bodies.size() + 1;  // ++gid
        shapeGetAabb(shapes[shape], b.xf, b.effMin, b.effMax);  // no +-threshold at creation
        b.leafMin = b.effMin; b.leafMax = b.effMax;
        b.inFixed = false;
        b.lastSetStep = step;
        if (mode == BP_DBVT_LITERAL) {
            literal.margin = dbvtMargin;
            literal.predictedframes = predictedFrames;
            b.lit = literal.createProxy(b.effMin, b.effMax, group, mask, world);
        }
        if (isSap()) {  // bp/AxisSweep3Internal.java:576-584 createProxy -> addHandle with the shape's AABB
            sapInit();
            V3 mn = b.effMin, mx = b.effMax;
            if (isSapLiteral()) {
                b.sapHandle = sapLit.addHandle(mn, mx, group, mask, world);
                if ((int)sapHandleUid.size() <= b.sapHandle) sapHandleUid.resize(b.sapHandle + 1, 0);
                sapHandleUid[b.sapHandle] = b.uid;
            }
            sapSet(b, mn, mx);
        }
        bodies.push_back(b);
        return b.uid;
    }
    void removeBody(int uid) {
        Body& b = bodies[uid - 1];
        if (b.alive && mode == BP_DBVT_LITERAL) literal.destroyProxy(b.lit);
        if (b.alive && isSapLiteral()) sapLit.removeHandle(b.sapHandle);
        b.alive = false;
    }

    static bool intersect(const V3& amin, const V3& amax, const V3& bmin, const V3& bmax) {  // bp/DbvtAabbMm.java:209-212
        return (amin.x <= bmax.x) && (amax.x >= bmin.x) && (amin.y <= bmax.y) && (amax.y >= bmin.y) &&
               (amin.z <= bmax.z) && (amax.z >= bmin.z);
    }

    // BroadphaseInterface.setAabb
    void setAabb(Body& b, const V3& mn, const V3& mx) {
        if (mode == BP_TIGHT) {  // bp/SimpleBroadphase.java:112-116
            b.effMin = mn; b.effMax = mx;
            b.lastSetStep = step;
            return;
        }
        if (mode == BP_DBVT_LITERAL) {
            literal.setAabb(b.lit, mn, mx);
            b.effMin = b.lit->aabb.mi; b.effMax = b.lit->aabb.mx;
            b.lastSetStep = step;
            return;
        }
        if (isSap()) {  // bp/AxisSweep3Internal.java:591-594 setAabb -> updateHandle
            if (isSapLiteral()) sapLit.updateHandle(b.sapHandle, mn, mx);
            sapSet(b, mn, mx);
            b.lastSetStep = step;
            return;
        }
        // bp/DbvtBroadphase.java:196-228
        V3 amin = mn, amax = mx;
        if (b.inFixed) {
            b.leafMin = amin; b.leafMax = amax;
            b.inFixed = false;
        } else {
            if (intersect(b.leafMin, b.leafMax, amin, amax)) {
                V3 delta; delta.set(mn).add(mx);
                delta.scl(0.5f);
                V3 center; center.set(b.effMin).add(b.effMax); center.scl(0.5f);  // bp/DbvtAabbMm.java:66-70
                delta.sub(center);
                delta.scl(predictedFrames);
                // bp/Dbvt.java:157-169 update(leaf, volume, velocity, margin): volume expanded IN PLACE
                bool contain = (b.leafMin.x <= amin.x) && (b.leafMin.y <= amin.y) && (b.leafMin.z <= amin.z) &&
                               (b.leafMax.x >= amax.x) && (b.leafMax.y >= amax.y) && (b.leafMax.z >= amax.z);
                if (!contain) {
                    V3 e(dbvtMargin, dbvtMargin, dbvtMargin);
                    amin.sub(e); amax.add(e);  // Expand
                    if (delta.x > 0) amax.x += delta.x; else amin.x += delta.x;  // SignedExpand
                    if (delta.y > 0) amax.y += delta.y; else amin.y += delta.y;
                    if (delta.z > 0) amax.z += delta.z; else amin.z += delta.z;
                    b.leafMin = amin; b.leafMax = amax;
                }
            } else {
                b.leafMin = amin; b.leafMax = amax;  // teleporting
            }
        }
        b.effMin = amin; b.effMax = amax;  // proxy.aabb.set(aabb) — aliasing quirk
        b.lastSetStep = step;
    }

    // disp/CollisionWorld.java:231-245 / :195-229
    void updateAabbs() {
        for (Body& b : bodies) {
            if (!b.alive || !b.active) continue;
            V3 mn, mx;
            shapeGetAabb(shapes[b.shape], b.xf, mn, mx);
            V3 ct(breakingThreshold, breakingThreshold, breakingThreshold);
            mn.sub(ct);
            mx.add(ct);
            V3 tmp; tmp.set(mx).sub(mn);
            if (b.isStatic || (tmp.len2() < 1e12f)) {
                setAabb(b, mn, mx);
            } else {
                b.aabbOverflow = true;  // reference sets DISABLE_SIMULATION
                b.active = false;
            }
        }
    }

    bool filter(const Body& a, const Body& b) const {  // bp/HashedOverlappingPairCache.java:179-188
        if (a.world != b.world) return false;  // separate worlds never share a pair cache
        bool collides = (a.group & b.mask) != 0;
        collides = collides && (b.group & a.mask) != 0;
        return collides;
    }

    // BroadphaseInterface.calculateOverlappingPairs: resulting pair set
    int calculateOverlappingPairs() {
        prevPairs = pairs;
        if (mode == BP_DBVT_LITERAL) {
            literal.collide();
            pairs.clear();
            for (auto& pr : literal.pairArray) pairs.push_back(std::make_pair(pr.first->uid, pr.second->uid));
            std::sort(pairs.begin(), pairs.end());
            step++;
            return (int)pairs.size();
        }
        if (isSapLiteral()) {
            // the reference's proxy uid IS the handle index (bp/AxisSweep3Internal.java:461), and freed handles are reused;
            // this oracle numbers bodies by creation, so handles are mapped back to its uids
            pairs.clear();
            for (auto& pr : sapLit.pairs) {
                int a = sapHandleUid[pr.first], c = sapHandleUid[pr.second];
                pairs.push_back(std::make_pair(std::min(a, c), std::max(a, c)));
            }
            std::sort(pairs.begin(), pairs.end());
            step++;
            return (int)pairs.size();
        }
        if (mode == BP_DBVT) {
            // bp/DbvtBroadphase.java:96-111: proxies not updated during this step (but updated the step
            // before) move to the fixed set; eff is kept, leaf volume becomes eff.
            for (Body& b : bodies) {
                if (!b.alive) continue;
                if (!b.inFixed && b.lastSetStep < step) {
                    b.inFixed = true;
                    b.leafMin = b.effMin; b.leafMax = b.effMax;
                }
            }
        }
        pairs.clear();
        std::vector<int> order;
        for (size_t i = 0; i < bodies.size(); i++)
            if (bodies[i].alive) order.push_back((int)i);
        if (bruteForcePairs) {
            for (size_t a = 0; a < order.size(); a++)
                for (size_t c = a + 1; c < order.size(); c++) {
                    const Body& A = bodies[order[a]];
                    const Body& B = bodies[order[c]];
                    if (filter(A, B) && (isSap() ? sapOverlap(A, B) : intersect(A.effMin, A.effMax, B.effMin, B.effMax)))
                        pairs.push_back(std::make_pair(A.uid, B.uid));
                }
        } else {
            // world-major order: separate worlds never share a pair cache (filter() rejects such pairs anyway), so each
            // world is swept on its own — 4096 batched worlds occupy the same coordinates and would make one sweep quadratic
            std::sort(order.begin(), order.end(), [&](int a, int c) {
                if (bodies[a].world != bodies[c].world) return bodies[a].world < bodies[c].world;
                if (bodies[a].effMin.x != bodies[c].effMin.x) return bodies[a].effMin.x < bodies[c].effMin.x;
                return a < c;
            });
            for (size_t a = 0; a < order.size(); a++) {
                const Body& A = bodies[order[a]];
                for (size_t c = a + 1; c < order.size(); c++) {
                    const Body& B = bodies[order[c]];
                    if (B.world != A.world || B.effMin.x > A.effMax.x) break;
                    // SAP modes: eff is a monotone image of the quantised box, so the float window is conservative
                    if (filter(A, B) && (isSap() ? sapOverlap(A, B) : intersect(A.effMin, A.effMax, B.effMin, B.effMax)))
                        pairs.push_back(std::make_pair(std::min(A.uid, B.uid), std::max(A.uid, B.uid)));
                }
            }
        }
        std::sort(pairs.begin(), pairs.end());
        step++;
        return (int)pairs.size();
    }

    // disp/CollisionWorld.java:260-356 rayTestSingle for one (object, shape, transform) against the running callback state
    void rayTestSingle(const V3& from, const V3& to, const Body& b, int shapeIndex, const Xf& xf, float& closest, RayHit& hit) const {
        const Shape& s = shapes[shapeIndex];
        if (s.isConvex()) {
            CastResult cr;
            cr.fraction = closest;
            if (rayConvexCast(from, to, s, xf, cr)) {
                if (cr.normal.len2() > 0.0001f) {
                    if (cr.fraction < closest) {
                        // castResult.normal.mul(rayFromTrans.basis) with the identity basis, then nor()
                        V3 nn(cr.normal.x * 1.f + cr.normal.y * 0.f + cr.normal.z * 0.f, cr.normal.x * 0.f + cr.normal.y * 1.f + cr.normal.z * 0.f,
                              cr.normal.x * 0.f + cr.normal.y * 0.f + cr.normal.z * 1.f);
                        nn.nor();
                        closest = cr.fraction;
                        hit.uid = b.uid;
                        hit.fraction = cr.fraction;
                        hit.normal.set(nn);
                        float sgl = 1.f - cr.fraction;
                        hit.point.set(sgl * from.x + cr.fraction * to.x, sgl * from.y + cr.fraction * to.y, sgl * from.z + cr.fraction * to.z);
                    }
                }
            }
        } else if (s.isConcave()) {
            Xf worldToObj; worldToObj.set(xf);
            worldToObj.inverse();
            TriangleRaycast rcb;
            rcb.from = from; worldToObj.transform(rcb.from);
            rcb.to = to; worldToObj.transform(rcb.to);
            rcb.hitFraction = closest;
            if (s.type == SH_MESH) {
                const MeshShapeData* md = meshes[shapeIndex].get();
                bvhReportRayOverlappingNodex(md->bvh, rcb.from, rcb.to, [&](int part, int tri) {
                    V3 t[3];
                    md->mesh.getTriangle(part, tri, t);
                    rcb.processTriangle(t, part, tri);
                });
            } else {
                V3 mn(jminf(rcb.from.x, rcb.to.x), jminf(rcb.from.y, rcb.to.y), jminf(rcb.from.z, rcb.to.z));
                V3 mx(jmaxf(rcb.from.x, rcb.to.x), jmaxf(rcb.from.y, rcb.to.y), jmaxf(rcb.from.z, rcb.to.z));
                planeProcessAllTriangles(s.planeNormal, s.planeConstant, mn, mx, [&](const V3* t, int part, int tri) { rcb.processTriangle(t, part, tri); });
            }
            if (rcb.hit) {  // ClosestRayResultCallback.addSingleResult of the last reported triangle (:708-729)
                closest = rcb.hitFraction;
                hit.uid = b.uid;
                hit.fraction = rcb.hitFraction;
                hit.normal.set(rcb.hitNormalLocal);
                v3mul(hit.normal, b.xf.basis);  // collisionObject.getWorldTransform().basis (the OBJECT's, also for compound children)
                float sgl = 1.f - rcb.hitFraction;
                hit.point.set(sgl * from.x + rcb.hitFraction * to.x, sgl * from.y + rcb.hitFraction * to.y, sgl * from.z + rcb.hitFraction * to.z);
            }
        } else if (s.isCompound()) {
            for (size_t i = 0; i < s.children.size(); i++) {
                Xf childWorld; childWorld.set(xf);
                childWorld.mul(s.children[i].transform);
                rayTestSingle(from, to, b, s.children[i].shape, childWorld, closest, hit);
            }
        }
    }

    // disp/CollisionWorld.java:553-590 rayTest with a ClosestRayResultCallback(group, mask)
    RayHit rayTestClosest(const V3& from, const V3& to, int group, int mask) const {
        RayHit hit;
        float closest = 1.f;  // RayResultCallback.closestHitFraction
        for (const Body& b : bodies) {
            if (!b.alive) continue;
            if (closest == 0.f) break;
            // RayResultCallback.needsCollision (disp/CollisionWorld.java:664-670)
            bool collides = ((int)b.group & mask) != 0;
            collides = collides && (group & (int)b.mask) != 0;
            if (!collides) continue;
            const Shape& s = shapes[b.shape];
            V3 mn, mx;
            shapeGetAabb(s, b.xf, mn, mx);
            float hitLambda = closest;
            V3 hitNormal;
            if (!rayAabb(from, to, mn, mx, hitLambda, hitNormal)) continue;
            rayTestSingle(from, to, b, b.shape, b.xf, closest, hit);
        }
        return hit;
    }

    // disp/CollisionWorld.java:392-551 objectQuerySingle against the running ClosestConvexResultCallback state
    // notMeVel != null: the callback is DiscreteDynamicsWorld's ClosestNotMeConvexResultCallback (dyn/DiscreteDynamicsWorld.java
    // :1143-1164), whose addSingleResult drops a result whose normal does not oppose the motion (allowedPenetration 0)
    void objectQuerySingle(const Shape& cast, const Xf& fromT, const Xf& toT, const Body& b, int shapeIndex, const Xf& xf,
                           float allowedPenetration, ConvexSweepHit& hit, const V3* notMeVel = nullptr) const {
        const Shape& s = shapes[shapeIndex];
        if (s.isConvex()) {
            ConvexCastResult cr;
            cr.allowedPenetration = allowedPenetration;
            cr.fraction = 1.f;
            if (gjkConvexCast(cast, fromT, toT, s, xf, cr)) {
                if (cr.normal.len2() > 0.0001f) {
                    if (cr.fraction < hit.fraction) {
                        cr.normal.nor();
                        if (notMeVel && cr.normal.dot(*notMeVel) >= -0.f) return;
                        hit.fraction = cr.fraction;           // ClosestConvexResultCallback.addSingleResult, normalInWorldSpace
                        hit.uid = b.uid;
                        hit.normal.set(cr.normal);
                        hit.point.set(cr.hitPoint);
                    }
                }
            }
        } else if (s.type == SH_MESH) {
            Xf worldToObj; worldToObj.set(xf);
            worldToObj.inverse();
            V3 fromLocal = fromT.origin; worldToObj.transform(fromLocal);
            V3 toLocal = toT.origin; worldToObj.transform(toLocal);
            Xf rotationXform;  // rotation of the cast box in mesh space = MeshRotation^-1 * ConvexToRotation, origin 0
            rotationXform.basis.set(worldToObj.basis);
            rotationXform.basis.mul(toT.basis);
            rotationXform.origin.set(0, 0, 0);
            V3 boxMin, boxMax;
            shapeGetAabb(cast, rotationXform, boxMin, boxMax);
            const MeshShapeData* md = meshes[shapeIndex].get();
            const float entry = hit.fraction;  // tccb.hitFraction = resultCallback.closestHitFraction, never updated afterwards
            bvhReportBoxCastOverlappingNodex(md->bvh, fromLocal, toLocal, boxMin, boxMax, [&](int part, int tri) {
                Shape tm;
                tm.type = SH_TRIANGLE;
                md->mesh.getTriangle(part, tri, tm.tri);
                tm.collisionMargin = s.getMargin();  // triangleMesh.getMargin()
                ConvexCastResult cr;
                cr.fraction = 1.f;
                if (subsimplexConvexCast(cast, fromT, toT, tm, xf, xf, cr)) {
                    if (cr.normal.len2() > 0.0001f) {
                        if (cr.fraction < entry) {
                            cr.normal.nor();
                            if (cr.fraction <= hit.fraction) {   // reportHit: hitFraction <= closestHitFraction
                                if (notMeVel && cr.normal.dot(*notMeVel) >= -0.f) return;
                                hit.fraction = cr.fraction;
                                hit.uid = b.uid;
                                hit.normal.set(cr.normal);       // normalInWorldSpace = true
                                hit.point.set(cr.hitPoint);
                            }
                        }
                    }
                }
            });
        } else if (s.type == SH_PLANE) {
            hit.unsupported = true;  // the reference calls calcTimeOfImpact on a null caster here (:470-477)
        } else if (s.isCompound()) {
            for (const CompoundChild& c : s.children) {
                Xf childWorld; childWorld.set(xf);
                childWorld.mul(c.transform);
                objectQuerySingle(cast, fromT, toT, b, c.shape, childWorld, allowedPenetration, hit, notMeVel);
            }
        }
    }

    // disp/CollisionWorld.java:596-651 convexSweepTest with a ClosestConvexResultCallback(group, mask), translational sweep
    ConvexSweepHit convexSweepClosest(int castShape, const Xf& fromT, const Xf& toT, int group, int mask, float allowedPenetration) const {
        ConvexSweepHit hit;
        const Shape& cast = shapes[castShape];
        // calculateTemporalAabb(R, linVel, angVel = 0, 1): R = the cast shape's rotation with a zero origin
        Xf R; R.basis.set(fromT.basis); R.origin.set(0, 0, 0);
        V3 castMin, castMax;
        shapeGetAabb(cast, R, castMin, castMax);
        V3 lin; lin.set(toT.origin).sub(fromT.origin);
        lin.scl(1.f / 1.f);
        lin.scl(1.f);
        if (lin.x > 0.f) castMax.x += lin.x; else castMin.x += lin.x;
        if (lin.y > 0.f) castMax.y += lin.y; else castMin.y += lin.y;
        if (lin.z > 0.f) castMax.z += lin.z; else castMin.z += lin.z;
        { V3 am(0, 0, 0); castMin.sub(am); castMax.add(am); }
        for (const Body& b : bodies) {
            if (!b.alive) continue;
            bool collides = ((int)b.group & mask & 0xFFFF) != 0;                     // :752-756
            collides = collides && ((group & (int)b.mask) & 0xFFFF) != 0;
            if (!collides) continue;
            V3 mn, mx;
            shapeGetAabb(shapes[b.shape], b.xf, mn, mx);
            mn.add(castMin);                                                         // AabbUtil2.aabbExpand
            mx.add(castMax);
            float hitLambda = 1.f;
            V3 hitNormal;
            if (!rayAabb(fromT.origin, toT.origin, mn, mx, hitLambda, hitNormal)) continue;
            objectQuerySingle(cast, fromT, toT, b, b.shape, b.xf, allowedPenetration, hit);
        }
        return hit;
    }

    // DiscreteDynamicsWorld.integrateTransforms' "CCD motion clamping" query (dyn/DiscreteDynamicsWorld.java:700-729): a sphere
    // of the body's ccdSweptSphereRadius swept from its world transform to the predicted origin, ClosestNotMeConvexResultCallback
    // (:1129-1199) with the body's own filter group / mask: not the body itself, not an object it already has contact points
    // with (any manifold of their pair's algorithm), not a hit whose normal does not oppose the motion.  The predicted
    // ROTATION only enters the reference through the angular term of the culling box (|w| * r * sqrt(3)), taken as 0 here.
    ConvexSweepHit ccdSweepNotMe(int meUid, float radius, const V3& to, float allowedPenetration) const {
        ConvexSweepHit hit;
        const Body& me = bodies[meUid - 1];
        Shape cast;
        initSphere(cast, radius);
        Xf fromT; fromT.set(me.xf);
        Xf toT; toT.set(me.xf); toT.origin.set(to);
        Xf R; R.basis.set(fromT.basis); R.origin.set(0, 0, 0);
        V3 castMin, castMax;
        shapeGetAabb(cast, R, castMin, castMax);
        V3 lin; lin.set(toT.origin).sub(fromT.origin);
        lin.scl(1.f / 1.f);
        lin.scl(1.f);
        if (lin.x > 0.f) castMax.x += lin.x; else castMin.x += lin.x;
        if (lin.y > 0.f) castMax.y += lin.y; else castMin.y += lin.y;
        if (lin.z > 0.f) castMax.z += lin.z; else castMin.z += lin.z;
        V3 linVelA; linVelA.set(to).sub(fromT.origin);
        V3 linVelB(0, 0, 0);
        V3 rel; rel.set(linVelA).sub(linVelB);
        const int group = me.group, mask = me.mask;
        for (const Body& b : bodies) {
            if (!b.alive) continue;
            if (b.uid == meUid) continue;                                              // :1169
            bool collides = ((int)b.group & mask & 0xFFFF) != 0;
            collides = collides && ((group & (int)b.mask) & 0xFFFF) != 0;
            if (!collides) continue;
            {   // :1181-1196 needsResponse (me is dynamic) -> skip objects with contact points already
                auto it = pairState.find(std::make_pair(std::min(meUid, b.uid), std::max(meUid, b.uid)));
                if (it != pairState.end() && std::binary_search(pairs.begin(), pairs.end(), it->first)) {
                    bool touching = it->second.hasManifold && it->second.manifold.cachedPoints > 0;
                    for (const PairState& k : it->second.kids) touching = touching || (k.hasManifold && k.manifold.cachedPoints > 0);
                    if (touching) continue;
                }
            }
            V3 mn, mx;
            shapeGetAabb(shapes[b.shape], b.xf, mn, mx);
            mn.add(castMin);
            mx.add(castMax);
            float hitLambda = 1.f;
            V3 hitNormal;
            if (!rayAabb(fromT.origin, toT.origin, mn, mx, hitLambda, hitNormal)) continue;
            objectQuerySingle(cast, fromT, toT, b, b.shape, b.xf, allowedPenetration, hit, &rel);
        }
        return hit;
    }

    // ---- narrowphase ------------------------------------------------------------------
    void initResult(ManifoldResult& r, const Body& b0, const Body& b1) {  // disp/ManifoldResult.java:70-75
        r.body0 = b0.uid; r.body1 = b1.uid;
        r.rootTransA.set(b0.xf); r.rootTransB.set(b1.xf);
        r.friction0 = b0.friction; r.friction1 = b1.friction;
        r.restitution0 = b0.restitution; r.restitution1 = b1.restitution;
        r.partId0 = r.partId1 = r.index0 = r.index1 = 0;
    }

    void convexConvex(const Shape* s0, const Shape* s1, const Body& b0, const Body& b1, ManifoldResult& res,
                      bool ownManifold, int tri) {  // disp/ConvexConvexAlgorithm.java:90-139
        float maxd = s0->getMargin() + s1->getMargin() + res.manifoldPtr->breakingThreshold;
        maxd *= maxd;
        GjkOut out;
        gjkChecks++;
        gjkGetClosestPoints(s0, s1, b0.xf, b1.xf, maxd, out);
        deepPenetrationChecks += out.deepPenetrationChecks;
        RawContact rc;
        rc.uid0 = res.body0; rc.uid1 = res.body1; rc.tri = tri;
        rc.hasContact = out.hasContact ? 1 : 0;
        rc.normal[0] = out.normalOnBInWorld.x; rc.normal[1] = out.normalOnBInWorld.y; rc.normal[2] = out.normalOnBInWorld.z;
        rc.point[0] = out.pointInWorld.x; rc.point[1] = out.pointInWorld.y; rc.point[2] = out.pointInWorld.z;
        rc.depth = out.depth; rc.method = out.lastUsedMethod; rc.iters = out.curIter;
        raw.push_back(rc);
        if (out.hasContact) res.addContactPoint(out.normalOnBInWorld, out.pointInWorld, out.depth);
        if (ownManifold) res.refreshContactPoints();
    }

    void sphereSphere(const Shape* s0, const Shape* s1, const Body& b0, const Body& b1, ManifoldResult& res) {
        // disp/SphereSphereCollisionAlgorithm.java:73-134
        V3 diff; diff.set(b0.xf.origin).sub(b1.xf.origin);
        float len = diff.len();
        float radius0 = s0->implicitDims.x * s0->localScaling.x;
        float radius1 = s1->implicitDims.x * s1->localScaling.x;
        RawContact rc;
        rc.uid0 = b0.uid; rc.uid1 = b1.uid; rc.tri = -1; rc.method = 10; rc.iters = 0;
        rc.hasContact = 0; rc.depth = 0;
        for (int k = 0; k < 3; k++) rc.normal[k] = rc.point[k] = 0;
        if (len > (radius0 + radius1)) {
            raw.push_back(rc);
            res.refreshContactPoints();
            return;
        }
        float dist = len - (radius0 + radius1);
        V3 normalOnSurfaceB(1, 0, 0);
        if (len > FLT_EPSILON_) normalOnSurfaceB.set(diff).scl(1.0f / len);
        V3 tmp, pos1;
        tmp.set(normalOnSurfaceB).scl(radius1);
        pos1.set(b1.xf.origin).add(tmp);
        rc.hasContact = 1;
        rc.normal[0] = normalOnSurfaceB.x; rc.normal[1] = normalOnSurfaceB.y; rc.normal[2] = normalOnSurfaceB.z;
        rc.point[0] = pos1.x; rc.point[1] = pos1.y; rc.point[2] = pos1.z;
        rc.depth = dist;
        raw.push_back(rc);
        res.addContactPoint(normalOnSurfaceB, pos1, dist);
        res.refreshContactPoints();
    }

    void convexPlane(const Shape* cs, const Shape* ps, const Body& convexObj, const Body& planeObj, ManifoldResult& res) {
        // disp/ConvexPlaneCollisionAlgorithm.java:75-136
        V3 planeNormal = ps->planeNormal;
        float planeConstant = ps->planeConstant;
        Xf planeInConvex; planeInConvex.set(convexObj.xf);
        planeInConvex.inverse();
        planeInConvex.mul(planeObj.xf);
        Xf convexInPlaneTrans; convexInPlaneTrans.set(planeObj.xf);
        convexInPlaneTrans.inverse();
        convexInPlaneTrans.mul(convexObj.xf);
        V3 tmp; tmp.set(planeNormal).scl(-1.0f);
        v3mul(tmp, planeInConvex.basis);
        V3 vtx;
        localGetSupportingVertex(*cs, tmp, vtx);
        V3 vtxInPlane = vtx;
        convexInPlaneTrans.transform(vtxInPlane);
        float distance = (planeNormal.dot(vtxInPlane) - planeConstant);
        V3 vtxInPlaneProjected;
        tmp.set(planeNormal).scl(distance);
        vtxInPlaneProjected.set(vtxInPlane).sub(tmp);
        V3 vtxInPlaneWorld = vtxInPlaneProjected;
        planeObj.xf.transform(vtxInPlaneWorld);
        bool hasCollision = distance < res.manifoldPtr->breakingThreshold;
        RawContact rc;
        rc.uid0 = res.body0; rc.uid1 = res.body1; rc.tri = -1; rc.method = 11; rc.iters = 0;
        rc.hasContact = hasCollision ? 1 : 0;
        V3 normalOnSurfaceB = planeNormal;
        v3mul(normalOnSurfaceB, planeObj.xf.basis);
        rc.normal[0] = normalOnSurfaceB.x; rc.normal[1] = normalOnSurfaceB.y; rc.normal[2] = normalOnSurfaceB.z;
        rc.point[0] = vtxInPlaneWorld.x; rc.point[1] = vtxInPlaneWorld.y; rc.point[2] = vtxInPlaneWorld.z;
        rc.depth = distance;
        raw.push_back(rc);
        if (hasCollision) res.addContactPoint(normalOnSurfaceB, vtxInPlaneWorld, distance);
        if (res.manifoldPtr->cachedPoints != 0) res.refreshContactPoints();
    }

    void convexConcave(const Shape* cs, const Shape* ms, const Body& convexBody, const Body& triBody, ManifoldResult& res) {
        // disp/ConvexConcaveCollisionAlgorithm.java:65-93 + disp/ConvexTriangleCallback.java:83-172
        float collisionMarginTriangle = ms->getMargin();
        Xf convexInTriangleSpace; convexInTriangleSpace.set(triBody.xf);
        convexInTriangleSpace.inverse();
        convexInTriangleSpace.mul(convexBody.xf);
        V3 aabbMin, aabbMax;
        shapeGetAabb(*cs, convexInTriangleSpace, aabbMin, aabbMax);
        V3 extra(collisionMarginTriangle, collisionMarginTriangle, collisionMarginTriangle);
        aabbMax.add(extra);
        aabbMin.sub(extra);
        const MeshShapeData* md = meshes[triBody.shape].get();
        int visited = 0;
        md->bvh.reportAabbOverlappingNodex(aabbMin, aabbMax, [&](int part, int triIndex) {
            // sh/BvhTriangleMeshShape.java:265-278 processNode -> ConvexTriangleCallback.processTriangle
            Shape tm;
            tm.type = SH_TRIANGLE;
            md->mesh.getTriangle(part, triIndex, tm.tri);
            tm.collisionMargin = collisionMarginTriangle;
            res.partId0 = -1; res.index0 = -1; res.partId1 = part; res.index1 = triIndex;  // :164
            trianglesTested++;
            // findAlgorithm(convexBody, triBody, sharedManifold): convex-convex with ownManifold=false
            convexConvexTri(cs, &tm, convexBody, triBody, res, (part << 21) | triIndex);  // raw-record key: partId << 21 | index
        }, &visited);
        bvhNodesVisited += visited;
        res.refreshContactPoints();
    }
    // ConvexConvexAlgorithm.processCollision(convexBody, triBody, ..., resultOut) with shared manifold:
    // the raw record is tagged with the broadphase pair (resultOut bodies), transforms are convex/tri.
    void convexConvexTri(const Shape* cs, const Shape* tm, const Body& bc, const Body& bt, ManifoldResult& res, int tri) {
        float maxd = cs->getMargin() + tm->getMargin() + res.manifoldPtr->breakingThreshold;
        maxd *= maxd;
        GjkOut out;
        gjkChecks++;
        gjkGetClosestPoints(cs, tm, bc.xf, bt.xf, maxd, out);
        deepPenetrationChecks += out.deepPenetrationChecks;
        RawContact rc;
        rc.uid0 = res.body0; rc.uid1 = res.body1; rc.tri = tri;
        rc.hasContact = out.hasContact ? 1 : 0;
        rc.normal[0] = out.normalOnBInWorld.x; rc.normal[1] = out.normalOnBInWorld.y; rc.normal[2] = out.normalOnBInWorld.z;
        rc.point[0] = out.pointInWorld.x; rc.point[1] = out.pointInWorld.y; rc.point[2] = out.pointInWorld.z;
        rc.depth = out.depth; rc.method = out.lastUsedMethod; rc.iters = out.curIter;
        raw.push_back(rc);
        if (out.hasContact) res.addContactPoint(out.normalOnBInWorld, out.pointInWorld, out.depth);
    }

    // One child algorithm of a compound pair: findAlgorithm(colObj with its temporary child shape, otherObj) picked from
    // the table (disp/DefaultCollisionConfiguration.java:149-213), with its own manifold (getNewManifold(body0, body1) of the
    // child algorithm: disp/SphereSphereCollisionAlgorithm.java:49-60, disp/ConvexPlaneCollisionAlgorithm.java:57-67,
    // disp/ConvexConvexAlgorithm.java:92-96).  A = the compound's child, B = the other object (or its child).
    void compoundLeaf(const Shape* sa, int childA, const Body& a, const Shape* sb, int childB, const Body& b, ManifoldResult& res,
                      PairState& ps, int& k) {
        if ((int)ps.kids.size() <= k) { ps.kids.resize(k + 1); ps.kidChild.resize(k + 1, std::make_pair(-1, -1)); }
        PairState& kid = ps.kids[k];
        ps.kidChild[k] = std::make_pair(childA, childB);
        const int code = -2 - k;
        k++;
        kid.manifold.breakingThreshold = breakingThreshold;
        for (int q = 0; q < 4; q++) kid.manifold.pointCache[q].srcSlot = (q < kid.manifold.cachedPoints) ? q : -1;
        res.manifoldPtr = &kid.manifold;
        size_t rawBefore = raw.size();
        if (sa->type == SH_SPHERE && sb->type == SH_SPHERE) {
            if (!kid.hasManifold) { kid.hasManifold = true; kid.manifold.body0 = a.uid; kid.manifold.body1 = b.uid; }
            sphereSphere(sa, sb, a, b, res);
        } else if (sa->isConvex() && sb->type == SH_PLANE) {
            if (!kid.hasManifold) { kid.hasManifold = true; kid.manifold.body0 = a.uid; kid.manifold.body1 = b.uid; }
            convexPlane(sa, sb, a, b, res);
        } else if (sa->isConvex() && sb->isConvex()) {
            if (!kid.hasManifold) { kid.hasManifold = true; kid.manifold.body0 = a.uid; kid.manifold.body1 = b.uid; }
            convexConvex(sa, sb, a, b, res, true, code);
        } else if (sa->isConvex() && sb->type == SH_MESH) {
            // ConvexConcaveCollisionAlgorithm(child, mesh) with the child algorithm's own ConvexTriangleCallback manifold
            // (disp/ConvexConcaveCollisionAlgorithm.java:47-93, disp/ConvexTriangleCallback.java:58-66): setBodies(convex, tri)
            kid.hasManifold = true;
            kid.manifold.body0 = a.uid; kid.manifold.body1 = b.uid;
            convexConcave(sa, sb, a, b, res);
            res.partId0 = res.partId1 = res.index0 = res.index1 = 0;
            // raw-record key of (child algorithm k, triangle t): -2 - (k << 21 | t)   (t < 2^21, sh/OptimizedBvh.java:65)
            for (size_t r = rawBefore; r < raw.size(); r++) { raw[r].uid0 = res.body0; raw[r].uid1 = res.body1; raw[r].tri = -2 - (((k - 1) << 21) | raw[r].tri); }
            return;
        }
        for (size_t r = rawBefore; r < raw.size(); r++) { raw[r].uid0 = res.body0; raw[r].uid1 = res.body1; raw[r].tri = code; }
    }
    // disp/CompoundCollisionAlgorithm.java:83-129 processCollision: every child in turn, the compound's object temporarily
    // carrying the child's shape and orgTrans * childTrans; contact points are still projected with the original transforms
    // (res keeps rootTransA / rootTransB of the pair's two objects).  `other` may itself be a compound: its child algorithm is
    // then the swapped compound algorithm over ITS children, against this child (:57-75 init -> findAlgorithm).
    // A child that is itself a CompoundShape gets a nested (never swapped: the compound comes first in findAlgorithm's
    // arguments, disp/DefaultCollisionConfiguration.java:198-200) CompoundCollisionAlgorithm whose colObj carries
    // orgTrans * childTrans as its world transform, so the LEAVES are visited depth first with transforms composed level by
    // level.  leaf(shape, worldTransform, index of the leaf in that depth-first order)
    template <class F>
    void visitCompoundLeaves(const Shape& cs, const Xf& org, int& leafIndex, F leaf) const {
        for (size_t i = 0; i < cs.children.size(); i++) {
            Xf childWorld;
            childWorld.set(org);
            childWorld.mul(cs.children[i].transform);  // newChildWorldTrans.mul(orgTrans, childTrans) (lm/Transform.java:122-131)
            const Shape& child = shapes[cs.children[i].shape];
            if (child.isCompound()) visitCompoundLeaves(child, childWorld, leafIndex, leaf);
            else leaf(&child, childWorld, leafIndex++);
        }
    }
    void compoundProcess(const Body& colObj, const Body& otherObj, const Shape* otherShape, int otherChild, ManifoldResult& res,
                         PairState& ps, int& k) {
        const Shape& cs = shapes[colObj.shape];
        int li = 0;
        visitCompoundLeaves(cs, colObj.xf, li, [&](const Shape* childShape, const Xf& childWorld, int i) {
            Body tmp = colObj;                 // colObj.setWorldTransform(newChildWorldTrans)
            tmp.xf.set(childWorld);
            if (otherShape->isCompound()) {
                // algorithm(leaf i of colObj, compound other) = swapped CompoundCollisionAlgorithm: its colObj is `other`
                int lj = 0;
                visitCompoundLeaves(*otherShape, otherObj.xf, lj, [&](const Shape* oShape, const Xf& oWorld, int j) {
                    Body tmpO = otherObj;
                    tmpO.xf.set(oWorld);
                    compoundLeaf(oShape, j, tmpO, childShape, i, tmp, res, ps, k);
                });
            } else {
                compoundLeaf(childShape, i, tmp, otherShape, otherChild, otherObj, res, ps, k);
            }
        });
    }

    // Dispatcher.dispatchAllCollisionPairs over the current pair set.  Returns number of manifolds.
    int dispatchAllPairs() {
        raw.clear();
        // pairs that left the cache lose their algorithm + manifold (bp/HashedOverlappingPairCache.java:129-174
        // cleanOverlappingPair -> algorithm.destroy -> releaseManifold)
        {
            std::map<std::pair<int, int>, PairState> keep;
            for (auto& p : pairs) {
                auto it = pairState.find(p);
                if (it != pairState.end()) keep.insert(*it);
            }
            pairState.swap(keep);
        }
        int before = 0;
        for (auto& p : pairs) {
            const Body& b0 = bodies[p.first - 1];
            const Body& b1 = bodies[p.second - 1];
            // disp/CollisionDispatcher.java:198-223 needsCollision
            if (!b0.active && !b1.active) continue;
            // ... else if (!body0.checkCollideWith(body1)): bodies linked by a constraint that disables their collision
            // (dynamics/RigidBody.java:624-639; the constraint is registered on both bodies, so the relation is symmetric)
            if (!noCollide.empty() && noCollide.count(p)) continue;
            const Shape* s0 = &shapes[b0.shape];
            const Shape* s1 = &shapes[b1.shape];
            PairState& ps = pairState[p];
            ManifoldResult res;
            initResult(res, b0, b1);
            res.manifoldPtr = &ps.manifold;
            ps.manifold.breakingThreshold = breakingThreshold;
            for (int k = 0; k < 4; k++) ps.manifold.pointCache[k].srcSlot = (k < ps.manifold.cachedPoints) ? k : -1;
            before = res.addedContacts;
            // disp/DefaultCollisionConfiguration.java:149-213
            if (s0->type == SH_SPHERE && s1->type == SH_SPHERE) {
                if (!ps.hasManifold) { ps.hasManifold = true; ps.manifold.body0 = b0.uid; ps.manifold.body1 = b1.uid; }
                sphereSphere(s0, s1, b0, b1, res);
            } else if (s0->isConvex() && s1->type == SH_PLANE) {
                if (!ps.hasManifold) { ps.hasManifold = true; ps.manifold.body0 = b0.uid; ps.manifold.body1 = b1.uid; }
                convexPlane(s0, s1, b0, b1, res);
            } else if (s1->isConvex() && s0->type == SH_PLANE) {
                if (!ps.hasManifold) { ps.hasManifold = true; ps.manifold.body0 = b1.uid; ps.manifold.body1 = b0.uid; }
                convexPlane(s1, s0, b1, b0, res);
            } else if (s0->isConvex() && s1->isConvex()) {
                if (!ps.hasManifold) { ps.hasManifold = true; ps.manifold.body0 = b0.uid; ps.manifold.body1 = b1.uid; }
                convexConvex(s0, s1, b0, b1, res, true, -1);
            } else if (s0->isConvex() && s1->type == SH_MESH) {
                if (!ps.hasManifold) { ps.hasManifold = true; }
                ps.manifold.body0 = b0.uid; ps.manifold.body1 = b1.uid;  // setBodies(convexBody, triBody)
                convexConcave(s0, s1, b0, b1, res);
            } else if (s1->isConvex() && s0->type == SH_MESH) {
                if (!ps.hasManifold) { ps.hasManifold = true; }
                ps.manifold.body0 = b1.uid; ps.manifold.body1 = b0.uid;
                convexConcave(s1, s0, b1, b0, res);
            } else if (s0->isCompound()) {          // compoundCreateFunc (disp/DefaultCollisionConfiguration.java:198-200)
                ps.isCompound = true;
                int k = 0;
                compoundProcess(b0, b1, s1, -1, res, ps, k);
            } else if (s1->isCompound()) {          // swappedCompoundCreateFunc (:202-204)
                ps.isCompound = true;
                int k = 0;
                compoundProcess(b1, b0, s0, -1, res, ps, k);
            } else {
                // EmptyAlgorithm: no manifold
            }
            addedContacts += res.addedContacts - before;
        }
        int n = 0;
        for (auto& kv : pairState) {
            if (kv.second.hasManifold) n++;
            for (auto& kid : kv.second.kids)
                if (kid.hasManifold) n++;
        }
        return n;
    }
};

}  // namespace orc
