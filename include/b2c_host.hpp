// b2c_host.hpp — C++ host-side mirror of the reference's collision interfaces over the C ABI (b2c.h).
//
// The reference is compiled Java and this image has no JDK, so the host side above the C ABI is written in
// C++ (and, for the pytest harness, in Python: libgdx-jbullet_b200/world.py).  Class and method names follow
//   bp/BroadphaseInterface.java:33-52   -> GpuBroadphase
//   bp/OverlappingPairCache.java:34-54  -> GpuPairCache
//   bp/Dispatcher.java:38-68            -> GpuDispatcher
//   disp/CollisionWorld.java:98-245     -> GpuCollisionWorld
// with the same argument meaning and call order, so a port of a reference test reads the same.  Header-only;
// link with libb2c.so.  Errors surface as std::runtime_error carrying b2c_last_error_string.
#pragma once
#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "b2c.h"

namespace b2c_host {

struct Vector3 { float x, y, z; };
struct Transform { float basis[9]; float origin[3]; };  // row-major basis (lm/Transform.java:46-49)
struct BroadphasePair { int32_t proxy0, proxy1; };      // uids (bp/BroadphasePair.java:37-40)

inline void check(int32_t rc, b2c_ctx* ctx) {
    if (rc != B2C_OK) throw std::runtime_error("b2c error " + std::to_string(rc) + ": " + b2c_last_error_string(ctx));
}

class GpuCollisionWorld;

class GpuPairCache {
public:
    explicit GpuPairCache(b2c_ctx* c) : ctx(c) {}
    int getNumOverlappingPairs() const { return num; }
    const std::vector<BroadphasePair>& getOverlappingPairArray() {
        int32_t n = 0;
        check(b2c_get_pairs(ctx, nullptr, 0, &n), ctx);
        pairs.resize((size_t)n);
        if (n) check(b2c_get_pairs(ctx, reinterpret_cast<int32_t*>(pairs.data()), n, &n), ctx);
        return pairs;
    }
    // The add / remove events the reference's cache sends to OverlappingPairCallback / GhostPairCallback
    // (bp/HashedOverlappingPairCache.java:135-137, 323-325) for the last calculateOverlappingPairs.
    void getPairDeltas(std::vector<BroadphasePair>& added, std::vector<BroadphasePair>& removed) {
        int32_t na = 0, nr = 0;
        check(b2c_get_pair_deltas(ctx, nullptr, 0, nullptr, 0, &na, &nr), ctx);
        added.resize((size_t)na);
        removed.resize((size_t)nr);
        check(b2c_get_pair_deltas(ctx, na ? reinterpret_cast<int32_t*>(added.data()) : nullptr, na,
                                  nr ? reinterpret_cast<int32_t*>(removed.data()) : nullptr, nr, &na, &nr), ctx);
    }
    int num = 0;
private:
    b2c_ctx* ctx;
    std::vector<BroadphasePair> pairs;
};

class GpuDispatcher {
public:
    explicit GpuDispatcher(b2c_ctx* c) : ctx(c) {}
    void dispatchAllCollisionPairs(GpuPairCache*, void* /*dispatchInfo*/, GpuDispatcher*) {
        check(b2c_dispatch_all_pairs(ctx, &numManifolds, &numContactsAdded), ctx);
        fetched = false;
    }
    int getNumManifolds() const { return numManifolds; }
    const b2c_manifold& getManifoldByIndexInternal(int i) {
        if (!fetched) {
            int32_t n = 0;
            check(b2c_get_manifolds(ctx, nullptr, 0, 0, &n), ctx);
            manifolds.resize((size_t)n);
            if (n) check(b2c_get_manifolds(ctx, manifolds.data(), n, 0, &n), ctx);
            fetched = true;
        }
        return manifolds.at((size_t)i);
    }
    // The touching manifolds as the packed contact stream (16-byte headers, 48-byte points): what a solver-side host reads
    // every step instead of walking getManifoldByIndexInternal.  uid0 / uid1 of a record = pairs[header.pair_index].
    void getPackedContacts(std::vector<b2c_packed_header>& headers, std::vector<b2c_packed_point>& points) {
        int32_t nh = 0, np = 0;
        check(b2c_get_packed_contacts(ctx, nullptr, 0, nullptr, 0, &nh, &np), ctx);
        headers.resize((size_t)nh);
        points.resize((size_t)np);
        if (nh) check(b2c_get_packed_contacts(ctx, headers.data(), nh, points.data(), np, &nh, &np), ctx);
    }
    // The same manifolds as full 96-byte points with uid headers: what the Java shim's GpuDispatcher reads (b2c_get_contacts).
    void getContacts(std::vector<b2c_contact_header>& headers, std::vector<b2c_manifold_point>& points) {
        int32_t nh = 0, np = 0;
        check(b2c_get_contacts(ctx, nullptr, 0, nullptr, 0, &nh, &np), ctx);
        headers.resize((size_t)nh);
        points.resize((size_t)np);
        if (nh) check(b2c_get_contacts(ctx, headers.data(), nh, points.data(), np, &nh, &np), ctx);
    }
    int32_t numManifolds = 0, numContactsAdded = 0;
private:
    b2c_ctx* ctx;
    std::vector<b2c_manifold> manifolds;
    bool fetched = false;
};

class GpuBroadphase {
public:
    explicit GpuBroadphase(b2c_ctx* c) : ctx(c), cache(c) {}
    // setAabb (bp/BroadphaseInterface.java:39) for hosts that compute AABBs themselves
    void setAabb(int32_t uid, const Vector3& mn, const Vector3& mx, GpuDispatcher* = nullptr) {
        float mm[6] = {mn.x, mn.y, mn.z, mx.x, mx.y, mx.z};
        check(b2c_set_aabbs(ctx, 1, &uid, mm), ctx);
    }
    void destroyProxy(int32_t uid, GpuDispatcher* = nullptr) { check(b2c_proxy_destroy(ctx, uid), ctx); }
    void calculateOverlappingPairs(GpuDispatcher* = nullptr) {
        int32_t n = 0;
        check(b2c_calculate_overlapping_pairs(ctx, &n), ctx);
        cache.num = n;
    }
    GpuPairCache* getOverlappingPairCache() { return &cache; }
    void getBroadphaseAabb(Vector3& mn, Vector3& mx) {
        float a[3], b[3];
        check(b2c_get_broadphase_aabb(ctx, a, b), ctx);
        mn = {a[0], a[1], a[2]};
        mx = {b[0], b[1], b[2]};
    }
private:
    b2c_ctx* ctx;
    GpuPairCache cache;
};

class GpuCollisionWorld {
public:
    explicit GpuCollisionWorld(const b2c_config& cfg) {
        int32_t rc = b2c_create(&cfg, &ctx);
        if (rc != B2C_OK) throw std::runtime_error("b2c_create failed (" + std::to_string(rc) + "): no sm_100 device? there is no CPU fallback");
        broadphase = new GpuBroadphase(ctx);
        dispatcher = new GpuDispatcher(ctx);
    }
    ~GpuCollisionWorld() { delete broadphase; delete dispatcher; b2c_destroy(ctx); }
    GpuCollisionWorld(const GpuCollisionWorld&) = delete;
    GpuCollisionWorld& operator=(const GpuCollisionWorld&) = delete;

    int32_t BoxShape(const Vector3& he) { int32_t s; float h[3] = {he.x, he.y, he.z}; check(b2c_shape_register_box(ctx, h, -1.f, &s), ctx); return s; }
    int32_t SphereShape(float r) { int32_t s; check(b2c_shape_register_sphere(ctx, r, &s), ctx); return s; }
    int32_t ConvexHullShape(const std::vector<Vector3>& pts) {
        int32_t s;
        check(b2c_shape_register_hull(ctx, &pts[0].x, (int32_t)pts.size(), -1.f, &s), ctx);
        return s;
    }
    // new CompoundShape() + addChildShape(childTransforms12[i], childShapes[i]) in order (sh/CompoundShape.java:50-82)
    int32_t CompoundShape(int32_t n, const int32_t* childShapes, const float* childTransforms12) {
        int32_t s; check(b2c_shape_register_compound(ctx, n, childShapes, childTransforms12, &s), ctx); return s;
    }
    int32_t StaticPlaneShape(const Vector3& n, float c) { int32_t s; float v[3] = {n.x, n.y, n.z}; check(b2c_shape_register_plane(ctx, v, c, &s), ctx); return s; }

    // disp/CollisionWorld.java:102-121
    int32_t addCollisionObject(int32_t shape, const Transform& t, int16_t group = 1, int16_t mask = -1, bool isStatic = false) {
        int32_t uid = 0;
        float t12[12];
        for (int i = 0; i < 9; i++) t12[i] = t.basis[i];
        for (int i = 0; i < 3; i++) t12[9 + i] = t.origin[i];
        check(b2c_proxy_create(ctx, shape, t12, group, mask, isStatic ? 1 : 0, 0, &uid), ctx);
        numBodies = uid;
        return uid;
    }
    void removeCollisionObject(int32_t uid) { check(b2c_proxy_destroy(ctx, uid), ctx); }
    // transforms of bodies 1..n as the ABI's 12 SoA planes
    void setWorldTransformPlanes(int32_t n, const float* planes12) { check(b2c_set_transforms(ctx, n, nullptr, planes12), ctx); }
    void updateAabbs() { check(b2c_update_aabbs(ctx), ctx); }                            // :231-245
    void performDiscreteCollisionDetection() {                                           // :123-151
        updateAabbs();
        broadphase->calculateOverlappingPairs(dispatcher);
        dispatcher->dispatchAllCollisionPairs(broadphase->getOverlappingPairCache(), nullptr, dispatcher);
    }
    // CollisionWorld.rayTest + ClosestRayResultCallback for n rays (disp/CollisionWorld.java:553-590, 697-729)
    void rayTestClosest(int32_t n, const float* fromXyz, const float* toXyz, int32_t* uidOut, float* fractionOut, float* normalOut,
                        float* pointOut, int16_t group = 1, int16_t mask = -1) {
        check(b2c_ray_test_closest(ctx, n, fromXyz, toXyz, group, mask, uidOut, fractionOut, normalOut, pointOut), ctx);
    }
    // CollisionWorld.convexSweepTest + ClosestConvexResultCallback for n translational sweeps (disp/CollisionWorld.java:596-651,
    // 765-800); allowedCcdPenetration = getDispatchInfo().allowedCcdPenetration (bp/DispatcherInfo.java:44)
    void convexSweepTestClosest(int32_t n, const int32_t* castShapeIds, const float* basis9, const float* fromXyz, const float* toXyz,
                                int32_t* uidOut, float* fractionOut, float* normalOut, float* pointOut, int16_t group = 1,
                                int16_t mask = -1, float allowedCcdPenetration = 0.04f) {
        check(b2c_convex_sweep_closest(ctx, n, castShapeIds, basis9, fromXyz, toXyz, group, mask, allowedCcdPenetration, uidOut,
                                       fractionOut, normalOut, pointOut), ctx);
    }
    // the CCD motion-clamping sweeps of DiscreteDynamicsWorld.integrateTransforms (dyn/DiscreteDynamicsWorld.java:700-729)
    void ccdSweepNotMe(int32_t n, const int32_t* bodyUids, const float* ccdRadius, const float* predictedXyz, int32_t* hitUidOut,
                       float* fractionOut, float* normalOut, float* pointOut, float allowedCcdPenetration = 0.04f) {
        check(b2c_ccd_sweep_not_me(ctx, n, bodyUids, ccdRadius, predictedXyz, allowedCcdPenetration, hitUidOut, fractionOut, normalOut,
                                   pointOut), ctx);
    }
    // SimulationIslandManager.updateActivationState + storeIslandActivationState (disp/SimulationIslandManager.java:57-110)
    int32_t computeIslands(std::vector<int32_t>& tags) {
        int32_t n = 0;
        tags.resize((size_t)numBodies);
        check(b2c_compute_islands(ctx, tags.data(), numBodies, &n), ctx);
        return n;
    }
    // RigidBody.checkCollideWithOverride (dynamics/RigidBody.java:624-639): constraint-linked bodies are not dispatched
    void setNoCollidePairs(const std::vector<BroadphasePair>& links) {
        check(b2c_set_no_collide_pairs(ctx, (int32_t)links.size(), links.empty() ? nullptr : &links[0].proxy0), ctx);
    }
    // ---- the fast path of INTEGRATION.md §4: a host that owns the step loop ---------------------------------------
    // once: the uid-keyed packed contact stream and the pair-cache events are produced inside the step itself
    void enableFastPath() {
        check(b2c_set_contact_prefetch(ctx, 3), ctx);
        check(b2c_set_pair_delta_prefetch(ctx, 1), ctx);
    }
    struct FastStep {
        int32_t numPairs = 0, numManifolds = 0, numContactsAdded = 0;
        std::vector<BroadphasePair> added, removed;          // pair-cache events of the step
        std::vector<b2c_packed_uid_header> headers;          // touching manifolds
        std::vector<b2c_packed_point> points;
    };
    // b2c_set_transforms + b2c_step_device + b2c_get_pair_deltas + b2c_begin_contact_download + b2c_sync_counts +
    // b2c_get_packed_contacts_uid, with buffers of capacity capPairs pairs (resized to the counts on return)
    void stepFast(int32_t n, const float* planes12, int32_t capPairs, FastStep& out) {
        out.added.resize((size_t)capPairs); out.removed.resize((size_t)capPairs);
        out.headers.resize((size_t)capPairs); out.points.resize((size_t)capPairs * 4);
        check(b2c_set_transforms(ctx, n, nullptr, planes12), ctx);
        check(b2c_step_device(ctx), ctx);
        int32_t na = 0, nr = 0, nh = 0, np = 0;
        check(b2c_get_pair_deltas(ctx, &out.added[0].proxy0, capPairs, &out.removed[0].proxy0, capPairs, &na, &nr), ctx);
        check(b2c_begin_contact_download(ctx, out.headers.data(), capPairs, out.points.data(), capPairs * 4), ctx);
        check(b2c_sync_counts(ctx, &out.numPairs, &out.numManifolds, &out.numContactsAdded), ctx);
        check(b2c_get_packed_contacts_uid(ctx, out.headers.data(), capPairs, out.points.data(), capPairs * 4, &nh, &np), ctx);
        out.added.resize((size_t)na); out.removed.resize((size_t)nr);
        out.headers.resize((size_t)nh); out.points.resize((size_t)np);
    }
    GpuBroadphase* getBroadphase() { return broadphase; }
    GpuPairCache* getPairCache() { return broadphase->getOverlappingPairCache(); }
    GpuDispatcher* getDispatcher() { return dispatcher; }
    b2c_ctx* handle() { return ctx; }
    int32_t numBodies = 0;
private:
    b2c_ctx* ctx = nullptr;
    GpuBroadphase* broadphase = nullptr;
    GpuDispatcher* dispatcher = nullptr;
};

}  // namespace b2c_host
