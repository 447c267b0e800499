// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// C entry points over the CPU restatement, loaded with ctypes by tests/, smoke() and
// bench.py's cpu_baseline / --impl reference legs.  Nothing in the product links this.
#include <chrono>
#include <cstring>
#include "world.h"
#include "islands.h"
#include <algorithm>
#include <iterator>

using namespace orc;

static Xf xfFrom12(const float* t) {  // 9 basis floats row-major + 3 origin floats
    Xf x;
    for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) x.basis.m[r][c] = t[r * 3 + c];
    x.origin.set(t[9], t[10], t[11]);
    return x;
}

extern "C" {

void* orc_create(int mode) {
    World* w = new World();
    w->mode = mode;
    return w;
}
void orc_destroy(void* h) { delete (World*)h; }
// AxisSweep3(worldAabbMin, worldAabbMax) bounds; call before the first body (bp/AxisSweep3.java:43-58)
void orc_set_world_aabb(void* h, const float* mn, const float* mx) {
    World* w = (World*)h;
    w->worldAabbMin.set(mn[0], mn[1], mn[2]);
    w->worldAabbMax.set(mx[0], mx[1], mx[2]);
}
// bp/AxisSweep3Internal.java:201-216 for one point: 3 quantised coordinates
void orc_sap_quantize(void* h, const float* p, int isMax, unsigned* out3) {
    World* w = (World*)h;
    w->sapInit();
    w->sapQ.quantize(out3, V3(p[0], p[1], p[2]), isMax);
}
// RigidBody.checkCollideWithOverride (dynamics/RigidBody.java:624-639): body pairs that must not be dispatched
void orc_set_no_collide_pairs(void* h, int n, const int* uidPairs) {
    World* w = (World*)h;
    w->noCollide.clear();
    for (int i = 0; i < n; i++) {
        int a = uidPairs[2 * i], b = uidPairs[2 * i + 1];
        w->noCollide.insert(std::make_pair(a < b ? a : b, a < b ? b : a));
    }
}
void orc_set_brute_force(void* h, int on) { ((World*)h)->bruteForcePairs = on != 0; }
void orc_set_params(void* h, float breaking, float dbvtMargin, float predictedFrames) {
    World* w = (World*)h;
    w->breakingThreshold = breaking;
    w->dbvtMargin = dbvtMargin;
    w->predictedFrames = predictedFrames;
}

int orc_shape_box(void* h, float hx, float hy, float hz) {
    Shape s;
    initBox(s, V3(hx, hy, hz));
    return ((World*)h)->addShape(s);
}
int orc_shape_sphere(void* h, float r) {
    Shape s;
    initSphere(s, r);
    return ((World*)h)->addShape(s);
}
int orc_shape_hull(void* h, const float* pts, int n) {
    Shape s;
    initHull(s, pts, n);
    return ((World*)h)->addShape(s);
}
int orc_shape_plane(void* h, float nx, float ny, float nz, float c) {
    Shape s;
    initPlane(s, V3(nx, ny, nz), c);
    return ((World*)h)->addShape(s);
}
// sh/CompoundShape.java:50-82: n children (shape ids registered before) under local transforms (n x 12 floats)
int orc_shape_compound(void* h, int n, const int* childShapes, const float* childXf12) {
    World* w = (World*)h;
    for (int i = 0; i < n; i++) {
        if (childShapes[i] < 0 || childShapes[i] >= (int)w->shapes.size()) return -1;
        const Shape& c = w->shapes[childShapes[i]];
        if (!c.isConvex() && !c.isCompound()) return -1;   // a child may itself be a CompoundShape
    }
    return w->addCompound(n, childShapes, childXf12);
}
int orc_shape_mesh(void* h, const float* verts, int nv, const int* idx, int ntri) {
    return ((World*)h)->addMesh(verts, nv, idx, ntri);
}
int orc_shape_mesh_parts(void* h, int nparts, const float* verts, const int* nv, const int* idx, const int* ntri) {
    return ((World*)h)->addMeshParts(nparts, verts, nv, idx, ntri);
}
int orc_mesh_num_nodes(void* h, int shape) { return (int)((World*)h)->meshes[shape]->bvh.nodes.size(); }
void orc_mesh_get_nodes(void* h, int shape, void* out16B) {
    Bvh& b = ((World*)h)->meshes[shape]->bvh;
    std::memcpy(out16B, b.nodes.data(), b.nodes.size() * sizeof(QNode));
}
void orc_mesh_get_quant(void* h, int shape, float* out9) {
    Bvh& b = ((World*)h)->meshes[shape]->bvh;
    out9[0] = b.bvhAabbMin.x; out9[1] = b.bvhAabbMin.y; out9[2] = b.bvhAabbMin.z;
    out9[3] = b.bvhAabbMax.x; out9[4] = b.bvhAabbMax.y; out9[5] = b.bvhAabbMax.z;
    out9[6] = b.bvhQuantization.x; out9[7] = b.bvhQuantization.y; out9[8] = b.bvhQuantization.z;
}

int orc_body_create(void* h, int shape, const float* xf12, int group, int mask, int isStatic, int world) {
    return ((World*)h)->addBody(shape, xfFrom12(xf12), group, mask, isStatic != 0, world);
}
void orc_body_destroy(void* h, int uid) { ((World*)h)->removeBody(uid); }
int orc_num_bodies(void* h) { return (int)((World*)h)->bodies.size(); }

// xf: n x 12 floats; uids may be NULL (= bodies 1..n)
void orc_set_transforms(void* h, int n, const int* uids, const float* xf) {
    World* w = (World*)h;
    for (int i = 0; i < n; i++) {
        int uid = uids ? uids[i] : i + 1;
        w->bodies[uid - 1].xf = xfFrom12(xf + 12 * i);
    }
}
void orc_set_active(void* h, int n, const int* uids, const unsigned char* active) {
    World* w = (World*)h;
    for (int i = 0; i < n; i++) {
        int uid = uids ? uids[i] : i + 1;
        w->bodies[uid - 1].active = active[i] != 0;
    }
}
void orc_set_material(void* h, int uid, float friction, float restitution) {
    World* w = (World*)h;
    w->bodies[uid - 1].friction = friction;
    w->bodies[uid - 1].restitution = restitution;
}
void orc_update_aabbs(void* h) { ((World*)h)->updateAabbs(); }
void orc_set_aabb(void* h, int uid, const float* mn, const float* mx) {
    World* w = (World*)h;
    w->setAabb(w->bodies[uid - 1], V3(mn[0], mn[1], mn[2]), V3(mx[0], mx[1], mx[2]));
}
// out: n x 6 floats (effective AABB min xyz, max xyz)
void orc_get_aabbs(void* h, float* out) {
    World* w = (World*)h;
    for (size_t i = 0; i < w->bodies.size(); i++) {
        const Body& b = w->bodies[i];
        out[6 * i + 0] = b.effMin.x; out[6 * i + 1] = b.effMin.y; out[6 * i + 2] = b.effMin.z;
        out[6 * i + 3] = b.effMax.x; out[6 * i + 4] = b.effMax.y; out[6 * i + 5] = b.effMax.z;
    }
}
int orc_calculate_overlapping_pairs(void* h) { return ((World*)h)->calculateOverlappingPairs(); }
void orc_get_pairs(void* h, int* out) {
    World* w = (World*)h;
    for (size_t i = 0; i < w->pairs.size(); i++) {
        out[2 * i] = w->pairs[i].first;
        out[2 * i + 1] = w->pairs[i].second;
    }
}
// Pairs added to / removed from the cache by the last calculateOverlappingPairs: what
// bp/HashedOverlappingPairCache.java:323-325 (add) and :135-137 (remove) report to the ghost pair callback.
// out arrays hold (uid0, uid1) rows, sorted; returns the counts through n2.
void orc_pair_deltas(void* h, int* addedOut, int capA, int* removedOut, int capR, int* n2) {
    World* w = (World*)h;
    std::vector<std::pair<int, int>> added, removed;
    std::set_difference(w->pairs.begin(), w->pairs.end(), w->prevPairs.begin(), w->prevPairs.end(), std::back_inserter(added));
    std::set_difference(w->prevPairs.begin(), w->prevPairs.end(), w->pairs.begin(), w->pairs.end(), std::back_inserter(removed));
    for (int i = 0; i < (int)added.size() && i < capA; i++) { addedOut[2 * i] = added[i].first; addedOut[2 * i + 1] = added[i].second; }
    for (int i = 0; i < (int)removed.size() && i < capR; i++) { removedOut[2 * i] = removed[i].first; removedOut[2 * i + 1] = removed[i].second; }
    n2[0] = (int)added.size();
    n2[1] = (int)removed.size();
}
// disp/SimulationIslandManager.java:57-110 over the current pair list; tags[i] for object i (uid i+1), -1 = static.
int orc_islands(void* h, int* tagsOut) {
    World* w = (World*)h;
    std::vector<char> merges(w->bodies.size());
    for (size_t i = 0; i < w->bodies.size(); i++) merges[i] = (w->bodies[i].alive && !w->bodies[i].isStatic) ? 1 : 0;
    return islandTags(w->pairs, merges, tagsOut);
}
// CollisionWorld.rayTest + ClosestRayResultCallback for n rays: from/to 3 floats each; out: uid (0 = miss), fraction,
// normal xyz, point xyz
void orc_ray_test_closest(void* h, int n, const float* from, const float* to, int group, int mask, int* uidOut, float* out7) {
    World* w = (World*)h;
    for (int i = 0; i < n; i++) {
        RayHit r = w->rayTestClosest(V3(from[3 * i], from[3 * i + 1], from[3 * i + 2]), V3(to[3 * i], to[3 * i + 1], to[3 * i + 2]), group, mask);
        uidOut[i] = r.uid;
        out7[7 * i] = r.fraction;
        out7[7 * i + 1] = r.normal.x; out7[7 * i + 2] = r.normal.y; out7[7 * i + 3] = r.normal.z;
        out7[7 * i + 4] = r.point.x; out7[7 * i + 5] = r.point.y; out7[7 * i + 6] = r.point.z;
    }
}
// CollisionWorld.convexSweepTest + ClosestConvexResultCallback for n translational sweeps of registered convex shapes:
// basis 9 floats (row-major) per sweep, from / to 3 floats each; out: uid (0 = miss, -1 = the sweep met a static plane),
// fraction, normal xyz, point xyz
void orc_convex_sweep_closest(void* h, int n, const int* shapes, const float* basis9, const float* from, const float* to, int group,
                              int mask, float allowedPenetration, int* uidOut, float* out7) {
    World* w = (World*)h;
    for (int i = 0; i < n; i++) {
        float f12[12], t12[12];
        for (int k = 0; k < 9; k++) f12[k] = t12[k] = basis9[9 * i + k];
        for (int k = 0; k < 3; k++) { f12[9 + k] = from[3 * i + k]; t12[9 + k] = to[3 * i + k]; }
        ConvexSweepHit r = w->convexSweepClosest(shapes[i], xfFrom12(f12), xfFrom12(t12), group, mask, allowedPenetration);
        uidOut[i] = r.unsupported ? -1 : r.uid;
        out7[7 * i] = r.fraction;
        out7[7 * i + 1] = r.normal.x; out7[7 * i + 2] = r.normal.y; out7[7 * i + 3] = r.normal.z;
        out7[7 * i + 4] = r.point.x; out7[7 * i + 5] = r.point.y; out7[7 * i + 6] = r.point.z;
    }
}
// The CCD motion-clamping sweeps of DiscreteDynamicsWorld.integrateTransforms for n bodies: out as orc_convex_sweep_closest
void orc_ccd_sweep_not_me(void* h, int n, const int* uids, const float* radius, const float* to, float allowedPenetration, int* uidOut,
                          float* out7) {
    World* w = (World*)h;
    for (int i = 0; i < n; i++) {
        ConvexSweepHit r = w->ccdSweepNotMe(uids[i], radius[i], V3(to[3 * i], to[3 * i + 1], to[3 * i + 2]), allowedPenetration);
        uidOut[i] = r.unsupported ? -1 : r.uid;
        out7[7 * i] = r.fraction;
        out7[7 * i + 1] = r.normal.x; out7[7 * i + 2] = r.normal.y; out7[7 * i + 3] = r.normal.z;
        out7[7 * i + 4] = r.point.x; out7[7 * i + 5] = r.point.y; out7[7 * i + 6] = r.point.z;
    }
}
int orc_dispatch_all_pairs(void* h) { return ((World*)h)->dispatchAllPairs(); }
int orc_num_raw(void* h) { return (int)((World*)h)->raw.size(); }
// raw record: 5 ints (uid0, uid1, tri, hasContact, method) + iters ; 7 floats (normal, point, depth)
void orc_get_raw(void* h, int* ints6, float* floats7) {
    World* w = (World*)h;
    for (size_t i = 0; i < w->raw.size(); i++) {
        const RawContact& r = w->raw[i];
        ints6[6 * i + 0] = r.uid0; ints6[6 * i + 1] = r.uid1; ints6[6 * i + 2] = r.tri;
        ints6[6 * i + 3] = r.hasContact; ints6[6 * i + 4] = r.method; ints6[6 * i + 5] = r.iters;
        for (int k = 0; k < 3; k++) { floats7[7 * i + k] = r.normal[k]; floats7[7 * i + 3 + k] = r.point[k]; }
        floats7[7 * i + 6] = r.depth;
    }
}
// Manifolds in pair order (the child manifolds of a compound pair in the order its child algorithms run).  Per manifold:
// hdr7 = (pairUid0, pairUid1, body0, body1, nContacts, child index in body0's compound shape or -1, the same for body1);
// per point (4 slots): 18 floats = localA(3) localB(3) worldA(3) worldB(3) normal(3) distance friction restitution,
// and 6 ints = lifeTime, srcSlot, partId0, partId1, index0, index1.
static void emitManifold(const std::pair<int, int>& key, const PersistentManifold& m, int child0, int child1, int n, int* hdr7,
                         float* pts, int* pint) {
    hdr7[7 * n + 0] = key.first; hdr7[7 * n + 1] = key.second;
    hdr7[7 * n + 2] = m.body0; hdr7[7 * n + 3] = m.body1; hdr7[7 * n + 4] = m.cachedPoints;
    hdr7[7 * n + 5] = child0; hdr7[7 * n + 6] = child1;
    for (int k = 0; k < 4; k++) {
        const ManifoldPoint& p = m.pointCache[k];
        float* f = pts + ((size_t)n * 4 + k) * 18;
        int* ii = pint + ((size_t)n * 4 + k) * 6;
        if (k >= m.cachedPoints) {
            for (int q = 0; q < 18; q++) f[q] = 0;
            for (int q = 0; q < 6; q++) ii[q] = 0;
            continue;
        }
        f[0] = p.localPointA.x; f[1] = p.localPointA.y; f[2] = p.localPointA.z;
        f[3] = p.localPointB.x; f[4] = p.localPointB.y; f[5] = p.localPointB.z;
        f[6] = p.positionWorldOnA.x; f[7] = p.positionWorldOnA.y; f[8] = p.positionWorldOnA.z;
        f[9] = p.positionWorldOnB.x; f[10] = p.positionWorldOnB.y; f[11] = p.positionWorldOnB.z;
        f[12] = p.normalWorldOnB.x; f[13] = p.normalWorldOnB.y; f[14] = p.normalWorldOnB.z;
        f[15] = p.distance1; f[16] = p.combinedFriction; f[17] = p.combinedRestitution;
        ii[0] = p.lifeTime; ii[1] = p.srcSlot; ii[2] = p.partId0; ii[3] = p.partId1;
        ii[4] = p.index0; ii[5] = p.index1;
    }
}
int orc_get_manifolds(void* h, int cap, int* hdr7, float* pts /*cap*4*18*/, int* pint /*cap*4*6*/) {
    World* w = (World*)h;
    int n = 0;
    for (auto& kv : w->pairState) {
        if (kv.second.hasManifold) {
            if (n < cap) emitManifold(kv.first, kv.second.manifold, -1, -1, n, hdr7, pts, pint);
            n++;
        }
        for (size_t k = 0; k < kv.second.kids.size(); k++) {
            if (!kv.second.kids[k].hasManifold) continue;
            if (n < cap) emitManifold(kv.first, kv.second.kids[k].manifold, kv.second.kidChild[k].first, kv.second.kidChild[k].second, n, hdr7, pts, pint);
            n++;
        }
    }
    return n;
}
void orc_get_counters(void* h, long* out5) {
    World* w = (World*)h;
    out5[0] = w->gjkChecks; out5[1] = w->deepPenetrationChecks; out5[2] = w->addedContacts;
    out5[3] = w->bvhNodesVisited; out5[4] = w->trianglesTested;
}

void orc_epa_debug(long* out /*8 + 36*/) {
    out[0] = g_epaStats.calls; out[1] = g_epaStats.epaIters; out[2] = g_epaStats.epaItersMax; out[3] = g_epaStats.faces;
    out[4] = g_epaStats.facesMax; out[5] = g_epaStats.gjkIters; out[6] = g_epaStats.gjkItersMax; out[7] = 0;
    for (int i = 0; i < 9; i++) for (int j = 0; j < 4; j++) out[8 + i * 4 + j] = g_epaStats.hist[i][j];
}

// ---- stand-alone kernels for known-answer tests ------------------------------------------
// shape AABB for a shape under a transform (no +-threshold)
void orc_shape_aabb(void* h, int shape, const float* xf12, float* out6) {
    World* w = (World*)h;
    V3 mn, mx;
    shapeGetAabb(w->shapes[shape], xfFrom12(xf12), mn, mx);
    out6[0] = mn.x; out6[1] = mn.y; out6[2] = mn.z; out6[3] = mx.x; out6[4] = mx.y; out6[5] = mx.z;
}
// one GjkPairDetector.getClosestPoints call; out: hasContact, method, iters, degenerate ; normal, point, depth
void orc_gjk_pair(void* h, int shapeA, const float* xfA, int shapeB, const float* xfB, int* outi4, float* outf7) {
    World* w = (World*)h;
    const Shape* a = &w->shapes[shapeA];
    const Shape* b = &w->shapes[shapeB];
    float maxd = a->getMargin() + b->getMargin() + w->breakingThreshold;
    maxd *= maxd;
    GjkOut out;
    gjkGetClosestPoints(a, b, xfFrom12(xfA), xfFrom12(xfB), maxd, out);
    outi4[0] = out.hasContact; outi4[1] = out.lastUsedMethod; outi4[2] = out.curIter; outi4[3] = out.degenerateSimplex;
    outf7[0] = out.normalOnBInWorld.x; outf7[1] = out.normalOnBInWorld.y; outf7[2] = out.normalOnBInWorld.z;
    outf7[3] = out.pointInWorld.x; outf7[4] = out.pointInWorld.y; outf7[5] = out.pointInWorld.z;
    outf7[6] = out.depth;
}
// support vertex (with or without margin)
void orc_support(void* h, int shape, const float* dir, int withMargin, float* out3) {
    World* w = (World*)h;
    V3 o;
    if (withMargin) localGetSupportingVertex(w->shapes[shape], V3(dir[0], dir[1], dir[2]), o);
    else localGetSupportingVertexWithoutMargin(w->shapes[shape], V3(dir[0], dir[1], dir[2]), o);
    out3[0] = o.x; out3[1] = o.y; out3[2] = o.z;
}
// BVH query: returns number of triangles reported (first cap written)
int orc_bvh_query(void* h, int shape, const float* mn, const float* mx, int* outTris, int cap) {
    World* w = (World*)h;
    int n = 0;
    w->meshes[shape]->bvh.reportAabbOverlappingNodex(V3(mn[0], mn[1], mn[2]), V3(mx[0], mx[1], mx[2]), [&](int, int tri) {
        if (n < cap) outTris[n] = tri;
        n++;
    });
    return n;
}
// the two trig constants EncloseOrigin's line case needs (np/GjkEpaSolver.java:449-451)
void orc_epa_constants(float* out2) {
    float angle = SIMD_2_PI_ / 3.0f;
    out2[0] = (float)std::sin((double)(angle * 0.5f));
    out2[1] = (float)std::cos((double)(angle * 0.5f));
}

// One full collision step timed on this thread (for bench.py's CPU arm): returns seconds.
double orc_timed_step(void* h, int n, const float* xf, int* npairs, int* nmanifolds) {
    World* w = (World*)h;
    auto t0 = std::chrono::steady_clock::now();
    orc_set_transforms(h, n, nullptr, xf);
    w->updateAabbs();
    *npairs = w->calculateOverlappingPairs();
    *nmanifolds = w->dispatchAllPairs();
    auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
