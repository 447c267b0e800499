// narrowphase.cuh — stage 2: per-pair contact generation and the persistent 4-point manifolds.
//
// Replaces, for one step:
//   disp/CollisionDispatcher.java:252-257 dispatchAllCollisionPairs + disp/DefaultNearCallback.java:39-66
//       + :198-223 needsCollision + disp/DefaultCollisionConfiguration.java:149-213    -> k_carry, k_classify
//   disp/SphereSphereCollisionAlgorithm.java:73-134                                      -> k_sphere_sphere
//   disp/ConvexPlaneCollisionAlgorithm.java:75-136                                       -> k_convex_plane
//   disp/ConvexConvexAlgorithm.java:90-139 + np/GjkPairDetector.java:73-316              -> k_gjk, k_epa
//   disp/ConvexConcaveCollisionAlgorithm.java:65-93 + disp/ConvexTriangleCallback.java:83-172
//       + sh/OptimizedBvh.java:709-740,940-997                                           -> k_mesh_query, k_gjk_tri, k_mesh_manifold
//   disp/ManifoldResult.java:92-191 + np/PersistentManifold.java:83-372                  -> manifoldAdd / manifoldRefresh
//
// Pairs are binned by algorithm and shape-type pair so that a warp runs one code path; each pair is one
// thread (the simplex lives in registers).  Manifolds are one 416-byte record per pair, in pair order,
// carried from step to step by matching the sorted pair keys.
#pragma once
#include "../../include/b2c.h"
#include "common.cuh"
#include "epa.cuh"
#include "gjk.cuh"
#ifndef B2C_HOST_EMULATION
#include "radix_sort.cuh"
#else
// tests/emu compiles this header for the HOST (one emulated lane, test infrastructure only): the look-back primitives of the
// sorter become plain loads and stores there, and its launch code is not needed
namespace b2c {
constexpr uint32_t RS_FLAG_AGG = 0x40000000u, RS_FLAG_INC = 0x80000000u, RS_VAL_MASK = 0x3fffffffu;
inline void rs_store_release(uint32_t* p, uint32_t v) { *p = v; }
inline uint32_t rs_load_relaxed(const uint32_t* p) { return *p; }
}
#endif

namespace b2c {

// Convex-convex pairs are binned by which side is a hull (the only support mapping with a loop; box and sphere are a few
// selects): BIN_GJK0 + 2*(A is hull) + (B is hull) for the pairs the prefilter looks at, BIN_PS0 + the same for the pairs that
// went past the prefilter last step (history byte) — those go straight to the survivor list.
// BIN_COMPOUND: pairs with a CompoundShape on either side (compound.cuh expands them into child work items);
// BIN_COMPOUND_KEEP: such pairs that are not dispatched this step but own child manifolds that must stay alive.
enum { BIN_SKIP = 0, BIN_SS = 1, BIN_CP = 2, BIN_MESH = 3, BIN_GJK0 = 4, BIN_PS0 = 8, BIN_COUNT = 12, BIN_COMPOUND = 12,
       BIN_COMPOUND_KEEP = 13 };

// Device-side split of b2c_manifold: the 32-byte header every kernel streams, and the point slots only the
// touching pairs read.  The ABI's 416-byte record is assembled when results are copied out.
struct ManifoldHdr {
    int pair_uid0, pair_uid1, body0, body1, num_contacts, algorithm, pad0, pad1;
};
struct MView {
    ManifoldHdr* h;
    b2c_manifold_point* p;
};

struct NpArgs {
    const int2* pairs;            // sorted (uid0, uid1)
    uint32_t* numPairs;           // device
    const float4* xf4;
    const int* shape;
    const uint8_t* flags;
    const float2* material;
    const ShapeDev* shapes;
    const float4* hullPts;
    const MeshDev* meshes;
    ManifoldHdr* mhdr;            // [maxPairs] this step: 32-byte manifold headers (SoA: streamed by every kernel)
    b2c_manifold_point* mpts;     // [4*maxPairs] this step: the 4 point slots of each manifold
    b2c_raw_contact* raw;         // [maxPairs]
    int8_t* rawFlag;              // [maxPairs] copy of raw[p].has_contact for the kernels that only need the flag
    int wantRaw;                  // b2c_set_raw_records: also write the raw records nobody on the device reads (inspection)
    uint8_t* hist;                // [maxPairs] GJK iterations each pair needed last step (0 = none / new pair, capped at 15):
                                  //   written by k_carry from the manifold header word k_gjk leaves there
    uint8_t* binOf;               // [maxPairs] bin of every pair (k_classify)
    uint32_t* binItems;           // [maxPairs] pair indices, stably partitioned by bin (k_partition16)
    uint32_t* binStart;           // [17] exclusive bin offsets; [b+1] = end of bin b
    uint32_t* binZero;            // cleared per dispatch: hist[16] | ticket | pad[15] | status[tiles][16]
    StepCounters* ctr;
    float threshold;
    uint32_t maxPairs;
    const uint64_t* noCollide;    // sorted (uid0 << uidBits | uid1) keys of body pairs that are never dispatched, or null
    uint32_t numNoCollide;
    int uidBits;
    int hasCompound;              // a CompoundShape is registered: k_classify also looks at the pairs it does not dispatch
};

__device__ __forceinline__ MView mview(const NpArgs& a, uint32_t p) {
    MView m;
    m.h = a.mhdr + p;
    m.p = a.mpts + 4 * (size_t)p;
    return m;
}

// ---- manifold (np/PersistentManifold.java, disp/ManifoldResult.java) ---------------------------------
__device__ __forceinline__ MView mview(const struct NpArgs& a, uint32_t p);
__device__ __forceinline__ f3 ld3(const float* p) { return mk3(p[0], p[1], p[2]); }
__device__ __forceinline__ void st3(float* p, f3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// np/PersistentManifold.java:83-156 sortCachedPoints + lm/VectorUtil.java:60-90 closestAxis4
__device__ __forceinline__ int manifoldSortCachedPoints(const MView& m, f3 newLocalA, float newDist) {
    int maxPenetrationIndex = -1;
    float maxPenetration = newDist;
    for (int i = 0; i < 4; i++)
        if (m.p[i].distance < maxPenetration) { maxPenetrationIndex = i; maxPenetration = m.p[i].distance; }
    f3 p0 = ld3(m.p[0].local_a), p1 = ld3(m.p[1].local_a), p2 = ld3(m.p[2].local_a), p3 = ld3(m.p[3].local_a);
    float res0 = 0.f, res1 = 0.f, res2 = 0.f, res3 = 0.f;
    if (maxPenetrationIndex != 0) res0 = len2_3(crs3(sub3(newLocalA, p1), sub3(p3, p2)));
    if (maxPenetrationIndex != 1) res1 = len2_3(crs3(sub3(newLocalA, p0), sub3(p3, p2)));
    if (maxPenetrationIndex != 2) res2 = len2_3(crs3(sub3(newLocalA, p0), sub3(p3, p1)));
    if (maxPenetrationIndex != 3) res3 = len2_3(crs3(sub3(newLocalA, p0), sub3(p2, p1)));
    res0 = fabsf(res0); res1 = fabsf(res1); res2 = fabsf(res2); res3 = fabsf(res3);
    int maxIndex = -1;
    float maxVal = -1e30f;
    if (res0 > maxVal) { maxIndex = 0; maxVal = res0; }
    if (res1 > maxVal) { maxIndex = 1; maxVal = res1; }
    if (res2 > maxVal) { maxIndex = 2; maxVal = res2; }
    if (res3 > maxVal) { maxIndex = 3; maxVal = res3; }
    return maxIndex < 0 ? 0 : maxIndex;
}

// disp/ManifoldResult.java:92-157 addContactPoint.  rootA/rootB are the transforms of the PAIR's
// body0/body1 (ManifoldResult.init, :70-75); pairBody0 = uid of the pair's first body.
__device__ __forceinline__ bool manifoldAdd(const MView& m, int pairBody0, const Xf& rootA, const Xf& rootB, f3 normal, f3 point,
                                            float depth, float threshold, float friction, float restitution, int partId1,
                                            int index1) {
    if (depth > threshold) return false;
    bool isSwapped = m.h->body0 != pairBody0;
    f3 pointA = add3(scl3(normal, depth), point);
    f3 localA, localB;
    if (isSwapped) { localA = invXfPoint(rootB, pointA); localB = invXfPoint(rootA, point); }
    else { localA = invXfPoint(rootA, pointA); localB = invXfPoint(rootB, point); }
    // np/PersistentManifold.java:214-233 getCacheEntry
    float shortest = threshold * threshold;
    int nearest = -1;
    int size = m.h->num_contacts;
    for (int i = 0; i < size; i++) {
        f3 d = sub3(ld3(m.p[i].local_a), localA);
        float dd = dot3(d, d);
        if (dd < shortest) { shortest = dd; nearest = i; }
    }
    int life = 0, src = -1;
    int idx;
    if (nearest >= 0) {  // :280-305 replaceContactPoint keeps lifetime (and the solver's cached impulses)
        idx = nearest;
        life = m.p[idx].life_time;
        src = m.p[idx].src_slot;
    } else {             // :235-257 addManifoldPoint
        idx = size;
        if (idx == 4) idx = manifoldSortCachedPoints(m, localA, depth);
        else m.h->num_contacts = size + 1;
    }
    b2c_manifold_point* p = &m.p[idx];
    st3(p->local_a, localA); st3(p->local_b, localB);
    st3(p->world_a, pointA); st3(p->world_b, point);
    st3(p->normal_on_b, normal);
    p->distance = depth;
    p->combined_friction = friction;
    p->combined_restitution = restitution;
    p->life_time = life;
    p->src_slot = src;
    p->part_id1 = partId1;
    p->index1 = index1;
    return true;
}

// np/PersistentManifold.java:312-372 refreshContactPoints(trA, trB) incl. :259-278 removeContactPoint
__device__ __forceinline__ void manifoldRefresh(const MView& m, const Xf& trA, const Xf& trB, float threshold) {
    for (int i = m.h->num_contacts - 1; i >= 0; i--) {
        b2c_manifold_point* p = &m.p[i];
        f3 wa = xfPoint(trA, ld3(p->local_a));
        f3 wb = xfPoint(trB, ld3(p->local_b));
        st3(p->world_a, wa); st3(p->world_b, wb);
        p->distance = dot3(sub3(wa, wb), ld3(p->normal_on_b));
        p->life_time++;
    }
    for (int i = m.h->num_contacts - 1; i >= 0; i--) {
        b2c_manifold_point* p = &m.p[i];
        bool remove = false;
        if (!(p->distance <= threshold)) {
            remove = true;
        } else {
            f3 n = ld3(p->normal_on_b);
            f3 projected = sub3(ld3(p->world_a), scl3(n, p->distance));
            f3 diff = sub3(ld3(p->world_b), projected);
            if (dot3(diff, diff) > threshold * threshold) remove = true;
        }
        if (remove) {
            int last = m.h->num_contacts - 1;
            if (i != last) {
                m.p[i] = m.p[last];
                m.p[last].life_time = 0;
                m.p[last].src_slot = -1;
            }
            m.h->num_contacts = last;
        }
    }
}
// disp/ManifoldResult.java:177-191
__device__ __forceinline__ void resultRefresh(const MView& m, int pairBody0, const Xf& rootA, const Xf& rootB, float threshold) {
    if (m.h->num_contacts == 0) return;
    if (m.h->body0 != pairBody0) manifoldRefresh(m, rootB, rootA, threshold);
    else manifoldRefresh(m, rootA, rootB, threshold);
}
// disp/ManifoldResult.java:160-175
__device__ __forceinline__ float combinedFriction(float f0, float f1) {
    float f = f0 * f1;
    if (f < -10.f) f = -10.f;
    if (f > 10.f) f = 10.f;
    return f;
}

// ---- the same manifold step with 128-bit accesses (single addContactPoint + refresh callers) ------------------
// One 96-byte point = six int4: q0 = (localA.xyz, localB.x) q1 = (localB.yz, worldA.xy) q2 = (worldA.z, worldB.xyz)
// q3 = (normal.xyz, distance) q4 = (friction, restitution, lifeTime, src_slot) q5 = (partId1, index1, pad, pad).
// manifoldStepV does exactly what "mark src_slot; manifoldAdd; resultRefresh" does above — same float operations in
// the same order — but reads the header once, fetches whole points with independent 16-byte loads and writes them
// back the same way, so a thread issues a handful of memory round trips instead of dozens of scalar ones.
struct PtV {
    int4 q0, q1, q2, q3, q4, q5;
};
__device__ __forceinline__ float i2f(int v) { return __int_as_float(v); }
__device__ __forceinline__ int f2i(float v) { return __float_as_int(v); }
__device__ __forceinline__ PtV ldPtV(const b2c_manifold_point* p) {
    const int4* s = reinterpret_cast<const int4*>(p);
    PtV r;
    r.q0 = s[0]; r.q1 = s[1]; r.q2 = s[2]; r.q3 = s[3]; r.q4 = s[4]; r.q5 = s[5];
    return r;
}
__device__ __forceinline__ void stPtV(b2c_manifold_point* p, const PtV& r) {
    int4* d = reinterpret_cast<int4*>(p);
    d[0] = r.q0; d[1] = r.q1; d[2] = r.q2; d[3] = r.q3; d[4] = r.q4; d[5] = r.q5;
}

// hdr0/hdr1: the manifold header already in registers (hdr1.x = num_contacts, hdr1.y = algorithm); the caller has set
// algorithm / bodies for a new manifold.  Returns true when a contact was added.  Writes the header back.
__device__ __forceinline__ bool manifoldStepV(ManifoldHdr* H, b2c_manifold_point* P, int4 hdr0, int4 hdr1, int pairBody0,
                                              const Xf& rootA, const Xf& rootB, bool has, f3 normal, f3 point, float depth,
                                              float threshold, float friction, float restitution) {
    int nc = hdr1.x;
    const bool isSwapped = hdr0.z != pairBody0;
    int newSlot = -1;      // slot that holds a point created this step (src_slot -1); replaced points keep their slot
    bool added = false;
    if (has && !(depth > threshold)) {  // disp/ManifoldResult.java:96
        added = true;
        f3 pointA = add3(scl3(normal, depth), point);
        f3 localA, localB;
        if (isSwapped) { localA = invXfPoint(rootB, pointA); localB = invXfPoint(rootA, point); }
        else { localA = invXfPoint(rootA, pointA); localB = invXfPoint(rootB, point); }
        // getCacheEntry (np/PersistentManifold.java:214-233): the first 16 bytes of every live point
        int4 a0 = make_int4(0, 0, 0, 0), a1 = a0, a2 = a0, a3 = a0;
        if (nc > 0) a0 = *reinterpret_cast<const int4*>(P);
        if (nc > 1) a1 = *reinterpret_cast<const int4*>(P + 1);
        if (nc > 2) a2 = *reinterpret_cast<const int4*>(P + 2);
        if (nc > 3) a3 = *reinterpret_cast<const int4*>(P + 3);
        float shortest = threshold * threshold;
        int nearest = -1;
        {
            f3 d;
            float dd;
            if (nc > 0) { d = sub3(mk3(i2f(a0.x), i2f(a0.y), i2f(a0.z)), localA); dd = dot3(d, d); if (dd < shortest) { shortest = dd; nearest = 0; } }
            if (nc > 1) { d = sub3(mk3(i2f(a1.x), i2f(a1.y), i2f(a1.z)), localA); dd = dot3(d, d); if (dd < shortest) { shortest = dd; nearest = 1; } }
            if (nc > 2) { d = sub3(mk3(i2f(a2.x), i2f(a2.y), i2f(a2.z)), localA); dd = dot3(d, d); if (dd < shortest) { shortest = dd; nearest = 2; } }
            if (nc > 3) { d = sub3(mk3(i2f(a3.x), i2f(a3.y), i2f(a3.z)), localA); dd = dot3(d, d); if (dd < shortest) { shortest = dd; nearest = 3; } }
        }
        int life = 0, src = -1, idx;
        if (nearest >= 0) {  // replaceContactPoint keeps the lifetime (:280-305); src_slot = the slot it continues
            idx = nearest;
            life = reinterpret_cast<const int4*>(P + idx)[4].z;
            src = idx;
        } else {
            idx = nc;
            if (idx == 4) {  // sortCachedPoints (:83-156) + closestAxis4 (lm/VectorUtil.java:60-90)
                const float d0 = i2f(reinterpret_cast<const int4*>(P)[3].w), d1 = i2f(reinterpret_cast<const int4*>(P + 1)[3].w);
                const float d2 = i2f(reinterpret_cast<const int4*>(P + 2)[3].w), d3 = i2f(reinterpret_cast<const int4*>(P + 3)[3].w);
                int maxPenetrationIndex = -1;
                float maxPenetration = depth;
                if (d0 < maxPenetration) { maxPenetrationIndex = 0; maxPenetration = d0; }
                if (d1 < maxPenetration) { maxPenetrationIndex = 1; maxPenetration = d1; }
                if (d2 < maxPenetration) { maxPenetrationIndex = 2; maxPenetration = d2; }
                if (d3 < maxPenetration) { maxPenetrationIndex = 3; maxPenetration = d3; }
                const f3 p0 = mk3(i2f(a0.x), i2f(a0.y), i2f(a0.z)), p1 = mk3(i2f(a1.x), i2f(a1.y), i2f(a1.z));
                const f3 p2 = mk3(i2f(a2.x), i2f(a2.y), i2f(a2.z)), p3 = mk3(i2f(a3.x), i2f(a3.y), i2f(a3.z));
                float res0 = 0.f, res1 = 0.f, res2 = 0.f, res3 = 0.f;
                if (maxPenetrationIndex != 0) res0 = len2_3(crs3(sub3(localA, p1), sub3(p3, p2)));
                if (maxPenetrationIndex != 1) res1 = len2_3(crs3(sub3(localA, p0), sub3(p3, p2)));
                if (maxPenetrationIndex != 2) res2 = len2_3(crs3(sub3(localA, p0), sub3(p3, p1)));
                if (maxPenetrationIndex != 3) res3 = len2_3(crs3(sub3(localA, p0), sub3(p2, p1)));
                res0 = fabsf(res0); res1 = fabsf(res1); res2 = fabsf(res2); res3 = fabsf(res3);
                int maxIndex = -1;
                float maxVal = -1e30f;
                if (res0 > maxVal) { maxIndex = 0; maxVal = res0; }
                if (res1 > maxVal) { maxIndex = 1; maxVal = res1; }
                if (res2 > maxVal) { maxIndex = 2; maxVal = res2; }
                if (res3 > maxVal) { maxIndex = 3; maxVal = res3; }
                idx = maxIndex < 0 ? 0 : maxIndex;
            } else {
                nc = nc + 1;
            }
            newSlot = idx;
        }
        PtV n;
        n.q0 = make_int4(f2i(localA.x), f2i(localA.y), f2i(localA.z), f2i(localB.x));
        n.q1 = make_int4(f2i(localB.y), f2i(localB.z), f2i(pointA.x), f2i(pointA.y));
        n.q2 = make_int4(f2i(pointA.z), f2i(point.x), f2i(point.y), f2i(point.z));
        n.q3 = make_int4(f2i(normal.x), f2i(normal.y), f2i(normal.z), f2i(depth));
        n.q4 = make_int4(f2i(friction), f2i(restitution), life, src);
        n.q5 = make_int4(0, 0, 0, 0);
        stPtV(P + idx, n);
    }
    // refreshContactPoints(trA, trB) (np/PersistentManifold.java:312-372) incl. removeContactPoint (:259-278); the
    // reference's two loops fused per point: point i is updated, then tested; everything above i is already final
    if (nc > 0) {
        const Xf& trA = isSwapped ? rootB : rootA;
        const Xf& trB = isSwapped ? rootA : rootB;
        const float thr2 = threshold * threshold;
        for (int i = nc - 1; i >= 0; i--) {
            PtV v = ldPtV(P + i);
            const f3 la = mk3(i2f(v.q0.x), i2f(v.q0.y), i2f(v.q0.z)), lb = mk3(i2f(v.q0.w), i2f(v.q1.x), i2f(v.q1.y));
            const f3 nrm = mk3(i2f(v.q3.x), i2f(v.q3.y), i2f(v.q3.z));
            const f3 wa = xfPoint(trA, la);
            const f3 wb = xfPoint(trB, lb);
            const float dist = dot3(sub3(wa, wb), nrm);
            bool remove = false;
            if (!(dist <= threshold)) {
                remove = true;
            } else {
                f3 projected = sub3(wa, scl3(nrm, dist));
                f3 diff = sub3(wb, projected);
                if (dot3(diff, diff) > thr2) remove = true;
            }
            if (!remove) {
                v.q1.z = f2i(wa.x); v.q1.w = f2i(wa.y); v.q2.x = f2i(wa.z);
                v.q2.y = f2i(wb.x); v.q2.z = f2i(wb.y); v.q2.w = f2i(wb.z);
                v.q3.w = f2i(dist);
                v.q4.z = v.q4.z + 1;                       // lifeTime++
                v.q4.w = (i == newSlot) ? -1 : i;          // src_slot: the slot this point held at the start of the step
                stPtV(P + i, v);
            } else {
                const int last = nc - 1;
                if (i != last) {
                    // p[i] = p[last] (already refreshed and kept); p[last].lifeTime = 0, src_slot = -1
                    PtV l = ldPtV(P + last);
                    stPtV(P + i, l);
                    l.q4.z = 0; l.q4.w = -1;
                    stPtV(P + last, l);
                } else {
                    // the reference leaves the refreshed values in the dead slot; keep the record identical
                    v.q1.z = f2i(wa.x); v.q1.w = f2i(wa.y); v.q2.x = f2i(wa.z);
                    v.q2.y = f2i(wb.x); v.q2.z = f2i(wb.y); v.q2.w = f2i(wb.z);
                    v.q3.w = f2i(dist);
                    v.q4.z = v.q4.z + 1;
                    v.q4.w = (i == newSlot) ? -1 : i;
                    stPtV(P + i, v);
                }
                nc = last;
            }
        }
    }
    hdr1.x = nc;
    int4* hd = reinterpret_cast<int4*>(H);
    hd[0] = hdr0;
    hd[1] = hdr1;
    return added;
}

// ---- k_carry: bring manifolds over from the previous step by pair key --------------------------------
// A pair that stayed in the cache keeps its algorithm and manifold (bp/BroadphasePair.java:37-40); a pair
// that left and came back starts empty (bp/HashedOverlappingPairCache.java:129-174 cleanOverlappingPair).
// prevFirst[uid0] is the first previous pair whose uid0 is >= the given one, so the search is confined to the
// handful of pairs that share uid0.
__global__ void __launch_bounds__(256)
k_carry(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ numPairs, const uint64_t* __restrict__ prevKeys,
        const uint32_t* __restrict__ prevNum, const uint32_t* __restrict__ prevFirst, const ManifoldHdr* __restrict__ prevH,
        const b2c_manifold_point* __restrict__ prevP, ManifoldHdr* __restrict__ H, b2c_manifold_point* __restrict__ P, int uidBits,
        StepCounters* ctr, uint8_t* __restrict__ hist, const uint32_t* __restrict__ curFirst) {
    const uint32_t n = *numPairs, pn = *prevNum;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        uint32_t p = base + lane;
        int found = -1, nc = 0;
        bool carriedManifold = false;
        if (p < n) {
            uint64_t k = keys[p];
            uint32_t uid0 = (uint32_t)(k >> uidBits);
            uint32_t a = pn ? prevFirst[uid0] : 0, b = pn ? prevFirst[uid0 + 1] : 0;
            // most rows are unchanged from one step to the next: the pair usually sits at the same rank in last step's row,
            // so one probe replaces the dependent loads of the binary search (which stays as the fallback)
            const uint32_t guess = a + (p - curFirst[uid0]);
            if (guess < b && prevKeys[guess] == k) {
                found = (int)guess;
            } else {
                while (a < b) {
                    uint32_t mid = (a + b) >> 1;
                    if (prevKeys[mid] < k) a = mid + 1; else b = mid;
                }
                if (a < pn && prevKeys[a] == k) found = (int)a;
            }
            int4 h0, h1;
            if (found >= 0) {
                const int4* src = reinterpret_cast<const int4*>(prevH + found);
                h0 = src[0]; h1 = src[1];
                nc = h1.x;
                // algorithm field: the pair already owns a manifold.  5 = compound pair: its child manifolds live in the
                // compound item arrays and are counted by k_compound_expand (header word pad0 = "counted this step")
                carriedManifold = h1.y != 0 && h1.y != 5;
                if (h1.y == 5) h1.z = 0;
            } else {
                h0 = make_int4((int)uid0, (int)(k & ((1ull << uidBits) - 1ull)), 0, 0);
                h1 = make_int4(0, 0, 0, 0);
            }
            int4* dst = reinterpret_cast<int4*>(H + p);   // adjacent lanes -> adjacent 32-byte headers
            dst[0] = h0; dst[1] = h1;
            {   // header word pad0 = last step's GJK iteration count | HDR_EPA_BIG
                const int trips = h1.y == 5 ? 0 : (h1.z & 0xff);
                hist[p] = (uint8_t)((trips > 15 ? 15 : trips) | ((h1.y != 5 && (h1.z & 0x100)) ? 0x80 : 0));
            }
        }
        uint32_t cm = __ballot_sync(0xffffffffu, carriedManifold);
        if (lane == 0 && cm) atomicAdd(&ctr->numManifolds, (uint32_t)__popc(cm));
        // live points: every lane copies its own manifold's points (96-byte records = 6 x int4, all loads of a point
        // in flight before its stores), so the lanes of a warp work in parallel instead of taking turns
        if (nc > 0) {
            const int4* src = reinterpret_cast<const int4*>(prevP + 4 * (size_t)found);
            int4* dst = reinterpret_cast<int4*>(P + 4 * (size_t)p);
            for (int q = 0; q < nc; q++) {
                int4 v0 = src[6 * q], v1 = src[6 * q + 1], v2 = src[6 * q + 2], v3 = src[6 * q + 3], v4 = src[6 * q + 4],
                     v5 = src[6 * q + 5];
                dst[6 * q] = v0; dst[6 * q + 1] = v1; dst[6 * q + 2] = v2; dst[6 * q + 3] = v3; dst[6 * q + 4] = v4;
                dst[6 * q + 5] = v5;
            }
        }
    }
}

// ---- one world partitioned over several GPUs: manifolds follow pairs that change owner ------------------
// Each rank emits the pairs whose first member (in sorted-AABB order) lies in its contiguous range of the
// sorted proxy list, so a pair near a range boundary can change owner from one step to the next.  Its manifold
// must follow it: k_export_departed lists last step's manifolds whose pair this rank no longer owns; after an
// all-gather every rank adopts the ones whose pair it owns now (k_import_arrivals).
__device__ __forceinline__ int findPairIndex(const uint64_t* __restrict__ keys, uint32_t n, const uint32_t* __restrict__ first,
                                             uint64_t k, int uidBits) {
    if (n == 0) return -1;
    uint32_t uid0 = (uint32_t)(k >> uidBits);
    uint32_t a = first[uid0], b = first[uid0 + 1];
    while (a < b) {
        uint32_t mid = (a + b) >> 1;
        if (keys[mid] < k) a = mid + 1; else b = mid;
    }
    return (a < n && keys[a] == k) ? (int)a : -1;
}

__global__ void __launch_bounds__(256)
k_export_departed(const uint64_t* __restrict__ prevKeys, const uint32_t* __restrict__ prevNum, const ManifoldHdr* __restrict__ prevH,
                  const b2c_manifold_point* __restrict__ prevP, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ numPairs,
                  const uint32_t* __restrict__ first, int uidBits, uint64_t* __restrict__ outKeys, ManifoldHdr* __restrict__ outH,
                  b2c_manifold_point* __restrict__ outP, uint32_t cap, uint32_t* __restrict__ outCount) {
    const uint32_t pn = *prevNum, n = *numPairs;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < pn; p += gridDim.x * blockDim.x) {
        ManifoldHdr h = prevH[p];
        // an empty manifold carries no state its next owner would not recreate identically (body order follows from the
        // pair and the shape types), so only touching manifolds travel
        if (h.algorithm == 0 || h.num_contacts == 0) continue;
        uint64_t k = prevKeys[p];
        if (findPairIndex(keys, n, first, k, uidBits) >= 0) continue;  // still ours: k_carry took care of it
        uint32_t slot = atomicAdd(outCount, 1u);
        if (slot >= cap) continue;  // the host checks the count against cap
        outKeys[slot] = k;
        outH[slot] = h;
        for (int q = 0; q < h.num_contacts && q < 4; q++) outP[4 * (size_t)slot + q] = prevP[4 * (size_t)p + q];
    }
}

__global__ void __launch_bounds__(256)
k_import_arrivals(const uint64_t* __restrict__ inKeys, const ManifoldHdr* __restrict__ inH, const b2c_manifold_point* __restrict__ inP,
                  uint32_t count, const uint64_t* __restrict__ keys, const uint32_t* __restrict__ numPairs,
                  const uint32_t* __restrict__ first, int uidBits, ManifoldHdr* __restrict__ H, b2c_manifold_point* __restrict__ P,
                  StepCounters* ctr) {
    const uint32_t n = *numPairs;
    for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < count; s += gridDim.x * blockDim.x) {
        int idx = findPairIndex(keys, n, first, inKeys[s], uidBits);
        if (idx < 0) continue;  // another rank's pair (or a pair that really vanished)
        ManifoldHdr h = inH[s];
        if (H[idx].algorithm != 0) continue;  // our own departed copy cannot come back; defensive
        H[idx] = h;
        for (int q = 0; q < h.num_contacts && q < 4; q++) P[4 * (size_t)idx + q] = inP[4 * (size_t)s + q];
        atomicAdd(&ctr->numManifolds, 1u);
    }
}

// Slot variant of the import: `nslots` fixed-size slots (one per rank, as an all-gather leaves them), each
// { count, pad[3], keys[cap], headers[cap], points[4*cap] }; counts are read on the device.
__host__ __device__ __forceinline__ size_t mgpuSlotBytes(uint32_t cap) {
    return 16 + (size_t)cap * (sizeof(uint64_t) + sizeof(ManifoldHdr) + 4 * sizeof(b2c_manifold_point));
}
__global__ void __launch_bounds__(256)
k_import_arrival_slots(const unsigned char* __restrict__ slots, uint32_t nslots, uint32_t cap, const uint64_t* __restrict__ keys,
                       const uint32_t* __restrict__ numPairs, const uint32_t* __restrict__ first, int uidBits, ManifoldHdr* __restrict__ H,
                       b2c_manifold_point* __restrict__ P, StepCounters* ctr) {
    const uint32_t n = *numPairs;
    const size_t slotBytes = mgpuSlotBytes(cap);
    for (uint32_t r = blockIdx.y; r < nslots; r += gridDim.y) {
        const unsigned char* slot = slots + (size_t)r * slotBytes;
        uint32_t count = *reinterpret_cast<const uint32_t*>(slot);
        if (count > cap) { if (threadIdx.x == 0 && blockIdx.x == 0) ctr->migrateOverflow = 1; count = cap; }
        const uint64_t* inKeys = reinterpret_cast<const uint64_t*>(slot + 16);
        const ManifoldHdr* inH = reinterpret_cast<const ManifoldHdr*>(slot + 16 + (size_t)cap * 8);
        const b2c_manifold_point* inP = reinterpret_cast<const b2c_manifold_point*>(slot + 16 + (size_t)cap * (8 + sizeof(ManifoldHdr)));
        for (uint32_t s = blockIdx.x * blockDim.x + threadIdx.x; s < count; s += gridDim.x * blockDim.x) {
            int idx = findPairIndex(keys, n, first, inKeys[s], uidBits);
            if (idx < 0) continue;
            ManifoldHdr h = inH[s];
            if (H[idx].algorithm != 0) continue;
            H[idx] = h;
            for (int q = 0; q < h.num_contacts && q < 4; q++) P[4 * (size_t)idx + q] = inP[4 * (size_t)s + q];
            atomicAdd(&ctr->numManifolds, 1u);
        }
    }
}

// ---- k_classify: needsCollision + algorithm table --------------------------------------------------
__device__ __forceinline__ bool isConvexType(int t) { return t == SH_BOX || t == SH_SPHERE || t == SH_HULL; }

// compound x {box, sphere, hull, plane, mesh, compound}: children are convex, so every child algorithm is one of
// sphere-sphere, convex-plane, convex-convex, convex-concave.
__device__ __forceinline__ bool compoundPairSupported(int t0, int t1) {
    if (t0 != SH_COMPOUND && t1 != SH_COMPOUND) return false;
    const int other = t0 == SH_COMPOUND ? t1 : t0;
    return other == SH_COMPOUND || other == SH_PLANE || other == SH_MESH || isConvexType(other);
}

__global__ void __launch_bounds__(256) k_classify(NpArgs a) {
    const uint32_t n = *a.numPairs;
    __shared__ uint32_t hist[16];
    if (threadIdx.x < 16) hist[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        int2 pr = a.pairs[p];
        int b0 = pr.x - 1, b1 = pr.y - 1;
        uint8_t f0 = a.flags[b0], f1 = a.flags[b1];
        int bin = BIN_SKIP;
        // disp/CollisionDispatcher.java:198-223: skip when both are inactive
        bool dispatch = (f0 & BF_ACTIVE) || (f1 & BF_ACTIVE);
        if (dispatch && a.numNoCollide) {  // :216-218 checkCollideWith: constraint-linked bodies
            const uint64_t k = ((uint64_t)(uint32_t)pr.x << a.uidBits) | (uint32_t)pr.y;
            uint32_t lo = 0, hi = a.numNoCollide;
            while (lo < hi) {
                uint32_t mid = (lo + hi) >> 1;
                if (a.noCollide[mid] < k) lo = mid + 1; else hi = mid;
            }
            if (lo < a.numNoCollide && a.noCollide[lo] == k) dispatch = false;
        }
        if (dispatch) {
            int t0 = a.shapes[a.shape[b0]].type, t1 = a.shapes[a.shape[b1]].type;
            if (t0 == SH_SPHERE && t1 == SH_SPHERE) bin = BIN_SS;
            else if ((isConvexType(t0) && t1 == SH_PLANE) || (isConvexType(t1) && t0 == SH_PLANE)) bin = BIN_CP;
            else if (isConvexType(t0) && isConvexType(t1))
                bin = ((a.hist[p] & 0x7f) >= 2 ? BIN_PS0 : BIN_GJK0) + (t0 == SH_HULL ? 2 : 0) + (t1 == SH_HULL ? 1 : 0);
            else if ((isConvexType(t0) && t1 == SH_MESH) || (isConvexType(t1) && t0 == SH_MESH)) bin = BIN_MESH;
            else if (compoundPairSupported(t0, t1)) bin = BIN_COMPOUND;  // disp/DefaultCollisionConfiguration.java:198-204
        } else if (a.hasCompound) {
            int t0 = a.shapes[a.shape[b0]].type, t1 = a.shapes[a.shape[b1]].type;
            if (compoundPairSupported(t0, t1)) bin = BIN_COMPOUND_KEEP;
        }
        a.binOf[p] = (uint8_t)bin;  // BIN_SKIP pairs are not dispatched: their raw record is not written this step
        // per-block histogram: one shared-memory atomic per group of lanes with the same bin
        uint32_t m = __match_any_sync(__activemask(), bin);
        if ((threadIdx.x & 31) == __ffs(m) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(m));
    }
    __syncthreads();
    if (threadIdx.x < 16 && hist[threadIdx.x]) atomicAdd(&a.binZero[threadIdx.x], hist[threadIdx.x]);
}

// k_partition16: partition of an index list by a 4-bit key in ONE pass.  The producer of the keys has already counted them
// (hist[16]), so every bin's range is known up front and a tile only has to reserve its share of each bin: it ranks its items
// per bin (match-any inside a warp, exclusive scan over the warps) and takes `count` consecutive slots with one atomicAdd per
// bin.  Inside a bin the items of one tile stay together in list order; the tiles themselves land in the order their atomics
// arrive — nothing downstream depends on the order inside a bin (every consumer works per item), only on items that were
// neighbours in the list staying neighbours, which a 2048-item tile preserves.  (A stable version with decoupled look-back
// over the tiles cost 17 us per call at 0.9 M items: the look-back chain, not the data, was the time.)
constexpr int BIN_ITEMS = 8, BIN_TILE = 256 * BIN_ITEMS;
// keys[i] < 16 for i < *nPtr; zero = (hist[16] | cursor[16] | ...) cleared by the caller, hist filled by the producer of the
// keys; out[...] = payload ? payload[i] : i grouped by key; startOut[0..16] = exclusive offsets.
__global__ void __launch_bounds__(256)
k_partition16(const uint8_t* __restrict__ keys, const uint32_t* __restrict__ nPtr, uint32_t* zero, uint32_t* __restrict__ startOut,
              const uint32_t* __restrict__ payload, uint32_t* __restrict__ out) {
    const uint32_t n = *nPtr;
    const uint32_t numTiles = (n + BIN_TILE - 1) / BIN_TILE;
    const uint32_t* hist = zero;
    uint32_t* cursor = zero + 16;
    __shared__ uint32_t warpCnt[8][16];
    __shared__ uint32_t base[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = (1u << lane) - 1u;
    uint32_t myStart = 0;  // threads 0..15: exclusive offset of bin threadIdx.x
    if (threadIdx.x < 16)
        for (int j = 0; j < (int)threadIdx.x; j++) myStart += hist[j];
    if (blockIdx.x == 0 && threadIdx.x <= 16) {
        uint32_t e = 0;
        for (int j = 0; j < (int)threadIdx.x; j++) e += hist[j];
        startOut[threadIdx.x] = e;
    }
    for (uint32_t tile = blockIdx.x; tile < numTiles; tile += gridDim.x) {
        __syncthreads();
        if (threadIdx.x < 128) (&warpCnt[0][0])[threadIdx.x] = 0;
        __syncthreads();
        const uint32_t tileStart = tile * BIN_TILE + warp * (BIN_ITEMS * 32) + lane;
        uint32_t bin[BIN_ITEMS], rank[BIN_ITEMS];
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            uint32_t idx = tileStart + k * 32;
            bin[k] = idx < n ? (uint32_t)keys[idx] : 0x100u + lane;
        }
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            const bool valid = bin[k] < 16u;
            uint32_t m = __match_any_sync(0xffffffffu, bin[k]);
            uint32_t lower = __popc(m & ltmask);
            uint32_t pre = 0;
            if (valid && lower == 0) {
                pre = warpCnt[warp][bin[k]];
                warpCnt[warp][bin[k]] = pre + __popc(m);
            }
            __syncwarp();
            pre = __shfl_sync(0xffffffffu, pre, __ffs(m) - 1);
            rank[k] = pre + lower;
        }
        __syncthreads();
        if (threadIdx.x < 16) {
            const int d = threadIdx.x;
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) {
                uint32_t c = warpCnt[w][d];
                warpCnt[w][d] = run;
                run += c;
            }
            base[d] = myStart + (run ? atomicAdd(cursor + d, run) : 0u);
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BIN_ITEMS; k++) {
            if (bin[k] < 16u) {
                const uint32_t idx = tileStart + k * 32;
                out[base[bin[k]] + warpCnt[warp][bin[k]] + rank[k]] = payload ? payload[idx] : idx;
            }
        }
    }
}
__device__ __forceinline__ uint32_t binItem(const NpArgs& a, uint32_t it) { return a.binItems[it]; }

__device__ __forceinline__ void writeRaw(b2c_raw_contact* r, int2 pr, int tri, int has, f3 n, f3 pt, float depth, int method,
                                         int iters) {
    r->uid0 = pr.x; r->uid1 = pr.y; r->tri = tri; r->has_contact = has;
    r->normal[0] = n.x; r->normal[1] = n.y; r->normal[2] = n.z;
    r->point[0] = pt.x; r->point[1] = pt.y; r->point[2] = pt.z;
    r->depth = depth; r->method = method; r->iters = iters; r->pad[0] = 0;
}

// ---- sphere-sphere (disp/SphereSphereCollisionAlgorithm.java:73-134) ---------------------------------
__global__ void __launch_bounds__(256) k_sphere_sphere(NpArgs a) {
    const uint32_t s = a.binStart[BIN_SS], e = a.binStart[BIN_SS + 1];
    uint32_t added = 0, created = 0;
    for (uint32_t it = s + blockIdx.x * blockDim.x + threadIdx.x; it < e; it += gridDim.x * blockDim.x) {
        uint32_t p = binItem(a, it);
        int2 pr = a.pairs[p];
        int b0 = pr.x - 1, b1 = pr.y - 1;
        Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
        float r0 = a.shapes[a.shape[b0]].dims[0], r1 = a.shapes[a.shape[b1]].dims[0];
        ManifoldHdr* H = a.mhdr + p;
        int4 h0 = reinterpret_cast<const int4*>(H)[0], h1 = reinterpret_cast<const int4*>(H)[1];
        bool fresh = false;
        if (h1.y == 0) { h1.y = 1; h0.z = pr.x; h0.w = pr.y; created++; fresh = true; }
        f3 diff = sub3(t0.o, t1.o);
        float len = len3(diff);
        if (len > (r0 + r1)) {
            if (a.wantRaw) writeRaw(a.raw + p, pr, -1, 0, mk3(0, 0, 0), mk3(0, 0, 0), 0.f, 10, 0);
            if (h1.x != 0) manifoldStepV(H, a.mpts + 4 * (size_t)p, h0, h1, pr.x, t0, t1, false, mk3(0, 0, 0), mk3(0, 0, 0), 0.f, a.threshold, 0.f, 0.f);
            else if (fresh) { reinterpret_cast<int4*>(H)[0] = h0; reinterpret_cast<int4*>(H)[1] = h1; }
            continue;
        }
        float dist = len - (r0 + r1);
        f3 n = mk3(1.f, 0.f, 0.f);
        if (len > B2C_FLT_EPSILON) n = scl3(diff, 1.f / len);
        f3 pos1 = add3(t1.o, scl3(n, r1));
        if (a.wantRaw) writeRaw(a.raw + p, pr, -1, 1, n, pos1, dist, 10, 0);   // the manifold is updated right here
        float2 m0 = a.material[b0], m1 = a.material[b1];
        if (manifoldStepV(H, a.mpts + 4 * (size_t)p, h0, h1, pr.x, t0, t1, true, n, pos1, dist, a.threshold, combinedFriction(m0.x, m1.x),
                          m0.y * m1.y))
            added++;
    }
    if (created) atomicAdd(&a.ctr->numManifolds, created);
    if (added) atomicAdd(&a.ctr->contactsAdded, added);
}

// ---- convex-plane (disp/ConvexPlaneCollisionAlgorithm.java:75-136) -----------------------------------
__device__ __forceinline__ AnyS makeAnyS(const ShapeDev& s, const float4* hullPts) {
    AnyS r;
    r.type = s.type;
    r.h = mk3(s.dims[0], s.dims[1], s.dims[2]);
    r.ta = r.tb = r.tc = mk3(0.f, 0.f, 0.f);
    r.pts = hullPts + s.pointOffset;
    r.n = s.numPoints;
    r.margin = s.margin;
    return r;
}

__global__ void __launch_bounds__(256) k_convex_plane(NpArgs a) {
    const uint32_t s = a.binStart[BIN_CP], e = a.binStart[BIN_CP + 1];
    uint32_t added = 0, created = 0;
    for (uint32_t it = s + blockIdx.x * blockDim.x + threadIdx.x; it < e; it += gridDim.x * blockDim.x) {
        uint32_t p = binItem(a, it);
        int2 pr = a.pairs[p];
        int b0 = pr.x - 1, b1 = pr.y - 1;
        ShapeDev s0 = a.shapes[a.shape[b0]], s1 = a.shapes[a.shape[b1]];
        bool swapped = (s0.type == SH_PLANE);  // planeConvexCF: convex is body1
        int bc = swapped ? b1 : b0, bp = swapped ? b0 : b1;
        const ShapeDev& cs = swapped ? s1 : s0;
        const ShapeDev& ps = swapped ? s0 : s1;
        Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
        const Xf& tc = swapped ? t1 : t0;
        const Xf& tp = swapped ? t0 : t1;
        MView m = mview(a, p);
        if (m.h->algorithm == 0) { m.h->algorithm = 2; m.h->body0 = bc + 1; m.h->body1 = bp + 1; created++; }
        for (int k = 0; k < m.h->num_contacts; k++) m.p[k].src_slot = k;
        f3 planeNormal = mk3(ps.plane[0], ps.plane[1], ps.plane[2]);
        float planeConstant = ps.plane[3];
        Xf planeInConvex = invMul(tc, tp);
        Xf convexInPlane = invMul(tp, tc);
        f3 dir = mulMV(planeInConvex.m, neg3(planeNormal));
        AnyS shp = makeAnyS(cs, a.hullPts);
        f3 vtx = shp.supportMargin(dir);
        f3 vtxInPlane = xfPoint(convexInPlane, vtx);
        float distance = dot3(planeNormal, vtxInPlane) - planeConstant;
        f3 projected = sub3(vtxInPlane, scl3(planeNormal, distance));
        f3 world = xfPoint(tp, projected);
        bool has = distance < a.threshold;
        f3 nW = mulMV(tp.m, planeNormal);
        if (a.wantRaw) writeRaw(a.raw + p, pr, -1, has ? 1 : 0, nW, world, distance, 11, 0);
        if (has) {
            float2 m0 = a.material[b0], m1 = a.material[b1];
            if (manifoldAdd(m, pr.x, t0, t1, nW, world, distance, a.threshold, combinedFriction(m0.x, m1.x), m0.y * m1.y, 0, 0)) added++;
        }
        if (m.h->num_contacts != 0) resultRefresh(m, pr.x, t0, t1, a.threshold);
    }
    if (created) atomicAdd(&a.ctr->numManifolds, created);
    if (added) atomicAdd(&a.ctr->contactsAdded, added);
}

// ---- CompoundShape work items (the kernels are in compound.cuh; the penetration bin below also serves them) ----------
struct CompoundCounters {
    uint32_t itemCount;   // items reserved by k_compound_expand (may exceed the capacity -> overflow)
    uint32_t overflow;
    uint32_t numItems;    // items the later kernels and the result calls may touch (0 after an overflow)
    uint32_t pad;
};

constexpr uint32_t CITEM_KEEP = 0x80000000u;  // itemCode flag: copy the manifold forward, do nothing else

struct CompoundArgs {
    const CompoundChildDev* children;
    CompoundCounters* cc;          // cleared per dispatch
    uint32_t* itemPair;            // [maxItems] pair index
    uint32_t* itemCode;            // [maxItems] k = i * n1 + j (index of the item inside its pair) | CITEM_KEEP
    int* itemPrev;                 // [maxItems] the same item in the previous dispatch's arrays, or -1
    b2c_raw_contact* raw;          // [maxItems] detector output (has_contact -3: child x mesh, records are in the mesh item array)
    uint32_t* meshStart;           // [maxItems] child x mesh items: first (child, triangle) record in GjkArgs.rawMesh
    uint32_t* meshCount;           // [maxItems]
    void* bigScratch;              // EpaScratch per thread of k_compound_mesh (global memory), or null
    uint32_t numBigScratch;
    ManifoldHdr* H;                // [maxItems] child manifolds of this dispatch (header word pad0 / pad1 = child index in
    b2c_manifold_point* P;         //            body0's / body1's compound shape, -1 = that object is not a compound)
    const ManifoldHdr* prevH;      // the previous dispatch's
    const b2c_manifold_point* prevP;
    uint32_t maxItems;
};

struct CompoundItem {
    int2 pr;            // the broadphase pair (uid0 < uid1)
    int bodyA, bodyB;   // 0-based body indices of the child algorithm's body0 / body1
    int shapeA, shapeB; // shape table indices (A: a compound's child; B: the other object's shape or its child)
    int childA, childB; // child index inside its compound, -1 = not a compound
    Xf tA, tB;          // world transforms the child algorithm sees
};

// Which child algorithm item k of pair p is, with its shapes and transforms.
__device__ __forceinline__ void decodeCompoundItem(const NpArgs& a, const CompoundArgs& c, uint32_t p, uint32_t k, CompoundItem& it) {
    it.pr = a.pairs[p];
    const int b0 = it.pr.x - 1, b1 = it.pr.y - 1;
    const int s0 = a.shape[b0], s1 = a.shape[b1];
    const ShapeDev& S0 = a.shapes[s0];
    const ShapeDev& S1 = a.shapes[s1];
    const bool c0 = S0.type == SH_COMPOUND, c1 = S1.type == SH_COMPOUND;
    const Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
    auto childXf = [&](const ShapeDev& S, int i, const Xf& org, int& shapeOut) {
        const CompoundChildDev& ch = c.children[S.pointOffset + i];
        shapeOut = ch.shape;
        return compoundChildWorld(org, c.children, ch);  // newChildWorldTrans.mul(orgTrans, childTrans) (disp/CompoundCollisionAlgorithm.java:107), per nesting level
    };
    if (c0 && c1) {
        const int n1 = S1.numPoints;
        const int i = (int)k / n1, j = (int)k % n1;
        it.bodyA = b1; it.childA = j; it.tA = childXf(S1, j, t1, it.shapeA);
        it.bodyB = b0; it.childB = i; it.tB = childXf(S0, i, t0, it.shapeB);
    } else if (c0) {
        it.bodyA = b0; it.childA = (int)k; it.tA = childXf(S0, (int)k, t0, it.shapeA);
        it.bodyB = b1; it.childB = -1; it.tB = t1; it.shapeB = s1;
    } else {
        it.bodyA = b1; it.childA = (int)k; it.tA = childXf(S1, (int)k, t1, it.shapeA);
        it.bodyB = b0; it.childB = -1; it.tB = t0; it.shapeB = s0;
    }
}

__device__ __forceinline__ uint32_t compoundItemCount(const NpArgs& a, int2 pr) {
    const ShapeDev& S0 = a.shapes[a.shape[pr.x - 1]];
    const ShapeDev& S1 = a.shapes[a.shape[pr.y - 1]];
    const uint32_t n0 = S0.type == SH_COMPOUND ? (uint32_t)S0.numPoints : 1u;
    const uint32_t n1 = S1.type == SH_COMPOUND ? (uint32_t)S1.numPoints : 1u;
    return n0 * n1;
}

// ---- GJK bins + EPA --------------------------------------------------------------------------------
struct EpaItem {        // a pair (or pair/triangle) whose detector asked for the penetration solver
    uint32_t pair;      // pair index
    int meshItem;       // index into the mesh item arrays, or -1; <= -2: compound child work item -2 - meshItem
    GjkResult g;
};

// Temporal coherence of the penetration bin: a pair whose item overflowed the small EPA pools keeps doing so while it stays
// in deep contact.  The pair remembers it (bit 8 of its manifold header word pad0; k_carry hands it to hist[] bit 7), and its
// items go straight to the large-pool tier, which then runs BESIDE the small-pool tier instead of after it.
constexpr int HDR_EPA_BIG = 0x100;      // ManifoldHdr.pad0 (bits 0-7: last step's GJK trip count)
constexpr uint8_t HIST_EPA_BIG = 0x80;  // NpArgs.hist[] (bits 0-6: trip count clamped to 15)
struct GjkArgs {
    EpaItem* epaItems;
    uint32_t maxEpa;
    uint32_t* epaRetry;      // items whose small pool overflowed
    uint32_t maxEpaRetry;
    uint32_t* epaBig;        // items predicted to need the large pools (capacity maxEpaRetry)
    // mesh work items
    uint32_t* meshPair;      // [maxMeshItems] pair index
    int* meshTri;            // [maxMeshItems] triangle index
    b2c_raw_contact* rawMesh;  // [maxMeshItems]
    uint32_t* meshStart;     // [maxPairs] first item of a mesh pair (indexed by pair)
    uint32_t* meshCount;     // [maxPairs]
    uint32_t maxMeshItems;
    CompoundArgs comp;       // compound child work items (EpaItem.meshItem <= -2 names item -2 - meshItem)
};

// The pair word of a new penetration item: an item of a pair that overflowed the small pools before also goes on the
// large-pool list and carries EPA_RETRY_BIT, so the small-pool tier leaves it alone.
__device__ __forceinline__ uint32_t epaRoute(const NpArgs& a, const GjkArgs& g, uint32_t p, uint32_t slot, bool big);

// A lane's convex shape, chosen at run time inside ONE code path (the whole bin shares the loop body; only
// the support mapping switches), so the instruction footprint stays small and every warp runs the same loop.
struct LaneShape {
    int type;
    f3 h;                 // box: implicitShapeDimensions
    const float4* pts;    // hull
    int n;
    __device__ __forceinline__ void load(const ShapeDev& s, const float4* hullPts) {
        type = s.type;
        h = mk3(s.dims[0], s.dims[1], s.dims[2]);
        pts = hullPts + s.pointOffset;
        n = s.numPoints;
    }
    __device__ __forceinline__ f3 support(f3 v) const {  // localGetSupportingVertexWithoutMargin
        if (type == SH_BOX) return mk3(fsel(v.x, h.x, -h.x), fsel(v.y, h.y, -h.y), fsel(v.z, h.z, -h.z));
        if (type == SH_HULL) { HullS hs; hs.pts = pts; hs.n = n; hs.margin = 0.f; return hs.support(v); }
        return mk3(0.f, 0.f, 0.f);  // sphere
    }
    __device__ __forceinline__ f3 supportWide(f3 v) const {  // same values; hull vertices fetched four at a time
        if (type == SH_BOX) return mk3(fsel(v.x, h.x, -h.x), fsel(v.y, h.y, -h.y), fsel(v.z, h.z, -h.z));
        if (type == SH_HULL) { HullS hs; hs.pts = pts; hs.n = n; hs.margin = 0.f; return hs.supportT<true>(v); }
        return mk3(0.f, 0.f, 0.f);
    }
};

// Warp-level refill: idle lanes take the next items of [0, end).  The warp reserves CHUNK items at a time from
// the global cursor (one same-address atomic per 32 items instead of one per round) and hands them out locally.
struct WarpQueue {
    uint32_t next, limit;  // warp-uniform: the warp's reserved range [next, limit)
    __device__ __forceinline__ void init() { next = limit = 0; }
    __device__ __forceinline__ uint32_t take(bool want, uint32_t* cursor, uint32_t end) {
        constexpr uint32_t CHUNK = 32;
        const int lane = threadIdx.x & 31;
        uint32_t m = __ballot_sync(0xffffffffu, want);
        uint32_t idx = 0xffffffffu;
        if (m == 0) return idx;
        uint32_t need = (uint32_t)__popc(m);
        if (limit - next < need && limit != 0xffffffffu) {
            // hand out what is left of the old chunk first, then reserve a new one
            uint32_t have = limit - next;
            uint32_t rank = __popc(m & ((1u << lane) - 1u));
            if (want && rank < have) idx = next + rank;
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(cursor, CHUNK);
            base = __shfl_sync(0xffffffffu, base, 0);
            if (want && rank >= have) idx = base + (rank - have);
            next = base + (need - have);
            limit = base + CHUNK;
            if (base >= end) limit = 0xffffffffu, next = 0xffffffffu;  // list exhausted: stop reserving
        } else if (limit != 0xffffffffu) {
            uint32_t rank = __popc(m & ((1u << lane) - 1u));
            if (want) idx = next + rank;
            next += need;
        }
        if (idx != 0xffffffffu && idx >= end) idx = 0xffffffffu;
        return idx;
    }
};

// k_gjk_prefilter: the first two trips of the GJK loop as straight-line code, one thread per pair, no
// divergence.  About two thirds of the AABB-overlapping pairs end in trip 2 through the separating-axis early
// out (np/GjkPairDetector.java:154: delta > 0 && delta^2 > squaredDistance * maximumDistanceSquared); for
// those the detector result is final here (no contact, lastUsedMethod -1, curIter 1).  Every other pair —
// including anything unusual in trip 1 — goes to the survivor list and is run from scratch by k_gjk, so this
// kernel only has to be exact when it says "done".
#ifndef PREF_MINB
#define PREF_MINB 3
#endif
__global__ void __launch_bounds__(256, PREF_MINB)
k_gjk_prefilter(NpArgs a, uint32_t* __restrict__ survivors, uint32_t* __restrict__ survCount, uint8_t* __restrict__ survKey,
                uint32_t* survHist /*[16], zeroed*/) {
    // [s0, mid): pairs to examine; [mid, e0): pairs predicted to survive (~95 % right; k_gjk runs survivors from scratch, so
    // a wrong prediction just ends there in trip 2) — appended to the survivor list with their history key, no arithmetic
    const uint32_t s0 = a.binStart[BIN_GJK0], mid = a.binStart[BIN_PS0], e0 = a.binStart[BIN_COUNT];
    __shared__ uint32_t warpCnt[8];
    __shared__ uint32_t blockBase;
    __shared__ uint32_t sHist[16];
    if (threadIdx.x < 16) sHist[threadIdx.x] = 0;
    __syncthreads();
    uint32_t checks = 0;
    for (uint32_t base = s0 + blockIdx.x * blockDim.x; base < e0; base += gridDim.x * blockDim.x) {
        const uint32_t it = base + threadIdx.x;
        bool survive = false;
        uint32_t p = 0;
        uint32_t key = 15u;  // examined pairs that survive have no usable history
        if (it >= mid && it < e0) {
            p = binItem(a, it);
            survive = true;
            const int last = a.hist[p] & 0x7f;
            key = last >= 15 ? 0u : (uint32_t)(15 - last);  // longest first
        } else if (it < mid) {
            p = binItem(a, it);
            int2 pr = a.pairs[p];
            const int b0 = pr.x - 1, b1 = pr.y - 1;
            const ShapeDev& sa = a.shapes[a.shape[b0]];
            const ShapeDev& sb = a.shapes[a.shape[b1]];
            LaneShape A, B;
            A.load(sa, a.hullPts);
            B.load(sb, a.hullPts);
            Xf ta = loadXf(a.xf4, b0), tb = loadXf(a.xf4, b1);
            float maxd = sa.margin + sb.margin + a.threshold;
            const float maxDistSq = maxd * maxd;
            f3 positionOffset = scl3(add3(ta.o, tb.o), 0.5f);
            f3 laO = sub3(ta.o, positionOffset), lbO = sub3(tb.o, positionOffset);
            // trip 1 (axis (0,1,0), empty simplex)
            f3 axis = mk3(0.f, 1.f, 0.f);
            f3 pW = add3(mulMV(ta.m, A.supportWide(mulMtV(ta.m, neg3(axis)))), laO);
            f3 qW = add3(mulMV(tb.m, B.supportWide(mulMtV(tb.m, axis))), lbO);
            f3 w = sub3(pW, qW);
            float delta = dot3(axis, w);
            float sq = B2C_SIMD_INFINITY;
            bool normal1 = !((delta > 0.f) && (delta * delta > sq * maxDistSq));
            normal1 = normal1 && !eq3bits(w, mk3(1e30f, 1e30f, 1e30f));          // inSimplex against lastW
            normal1 = normal1 && !((sq - delta) <= (sq * GJK_REL_ERROR2));        // f0 <= f1
            axis = w;                                                             // one-vertex simplex: v = p - q
            float sq1 = len2_3(axis);
            normal1 = normal1 && !(sq1 < GJK_REL_ERROR2);
            normal1 = normal1 && !((sq - sq1) <= B2C_FLT_EPSILON * sq);
            // trip 2
            pW = add3(mulMV(ta.m, A.supportWide(mulMtV(ta.m, neg3(axis)))), laO);
            qW = add3(mulMV(tb.m, B.supportWide(mulMtV(tb.m, axis))), lbO);
            w = sub3(pW, qW);
            delta = dot3(axis, w);
            bool done = normal1 && (delta > 0.f) && (delta * delta > sq1 * maxDistSq);
            if (done) {
                if (a.wantRaw) writeRaw(a.raw + p, pr, -1, 0, mk3(0, 0, 0), mk3(0, 0, 0), 0.f, -1, 1);
                a.rawFlag[p] = 0;   // all k_manifold_cc needs of a pair without a contact
                checks++;
            }
            survive = !done;
        }
        // block-aggregated append: one global atomic per block round
        uint32_t m = __ballot_sync(0xffffffffu, survive);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) warpCnt[warp] = __popc(m);
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int k = 0; k < 8; k++) { uint32_t c = warpCnt[k]; warpCnt[k] = tot; tot += c; }
            blockBase = tot ? atomicAdd(survCount, tot) : 0;
        }
        __syncthreads();
        if (survive) {
            const uint32_t slot = blockBase + warpCnt[warp] + __popc(m & ((1u << lane) - 1u));
            survivors[slot] = p;
            // Temporal coherence: the iteration count a pair needed last step (kept in its manifold header by k_gjk) is the
            // count it needs now for ~95 % of the pairs.  Survivors are ordered by it, longest first, so the lanes of a
            // warp start, iterate and finish together instead of each being in a different phase of the detector.
            survKey[slot] = (uint8_t)key;
            const uint32_t same = __match_any_sync(__activemask(), key);  // one shared-memory atomic per key per warp
            if (lane == __ffs(same) - 1) atomicAdd(&sHist[key], (uint32_t)__popc(same));
        }
        __syncthreads();
    }
    if (threadIdx.x < 16 && sHist[threadIdx.x]) atomicAdd(&survHist[threadIdx.x], sHist[threadIdx.x]);
    if (checks) atomicAdd(&a.ctr->gjkChecks, checks);
}

// k_gjk: convex-convex detector for the 8 type pairs of {box, sphere, hull}^2 minus sphere-sphere.
// Persistent warps; each lane owns one pair at a time and all lanes step through GjkLane::iterate together.
// Output is the raw detector record (or an EPA work item); manifolds are updated by k_manifold_cc.
#ifndef GJK_MINB
#define GJK_MINB 4
#endif
__global__ void __launch_bounds__(128, GJK_MINB)
k_gjk(NpArgs a, GjkArgs g, uint32_t* cursor, const uint32_t* __restrict__ survivors, const uint32_t* __restrict__ survCount,
      const uint32_t* __restrict__ survStart /*[17] offsets of the 16 history buckets; bucket 15 = no history*/) {
    const uint32_t count = *survCount;
    const uint32_t unknownStart = survStart[15];
    uint32_t deep = 0, checks = 0;
    GjkLane L;
    LaneShape A, B;
    Xf ta, tb;
    uint32_t p = 0;
    int2 pr = make_int2(0, 0);
    bool busy = false, more = true;
    WarpQueue wq;
    wq.init();
    while (true) {
                // Pairs with a known history come in runs of equal predicted iteration count: the warp takes 32 of them at a time
        // and refills only when all of its lanes are done, so the lanes stay in phase (same trip, mostly the same simplex
        // size).  Pairs without history (the tail of the list) have unrelated iteration counts: there every lane refills
        // as soon as it is free.
        const bool inKnownPart = wq.next < unknownStart;
        const bool anyBusy = __any_sync(0xffffffffu, busy);  // every lane votes (never behind a short-circuit)
        const bool want = !busy && more && (!inKnownPart || !anyBusy);
        const uint32_t it = wq.take(want, cursor, count);  // one convergent call site for the whole warp
        if (want) {
            if (it == 0xffffffffu) {
                more = false;
            } else {
                p = survivors[it];
                pr = a.pairs[p];
                const int b0 = pr.x - 1, b1 = pr.y - 1;
                const ShapeDev& sa = a.shapes[a.shape[b0]];
                const ShapeDev& sb = a.shapes[a.shape[b1]];
                A.load(sa, a.hullPts);
                B.load(sb, a.hullPts);
                ta = loadXf(a.xf4, b0);
                tb = loadXf(a.xf4, b1);
                float mA = sa.margin, mB = sb.margin;
                float maxd = mA + mB + a.threshold;  // disp/ConvexConvexAlgorithm.java:122-123
                L.begin(ta, tb, mA, mB, maxd * maxd);
                busy = true;
                checks++;
            }
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy) {
            f3 pW = add3(mulMV(ta.m, A.support(L.dirA(ta))), L.laO);
            f3 qW = add3(mulMV(tb.m, B.support(L.dirB(tb))), L.lbO);
            if (L.iterate(pW, qW)) {
                GjkResult r;
                L.finish(r);
                busy = false;
                const bool big = (a.hist[p] & HIST_EPA_BIG) != 0;
                a.mhdr[p].pad0 = (r.curIter > 255 ? 255 : r.curIter) | (big ? HDR_EPA_BIG : 0);  // next step's ordering key (k_gjk_prefilter); travels with the manifold
                bool queued = false;
                if (r.needEpa) {
                    deep++;
                    uint32_t slot = atomicAdd(&a.ctr->epaCount, 1u);
                    if (slot < g.maxEpa) {
                        g.epaItems[slot].pair = epaRoute(a, g, p, slot, big);
                        g.epaItems[slot].meshItem = -1;
                        g.epaItems[slot].g = r;
                        a.raw[p].has_contact = -2;  // pending in the penetration bin
                        a.rawFlag[p] = -2;
                        queued = true;
                    } else {
                        a.ctr->epaFailed = 0x7fffffffu;  // capacity: reported by the host as B2C_ERR_CAPACITY
                    }
                }
                if (!queued) {
                    f3 pt = add3(r.pointOnB, r.positionOffset);
                    if (a.wantRaw || r.isValid)   // the record carries the contact to k_manifold_cc; without one the flag is enough
                    writeRaw(a.raw + p, pr, -1, r.isValid ? 1 : 0, r.isValid ? r.normalInB : mk3(0, 0, 0), r.isValid ? pt : mk3(0, 0, 0),
                             r.isValid ? r.distance : 0.f, r.lastUsedMethod, r.curIter);
                    a.rawFlag[p] = r.isValid ? 1 : 0;
                }
            }
        }
    }
    if (deep) atomicAdd(&a.ctr->deepChecks, deep);
    if (checks) atomicAdd(&a.ctr->gjkChecks, checks);
}

// ConvexConvexAlgorithm's manifold side for one pair of the GJK bins, after the detector has produced the raw
// record: getNewManifold on first use (disp/ConvexConvexAlgorithm.java:92-96), ManifoldResult.addContactPoint,
// refreshContactPoints (:136-138).
__device__ __forceinline__ void manifoldCcOne(const NpArgs& a, uint32_t p, bool has, uint32_t& added, uint32_t& created) {
    const int2 pr = a.pairs[p];
    ManifoldHdr* H = a.mhdr + p;
    int4 h0 = reinterpret_cast<const int4*>(H)[0], h1 = reinterpret_cast<const int4*>(H)[1];
    bool fresh = false;
    if (h1.y == 0) { h1.y = 3; h0.z = pr.x; h0.w = pr.y; created++; fresh = true; }
    if (h1.x == 0 && !has) {  // nothing to add, nothing to refresh
        if (fresh) { reinterpret_cast<int4*>(H)[0] = h0; reinterpret_cast<int4*>(H)[1] = h1; }
        return;
    }
    const int b0 = pr.x - 1, b1 = pr.y - 1;
    const Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
    f3 n = mk3(0.f, 0.f, 0.f), pt = n;
    float depth = 0.f, fr = 0.f, re = 0.f;
    if (has) {
        const b2c_raw_contact* r = a.raw + p;
        n = mk3(r->normal[0], r->normal[1], r->normal[2]);
        pt = mk3(r->point[0], r->point[1], r->point[2]);
        depth = r->depth;
        const float2 m0 = a.material[b0], m1 = a.material[b1];
        fr = combinedFriction(m0.x, m1.x);
        re = m0.y * m1.y;
    }
    if (manifoldStepV(H, a.mpts + 4 * (size_t)p, h0, h1, pr.x, t0, t1, has, n, pt, depth, a.threshold, fr, re)) added++;
}

// k_manifold_cc: every pair of the GJK bins whose detector finished in k_gjk / k_gjk_prefilter.  Pairs waiting in the
// penetration bin (rawFlag == -2, set by k_gjk and never touched by k_epa) are left to the manifold loop of k_epa<1>, so this kernel
// can run concurrently with the EPA kernels on another stream.
#ifndef MCC_MINB
#define MCC_MINB 4
#endif
__global__ void __launch_bounds__(256, MCC_MINB) k_manifold_cc(NpArgs a) {
    const uint32_t s0 = a.binStart[BIN_GJK0], e0 = a.binStart[BIN_COUNT];
    uint32_t added = 0, created = 0;
    for (uint32_t it = s0 + blockIdx.x * blockDim.x + threadIdx.x; it < e0; it += gridDim.x * blockDim.x) {
        uint32_t p = binItem(a, it);
        const int8_t f = a.rawFlag[p];
        if (f == -2) continue;
        manifoldCcOne(a, p, f == 1, added, created);
    }
    if (created) atomicAdd(&a.ctr->numManifolds, created);
    if (added) atomicAdd(&a.ctr->contactsAdded, added);
}

// ---- convex vs BvhTriangleMeshShape ------------------------------------------------------------------
// sh/OptimizedBvh.java:1038-1056 quantizeWithClamp (round-to-nearest for both min and max)
__device__ __forceinline__ void quantizeClamp(const MeshDev& md, f3 pt, uint32_t q[3]) {
    f3 c = mk3(jminf(jmaxf(pt.x, md.qmin[0]), md.qmax[0]), jminf(jmaxf(pt.y, md.qmin[1]), md.qmax[1]),
               jminf(jmaxf(pt.z, md.qmin[2]), md.qmax[2]));
    f3 v = mk3((c.x - md.qmin[0]) * md.quant[0], (c.y - md.qmin[1]) * md.quant[1], (c.z - md.qmin[2]) * md.quant[2]);
    q[0] = (uint32_t)__float2int_rz(v.x + 0.5f) & 0xFFFFu;
    q[1] = (uint32_t)__float2int_rz(v.y + 0.5f) & 0xFFFFu;
    q[2] = (uint32_t)__float2int_rz(v.z + 0.5f) & 0xFFFFu;
}

// convex shape AABB in mesh space (disp/ConvexTriangleCallback.java:83-106) — same float sequences as
// the world-space AABBs in broadphase.cuh, restated here to keep this header self-contained.
__device__ __forceinline__ void convexAabbIn(const ShapeDev& s, const Xf& t, f3& mn, f3& mx) {
    if (s.type == SH_SPHERE) {
        f3 e = mk3(s.margin, s.margin, s.margin);
        mn = sub3(t.o, e);
        mx = add3(t.o, e);
        return;
    }
    f3 he, c;
    if (s.type == SH_BOX) {
        he = mk3(s.dims[0] + s.margin, s.dims[1] + s.margin, s.dims[2] + s.margin);
        c = t.o;
    } else {
        f3 lmin = mk3(s.aabbMin[0], s.aabbMin[1], s.aabbMin[2]), lmax = mk3(s.aabbMax[0], s.aabbMax[1], s.aabbMax[2]);
        he = scl3(sub3(lmax, lmin), 0.5f);
        he = mk3(he.x + s.margin, he.y + s.margin, he.z + s.margin);
        c = xfPoint(t, scl3(add3(lmax, lmin), 0.5f));
    }
    f3 ext = mk3(dot3(mk3(fabsf(t.m[0][0]), fabsf(t.m[0][1]), fabsf(t.m[0][2])), he),
                 dot3(mk3(fabsf(t.m[1][0]), fabsf(t.m[1][1]), fabsf(t.m[1][2])), he),
                 dot3(mk3(fabsf(t.m[2][0]), fabsf(t.m[2][1]), fabsf(t.m[2][2])), he));
    mn = sub3(c, ext);
    mx = add3(c, ext);
}

// sh/OptimizedBvh.java:940-997 walkStacklessQuantizedTree; calls f(triIndex) for each overlapping leaf in
// array order (the order the reference folds contacts in).
template <class F>
__device__ __forceinline__ uint32_t walkBvh(const MeshDev& md, const uint32_t qmin[3], const uint32_t qmax[3], F f) {
    int cur = 0;
    uint32_t visited = 0;
    const int end = md.numNodes;
    while (cur < end) {
        int4 nd = __ldg(md.nodes + cur);
        visited++;
        uint32_t nminx = (uint32_t)nd.x & 0xFFFFu, nminy = ((uint32_t)nd.x >> 16) & 0xFFFFu, nminz = (uint32_t)nd.y & 0xFFFFu;
        uint32_t nmaxx = ((uint32_t)nd.y >> 16) & 0xFFFFu, nmaxy = (uint32_t)nd.z & 0xFFFFu, nmaxz = ((uint32_t)nd.z >> 16) & 0xFFFFu;
        bool overlap = !(qmin[0] > nmaxx || qmax[0] < nminx) && !(qmin[2] > nmaxz || qmax[2] < nminz) &&
                       !(qmin[1] > nmaxy || qmax[1] < nminy);
        bool leaf = nd.w >= 0;
        if (leaf && overlap) f(nd.w);  // partId << 21 | triangle index (sh/QuantizedBvhNodes.java:186-198)
        if (overlap || leaf) cur++;
        else cur += -nd.w;
    }
    return visited;
}

__global__ void __launch_bounds__(128) k_mesh_query(NpArgs a, GjkArgs g) {
    const uint32_t s = a.binStart[BIN_MESH], e = a.binStart[BIN_MESH + 1];
    uint32_t created = 0;
    for (uint32_t it = s + blockIdx.x * blockDim.x + threadIdx.x; it < e; it += gridDim.x * blockDim.x) {
        uint32_t p = binItem(a, it);
        int2 pr = a.pairs[p];
        int b0 = pr.x - 1, b1 = pr.y - 1;
        ShapeDev s0 = a.shapes[a.shape[b0]], s1 = a.shapes[a.shape[b1]];
        bool swapped = (s0.type == SH_MESH);
        int bc = swapped ? b1 : b0, bt = swapped ? b0 : b1;
        const ShapeDev& cs = swapped ? s1 : s0;
        const ShapeDev& ms = swapped ? s0 : s1;
        Xf tc = loadXf(a.xf4, bc), tt = loadXf(a.xf4, bt);
        MView m = mview(a, p);
        if (m.h->algorithm == 0) created++;
        m.h->algorithm = 4;
        m.h->body0 = bc + 1; m.h->body1 = bt + 1;  // manifoldPtr.setBodies(convexBody, triBody)
        for (int k = 0; k < m.h->num_contacts; k++) m.p[k].src_slot = k;
        MeshDev md = a.meshes[ms.mesh];
        Xf convexInTri = invMul(tt, tc);
        f3 mn, mx;
        convexAabbIn(cs, convexInTri, mn, mx);
        f3 extra = mk3(ms.margin, ms.margin, ms.margin);
        mx = add3(mx, extra);
        mn = sub3(mn, extra);
        uint32_t qmin[3], qmax[3];
        quantizeClamp(md, mn, qmin);
        quantizeClamp(md, mx, qmax);
        uint32_t count = 0;
        walkBvh(md, qmin, qmax, [&](int) { count++; });
        uint32_t start = atomicAdd(&a.ctr->meshItems, count);
        if (start + count > g.maxMeshItems) {
            a.ctr->meshOverflow = 1;
            g.meshStart[p] = 0;
            g.meshCount[p] = 0;
            continue;
        }
        g.meshStart[p] = start;
        g.meshCount[p] = count;
        uint32_t k = start;
        walkBvh(md, qmin, qmax, [&](int tri) { g.meshPair[k] = p; g.meshTri[k] = tri; k++; });
    }
    if (created) atomicAdd(&a.ctr->numManifolds, created);
}

// `leaf` is the BVH leaf word partId << 21 | triangleIndex (sh/OptimizedBvh.java:278); for a one-part mesh it IS the index
__device__ __forceinline__ TriS loadTri(const MeshDev& md, int leaf, float margin) {
    TriS t;
    const int tri = md.numParts > 1 ? __ldg(md.partStart + (leaf >> 21)) + (leaf & 0x1FFFFF) : leaf;
    int i0 = __ldg(md.idx + 3 * tri), i1 = __ldg(md.idx + 3 * tri + 1), i2 = __ldg(md.idx + 3 * tri + 2);
    t.a = mk3(__ldg(md.verts + 3 * i0), __ldg(md.verts + 3 * i0 + 1), __ldg(md.verts + 3 * i0 + 2));
    t.b = mk3(__ldg(md.verts + 3 * i1), __ldg(md.verts + 3 * i1 + 1), __ldg(md.verts + 3 * i1 + 2));
    t.c = mk3(__ldg(md.verts + 3 * i2), __ldg(md.verts + 3 * i2 + 1), __ldg(md.verts + 3 * i2 + 2));
    t.margin = margin;
    return t;
}

// k_gjk_tri: one lane per (pair, triangle) item: ConvexConvexAlgorithm on (convex, TriangleShape) with the
// shared manifold (disp/ConvexTriangleCallback.java:111-172).  Same persistent-lane loop as k_gjk.
__global__ void __launch_bounds__(128) k_gjk_tri(NpArgs a, GjkArgs g, uint32_t* cursor) {
    uint32_t nItems = a.ctr->meshItems < g.maxMeshItems ? a.ctr->meshItems : g.maxMeshItems;
    if (a.ctr->meshOverflow) nItems = 0;
    uint32_t deep = 0, checks = 0;
    GjkLane L;
    LaneShape A;
    TriS T;
    Xf ta, tb;
    uint32_t p = 0, item = 0;
    int tri = 0;
    int2 pr = make_int2(0, 0);
    bool busy = false, more = true;
    WarpQueue wq;
    wq.init();
    while (true) {
        const bool want = !busy && more;
        const uint32_t it = wq.take(want, cursor, nItems);
        if (want) {
            if (it == 0xffffffffu) {
                more = false;
            } else {
                item = it;
                p = g.meshPair[it];
                tri = g.meshTri[it];
                pr = a.pairs[p];
                const int b0 = pr.x - 1, b1 = pr.y - 1;
                const ShapeDev& s0 = a.shapes[a.shape[b0]];
                const ShapeDev& s1 = a.shapes[a.shape[b1]];
                const bool swapped = (s0.type == SH_MESH);
                const ShapeDev& cs = swapped ? s1 : s0;
                const ShapeDev& ms = swapped ? s0 : s1;
                ta = loadXf(a.xf4, swapped ? b1 : b0);
                tb = loadXf(a.xf4, swapped ? b0 : b1);
                A.load(cs, a.hullPts);
                T = loadTri(a.meshes[ms.mesh], tri, ms.margin);
                float mA = cs.margin, mB = ms.margin;
                float maxd = mA + mB + a.threshold;
                L.begin(ta, tb, mA, mB, maxd * maxd);
                busy = true;
                checks++;
            }
        }
        if (!__any_sync(0xffffffffu, busy)) break;
        if (busy) {
            f3 pW = add3(mulMV(ta.m, A.support(L.dirA(ta))), L.laO);
            f3 qW = add3(mulMV(tb.m, T.support(L.dirB(tb))), L.lbO);
            if (L.iterate(pW, qW)) {
                GjkResult r;
                L.finish(r);
                busy = false;
                b2c_raw_contact* rw = g.rawMesh + item;
                bool queued = false;
                if (r.needEpa) {
                    deep++;
                    uint32_t slot = atomicAdd(&a.ctr->epaCount, 1u);
                    if (slot < g.maxEpa) {
                        g.epaItems[slot].pair = epaRoute(a, g, p, slot, (a.hist[p] & HIST_EPA_BIG) != 0);
                        g.epaItems[slot].meshItem = (int)item;
                        g.epaItems[slot].g = r;
                        rw->has_contact = -2;  // pending
                        queued = true;
                    } else {
                        a.ctr->epaFailed = 0x7fffffffu;
                    }
                }
                if (!queued) {
                    f3 pt = add3(r.pointOnB, r.positionOffset);
                    writeRaw(rw, pr, tri, r.isValid ? 1 : 0, r.isValid ? r.normalInB : mk3(0, 0, 0), r.isValid ? pt : mk3(0, 0, 0),
                             r.isValid ? r.distance : 0.f, r.lastUsedMethod, r.curIter);
                }
            }
        }
    }
    if (deep) atomicAdd(&a.ctr->deepChecks, deep);
    if (checks) atomicAdd(&a.ctr->gjkChecks, checks);
}

// EPA bin: finishes np/GjkPairDetector.java:265-303 for the pairs that asked for it.  Three variants share the
// code; which of TIER 0 / TIER 2 does the work is decided ON THE DEVICE from the size of the bin (the other one
// returns at once), because the two regimes want opposite trade-offs (measured, profiles/):
//   TIER 0  small bin (<= EPA_SMEM_LANES items): latency matters.  Per-lane pools in SHARED memory (one warp per
//           block, ~107 KB, two blocks per SM, odd word stride between lanes): an EPA run is a long chain of
//           dependent small accesses, and in local memory each one is an L1-miss-prone interleaved line.
//   TIER 2  large bin: throughput matters.  Pools in local memory, 64-thread blocks, ~75 k threads in flight.
//   TIER 1  retry of the items that overflowed their pool, with the large pool in global memory (pool exhaustion
//           there = EPA failed, like the reference's EPA_Failed).
constexpr uint32_t EPA_SMEM_LANES = 148u * 2u * 32u;

constexpr uint32_t EPA_RETRY_BIT = 0x80000000u;  // set in EpaItem.pair while the item waits for (or belongs to) the large-pool tier
__device__ __forceinline__ uint32_t epaRoute(const NpArgs& a, const GjkArgs& g, uint32_t p, uint32_t slot, bool big) {
    if (!big) return p;
    const uint32_t k = atomicAdd(&a.ctr->epaBig, 1u);
    if (k >= g.maxEpaRetry) return p;   // list full: an ordinary item (it will overflow and be retried as before)
    g.epaBig[k] = slot;
    return p | EPA_RETRY_BIT;
}

template <int TIER>
__global__ void __launch_bounds__(TIER == 0 ? 256 : 64) k_epa(NpArgs a, GjkArgs g, int solo, int lpw) {
    uint32_t nItems;
    if (TIER != 1) {
        nItems = a.ctr->epaCount < g.maxEpa ? a.ctr->epaCount : g.maxEpa;
        // solo: the host launched only this variant (it knows the size of last step's bin); any count is handled
        if (!solo && TIER == 0 && nItems > EPA_SMEM_LANES) return;
        if (!solo && TIER == 2 && nItems <= EPA_SMEM_LANES) return;
    } else {
        // TIER 1 runs twice: solo = 1 over the PREDICTED list, beside the small-pool tier; solo = 0 afterwards over the items
        // that overflowed there, followed by the manifold side of the whole bin
        const uint32_t cnt = solo ? a.ctr->epaBig : a.ctr->epaRetry;
        nItems = cnt < g.maxEpaRetry ? cnt : g.maxEpaRetry;
    }
    uint32_t failed = 0;
    extern __shared__ __align__(16) unsigned char epaSmem[];
    const uint32_t first = blockIdx.x * blockDim.x + threadIdx.x, step = gridDim.x * blockDim.x;
    // TIER 1: one item per WARP (lane 0), its large pool in this warp's slice of shared memory — a retried item is a long
    // run (up to 256 polytope expansions), and in global memory every one of its dependent accesses would cost ~1 us
    const uint32_t warpsPerBlock = blockDim.x >> 5;
    // TIER 0: `lpw` active lanes per warp (32 pool slices per block whatever the block size): the fewer lanes share a
    // warp, the less one item's iterations wait behind the divergent paths of its neighbours
    const uint32_t slot0 = (threadIdx.x >> 5) * (uint32_t)lpw + (threadIdx.x & 31);
    const bool active0 = (int)(threadIdx.x & 31) < lpw;
    const uint32_t itFirst = TIER == 2 ? first
                           : TIER == 0 ? (active0 ? blockIdx.x * 32u + slot0 : 0xffffffffu)
                           : ((threadIdx.x & 31) == 0 ? blockIdx.x * warpsPerBlock + (threadIdx.x >> 5) : 0xffffffffu);
    const uint32_t itStep = TIER == 2 ? step : TIER == 0 ? gridDim.x * 32u : gridDim.x * warpsPerBlock;
    for (uint32_t it0 = itFirst; it0 < nItems; it0 += itStep) {
        const uint32_t it = TIER != 1 ? it0 : (solo ? g.epaBig[it0] : g.epaRetry[it0]);
        EpaItem item = g.epaItems[it];
        if (TIER != 1 && (item.pair & EPA_RETRY_BIT)) continue;   // the large-pool tier has it (epaRoute)
        uint32_t p = item.pair & ~EPA_RETRY_BIT;
        int2 pr = a.pairs[p];
        int b0 = pr.x - 1, b1 = pr.y - 1;
        ShapeDev s0 = a.shapes[a.shape[b0]], s1 = a.shapes[a.shape[b1]];
        Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
        AnyS A, B;
        Xf ta, tb;
        int tri = -1;
        if (item.meshItem >= 0) {
            bool swapped = (s0.type == SH_MESH);
            const ShapeDev& cs = swapped ? s1 : s0;
            const ShapeDev& ms = swapped ? s0 : s1;
            ta = swapped ? t1 : t0;
            tb = swapped ? t0 : t1;
            tri = g.meshTri[item.meshItem];
            TriS T = loadTri(a.meshes[ms.mesh], tri, ms.margin);
            A = makeAnyS(cs, a.hullPts);
            B.type = SH_TRIANGLE; B.h = mk3(0, 0, 0); B.ta = T.a; B.tb = T.b; B.tc = T.c; B.pts = nullptr; B.n = 0; B.margin = ms.margin;
        } else if (item.meshItem <= -2) {  // a child algorithm of a compound pair (compound.cuh)
            const uint32_t ci = (uint32_t)(-2 - item.meshItem);
            const uint32_t code = g.comp.itemCode[ci];
            CompoundItem cit;
            decodeCompoundItem(a, g.comp, p, code, cit);
            A = makeAnyS(a.shapes[cit.shapeA], a.hullPts);
            B = makeAnyS(a.shapes[cit.shapeB], a.hullPts);
            ta = cit.tA; tb = cit.tB;
            tri = -2 - (int)code;
        } else {
            A = makeAnyS(s0, a.hullPts);
            B = makeAnyS(s1, a.hullPts);
            ta = t0; tb = t1;
        }
        GjkResult r = item.g;
        Xf la = ta, lb = tb;
        la.o = sub3(ta.o, r.positionOffset);
        lb.o = sub3(tb.o, r.positionOffset);
        f3 wA, wB;
        bool epaFail = false, poolOverflow = false;
        bool ok;
        if (TIER == 0) {
            EpaScratchSmall* sc = reinterpret_cast<EpaScratchSmall*>(epaSmem + (size_t)slot0 * EPA_SMALL_STRIDE);
            ok = epaPenetration(A, B, la, lb, sc, wA, wB, epaFail, poolOverflow);
        } else if (TIER == 2) {
            EpaScratchLocal sc;
            ok = epaPenetration(A, B, la, lb, &sc, wA, wB, epaFail, poolOverflow);
        }
        if (TIER != 1) {
            if (poolOverflow) {
                uint32_t slot = atomicAdd(&a.ctr->epaRetry, 1u);
                if (slot < g.maxEpaRetry) {
                    g.epaRetry[slot] = it;
                    g.epaItems[it].pair = p | EPA_RETRY_BIT;
                    if (item.meshItem > -2) atomicOr(&a.mhdr[p].pad0, HDR_EPA_BIG);   // remembered by the pair (not by compound pairs: their pad0 is a counter flag)
                    continue;
                }
                epaFail = true;  // retry list full: report as failure below
            }
        } else {
            EpaScratch* sc = reinterpret_cast<EpaScratch*>(epaSmem + (size_t)(threadIdx.x >> 5) * sizeof(EpaScratch));
            ok = epaPenetration(A, B, la, lb, sc, wA, wB, epaFail, poolOverflow);
            if (poolOverflow) epaFail = true;
        }
        if (epaFail) failed++;
        bool isValid = r.isValid;
        float distance = r.distance;
        f3 pointOnB = r.pointOnB, normalInB = r.normalInB;
        int method = r.lastUsedMethod;
        if (ok) {
            f3 nrm = sub3(wB, wA);
            float lenSqr = len2_3(nrm);
            if (lenSqr > (B2C_FLT_EPSILON * B2C_FLT_EPSILON)) {
                nrm = scl3(nrm, 1.f / jsqrtf(lenSqr));
                float distance2 = -len3(sub3(wA, wB));
                if (!isValid || (distance2 < distance)) {
                    distance = distance2;
                    pointOnB = wB;
                    normalInB = nrm;
                    isValid = true;
                    method = 3;
                }
            } else {
                method = 4;
            }
        } else {
            method = 5;
        }
        f3 pt = add3(pointOnB, r.positionOffset);
        b2c_raw_contact* rw = item.meshItem >= 0 ? g.rawMesh + item.meshItem : (item.meshItem <= -2 ? g.comp.raw + (-2 - item.meshItem) : a.raw + p);
        writeRaw(rw, pr, tri, isValid ? 1 : 0, isValid ? normalInB : mk3(0, 0, 0), isValid ? pt : mk3(0, 0, 0),
                 isValid ? distance : 0.f, method, r.curIter);
        if (TIER == 1 && item.meshItem == -1) {  // the retried pair's manifold, by the thread that finished its detector
            uint32_t added = 0, created = 0;
            manifoldCcOne(a, p, isValid, added, created);
            if (created) atomicAdd(&a.ctr->numManifolds, created);
            if (added) atomicAdd(&a.ctr->contactsAdded, added);
        }
    }
    if (failed) atomicAdd(&a.ctr->epaFailed, failed);
    if (TIER == 1 && !solo) {
        // manifold side of every convex-convex pair that went through the penetration bin (ConvexConvexAlgorithm,
        // disp/ConvexConvexAlgorithm.java:92-139); retried items were done above, (pair, triangle) items are folded by
        // k_mesh_manifold
        const uint32_t nAll = a.ctr->epaCount < g.maxEpa ? a.ctr->epaCount : g.maxEpa;
        uint32_t added = 0, created = 0;
        for (uint32_t it = first; it < nAll; it += step) {
            const uint32_t pp = g.epaItems[it].pair;
            if ((pp & EPA_RETRY_BIT) || g.epaItems[it].meshItem != -1) continue;
            manifoldCcOne(a, pp, a.raw[pp].has_contact == 1, added, created);
        }
        if (created) atomicAdd(&a.ctr->numManifolds, created);
        if (added) atomicAdd(&a.ctr->contactsAdded, added);
    }
}

// per mesh pair: fold the per-triangle contacts into the shared manifold in BVH order, then one refresh
// (disp/ConvexTriangleCallback.java:160-169, disp/ConvexConcaveCollisionAlgorithm.java:89)
__global__ void __launch_bounds__(128) k_mesh_manifold(NpArgs a, GjkArgs g) {
    const uint32_t s = a.binStart[BIN_MESH], e = a.binStart[BIN_MESH + 1];
    uint32_t added = 0;
    for (uint32_t it = s + blockIdx.x * blockDim.x + threadIdx.x; it < e; it += gridDim.x * blockDim.x) {
        uint32_t p = binItem(a, it);
        int2 pr = a.pairs[p];
        int b0 = pr.x - 1, b1 = pr.y - 1;
        Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
        MView m = mview(a, p);
        float2 m0 = a.material[b0], m1 = a.material[b1];
        float fr = combinedFriction(m0.x, m1.x), re = m0.y * m1.y;
        uint32_t st = g.meshStart[p], cn = g.meshCount[p];
        for (uint32_t k = st; k < st + cn; k++) {
            const b2c_raw_contact* r = g.rawMesh + k;
            if (r->has_contact == 1) {
                if (manifoldAdd(m, pr.x, t0, t1, mk3(r->normal[0], r->normal[1], r->normal[2]), mk3(r->point[0], r->point[1], r->point[2]),
                                r->depth, a.threshold, fr, re, r->tri >> 21, r->tri & 0x1FFFFF))  // partId1, index1
                    added++;
            }
        }
        resultRefresh(m, pr.x, t0, t1, a.threshold);
        a.raw[p].has_contact = -3;  // raw records of this pair live in the mesh item array
    }
    if (added) atomicAdd(&a.ctr->contactsAdded, added);
}

// Compact the touching manifolds (header + live points) for the D2H contact stream.
// MODE 0: 96-byte b2c_manifold_point records; MODE 1: 64-byte b2c_solver_point records (what the constraint solver reads);
// MODE 2: 16-byte b2c_packed_header + 48-byte b2c_packed_point (nothing the host can derive itself crosses PCIe).
// itemPair: null for the pair manifolds; for the child manifolds of compound pairs the pair index of every work item.
template <int MODE>
__global__ void __launch_bounds__(256)
k_compact_contacts(NpArgs a, b2c_contact_header* __restrict__ hdr, void* __restrict__ ptsOut, uint32_t capH,
                   uint32_t capP, uint32_t* __restrict__ counts /*[2]: headers, points*/, const uint32_t* __restrict__ itemPair, int phase) {
    // phase 0: every touching manifold.  phase 1 / 2 split the pair manifolds in two so that the download of the first part can
    // start while the penetration bin is still running: 1 = manifolds that are final once k_manifold_cc and the closed-form bins
    // are done (everything but 2), 2 = pairs waiting in the penetration bin (rawFlag -2) and mesh pairs (folded at the very end).
    constexpr bool SLIM = MODE == 1;
    b2c_manifold_point* __restrict__ pts = reinterpret_cast<b2c_manifold_point*>(ptsOut);
    const uint32_t n = *a.numPairs;
    const int lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        uint32_t p = base + lane;
        int nc = 0;
        if (p < n && a.mhdr[p].algorithm != 0) nc = a.mhdr[p].num_contacts;
        if (phase != 0 && nc > 0) {
            const int bin = a.binOf[p];
            const bool late = bin == BIN_MESH || (bin >= BIN_GJK0 && bin < BIN_COUNT && a.rawFlag[p] == -2);
            if (late != (phase == 2)) nc = 0;
        }
        uint32_t m = __ballot_sync(0xffffffffu, nc > 0);
        if (m == 0) continue;
        // warp totals: headers = popc(m), points = sum nc (inclusive scan by shuffles)
        int incl = nc;
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        int totalP = __shfl_sync(0xffffffffu, incl, 31);
        uint32_t baseH = 0, baseP = 0;
        if (lane == 0) {
            // ONE 64-bit atomic reserves headers (low word) and points (high word) together, so the points of consecutive
            // headers are consecutive: first_point is the running sum of num_contacts in header order (format 3 relies on it)
            const unsigned long long old = atomicAdd(reinterpret_cast<unsigned long long*>(counts),
                                                     (unsigned long long)__popc(m) | ((unsigned long long)(uint32_t)totalP << 32));
            baseH = (uint32_t)old;
            baseP = (uint32_t)(old >> 32);
        }
        baseH = __shfl_sync(0xffffffffu, baseH, 0);
        baseP = __shfl_sync(0xffffffffu, baseP, 0);
        if (nc > 0) {
            uint32_t h = baseH + __popc(m & ((1u << lane) - 1u));
            uint32_t fp = baseP + (uint32_t)(incl - nc);
            if (h < capH && fp + nc <= capP) {
                const ManifoldHdr* mf = a.mhdr + p;
                if (MODE == 3) {
                    // b2c_packed_uid_header: the host needs no pair list to read it
                    int4 ph;
                    ph.x = mf->pair_uid0;
                    ph.y = mf->pair_uid1;
                    ph.z = nc | (mf->algorithm << 8) | ((mf->body0 != mf->pair_uid0) ? 0x10000 : 0);
                    ph.w = itemPair ? ((mf->pad0 & 0xffff) | (mf->pad1 << 16)) : -1;
                    reinterpret_cast<int4*>(hdr)[h] = ph;
                } else if (MODE == 2) {
                    int4 ph;
                    ph.x = itemPair ? (int)itemPair[p] : (int)p;
                    ph.y = (int)fp;
                    ph.z = nc | (mf->algorithm << 8) | ((mf->body0 != mf->pair_uid0) ? 0x10000 : 0);
                    ph.w = itemPair ? ((mf->pad0 & 0xffff) | (mf->pad1 << 16)) : -1;  // child0 | child1 << 16 (int16 each, -1 = none)
                    reinterpret_cast<int4*>(hdr)[h] = ph;
                } else {
                    b2c_contact_header hh;
                    hh.pair_uid0 = mf->pair_uid0; hh.pair_uid1 = mf->pair_uid1; hh.body0 = mf->body0; hh.body1 = mf->body1;
                    hh.num_contacts = nc; hh.algorithm = mf->algorithm; hh.first_point = (int)fp;
                    // child manifold of a compound pair: which child algorithm it is, as a negative pair_index
                    hh.pair_index = itemPair ? -1 - (((mf->pad0 + 1) & 0x7fff) | (((mf->pad1 + 1) & 0x7fff) << 15)) : (int)p;
                    hdr[h] = hh;
                }
                for (int k = 0; k < nc; k++) {
                    const int4* src = reinterpret_cast<const int4*>(a.mpts + 4 * (size_t)p + k);
                    if (MODE >= 2) {
                        // words of the 96-byte record: 6-8 world_a, 9-11 world_b, 12-14 normal, 15 distance, 18 life, 19 src_slot, 21 index1
                        const int4 v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4], v5 = src[5];
                        int4* dst = reinterpret_cast<int4*>(reinterpret_cast<b2c_packed_point*>(ptsOut) + fp + k);
                        dst[0] = make_int4(v1.z, v1.w, v2.x, v2.y);  // world_a xyz, world_b x
                        dst[1] = make_int4(v2.z, v2.w, v3.x, v3.y);  // world_b yz, normal xy
                        const int life = v4.z > 0xffffff ? 0xffffff : v4.z;
                        dst[2] = make_int4(v3.z, v3.w, (life << 8) | ((v4.w + 1) & 0xff), (v5.x << 21) | v5.y);  // normal z, distance, life | src_slot+1, partId1 << 21 | index1
                    } else if (!SLIM) {
                        int4* dst = reinterpret_cast<int4*>(pts + fp + k);
                        for (int q = 0; q < 6; q++) dst[q] = src[q];
                    } else {
                        // 96-byte record words: 0-2 local_a, 3-5 local_b, 6-8 world_a, 9-11 world_b, 12-14 normal, 15 distance,
                        // 16 friction, 17 restitution, 18 life, 19 src_slot, 20 part_id1, 21 index1, 22-23 pad
                        const int4 v1 = src[1], v2 = src[2], v3 = src[3], v4 = src[4], v5 = src[5];
                        int4* dst = reinterpret_cast<int4*>(reinterpret_cast<b2c_solver_point*>(ptsOut) + fp + k);
                        dst[0] = make_int4(v1.z, v1.w, v2.x, v2.y);  // world_a xyz, world_b x
                        dst[1] = make_int4(v2.z, v2.w, v3.x, v3.y);  // world_b yz, normal xy
                        dst[2] = make_int4(v3.z, v3.w, v4.x, v4.y);  // normal z, distance, friction, restitution
                        dst[3] = make_int4(v4.z, v4.w, v5.x, v5.y);  // life, src_slot, part_id1, index1
                    }
                }
            }
        }
    }
}

}  // namespace b2c
