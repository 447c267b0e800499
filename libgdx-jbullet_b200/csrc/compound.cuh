// compound.cuh — CompoundShape pairs: the last entry of the default dispatch table
// (disp/DefaultCollisionConfiguration.java:198-204) that the device path did not cover.
//
// Replaces, for one step:
//   disp/CompoundCollisionAlgorithm.java:49-75   init: one child algorithm (and so one manifold) per child
//   disp/CompoundCollisionAlgorithm.java:83-129  processCollision: every child in turn, the compound's object temporarily
//       carrying the child's shape and orgTrans * childTrans; no culling of children (the reference has none either)
//   sh/CompoundShape.java:50-82, 124-160          local AABB of the children, getAabb (broadphase.cuh shapeAabb)
//
// GPU structure.  The reference walks a pair's children sequentially, but every child algorithm owns its own manifold and
// reads only the two world transforms, so the (pair, child[, child]) combinations are independent work items:
//   k_compound_expand    one thread per compound pair: reserves count = n0 * n1 consecutive items, remembers where the pair's
//                        items of the previous dispatch start (kept in the pair's manifold header, which k_carry moves with
//                        the pair), so child manifolds persist exactly as long as the pair stays in the cache
//   k_compound_gjk       the child algorithm's detector — sphere-sphere, convex-plane, or GJK on (child shape,
//                        orgTrans * childTrans) — into a raw record; persistent lanes that refill as their item ends
//   k_epa (narrowphase.cuh) items whose GJK asks for the penetration solver join the pairs' penetration bin: same tiers
//   k_compound_manifold  one thread per item: manifold of the previous dispatch -> this dispatch's slot, then
//                        ManifoldResult.addContactPoint / refreshContactPoints against the ORIGINAL transforms of the pair's
//                        two objects ("the contactpoint is still projected back using the original inverted worldtrans", :112)
//   k_compound_mesh      child x BvhTriangleMeshShape (ConvexConcave per child): BVH query, per-triangle detector and the
//                        order-dependent fold into the child's manifold, one thread per child work item
// Pairs that are in the cache but not dispatched this step (both objects asleep, or linked by a constraint) keep their child
// manifolds untouched: their items are copied forward and nothing else (BIN_COMPOUND_KEEP).
//
// Child algorithm of item (i[, j]) — A is always the compound's child, B the other object (or its child):
//   compound(body0) x other(body1)      : A = child i of body0, B = body1                    (compoundCreateFunc)
//   other(body0) x compound(body1)      : A = child i of body1, B = body0                    (swappedCompoundCreateFunc)
//   compound(body0) x compound(body1)   : outer i over body0's children, inner j over body1's: the child algorithm of
//                                         (child i, body1) is the swapped compound algorithm, so A = child j of body1,
//                                         B = child i of body0 (:57-75 findAlgorithm(colObj, otherObj) per child)
// and the child manifold's bodies are (A's object, B's object) — getNewManifold(body0, body1) of the child algorithm.
#pragma once
#include "broadphase.cuh"
#include "narrowphase.cuh"

namespace b2c {

__global__ void __launch_bounds__(128) k_compound_expand(NpArgs a, CompoundArgs c) {
    const uint32_t s = a.binStart[BIN_COMPOUND], mid = a.binStart[BIN_COMPOUND_KEEP], e = a.binStart[BIN_COMPOUND_KEEP + 1];
    uint32_t counted = 0;
    for (uint32_t it = s + blockIdx.x * blockDim.x + threadIdx.x; it < e; it += gridDim.x * blockDim.x) {
        const uint32_t p = binItem(a, it);
        const bool keep = it >= mid;
        const int2 pr = a.pairs[p];
        ManifoldHdr* H = a.mhdr + p;
        int4 h0 = reinterpret_cast<const int4*>(H)[0], h1 = reinterpret_cast<const int4*>(H)[1];
        const bool had = h1.y == 5;     // the pair already owns child algorithms (bp/BroadphasePair.java:37-40 caches them)
        if (keep && !had) continue;     // never dispatched: no algorithm, no manifolds
        const uint32_t count = compoundItemCount(a, pr);
        const uint32_t start = atomicAdd(&c.cc->itemCount, count);
        if (start + count > c.maxItems) {
            c.cc->overflow = 1;
            h1.y = 0; h1.w = 0;         // the pair starts afresh once the capacity problem is fixed
            reinterpret_cast<int4*>(H)[1] = h1;
            continue;
        }
        const int prev = had ? h1.w : -1;
        if (h1.z == 0) counted += count;  // getNumManifolds: once per step (k_carry clears the word when it moves the header)
        h0 = make_int4(pr.x, pr.y, pr.x, pr.y);
        h1 = make_int4(0, 5, 1, (int)start);
        reinterpret_cast<int4*>(H)[0] = h0;
        reinterpret_cast<int4*>(H)[1] = h1;
        for (uint32_t k = 0; k < count; k++) {
            c.itemPair[start + k] = p;
            c.itemCode[start + k] = k | (keep ? CITEM_KEEP : 0u);
            c.itemPrev[start + k] = prev >= 0 ? prev + (int)k : -1;
        }
    }
    if (counted) atomicAdd(&a.ctr->numManifolds, counted);
}

// k_compound_gjk: the detector of every child algorithm.  Persistent warps as in k_gjk: a lane owns one item at a time and
// refills from the work list as soon as its item ends (most (child, other) combinations are far apart and end in trip 2);
// the closed-form child algorithms (sphere-sphere, convex-plane) are finished inside the refill.  Items whose detector asks
// for the penetration solver join the pairs' penetration bin (EpaItem.meshItem = -2 - item) and are finished by k_epa.
__global__ void __launch_bounds__(128, 4) k_compound_gjk(NpArgs a, GjkArgs g, uint32_t* cursor) {
    const CompoundArgs& c = g.comp;
    const uint32_t total = c.cc->itemCount;
    const uint32_t n = (c.cc->overflow || total > c.maxItems) ? 0u : total;
    if (blockIdx.x == 0 && threadIdx.x == 0) c.cc->numItems = n;
    uint32_t checks = 0, deep = 0;
    GjkLane L;
    LaneShape A, B;
    Xf ta, tb;
    uint32_t item = 0, p = 0;
    int tag = 0;
    int2 pr = make_int2(0, 0);
    bool busy = false, more = true;
    WarpQueue wq;
    wq.init();
    while (true) {
        const bool want = !busy && more;
        const uint32_t idx = wq.take(want, cursor, n);
        if (want) {
            if (idx == 0xffffffffu) {
                more = false;
            } else {
                const uint32_t code = c.itemCode[idx];
                if (!(code & CITEM_KEEP)) {
                    item = idx;
                    p = c.itemPair[idx];
                    CompoundItem it;
                    decodeCompoundItem(a, c, p, code, it);
                    pr = it.pr;
                    tag = -2 - (int)code;  // raw-record key of a child algorithm: -2 - k
                    const ShapeDev& sa = a.shapes[it.shapeA];
                    const ShapeDev& sb = a.shapes[it.shapeB];
                    b2c_raw_contact* rw = c.raw + item;
                    if (sb.type == SH_MESH) {
                        rw->has_contact = -3;  // child x mesh: k_compound_mesh walks the BVH and writes per-triangle records
                    } else if (sa.type == SH_SPHERE && sb.type == SH_SPHERE) {
                        // disp/SphereSphereCollisionAlgorithm.java:73-134 on (child, other)
                        const float r0 = sa.dims[0], r1 = sb.dims[0];
                        f3 diff = sub3(it.tA.o, it.tB.o);
                        float len = len3(diff);
                        if (len > (r0 + r1)) {
                            writeRaw(rw, pr, tag, 0, mk3(0, 0, 0), mk3(0, 0, 0), 0.f, 10, 0);
                        } else {
                            float dist = len - (r0 + r1);
                            f3 nrm = mk3(1.f, 0.f, 0.f);
                            if (len > B2C_FLT_EPSILON) nrm = scl3(diff, 1.f / len);
                            f3 pos1 = add3(it.tB.o, scl3(nrm, r1));
                            writeRaw(rw, pr, tag, 1, nrm, pos1, dist, 10, 0);
                        }
                    } else if (sb.type == SH_PLANE) {
                        // disp/ConvexPlaneCollisionAlgorithm.java:75-136, convex = the child (convexPlaneCF, never swapped here)
                        const f3 planeNormal = mk3(sb.plane[0], sb.plane[1], sb.plane[2]);
                        const float planeConstant = sb.plane[3];
                        Xf planeInConvex = invMul(it.tA, it.tB);
                        Xf convexInPlane = invMul(it.tB, it.tA);
                        f3 dir = mulMV(planeInConvex.m, neg3(planeNormal));
                        AnyS shp = makeAnyS(sa, a.hullPts);
                        f3 vtx = shp.supportMargin(dir);
                        f3 vtxInPlane = xfPoint(convexInPlane, vtx);
                        float distance = dot3(planeNormal, vtxInPlane) - planeConstant;
                        f3 projected = sub3(vtxInPlane, scl3(planeNormal, distance));
                        f3 world = xfPoint(it.tB, projected);
                        bool has = distance < a.threshold;
                        f3 nW = mulMV(it.tB.m, planeNormal);
                        writeRaw(rw, pr, tag, has ? 1 : 0, nW, world, distance, 11, 0);
                    } else {
                        // disp/ConvexConvexAlgorithm.java:90-139: GjkPairDetector on (child, other)
                        A.load(sa, a.hullPts);
                        B.load(sb, a.hullPts);
                        ta = it.tA;
                        tb = it.tB;
                        const float mA = sa.margin, mB = sb.margin;
                        const float maxd = mA + mB + a.threshold;
                        L.begin(ta, tb, mA, mB, maxd * maxd);
                        busy = true;
                        checks++;
                    }
                }
            }
        }
        if (!__any_sync(0xffffffffu, busy)) {
            if (!__any_sync(0xffffffffu, more)) break;
            continue;
        }
        if (busy) {
            f3 pW = add3(mulMV(ta.m, A.support(L.dirA(ta))), L.laO);
            f3 qW = add3(mulMV(tb.m, B.support(L.dirB(tb))), L.lbO);
            if (L.iterate(pW, qW)) {
                GjkResult r;
                L.finish(r);
                busy = false;
                b2c_raw_contact* rw = c.raw + item;
                bool queued = false;
                if (r.needEpa) {
                    deep++;
                    uint32_t slot = atomicAdd(&a.ctr->epaCount, 1u);
                    if (slot < g.maxEpa) {
                        g.epaItems[slot].pair = p;
                        g.epaItems[slot].meshItem = -2 - (int)item;
                        g.epaItems[slot].g = r;
                        rw->has_contact = -2;  // pending in the penetration bin
                        queued = true;
                    } else {
                        a.ctr->epaFailed = 0x7fffffffu;  // capacity: reported by the host as B2C_ERR_CAPACITY
                    }
                }
                if (!queued) {
                    f3 pt = add3(r.pointOnB, r.positionOffset);
                    writeRaw(rw, pr, tag, r.isValid ? 1 : 0, r.isValid ? r.normalInB : mk3(0, 0, 0), r.isValid ? pt : mk3(0, 0, 0),
                             r.isValid ? r.distance : 0.f, r.lastUsedMethod, r.curIter);
                }
            }
        }
    }
    if (checks) atomicAdd(&a.ctr->gjkChecks, checks);
    if (deep) atomicAdd(&a.ctr->deepChecks, deep);
}

__global__ void __launch_bounds__(128) k_compound_manifold(NpArgs a, CompoundArgs c) {
    const uint32_t n = c.cc->numItems;
    uint32_t added = 0;
    for (uint32_t item = blockIdx.x * blockDim.x + threadIdx.x; item < n; item += gridDim.x * blockDim.x) {
        const uint32_t code = c.itemCode[item];
        const int prev = c.itemPrev[item];
        MView m;
        m.h = c.H + item;
        m.p = c.P + 4 * (size_t)item;
        int4* hd = reinterpret_cast<int4*>(m.h);
        if (prev >= 0) {  // the child algorithm's manifold lives as long as the pair's algorithm
            const int4* hs = reinterpret_cast<const int4*>(c.prevH + prev);
            const int4 g0 = hs[0], g1 = hs[1];
            hd[0] = g0; hd[1] = g1;
            const int4* src = reinterpret_cast<const int4*>(c.prevP + 4 * (size_t)prev);
            int4* dst = reinterpret_cast<int4*>(m.p);
            for (int q = 0; q < g1.x && q < 4; q++)
                for (int w = 0; w < 6; w++) dst[6 * q + w] = src[6 * q + w];
        } else {
            hd[0] = make_int4(0, 0, 0, 0);
            hd[1] = make_int4(0, 0, 0, 0);
        }
        if (code & CITEM_KEEP) continue;
        CompoundItem it;
        decodeCompoundItem(a, c, c.itemPair[item], code, it);
        if (m.h->algorithm == 0) {  // getNewManifold(body0, body1) of the child algorithm
            const int ta = a.shapes[it.shapeA].type, tb = a.shapes[it.shapeB].type;
            m.h->pair_uid0 = it.pr.x; m.h->pair_uid1 = it.pr.y;
            m.h->body0 = it.bodyA + 1; m.h->body1 = it.bodyB + 1;
            m.h->num_contacts = 0;
            m.h->algorithm = (ta == SH_SPHERE && tb == SH_SPHERE) ? 1 : (tb == SH_PLANE ? 2 : (tb == SH_MESH ? 4 : 3));
            m.h->pad0 = it.childA; m.h->pad1 = it.childB;
        }
        for (int q = 0; q < m.h->num_contacts; q++) m.p[q].src_slot = q;
        if (a.shapes[it.shapeB].type == SH_MESH) continue;  // contacts are folded in by k_compound_mesh
        // ManifoldResult of the PAIR (disp/ManifoldResult.java:70-75): root transforms and materials of its two objects
        const int b0 = it.pr.x - 1, b1 = it.pr.y - 1;
        const Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
        const b2c_raw_contact* r = c.raw + item;
        if (r->has_contact == 1) {
            const float2 m0 = a.material[b0], m1 = a.material[b1];
            if (manifoldAdd(m, it.pr.x, t0, t1, mk3(r->normal[0], r->normal[1], r->normal[2]), mk3(r->point[0], r->point[1], r->point[2]),
                            r->depth, a.threshold, combinedFriction(m0.x, m1.x), m0.y * m1.y, 0, 0))
                added++;
        }
        resultRefresh(m, it.pr.x, t0, t1, a.threshold);
    }
    if (added) atomicAdd(&a.ctr->contactsAdded, added);
}

// np/GjkPairDetector.java:265-303: fold the penetration solver's answer into the detector result
__device__ __forceinline__ void mergeEpaResult(const GjkResult& r, bool ok, f3 wA, f3 wB, bool& isValid, float& distance, f3& pointOnB,
                                               f3& normalInB, int& method) {
    isValid = r.isValid;
    distance = r.distance;
    pointOnB = r.pointOnB;
    normalInB = r.normalInB;
    method = r.lastUsedMethod;
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
}

// k_compound_mesh: child x BvhTriangleMeshShape = ConvexConcaveCollisionAlgorithm per child with the child's own manifold
// (disp/ConvexConcaveCollisionAlgorithm.java:65-93, disp/ConvexTriangleCallback.java:83-172).  One thread per child work item:
// the BVH query with the child's box in mesh space, then for every triangle in BVH order the convex-convex detector
// (child, TriangleShape) and ManifoldResult.addContactPoint into the shared manifold (order-dependent 4-point reduction), one
// refresh at the end.  Sequential per item — compounds resting on a mesh touch a handful of triangles per child — with the
// penetration solver in the thread (medium pool in local memory, full-size pool in this thread's global-memory slot on overflow).
__global__ void __launch_bounds__(64) k_compound_mesh(NpArgs a, GjkArgs g) {
    const CompoundArgs& c = g.comp;
    const uint32_t n = c.cc->numItems;
    uint32_t checks = 0, deep = 0, failed = 0, added = 0;
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    for (uint32_t item = tid; item < n; item += gridDim.x * blockDim.x) {
        const uint32_t code = c.itemCode[item];
        if (code & CITEM_KEEP) continue;
        CompoundItem it;
        decodeCompoundItem(a, c, c.itemPair[item], code, it);
        const ShapeDev& ms = a.shapes[it.shapeB];
        if (ms.type != SH_MESH) continue;
        const ShapeDev& cs = a.shapes[it.shapeA];
        const MeshDev md = a.meshes[ms.mesh];
        const Xf convexInTri = invMul(it.tB, it.tA);
        f3 mn, mx;
        convexAabbIn(cs, convexInTri, mn, mx);
        const f3 extra = mk3(ms.margin, ms.margin, ms.margin);
        mx = add3(mx, extra);
        mn = sub3(mn, extra);
        uint32_t qmin[3], qmax[3];
        quantizeClamp(md, mn, qmin);
        quantizeClamp(md, mx, qmax);
        uint32_t count = 0;
        walkBvh(md, qmin, qmax, [&](int) { count++; });
        const uint32_t start = atomicAdd(&a.ctr->meshItems, count);
        if (start + count > g.maxMeshItems) {
            a.ctr->meshOverflow = 1;
            c.meshStart[item] = 0;
            c.meshCount[item] = 0;
            continue;
        }
        c.meshStart[item] = start;
        c.meshCount[item] = count;
        MView m;
        m.h = c.H + item;
        m.p = c.P + 4 * (size_t)item;
        const int b0 = it.pr.x - 1, b1 = it.pr.y - 1;
        const Xf t0 = loadXf(a.xf4, b0), t1 = loadXf(a.xf4, b1);
        const float2 m0 = a.material[b0], m1 = a.material[b1];
        const float fr = combinedFriction(m0.x, m1.x), re = m0.y * m1.y;
        LaneShape A;
        A.load(cs, a.hullPts);
        const float mA = cs.margin, mB = ms.margin;
        const float maxd = mA + mB + a.threshold;
        uint32_t k = start;
        walkBvh(md, qmin, qmax, [&](int tri) {
            const TriS T = loadTri(md, tri, ms.margin);
            GjkLane L;
            L.begin(it.tA, it.tB, mA, mB, maxd * maxd);
            checks++;
            for (;;) {
                f3 pW = add3(mulMV(it.tA.m, A.support(L.dirA(it.tA))), L.laO);
                f3 qW = add3(mulMV(it.tB.m, T.support(L.dirB(it.tB))), L.lbO);
                if (L.iterate(pW, qW)) break;
            }
            GjkResult r;
            L.finish(r);
            bool isValid = r.isValid;
            float distance = r.distance;
            f3 pointOnB = r.pointOnB, normalInB = r.normalInB;
            int method = r.lastUsedMethod;
            if (r.needEpa) {
                deep++;
                AnyS EA = makeAnyS(cs, a.hullPts), EB;
                EB.type = SH_TRIANGLE; EB.h = mk3(0, 0, 0); EB.ta = T.a; EB.tb = T.b; EB.tc = T.c; EB.pts = nullptr; EB.n = 0; EB.margin = ms.margin;
                Xf la = it.tA, lb = it.tB;
                la.o = sub3(it.tA.o, r.positionOffset);
                lb.o = sub3(it.tB.o, r.positionOffset);
                f3 wA, wB;
                bool epaFail = false, poolOverflow = false, ok;
                {
                    EpaScratchLocal sc;
                    ok = epaPenetration(EA, EB, la, lb, &sc, wA, wB, epaFail, poolOverflow);
                }
                if (poolOverflow) {
                    if (c.bigScratch && tid < c.numBigScratch) {
                        EpaScratch* big = reinterpret_cast<EpaScratch*>(c.bigScratch) + tid;
                        ok = epaPenetration(EA, EB, la, lb, big, wA, wB, epaFail, poolOverflow);
                    }
                    if (poolOverflow) { epaFail = true; ok = false; }
                }
                if (epaFail) failed++;
                mergeEpaResult(r, ok, wA, wB, isValid, distance, pointOnB, normalInB, method);
            }
            const f3 pt = add3(pointOnB, r.positionOffset);
            // raw-record key of (child algorithm, triangle): -2 - (k << 21 | t)
            writeRaw(g.rawMesh + k, it.pr, -2 - (int)((code << 21) | ((uint32_t)tri & 0x1FFFFFu)), isValid ? 1 : 0, isValid ? normalInB : mk3(0, 0, 0),
                     isValid ? pt : mk3(0, 0, 0), isValid ? distance : 0.f, method, r.curIter);
            if (isValid) {
                if (manifoldAdd(m, it.pr.x, t0, t1, normalInB, pt, distance, a.threshold, fr, re, tri >> 21, tri & 0x1FFFFF)) added++;
            }
            k++;
        });
        resultRefresh(m, it.pr.x, t0, t1, a.threshold);
    }
    if (checks) atomicAdd(&a.ctr->gjkChecks, checks);
    if (deep) atomicAdd(&a.ctr->deepChecks, deep);
    if (failed) atomicAdd(&a.ctr->epaFailed, failed);
    if (added) atomicAdd(&a.ctr->contactsAdded, added);
}

// registration helper: the compound's local AABB from its children, with the same device code the per-step AABB kernel uses
// (sh/CompoundShape.java:50-82: running Math.min / Math.max of child.getAabb(childTransform))
__global__ void k_compound_local_aabb(const ShapeDev* __restrict__ shapes, const CompoundChildDev* __restrict__ children, int first, int n,
                                      float* __restrict__ out6) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    f3 lo = mk3(1e30f, 1e30f, 1e30f), hi = mk3(-1e30f, -1e30f, -1e30f);
    for (int i = 0; i < n; i++) {
        const CompoundChildDev& ch = children[first + i];
        Xf l;
        l.m[0][0] = ch.m[0]; l.m[0][1] = ch.m[1]; l.m[0][2] = ch.m[2];
        l.m[1][0] = ch.m[3]; l.m[1][1] = ch.m[4]; l.m[1][2] = ch.m[5];
        l.m[2][0] = ch.m[6]; l.m[2][1] = ch.m[7]; l.m[2][2] = ch.m[8];
        l.o = mk3(ch.o[0], ch.o[1], ch.o[2]);
        f3 mn, mx;
        shapeAabb(shapes[ch.shape], l, mn, mx);
        lo = mk3(jminf(lo.x, mn.x), jminf(lo.y, mn.y), jminf(lo.z, mn.z));  // lm/VectorUtil.java:176-180
        hi = mk3(jmaxf(hi.x, mx.x), jmaxf(hi.y, mx.y), jmaxf(hi.z, mx.z));  // :182-186
    }
    out6[0] = lo.x; out6[1] = lo.y; out6[2] = lo.z;
    out6[3] = hi.x; out6[4] = hi.y; out6[5] = hi.z;
}

}  // namespace b2c
