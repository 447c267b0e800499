// gjk.cuh — per-thread GJK closest points (Voronoi simplex solver in registers) for the narrowphase.
//
// Computes what np/GjkPairDetector.java:73-316 computes, with np/VoronoiSimplexSolver.java:79-633 as
// the sub-simplex solver: same operation order, same epsilons, same branch decisions, because contact
// parity within 1e-4 needs identical branches on degenerate (face-face) configurations.
// Structure is GPU-first: the simplex (<= 4 vertices of w,p,q) lives in named registers with static
// indexing only, the solver is a handful of inlined functions with no virtual dispatch, and shapes are
// template parameters so a warp executes one support mapping.
#pragma once
#include "common.cuh"

namespace b2c {

constexpr float B2C_FLT_EPSILON = 1.19209290e-07f;            // BulletGlobals.java:40
constexpr float B2C_SIMD_INFINITY = 3.4028234663852886e38f;   // BulletGlobals.java:48 (Float.MAX_VALUE)
constexpr float GJK_REL_ERROR2 = 1.0e-6f;                     // np/GjkPairDetector.java:44

// ---- support mappings (sh/*Shape.java) --------------------------------------------------------------
__device__ __forceinline__ float fsel(float a, float b, float c) { return a >= 0 ? b : c; }  // lm/ScalarUtil.java:35-37

// sh/ConvexInternalShape.java:85-100: add the margin along the normalised direction
__device__ __forceinline__ f3 addMarginDir(f3 sup, f3 dir, float margin) {
    if (margin != 0.0f) {
        f3 vn = dir;
        if (len2_3(vn) < (B2C_FLT_EPSILON * B2C_FLT_EPSILON)) vn = mk3(-1.f, -1.f, -1.f);
        vn = nor3(vn);
        vn = scl3(vn, margin);
        sup = add3(sup, vn);
    }
    return sup;
}

struct BoxS {  // sh/BoxShape.java:73-97
    f3 h;      // implicitShapeDimensions
    float margin;
    __device__ __forceinline__ f3 support(f3 v) const { return mk3(fsel(v.x, h.x, -h.x), fsel(v.y, h.y, -h.y), fsel(v.z, h.z, -h.z)); }
    __device__ __forceinline__ f3 supportMargin(f3 v) const {
        f3 e = mk3(h.x + margin, h.y + margin, h.z + margin);
        return mk3(fsel(v.x, e.x, -e.x), fsel(v.y, e.y, -e.y), fsel(v.z, e.z, -e.z));
    }
};
struct SphereS {  // sh/SphereShape.java:44-47, margin = radius (:93-97)
    float margin;
    __device__ __forceinline__ f3 support(f3) const { return mk3(0.f, 0.f, 0.f); }
    __device__ __forceinline__ f3 supportMargin(f3 v) const { return addMarginDir(mk3(0.f, 0.f, 0.f), v, margin); }
};
struct HullS {  // sh/ConvexHullShape.java:75-102,142-157; points are pre-multiplied by localScaling
    const float4* __restrict__ pts;
    int n;
    float margin;
    template <bool WIDE>
    __device__ __forceinline__ f3 supportT(f3 v0) const {
        f3 sup = mk3(0.f, 0.f, 0.f);
        float maxDot = -1e30f;
        f3 v = v0;
        float l2 = len2_3(v);
        if (l2 < 0.0001f) v = mk3(1.f, 0.f, 0.f);
        else v = scl3(v, 1.0f / jsqrtf(l2));
        if (!WIDE) {
            for (int i = 0; i < n; i++) {
                float4 p = __ldg(pts + i);
                float d = v.x * p.x + v.y * p.y + v.z * p.z;
                if (d > maxDot) { maxDot = d; sup = mk3(p.x, p.y, p.z); }  // first strict max
            }
            return sup;
        }
        // WIDE (straight-line callers with registers to spare): four vertex loads in flight per round (the pool is
        // padded, so reading up to 3 entries past n is safe); the comparisons still run in vertex order, so the first
        // strict maximum wins exactly as in the reference loop
        for (int i = 0; i < n; i += 4) {
            const float4 p0 = __ldg(pts + i), p1 = __ldg(pts + i + 1), p2 = __ldg(pts + i + 2), p3 = __ldg(pts + i + 3);
            const float d0 = v.x * p0.x + v.y * p0.y + v.z * p0.z;
            const float d1 = v.x * p1.x + v.y * p1.y + v.z * p1.z;
            const float d2 = v.x * p2.x + v.y * p2.y + v.z * p2.z;
            const float d3 = v.x * p3.x + v.y * p3.y + v.z * p3.z;
            if (d0 > maxDot) { maxDot = d0; sup = mk3(p0.x, p0.y, p0.z); }
            if (i + 1 < n && d1 > maxDot) { maxDot = d1; sup = mk3(p1.x, p1.y, p1.z); }
            if (i + 2 < n && d2 > maxDot) { maxDot = d2; sup = mk3(p2.x, p2.y, p2.z); }
            if (i + 3 < n && d3 > maxDot) { maxDot = d3; sup = mk3(p3.x, p3.y, p3.z); }
        }
        return sup;
    }
    __device__ __forceinline__ f3 support(f3 v0) const { return supportT<false>(v0); }
    __device__ __forceinline__ f3 supportMargin(f3 v) const { return addMarginDir(support(v), v, margin); }
};
struct TriS {  // sh/TriangleShape.java:93-100 with lm/VectorUtil.java:41-58 maxAxis
    f3 a, b, c;
    float margin;
    __device__ __forceinline__ f3 support(f3 d) const {
        float da = dot3(d, a), db = dot3(d, b), dc = dot3(d, c);
        int k = 0;
        float mv = -1e30f;
        if (da > mv) { k = 0; mv = da; }
        if (db > mv) { k = 1; mv = db; }
        if (dc > mv) { k = 2; mv = dc; }
        return k == 0 ? a : (k == 1 ? b : c);
    }
    __device__ __forceinline__ f3 supportMargin(f3 v) const { return addMarginDir(support(v), v, margin); }
};

// ---- Voronoi simplex solver -------------------------------------------------------------------------
struct SubRes {       // np/VoronoiSimplexSolver.java:649-676
    float bary[4];
    uint32_t used;    // bit0..3 = usedVertexA..D
    bool degenerate;
};
__device__ __forceinline__ bool baryValid(const SubRes& r) {
    return r.bary[0] >= 0.f && r.bary[1] >= 0.f && r.bary[2] >= 0.f && r.bary[3] >= 0.f;
}

// np/VoronoiSimplexSolver.java:267-389 with p = origin.  Writes bary[0..2] (and 0 to bary[3]) + used mask;
// returns the closest point (needed by the tetrahedron case for the squared distance).
// The reference tests the Voronoi regions one after the other with early returns (A, B, AB, C, AC, BC, face).
// Here every lane first evaluates all the dot products and region predicates (no branches: lanes of a warp sit
// in different regions), the region is chosen with the reference's priority, and only the few operations that
// are specific to a region are branched on.  Each value is produced by the same operations in the same order as
// in the reference, so the results are bit-identical.
__device__ __noinline__ f3 closestPtOriginTriangle(f3 a, f3 b, f3 c, float& u0, float& u1, float& u2, uint32_t& used) {
    const f3 p = mk3(0.f, 0.f, 0.f);
    const f3 ab = sub3(b, a), ac = sub3(c, a), ap = sub3(p, a), bp = sub3(p, b), cp = sub3(p, c);
    const float d1 = dot3(ab, ap), d2 = dot3(ac, ap);
    const float d3 = dot3(ab, bp), d4 = dot3(ac, bp);
    const float d5 = dot3(ab, cp), d6 = dot3(ac, cp);
    const float vc = d1 * d4 - d3 * d2;
    const float vb = d5 * d2 - d1 * d6;
    const float va = d3 * d6 - d5 * d4;
    int region;  // 0 A, 1 B, 2 AB, 3 C, 4 AC, 5 BC, 6 face
    if (d1 <= 0.f && d2 <= 0.f) region = 0;
    else if (d3 >= 0.f && d4 <= d3) region = 1;
    else if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) region = 2;
    else if (d6 >= 0.f && d5 <= d6) region = 3;
    else if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) region = 4;
    else if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) region = 5;
    else region = 6;
    f3 r = a;
    u0 = 1.f; u1 = 0.f; u2 = 0.f; used = 1u;
    if (region == 1) { r = b; u0 = 0.f; u1 = 1.f; used = 2u; }
    else if (region == 3) { r = c; u0 = 0.f; u2 = 1.f; used = 4u; }
    else if (region == 2) {
        float v = d1 / (d1 - d3);
        r = mk3(v * ab.x + a.x, v * ab.y + a.y, v * ab.z + a.z);
        u0 = 1.f - v; u1 = v; used = 3u;
    } else if (region == 4) {
        float w = d2 / (d2 - d6);
        r = mk3(w * ac.x + a.x, w * ac.y + a.y, w * ac.z + a.z);
        u0 = 1.f - w; u2 = w; used = 5u;
    } else if (region == 5) {
        float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
        f3 t = sub3(c, b);
        r = mk3(w * t.x + b.x, w * t.y + b.y, w * t.z + b.z);
        u0 = 0.f; u1 = 1.f - w; u2 = w; used = 6u;
    } else if (region == 6) {
        float denom = 1.f / (va + vb + vc);
        float v = vb * denom, w = vc * denom;
        f3 t1 = scl3(ab, v), t2 = scl3(ac, w);
        r = mk3(a.x + t1.x + t2.x, a.y + t1.y + t2.y, a.z + t1.z + t2.z);
        u0 = 1.f - v - w; u1 = v; u2 = w; used = 7u;
    }
    return r;
}

// np/VoronoiSimplexSolver.java:393-425 with p = origin
__device__ __forceinline__ int originOutsideOfPlane(f3 a, f3 b, f3 c, f3 d) {
    f3 normal = crs3(sub3(b, a), sub3(c, a));
    float signp = dot3(sub3(mk3(0.f, 0.f, 0.f), a), normal);
    float signd = dot3(sub3(d, a), normal);
    if (signd * signd < ((1e-4f) * (1e-4f))) return -1;
    return (signp * signd < 0.f) ? 1 : 0;
}

struct Simplex {
    f3 W0, W1, W2, W3, P0, P1, P2, P3, Q0, Q1, Q2, Q3;
    f3 cachedP1, cachedP2, cachedV, lastW;
    int n;
    bool valid, needsUpdate;

    __device__ __forceinline__ void reset() {  // :564-570
        valid = false; n = 0; needsUpdate = true;
        lastW = mk3(1e30f, 1e30f, 1e30f);
        cachedP1 = cachedP2 = cachedV = mk3(0.f, 0.f, 0.f);
        W0 = W1 = W2 = W3 = P0 = P1 = P2 = P3 = Q0 = Q1 = Q2 = Q3 = mk3(0.f, 0.f, 0.f);
    }
    __device__ __forceinline__ void addVertex(f3 w, f3 p, f3 q) {  // :572-581
        lastW = w; needsUpdate = true;
        switch (n) {
        case 0: W0 = w; P0 = p; Q0 = q; break;
        case 1: W1 = w; P1 = p; Q1 = q; break;
        case 2: W2 = w; P2 = p; Q2 = q; break;
        default: W3 = w; P3 = p; Q3 = q; break;
        }
        n++;
    }
    __device__ __forceinline__ bool inSimplex(f3 w) const {  // :615-633
        bool found = false;
        if (n > 0 && eq3bits(W0, w)) found = true;
        if (n > 1 && eq3bits(W1, w)) found = true;
        if (n > 2 && eq3bits(W2, w)) found = true;
        if (n > 3 && eq3bits(W3, w)) found = true;
        if (eq3bits(w, lastW)) return true;
        return found;
    }
    // removeVertex(index): slot[index] = slot[--n]  (:79-85), unrolled to static register moves
    __device__ __forceinline__ void reduce(uint32_t used) {  // :87-95
        if (n >= 4 && !(used & 8u)) { n--; /* slot3 = slot3 */ }
        if (n >= 3 && !(used & 4u)) {
            n--;
            if (n == 3) { W2 = W3; P2 = P3; Q2 = Q3; }
        }
        if (n >= 2 && !(used & 2u)) {
            n--;
            if (n == 3) { W1 = W3; P1 = P3; Q1 = Q3; }
            else if (n == 2) { W1 = W2; P1 = P2; Q1 = Q2; }
        }
        if (n >= 1 && !(used & 1u)) {
            n--;
            if (n == 3) { W0 = W3; P0 = P3; Q0 = Q3; }
            else if (n == 2) { W0 = W2; P0 = P2; Q0 = Q2; }
            else if (n == 1) { W0 = W1; P0 = P1; Q0 = Q1; }
        }
    }

    // :98-264
    __device__ __forceinline__ bool update() {
        if (!needsUpdate) return valid;
        needsUpdate = false;
        SubRes r;
        r.bary[0] = r.bary[1] = r.bary[2] = r.bary[3] = 0.f;
        r.used = 0; r.degenerate = false;
        if (n == 0) {
            valid = false;
        } else if (n == 1) {
            cachedP1 = P0; cachedP2 = Q0;
            cachedV = sub3(cachedP1, cachedP2);
            valid = true;  // barycentric (1,0,0,0)
        } else if (n == 2) {
            f3 from = W0, to = W1;
            f3 diff = sub3(mk3(0.f, 0.f, 0.f), from);
            f3 v = sub3(to, from);
            float t = dot3(v, diff);
            if (t > 0) {
                float dotVV = dot3(v, v);
                if (t < dotVV) { t /= dotVV; r.used = 3u; }
                else { t = 1; r.used = 2u; }
            } else {
                t = 0; r.used = 1u;
            }
            r.bary[0] = 1.f - t; r.bary[1] = t;
            cachedP1 = add3(P0, scl3(sub3(P1, P0), t));
            cachedP2 = add3(Q0, scl3(sub3(Q1, Q0), t));
            cachedV = sub3(cachedP1, cachedP2);
            reduce(r.used);
            valid = baryValid(r);
        } else if (n == 3) {
            closestPtOriginTriangle(W0, W1, W2, r.bary[0], r.bary[1], r.bary[2], r.used);
            f3 t1 = scl3(P0, r.bary[0]), t2 = scl3(P1, r.bary[1]), t3 = scl3(P2, r.bary[2]);
            cachedP1 = mk3(t1.x + t2.x + t3.x, t1.y + t2.y + t3.y, t1.z + t2.z + t3.z);
            t1 = scl3(Q0, r.bary[0]); t2 = scl3(Q1, r.bary[1]); t3 = scl3(Q2, r.bary[2]);
            cachedP2 = mk3(t1.x + t2.x + t3.x, t1.y + t2.y + t3.y, t1.z + t2.z + t3.z);
            cachedV = sub3(cachedP1, cachedP2);
            reduce(r.used);
            valid = baryValid(r);
        } else if (n == 4) {
            // :428-561 closestPtPointTetrahedron
            int oABC = originOutsideOfPlane(W0, W1, W2, W3);
            int oACD = originOutsideOfPlane(W0, W2, W3, W1);
            int oADB = originOutsideOfPlane(W0, W3, W1, W2);
            int oBDC = originOutsideOfPlane(W1, W3, W2, W0);
            bool hasSeparation;
            if (oABC < 0 || oACD < 0 || oADB < 0 || oBDC < 0) {
                r.degenerate = true;
                hasSeparation = false;
            } else if (oABC == 0 && oACD == 0 && oADB == 0 && oBDC == 0) {
                hasSeparation = false;
            } else {
                hasSeparation = true;
                float best = B2C_SIMD_INFINITY;
                r.used = 15u;
                float u0, u1, u2;
                uint32_t us;
                // the four face tests run for every lane of the warp that is in this case (convergent calls);
                // a face the origin is not outside of simply does not update the best candidate
                {
                    f3 q = closestPtOriginTriangle(W0, W1, W2, u0, u1, u2, us);
                    float sq = dot3(q, q);  // (q - 0).(q - 0): q - p with p = 0 keeps q bit-for-bit
                    if (oABC != 0 && sq < best) {
                        best = sq;
                        r.used = (us & 1u) | (us & 2u) | (us & 4u);
                        r.bary[0] = u0; r.bary[1] = u1; r.bary[2] = u2; r.bary[3] = 0.f;
                    }
                }
                {
                    f3 q = closestPtOriginTriangle(W0, W2, W3, u0, u1, u2, us);
                    float sq = dot3(q, q);
                    if (oACD != 0 && sq < best) {
                        best = sq;
                        r.used = (us & 1u) | ((us & 2u) << 1) | ((us & 4u) << 1);
                        r.bary[0] = u0; r.bary[1] = 0.f; r.bary[2] = u1; r.bary[3] = u2;
                    }
                }
                {
                    f3 q = closestPtOriginTriangle(W0, W3, W1, u0, u1, u2, us);
                    float sq = dot3(q, q);
                    if (oADB != 0 && sq < best) {
                        best = sq;
                        r.used = (us & 1u) | ((us & 4u) >> 1) | ((us & 2u) << 2);
                        r.bary[0] = u0; r.bary[1] = u2; r.bary[2] = 0.f; r.bary[3] = u1;
                    }
                }
                {
                    f3 q = closestPtOriginTriangle(W1, W3, W2, u0, u1, u2, us);
                    float sq = dot3(q, q);
                    if (oBDC != 0 && sq < best) {
                        best = sq;
                        r.used = ((us & 1u) << 1) | (us & 4u) | ((us & 2u) << 2);
                        r.bary[0] = 0.f; r.bary[1] = u0; r.bary[2] = u2; r.bary[3] = u1;
                    }
                }
            }
            if (hasSeparation) {
                f3 t1 = scl3(P0, r.bary[0]), t2 = scl3(P1, r.bary[1]), t3 = scl3(P2, r.bary[2]), t4 = scl3(P3, r.bary[3]);
                cachedP1 = mk3(t1.x + t2.x + t3.x + t4.x, t1.y + t2.y + t3.y + t4.y, t1.z + t2.z + t3.z + t4.z);
                t1 = scl3(Q0, r.bary[0]); t2 = scl3(Q1, r.bary[1]); t3 = scl3(Q2, r.bary[2]); t4 = scl3(Q3, r.bary[3]);
                cachedP2 = mk3(t1.x + t2.x + t3.x + t4.x, t1.y + t2.y + t3.y + t4.y, t1.z + t2.z + t3.z + t4.z);
                cachedV = sub3(cachedP1, cachedP2);
                reduce(r.used);
                valid = baryValid(r);
            } else if (r.degenerate) {
                valid = false;
            } else {
                valid = true;
                cachedV = mk3(0.f, 0.f, 0.f);
            }
        } else {
            valid = false;
        }
        return valid;
    }
};

// State a pair carries from the GJK kernel to the EPA kernel (np/GjkPairDetector.java:260-303).
struct GjkResult {
    f3 pointOnA, pointOnB, normalInB, positionOffset;
    float distance;
    int lastUsedMethod, curIter, degenerate;
    bool isValid, needEpa;
};

// np/GjkPairDetector.java:73-258 as a resumable per-lane state machine: begin() is the prologue (:76-117),
// iterate() is one trip of the for(;;) loop (:132-228) and returns true when the loop ends, finish() is
// the epilogue up to (not including) the penetration-depth call (:230-264).  Splitting it this way lets a
// warp keep all 32 lanes inside the same loop body while each lane works on its own pair and refills from
// the work list as soon as its pair terminates (iteration counts differ wildly between pairs).
struct GjkLane {
    Simplex S;
    f3 axis, positionOffset, laO, lbO;
    float squaredDistance, maxDistSq, marginA, marginB;
    int curIter, degenerate;
    bool checkSimplex, checkPenetration;

    __device__ __forceinline__ void begin(const Xf& ta, const Xf& tb, float mA, float mB, float maxd) {
        positionOffset = scl3(add3(ta.o, tb.o), 0.5f);
        laO = sub3(ta.o, positionOffset);
        lbO = sub3(tb.o, positionOffset);
        marginA = mA; marginB = mB; maxDistSq = maxd;
        curIter = 0;
        axis = mk3(0.f, 1.f, 0.f);  // :102
        checkSimplex = false; checkPenetration = true;
        degenerate = 0;
        squaredDistance = B2C_SIMD_INFINITY;
        S.reset();
    }
    __device__ __forceinline__ f3 dirA(const Xf& ta) const { return mulMtV(ta.m, neg3(axis)); }
    __device__ __forceinline__ f3 dirB(const Xf& tb) const { return mulMtV(tb.m, axis); }
    // pW / qW: the two support points already mapped through the recentred transforms
    __device__ __forceinline__ bool iterate(f3 pW, f3 qW) {
        f3 w = sub3(pW, qW);
        float delta = dot3(axis, w);
        if ((delta > 0.f) && (delta * delta > squaredDistance * maxDistSq)) { checkPenetration = false; return true; }
        if (S.inSimplex(w)) { degenerate = 1; checkSimplex = true; return true; }
        float f0 = squaredDistance - delta;
        float f1 = squaredDistance * GJK_REL_ERROR2;
        if (f0 <= f1) {
            if (f0 <= 0.f) degenerate = 2;
            checkSimplex = true;
            return true;
        }
        S.addVertex(w, pW, qW);
        bool ok = S.update();
        axis = S.cachedV;
        if (!ok) { degenerate = 3; checkSimplex = true; return true; }
        if (len2_3(axis) < GJK_REL_ERROR2) { degenerate = 6; checkSimplex = true; return true; }
        float prev = squaredDistance;
        squaredDistance = len2_3(axis);
        if (prev - squaredDistance <= B2C_FLT_EPSILON * prev) { checkSimplex = true; return true; }  // backup_closest: axis is cachedV
        if (curIter++ > 1000) return true;
        if (S.n == 4) return true;  // fullSimplex (backup_closest is a no-op: axis is cachedV)
        return false;
    }
    __device__ __forceinline__ void finish(GjkResult& out) {
        float distance = 0.f;
        f3 normalInB = mk3(0.f, 0.f, 0.f);
        f3 pointOnA = mk3(0.f, 0.f, 0.f), pointOnB = mk3(0.f, 0.f, 0.f);
        bool isValid = false;
        int lastUsedMethod = -1;
        const float margin = marginA + marginB;
        if (checkSimplex) {
            pointOnA = S.cachedP1;  // compute_points (:643-647): the cache is current
            pointOnB = S.cachedP2;
            normalInB = sub3(pointOnA, pointOnB);
            float lenSqr = len2_3(axis);
            if (lenSqr < 0.0001f) degenerate = 5;
            if (lenSqr > B2C_FLT_EPSILON * B2C_FLT_EPSILON) {
                float rlen = 1.f / jsqrtf(lenSqr);
                normalInB = scl3(normalInB, rlen);
                float s = jsqrtf(squaredDistance);
                pointOnA = sub3(pointOnA, scl3(axis, marginA / s));
                pointOnB = add3(pointOnB, scl3(axis, marginB / s));
                distance = ((1.f / rlen) - margin);
                isValid = true;
                lastUsedMethod = 1;
            } else {
                lastUsedMethod = 2;
            }
        }
        bool catchDegenerate = (degenerate != 0) && ((distance + margin) < 0.01f);
        out.needEpa = checkPenetration && (!isValid || catchDegenerate);
        out.isValid = isValid;
        out.distance = distance;
        out.pointOnA = pointOnA;
        out.pointOnB = pointOnB;
        out.normalInB = normalInB;
        out.positionOffset = positionOffset;
        out.lastUsedMethod = lastUsedMethod;
        out.curIter = curIter;
        out.degenerate = degenerate;
    }
};

}  // namespace b2c
