// epa.cuh — penetration depth for pairs whose cores interpenetrate (the rare narrowphase bin).
//
// Computes what np/GjkEpaPenetrationDepthSolver.java:41-63 -> np/GjkEpaSolver.java:864-911 computes:
// a second GJK on the margin-inflated shapes (SearchOrigin, :363-420, <=128 iterations, revisited-ray
// check), EncloseOrigin (:422-498) and the polytope expansion EvaluatePD (:712-856, <=256 iterations).
// GPU structure: one thread per pair, all dynamic storage replaced by a fixed per-thread pool in global
// memory addressed by index (no pointers, no recursion): Minkowski vertices, faces, the visited-ray list
// and an explicit stack for the horizon walk.  Pool exhaustion is reported as EPA failure, like the
// reference's own EPA_Failed status, never as undefined behaviour.
#pragma once
#include "common.cuh"
#include "gjk.cuh"

namespace b2c {

constexpr int EPA_MAXV = 272;     // 5 base + one support vertex per EPA iteration (<=256) + slack
constexpr int EPA_MAXF = 640;     // faces ever created for one pair
constexpr int EPA_MAXRAY = 132;   // rays seen by SearchOrigin (<=129)
constexpr int EPA_MAXSTK = 192;   // horizon walk stack
constexpr int EPA_GJK_MAXIT = 128;       // np/GjkEpaSolver.java:108
constexpr float EPA_INSIMPLEX_EPS = 0.0001f;  // :111
constexpr float EPA_SQINSIMPLEX_EPS = EPA_INSIMPLEX_EPS * EPA_INSIMPLEX_EPS;
constexpr int EPA_MAXIT = 256;           // :113
constexpr float EPA_INFACE_EPS = 0.01f;  // :114
constexpr float EPA_ACCURACY = 0.001f;   // :115

// Pools are templated: the common case (a few iterations, < 80 faces) runs out of a small per-lane pool with
// one-byte links in SHARED memory; a pair that overflows it is redone by a second kernel with the large pool in
// global memory.
template <typename IdxT>
struct EpaFaceT {  // np/GjkEpaSolver.java:515-525, pointers replaced by pool indices (NIL = null)
    float nx, ny, nz, d;
    IdxT v[3];
    IdxT f[3];
    IdxT prev, next;
    uint8_t e[3];
    uint8_t pad;
    uint16_t mark;
};
struct EpaMkv { f3 w, r; };  // :119-127

template <typename IdxT, int MAXF, int MAXV, int MAXRAY, int MAXSTK>
struct EpaScratchT {
    typedef IdxT Idx;
    typedef EpaFaceT<IdxT> Face;
    static constexpr int kMaxF = MAXF, kMaxV = MAXV, kMaxRay = MAXRAY, kMaxStk = MAXSTK;
    EpaMkv mkv[MAXV];
    Face face[MAXF];
    f3 ray[MAXRAY];
    IdxT stkF[MAXSTK];
    uint8_t stkE[MAXSTK];
};
typedef EpaScratchT<int16_t, EPA_MAXF, EPA_MAXV, EPA_MAXRAY, EPA_MAXSTK> EpaScratch;   // large pool (global memory)
typedef EpaScratchT<uint8_t, 80, 28, 12, 24> EpaScratchSmall;                          // small pool (shared memory)
constexpr int EPA_SMALL_STRIDE = (int)sizeof(EpaScratchSmall) + 4;                        // odd word stride between lanes
typedef EpaScratchT<uint8_t, 120, 40, 24, 48> EpaScratchLocal;                         // medium pool (local memory)

// A convex shape chosen at run time (the EPA bin is small and mixed, so no template split here).
struct AnyS {
    int type;
    f3 h;             // box half extents (core)
    f3 ta, tb, tc;    // triangle
    const float4* pts;
    int n;
    float margin;
    __device__ __forceinline__ f3 support(f3 v) const {
        if (type == SH_BOX) { BoxS b; b.h = h; b.margin = margin; return b.support(v); }
        if (type == SH_SPHERE) return mk3(0.f, 0.f, 0.f);
        if (type == SH_HULL) { HullS s; s.pts = pts; s.n = n; s.margin = margin; return s.support(v); }
        TriS t; t.a = ta; t.b = tb; t.c = tc; t.margin = margin;
        return t.support(v);
    }
    __device__ __forceinline__ f3 supportMargin(f3 v) const {
        if (type == SH_BOX) { BoxS b; b.h = h; b.margin = margin; return b.supportMargin(v); }
        return addMarginDir(support(v), v, margin);
    }
};

template <class SC>
struct EpaCtx {
    typedef typename SC::Face Face;
    static constexpr int NIL = -1;
    const AnyS* A;
    const AnyS* B;
    Xf ta, tb;          // localTransA / localTransB (recentred)
    float margin;       // radialmargin + EPA_accuracy (:880)
    SC* s;
    // GJK part
    EpaMkv simplex[5];
    f3 ray;
    int order, iterations, nrays;
    bool failed;
    // EPA part
    int root, nfaces, nface_alloc, nmkv;
    bool overflow;

    __device__ f3 localSupport(f3 d, int i) const {  // :193-203
        if (i == 0) return add3(mulMV(ta.m, A->supportMargin(mulMtV(ta.m, d))), ta.o);
        return add3(mulMV(tb.m, B->supportMargin(mulMtV(tb.m, d))), tb.o);
    }
    __device__ void support(f3 d, EpaMkv& v) const {  // :205-220
        v.r = d;
        f3 t1 = localSupport(d, 0);
        f3 t2 = localSupport(neg3(d), 1);
        v.w = sub3(t1, t2);
        v.w.x += margin * d.x;
        v.w.y += margin * d.y;
        v.w.z += margin * d.z;
    }
    __device__ bool fetchSupport() {  // :222-241 (hash chain == list membership)
        for (int i = 0; i < nrays; i++)
            if (eq3bits(s->ray[i], ray)) { --order; return false; }
        if (nrays < SC::kMaxRay) s->ray[nrays++] = ray;
        else overflow = true;
        ++order;
        support(ray, simplex[order]);
        return dot3(ray, simplex[order].w) > 0;
    }
    __device__ bool solveSimplex2(f3 ao, f3 ab) {  // :243-260
        if (dot3(ab, ao) >= 0) {
            f3 cabo = crs3(ab, ao);
            if (len2_3(cabo) > EPA_SQINSIMPLEX_EPS) ray = crs3(cabo, ab);
            else return true;
        } else {
            order = 0;
            simplex[0] = simplex[1];
            ray = ao;
        }
        return false;
    }
    __device__ bool solveSimplex3a(f3 ao, f3 ab, f3 ac, f3 cabc) {  // :271-311
        f3 t = crs3(cabc, ab), t2 = crs3(cabc, ac);
        if (dot3(t, ao) < -EPA_INSIMPLEX_EPS) {
            order = 1;
            simplex[0] = simplex[1];
            simplex[1] = simplex[2];
            return solveSimplex2(ao, ab);
        } else if (dot3(t2, ao) > +EPA_INSIMPLEX_EPS) {
            order = 1;
            simplex[1] = simplex[2];
            return solveSimplex2(ao, ac);
        } else {
            float d = dot3(cabc, ao);
            if (fabsf(d) > EPA_INSIMPLEX_EPS) {
                if (d > 0) {
                    ray = cabc;
                } else {
                    ray = neg3(cabc);
                    EpaMkv sw = simplex[0];
                    simplex[0] = simplex[1];
                    simplex[1] = sw;
                }
                return false;
            }
            return true;
        }
    }
    __device__ bool solveSimplex4(f3 ao, f3 ab, f3 ac, f3 ad) {  // :313-352
        f3 t = crs3(ab, ac), t2 = crs3(ac, ad), t3 = crs3(ad, ab);
        if (dot3(t, ao) > EPA_INSIMPLEX_EPS) {
            order = 2;
            simplex[0] = simplex[1];
            simplex[1] = simplex[2];
            simplex[2] = simplex[3];
            return solveSimplex3a(ao, ab, ac, t);
        } else if (dot3(t2, ao) > EPA_INSIMPLEX_EPS) {
            order = 2;
            simplex[2] = simplex[3];
            return solveSimplex3a(ao, ac, ad, t2);
        } else if (dot3(t3, ao) > EPA_INSIMPLEX_EPS) {
            order = 2;
            simplex[1] = simplex[0];
            simplex[0] = simplex[2];
            simplex[2] = simplex[3];
            return solveSimplex3a(ao, ad, ab, t3);
        }
        return true;
    }
    __device__ bool searchOrigin() {  // :354-420
        iterations = 0;
        order = -1;
        failed = false;
        ray = nor3(mk3(1.f, 0.f, 0.f));
        nrays = 0;
        fetchSupport();
        ray = neg3(simplex[0].w);
        for (; iterations < EPA_GJK_MAXIT; ++iterations) {
            float rl = len3(ray);
            ray = scl3(ray, 1.f / (rl > 0.f ? rl : 1.f));
            if (fetchSupport()) {
                bool found = false;
                switch (order) {
                case 1:
                    found = solveSimplex2(neg3(simplex[1].w), sub3(simplex[0].w, simplex[1].w));
                    break;
                case 2:
                    found = solveSimplex3a(neg3(simplex[2].w), sub3(simplex[1].w, simplex[2].w), sub3(simplex[0].w, simplex[2].w),
                                           crs3(sub3(simplex[1].w, simplex[2].w), sub3(simplex[0].w, simplex[2].w)));
                    break;
                case 3:
                    found = solveSimplex4(neg3(simplex[3].w), sub3(simplex[2].w, simplex[3].w), sub3(simplex[1].w, simplex[3].w),
                                          sub3(simplex[0].w, simplex[3].w));
                    break;
                }
                if (found) return true;
            } else {
                return false;
            }
        }
        failed = true;
        return false;
    }
    __device__ bool encloseOrigin() {  // :422-498
        switch (order) {
        case 0:
            break;
        case 1: {
            f3 ab = sub3(simplex[1].w, simplex[0].w);
            f3 b0 = crs3(ab, mk3(1.f, 0.f, 0.f)), b1 = crs3(ab, mk3(0.f, 1.f, 0.f)), b2 = crs3(ab, mk3(0.f, 0.f, 1.f));
            float m0 = len2_3(b0), m1 = len2_3(b1), m2 = len2_3(b2);
            // QuaternionUtil.setRotation(q, nor(ab), 2pi/3) (lm/QuaternionUtil.java:41-46); the two constants
            // are (float)Math.sin((2pi/3)*0.5f) and (float)Math.cos((2pi/3)*0.5f)
            f3 axis = nor3(ab);
            float dl = len3(axis);
            float sn = __uint_as_float(0x3f5db3d8u) / dl;
            float qx = axis.x * sn, qy = axis.y * sn, qz = axis.z * sn, qw = __uint_as_float(0x3effffffu);
            // MatrixUtil.setRotation(r, q) (lm/MatrixUtil.java:318-335)
            float dd = qx * qx + qy * qy + qz * qz + qw * qw;
            float sc = 2.f / dd;
            float xs = qx * sc, ys = qy * sc, zs = qz * sc;
            float wx = qw * xs, wy = qw * ys, wz = qw * zs;
            float xx = qx * xs, xy = qx * ys, xz = qx * zs;
            float yy = qy * ys, yz = qy * zs, zz = qz * zs;
            float r[3][3];
            r[0][0] = 1.f - (yy + zz); r[0][1] = xy - wz; r[0][2] = xz + wy;
            r[1][0] = xy + wz; r[1][1] = 1.f - (xx + zz); r[1][2] = yz - wx;
            r[2][0] = xz - wy; r[2][1] = yz + wx; r[2][2] = 1.f - (xx + yy);
            f3 w = m0 > m1 ? (m0 > m2 ? b0 : b2) : (m1 > m2 ? b1 : b2);
            support(nor3(w), simplex[4]);
            w = mulMV(r, w);
            support(nor3(w), simplex[2]);
            w = mulMV(r, w);
            support(nor3(w), simplex[3]);
            order = 4;
            return true;
        }
        case 2: {
            f3 n = nor3(crs3(sub3(simplex[1].w, simplex[0].w), sub3(simplex[2].w, simplex[0].w)));
            support(n, simplex[3]);
            support(neg3(n), simplex[4]);
            order = 4;
            return true;
        }
        case 3:
        case 4:
            return true;
        }
        return false;
    }

    // ---- EPA -----------------------------------------------------------------------------------------
    typedef typename SC::Idx Idx;
    __device__ __forceinline__ static int rd(Idx x) { return (x == (Idx)-1) ? -1 : (int)x; }  // link -> index, NIL -> -1

    __device__ bool setFace(int fi, int a, int b, int c) {  // :604-631
        Face& f = s->face[fi];
        f3 aw = s->mkv[a].w, bw = s->mkv[b].w, cw = s->mkv[c].w;
        f3 nrm = crs3(sub3(bw, aw), sub3(cw, aw));
        float len = len3(nrm);
        f3 t1 = crs3(aw, bw), t2 = crs3(bw, cw), t3 = crs3(cw, aw);
        bool valid = (dot3(t1, nrm) >= -EPA_INFACE_EPS) && (dot3(t2, nrm) >= -EPA_INFACE_EPS) && (dot3(t3, nrm) >= -EPA_INFACE_EPS);
        f.v[0] = (Idx)a; f.v[1] = (Idx)b; f.v[2] = (Idx)c;
        f.mark = 0;
        f3 n = scl3(nrm, 1.f / (len > 0.f ? len : B2C_SIMD_INFINITY));
        f.nx = n.x; f.ny = n.y; f.nz = n.z;
        f.d = jmaxf(0.f, -dot3(n, aw));
        return valid;
    }
    __device__ int newFace(int a, int b, int c) {  // :633-647
        if (nface_alloc >= SC::kMaxF) { overflow = true; return -1; }
        int pf = nface_alloc++;
        Face& f = s->face[pf];
        f.f[0] = f.f[1] = f.f[2] = (Idx)-1;
        f.e[0] = f.e[1] = f.e[2] = 0;
        if (setFace(pf, a, b, c)) {
            if (root >= 0) s->face[root].prev = (Idx)pf;
            f.prev = (Idx)-1;
            f.next = (Idx)root;
            root = pf;
            ++nfaces;
        } else {
            f.prev = f.next = (Idx)-1;
        }
        return pf;
    }
    __device__ void detach(int fi) {  // :649-666
        Face& f = s->face[fi];
        const int fprev = rd(f.prev), fnext = rd(f.next);
        if (fprev >= 0 || fnext >= 0) {
            --nfaces;
            if (fi == root) {
                root = fnext;
                s->face[root].prev = (Idx)-1;
            } else {
                if (fnext < 0) {
                    s->face[fprev].next = (Idx)-1;
                } else {
                    s->face[fprev].next = (Idx)fnext;
                    s->face[fnext].prev = (Idx)fprev;
                }
            }
            f.prev = f.next = (Idx)-1;
        }
    }
    __device__ void link(int f0, int e0, int f1, int e1) {  // :668-673
        s->face[f0].f[e0] = (Idx)f1;
        s->face[f1].e[e1] = (uint8_t)e0;
        s->face[f1].f[e1] = (Idx)f0;
        s->face[f0].e[e0] = (uint8_t)e1;
    }
    // :683-706 BuildHorizon, recursion unrolled onto an explicit stack (children pushed in reverse so the
    // visiting order, and with it the cf/ff chaining, is the reference's depth-first order)
    __device__ int buildHorizon(int markid, int w, int f0, int e0, int& cf, int& ff) {
        int ne = 0, sp = 0;
        s->stkF[sp] = (Idx)f0; s->stkE[sp] = (uint8_t)e0; sp++;
        const f3 ww = s->mkv[w].w;
        while (sp > 0) {
            sp--;
            int fi = rd(s->stkF[sp]), e = s->stkE[sp];
            if (fi < 0) { overflow = true; continue; }
            Face& f = s->face[fi];
            if (f.mark == markid) continue;
            int e1 = (e + 1) % 3;
            if ((dot3(mk3(f.nx, f.ny, f.nz), ww) + f.d) > 0) {
                int nf = newFace(rd(f.v[e1]), rd(f.v[e]), w);
                if (nf < 0) return ne;
                link(nf, 0, fi, e);
                if (cf >= 0) link(cf, 1, nf, 2);
                else ff = nf;
                cf = nf;
                ne += 1;
            } else {
                int e2 = (e + 2) % 3;
                detach(fi);
                f.mark = (uint16_t)markid;
                if (sp + 2 > SC::kMaxStk) { overflow = true; return ne; }
                s->stkF[sp] = f.f[e2]; s->stkE[sp] = f.e[e2]; sp++;
                s->stkF[sp] = f.f[e1]; s->stkE[sp] = f.e[e1]; sp++;
            }
        }
        return ne;
    }
    __device__ f3 getCoordinates(int fi) const {  // :553-587
        const Face& f = s->face[fi];
        f3 o = scl3(mk3(f.nx, f.ny, f.nz), -f.d);
        f3 w0 = s->mkv[rd(f.v[0])].w, w1 = s->mkv[rd(f.v[1])].w, w2 = s->mkv[rd(f.v[2])].w;
        float a0 = len3(crs3(sub3(w0, o), sub3(w1, o)));
        float a1 = len3(crs3(sub3(w1, o), sub3(w2, o)));
        float a2 = len3(crs3(sub3(w2, o), sub3(w0, o)));
        float sm = a0 + a1 + a2;
        return scl3(mk3(a1, a2, a0), 1.f / (sm > 0.f ? sm : 1.f));
    }

    // :712-856 EvaluatePD.  Returns depth; nearest[0/1] and epaFailed are outputs.
    __device__ float evaluatePD(f3& near0, f3& near1, bool& epaFailed) {
        const int tetF[4][3] = {{2, 1, 0}, {3, 0, 1}, {3, 1, 2}, {3, 2, 0}};
        const int tetE[6][4] = {{0, 0, 2, 1}, {0, 1, 1, 1}, {0, 2, 3, 1}, {1, 0, 3, 2}, {2, 0, 1, 2}, {3, 0, 2, 2}};
        const int hexF[6][3] = {{2, 0, 4}, {4, 1, 2}, {1, 4, 0}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
        const int hexE[9][4] = {{0, 0, 4, 0}, {0, 1, 2, 1}, {0, 2, 1, 2}, {1, 1, 5, 2}, {1, 0, 2, 0},
                                {2, 2, 3, 2}, {3, 1, 5, 0}, {3, 0, 4, 2}, {5, 1, 4, 1}};
        int bestface = -1;
        int markid = 1;
        float depth = -B2C_SIMD_INFINITY;
        root = -1; nfaces = 0; nface_alloc = 0; nmkv = 0;
        int iters = 0;
        epaFailed = false;
        if (encloseOrigin()) {
            int basefaces[6];
            int nfidx = 0, neidx = 0;
            bool tet = (order == 3);
            if (order == 3) { nfidx = 4; neidx = 6; }
            else if (order == 4) { nfidx = 6; neidx = 9; }
            for (int i = 0; i <= order; ++i) s->mkv[nmkv++] = simplex[i];
            for (int i = 0; i < nfidx; ++i)
                basefaces[i] = tet ? newFace(tetF[i][0], tetF[i][1], tetF[i][2]) : newFace(hexF[i][0], hexF[i][1], hexF[i][2]);
            for (int i = 0; i < neidx; ++i) {
                if (tet) link(basefaces[tetE[i][0]], tetE[i][1], basefaces[tetE[i][2]], tetE[i][3]);
                else link(basefaces[hexE[i][0]], hexE[i][1], basefaces[hexE[i][2]], hexE[i][3]);
            }
        }
        if (0 == nfaces) return depth;
        for (; iters < EPA_MAXIT; ++iters) {
            // FindBest :589-602
            int bf = -1;
            {
                float bd = B2C_SIMD_INFINITY;
                for (int cf = root; cf >= 0; cf = rd(s->face[cf].next))
                    if (s->face[cf].d < bd) { bd = s->face[cf].d; bf = cf; }
            }
            if (bf < 0) break;
            if (nmkv >= SC::kMaxV) { overflow = true; break; }
            int w = nmkv++;
            f3 bn = mk3(s->face[bf].nx, s->face[bf].ny, s->face[bf].nz);
            support(neg3(bn), s->mkv[w]);
            float d = dot3(bn, s->mkv[w].w) + s->face[bf].d;
            bestface = bf;
            if (d < -EPA_ACCURACY) {
                int cf = -1, ff = -1, nf = 0;
                detach(bf);
                s->face[bf].mark = (uint16_t)(++markid);
                for (int i = 0; i < 3 && !overflow; ++i)
                    nf += buildHorizon(markid, w, rd(s->face[bf].f[i]), s->face[bf].e[i], cf, ff);
                if (overflow) break;
                if (nf <= 2) break;
                link(cf, 1, ff, 2);
            } else {
                break;
            }
        }
        if (overflow) return -B2C_SIMD_INFINITY;
        if (bestface >= 0) {
            f3 b = getCoordinates(bestface);
            const Face& f = s->face[bestface];
            depth = jmaxf(0.f, f.d);
            f3 fa[3], fb[3];
            for (int j = 0; j < 3; ++j) {
                f3 r = s->mkv[rd(f.v[j])].r;
                fa[j] = localSupport(scl3(r, 1.f), 0);
                fb[j] = localSupport(scl3(r, -1.f), 1);
            }
            f3 t1 = scl3(fa[0], b.x), t2 = scl3(fa[1], b.y), t3 = scl3(fa[2], b.z);
            near0 = mk3(t1.x + t2.x + t3.x, t1.y + t2.y + t3.y, t1.z + t2.z + t3.z);
            t1 = scl3(fb[0], b.x); t2 = scl3(fb[1], b.y); t3 = scl3(fb[2], b.z);
            near1 = mk3(t1.x + t2.x + t3.x, t1.y + t2.y + t3.y, t1.z + t2.z + t3.z);
        } else {
            epaFailed = true;
        }
        return depth;
    }
};

// np/GjkEpaSolver.java:864-911 collide + np/GjkEpaPenetrationDepthSolver.java:41-63 calcPenDepth.
// Returns true with witnesses when penetrating.  poolOverflow: the pool was too small, nothing is decided.
template <class SC>
__device__ __noinline__ bool epaPenetration(const AnyS& A, const AnyS& B, const Xf& la, const Xf& lb, SC* scratch, f3& wOnA,
                                            f3& wOnB, bool& epaFailed, bool& poolOverflow) {
    EpaCtx<SC> c;
    c.A = &A; c.B = &B; c.ta = la; c.tb = lb;
    c.margin = 0.f + EPA_ACCURACY;
    c.s = scratch;
    c.overflow = false;
    epaFailed = false;
    poolOverflow = false;
    bool collide = c.searchOrigin();
    if (c.overflow) { poolOverflow = true; return false; }
    if (collide) {
        f3 n0, n1;
        float pd = c.evaluatePD(n0, n1, epaFailed);
        if (c.overflow) { poolOverflow = true; return false; }
        if (pd > 0) {
            wOnA = n0;
            wOnB = n1;
            return true;
        }
    }
    return false;
}

}  // namespace b2c
