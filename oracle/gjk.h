// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// gjk.h — restatement of np/GjkPairDetector.java, np/GjkEpaSolver.java and
// np/GjkEpaPenetrationDepthSolver.java (np/ = collision/narrowphase/).
#pragma once
#include <vector>
#include "jmath.h"
#include "shapes.h"
#include "voronoi.h"

namespace orc {

struct EpaDebugStats { long calls = 0, epaIters = 0, epaItersMax = 0, faces = 0, facesMax = 0, gjkIters = 0, gjkItersMax = 0, hist[9][4] = {}; int curType = 0; };
static EpaDebugStats g_epaStats;

// ---------------------------------------------------------------------------------------
// np/GjkEpaSolver.java
// ---------------------------------------------------------------------------------------
struct EpaResults {
    int status = 0;  // 0 Separated, 1 Penetrating, 2 GJK_Failed, 3 EPA_Failed (:87-92)
    V3 witnesses[2];
    V3 normal;
    float depth = 0;
    int epa_iterations = 0, gjk_iterations = 0;
};

struct GjkEpaSolver {
    static constexpr int GJK_maxiterations = 128;        // :108
    static constexpr float GJK_insimplex_eps = 0.0001f;  // :111
    static constexpr float GJK_sqinsimplex_eps = GJK_insimplex_eps * GJK_insimplex_eps;
    static constexpr int EPA_maxiterations = 256;        // :113
    static constexpr float EPA_inface_eps = 0.01f;       // :114
    static constexpr float EPA_accuracy = 0.001f;        // :115

    struct Mkv { V3 w, r; };  // :119-127

    struct GJK {
        // The reference keeps visited rays in a 64-bucket chained hash (:188-191, :222-241);
        // membership in it is exactly "bitwise-equal ray seen before", restated as a list.
        std::vector<V3> seen;
        M3 wrot[2];
        V3 pos[2];
        const Shape* shapes[2];
        Mkv simplex[5];
        V3 ray;
        int order = 0, iterations = 0;
        float margin = 0;
        bool failed = false;

        void init(const M3& r0, const V3& p0, const Shape* s0, const M3& r1, const V3& p1, const Shape* s1, float pm) {
            wrot[0].set(r0); pos[0].set(p0); shapes[0] = s0;
            wrot[1].set(r1); pos[1].set(p1); shapes[1] = s1;
            margin = pm;
            failed = false;
        }
        V3 LocalSupport(const V3& d, int i) const {  // :193-203
            V3 tmp, out;
            transposeTransform(tmp, d, wrot[i]);
            localGetSupportingVertex(*shapes[i], tmp, out);
            v3mul(out, wrot[i]);
            out.add(pos[i]);
            return out;
        }
        void Support(const V3& d, Mkv& v) const {  // :205-220
            v.r.set(d);
            V3 tmp1 = LocalSupport(d, 0);
            V3 tmp; tmp.set(d); tmp.scl(-1.0f);
            V3 tmp2 = LocalSupport(tmp, 1);
            v.w.set(tmp1).sub(tmp2);
            v.w.x += margin * d.x;
            v.w.y += margin * d.y;
            v.w.z += margin * d.z;
        }
        bool FetchSupport() {  // :222-241
            for (size_t i = 0; i < seen.size(); i++)
                if (seen[i].equals(ray)) { --order; return false; }
            seen.push_back(ray);
            Support(ray, simplex[++order]);
            return ray.dot(simplex[order].w) > 0;
        }
        bool SolveSimplex2(const V3& ao, const V3& ab) {  // :243-260
            if (ab.dot(ao) >= 0) {
                V3 cabo; cabo.set(ab).crs(ao);
                if (cabo.len2() > GJK_sqinsimplex_eps) {
                    ray.set(cabo).crs(ab);
                } else {
                    return true;
                }
            } else {
                order = 0;
                simplex[0] = simplex[1];
                ray.set(ao);
            }
            return false;
        }
        bool SolveSimplex3(const V3& ao, const V3& ab, const V3& ac) {  // :262-269
            V3 tmp; tmp.set(ab).crs(ac);
            return SolveSimplex3a(ao, ab, ac, tmp);
        }
        bool SolveSimplex3a(const V3& ao, const V3& ab, const V3& ac, const V3& cabc) {  // :271-311
            V3 tmp; tmp.set(cabc).crs(ab);
            V3 tmp2; tmp2.set(cabc).crs(ac);
            bool result;
            if (tmp.dot(ao) < -GJK_insimplex_eps) {
                order = 1;
                simplex[0] = simplex[1];
                simplex[1] = simplex[2];
                result = SolveSimplex2(ao, ab);
            } else if (tmp2.dot(ao) > +GJK_insimplex_eps) {
                order = 1;
                simplex[1] = simplex[2];
                result = SolveSimplex2(ao, ac);
            } else {
                float d = cabc.dot(ao);
                if (jabsf(d) > GJK_insimplex_eps) {
                    if (d > 0) {
                        ray.set(cabc);
                    } else {
                        ray.set(cabc).scl(-1.0f);
                        Mkv swapTmp = simplex[0];
                        simplex[0] = simplex[1];
                        simplex[1] = swapTmp;
                    }
                    result = false;
                } else {
                    result = true;
                }
            }
            return result;
        }
        bool SolveSimplex4(const V3& ao, const V3& ab, const V3& ac, const V3& ad) {  // :313-352
            V3 crs;
            V3 tmp; tmp.set(ab).crs(ac);
            V3 tmp2; tmp2.set(ac).crs(ad);
            V3 tmp3; tmp3.set(ad).crs(ab);
            bool result;
            if (tmp.dot(ao) > GJK_insimplex_eps) {
                crs.set(tmp);
                order = 2;
                simplex[0] = simplex[1];
                simplex[1] = simplex[2];
                simplex[2] = simplex[3];
                result = SolveSimplex3a(ao, ab, ac, crs);
            } else if (tmp2.dot(ao) > GJK_insimplex_eps) {
                crs.set(tmp2);
                order = 2;
                simplex[2] = simplex[3];
                result = SolveSimplex3a(ao, ac, ad, crs);
            } else if (tmp3.dot(ao) > GJK_insimplex_eps) {
                crs.set(tmp3);
                order = 2;
                simplex[1] = simplex[0];
                simplex[0] = simplex[2];
                simplex[2] = simplex[3];
                result = SolveSimplex3a(ao, ad, ab, crs);
            } else {
                result = true;
            }
            return result;
        }
        bool SearchOrigin() {  // :354-420, initray (1,0,0)
            V3 tmp1, tmp2, tmp3, tmp4;
            iterations = 0;
            order = -1;
            failed = false;
            ray.set(1, 0, 0);
            ray.nor();
            seen.clear();
            FetchSupport();
            ray.set(simplex[0].w).scl(-1.0f);
            for (; iterations < GJK_maxiterations; ++iterations) {
                float rl = ray.len();
                ray.scl(1.0f / (rl > 0.0f ? rl : 1.0f));
                if (FetchSupport()) {
                    bool found = false;
                    switch (order) {
                    case 1:
                        tmp1.set(simplex[1].w).scl(-1.0f);
                        tmp2.set(simplex[0].w).sub(simplex[1].w);
                        found = SolveSimplex2(tmp1, tmp2);
                        break;
                    case 2:
                        tmp1.set(simplex[2].w).scl(-1.0f);
                        tmp2.set(simplex[1].w).sub(simplex[2].w);
                        tmp3.set(simplex[0].w).sub(simplex[2].w);
                        found = SolveSimplex3(tmp1, tmp2, tmp3);
                        break;
                    case 3:
                        tmp1.set(simplex[3].w).scl(-1.0f);
                        tmp2.set(simplex[2].w).sub(simplex[3].w);
                        tmp3.set(simplex[1].w).sub(simplex[3].w);
                        tmp4.set(simplex[0].w).sub(simplex[3].w);
                        found = SolveSimplex4(tmp1, tmp2, tmp3, tmp4);
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
        bool EncloseOrigin() {  // :422-498
            V3 tmp, tmp1, tmp2;
            switch (order) {
            case 0:
                break;
            case 1: {
                V3 ab; ab.set(simplex[1].w).sub(simplex[0].w);
                V3 b[3] = {V3(1, 0, 0), V3(0, 1, 0), V3(0, 0, 1)};
                for (int k = 0; k < 3; k++) { V3 t = ab; t.crs(b[k]); b[k] = t; }
                float m[3] = {b[0].len2(), b[1].len2(), b[2].len2()};
                Quat q;
                tmp.set(ab).nor();
                quatSetRotation(q, tmp, SIMD_2_PI_ / 3.0f);
                M3 r;
                setRotation(r, q);
                V3 w;
                w.set(b[m[0] > m[1] ? (m[0] > m[2] ? 0 : 2) : (m[1] > m[2] ? 1 : 2)]);
                tmp.set(w).nor();
                Support(tmp, simplex[4]);
                v3mul(w, r);
                tmp.set(w).nor();
                Support(tmp, simplex[2]);
                v3mul(w, r);
                tmp.set(w).nor();
                Support(tmp, simplex[3]);
                v3mul(w, r);
                order = 4;
                return true;
            }
            case 2: {
                tmp1.set(simplex[1].w).sub(simplex[0].w);
                tmp2.set(simplex[2].w).sub(simplex[0].w);
                V3 n; n.set(tmp1).crs(tmp2);
                n.nor();
                Support(n, simplex[3]);
                tmp.set(n).scl(-1.0f);
                Support(tmp, simplex[4]);
                order = 4;
                return true;
            }
            case 3:
                return true;
            case 4:
                return true;
            }
            return false;
        }
    };

    struct Face {  // :515-525
        Mkv* v[3];
        Face* f[3];
        int e[3];
        V3 n;
        float d;
        int mark;
        Face* prev;
        Face* next;
    };

    struct EPA {
        GJK* gjk;
        Face* root = nullptr;
        int nfaces = 0, iterations = 0;
        V3 features[2][3];
        V3 nearest[2];
        V3 normal;
        float depth = 0;
        bool failed = false;
        // object pools (the reference draws these from ObjectStackList; addresses must be stable)
        std::vector<Face*> facePool;
        std::vector<Mkv*> mkvPool;
        ~EPA() {
            for (Face* f : facePool) delete f;
            for (Mkv* m : mkvPool) delete m;
        }
        Face* allocFace() {
            Face* f = new Face();
            f->v[0] = f->v[1] = f->v[2] = nullptr;
            f->f[0] = f->f[1] = f->f[2] = nullptr;
            f->e[0] = f->e[1] = f->e[2] = 0;
            f->d = 0; f->mark = 0; f->prev = f->next = nullptr;
            facePool.push_back(f);
            return f;
        }
        Mkv* allocMkv() { Mkv* m = new Mkv(); mkvPool.push_back(m); return m; }

        V3 GetCoordinates(const Face* face) const {  // :553-587
            V3 tmp, tmp1, tmp2, o;
            o.set(face->n).scl(-face->d);
            float a[3];
            tmp1.set(face->v[0]->w).sub(o);
            tmp2.set(face->v[1]->w).sub(o);
            tmp.set(tmp1).crs(tmp2);
            a[0] = tmp.len();
            tmp1.set(face->v[1]->w).sub(o);
            tmp2.set(face->v[2]->w).sub(o);
            tmp.set(tmp1).crs(tmp2);
            a[1] = tmp.len();
            tmp1.set(face->v[2]->w).sub(o);
            tmp2.set(face->v[0]->w).sub(o);
            tmp.set(tmp1).crs(tmp2);
            a[2] = tmp.len();
            float sm = a[0] + a[1] + a[2];
            V3 out(a[1], a[2], a[0]);
            out.scl(1.0f / (sm > 0.0f ? sm : 1.0f));
            return out;
        }
        Face* FindBest() const {  // :589-602
            Face* bf = nullptr;
            if (root) {
                Face* cf = root;
                float bd = SIMD_INFINITY_;
                do {
                    if (cf->d < bd) { bd = cf->d; bf = cf; }
                } while (nullptr != (cf = cf->next));
            }
            return bf;
        }
        bool Set(Face* f, Mkv* a, Mkv* b, Mkv* c) const {  // :604-631
            V3 tmp1, tmp2, tmp3, nrm;
            tmp1.set(b->w).sub(a->w);
            tmp2.set(c->w).sub(a->w);
            nrm.set(tmp1).crs(tmp2);
            float len = nrm.len();
            tmp1.set(a->w).crs(b->w);
            tmp2.set(b->w).crs(c->w);
            tmp3.set(c->w).crs(a->w);
            bool valid = (tmp1.dot(nrm) >= -EPA_inface_eps) && (tmp2.dot(nrm) >= -EPA_inface_eps) &&
                         (tmp3.dot(nrm) >= -EPA_inface_eps);
            f->v[0] = a; f->v[1] = b; f->v[2] = c;
            f->mark = 0;
            f->n.set(nrm).scl(1.0f / (len > 0.0f ? len : SIMD_INFINITY_));
            f->d = jmaxf(0.0f, -f->n.dot(a->w));
            return valid;
        }
        Face* NewFace(Mkv* a, Mkv* b, Mkv* c) {  // :633-647
            Face* pf = allocFace();
            if (Set(pf, a, b, c)) {
                if (root) root->prev = pf;
                pf->prev = nullptr;
                pf->next = root;
                root = pf;
                ++nfaces;
            } else {
                pf->prev = pf->next = nullptr;
            }
            return pf;
        }
        void Detach(Face* face) {  // :649-666
            if (face->prev != nullptr || face->next != nullptr) {
                --nfaces;
                if (face == root) {
                    root = face->next;
                    root->prev = nullptr;
                } else {
                    if (face->next == nullptr) {
                        face->prev->next = nullptr;
                    } else {
                        face->prev->next = face->next;
                        face->next->prev = face->prev;
                    }
                }
                face->prev = face->next = nullptr;
            }
        }
        static void Link(Face* f0, int e0, Face* f1, int e1) {  // :668-673
            f0->f[e0] = f1;
            f1->e[e1] = e0;
            f1->f[e1] = f0;
            f0->e[e0] = e1;
        }
        Mkv* Support(const V3& w) {  // :675-680
            Mkv* v = allocMkv();
            gjk->Support(w, *v);
            return v;
        }
        int BuildHorizon(int markid, Mkv* w, Face* f, int e, Face** cf, Face** ff) {  // :683-706
            static const int mod3[5] = {0, 1, 2, 0, 1};
            int ne = 0;
            if (f->mark != markid) {
                int e1 = mod3[e + 1];
                if ((f->n.dot(w->w) + f->d) > 0) {
                    Face* nf = NewFace(f->v[e1], f->v[e], w);
                    Link(nf, 0, f, e);
                    if (*cf != nullptr) Link(*cf, 1, nf, 2);
                    else *ff = nf;
                    *cf = nf;
                    ne = 1;
                } else {
                    int e2 = mod3[e + 2];
                    Detach(f);
                    f->mark = markid;
                    ne += BuildHorizon(markid, w, f->f[e1], f->e[e1], cf, ff);
                    ne += BuildHorizon(markid, w, f->f[e2], f->e[e2], cf, ff);
                }
            }
            return ne;
        }
        float EvaluatePD(float accuracy = EPA_accuracy) {  // :712-856
            static const int tetrahedron_fidx[4][3] = {{2, 1, 0}, {3, 0, 1}, {3, 1, 2}, {3, 2, 0}};
            static const int tetrahedron_eidx[6][4] = {{0, 0, 2, 1}, {0, 1, 1, 1}, {0, 2, 3, 1},
                                                       {1, 0, 3, 2}, {2, 0, 1, 2}, {3, 0, 2, 2}};
            static const int hexahedron_fidx[6][3] = {{2, 0, 4}, {4, 1, 2}, {1, 4, 0}, {0, 3, 1}, {0, 2, 3}, {1, 3, 2}};
            static const int hexahedron_eidx[9][4] = {{0, 0, 4, 0}, {0, 1, 2, 1}, {0, 2, 1, 2}, {1, 1, 5, 2}, {1, 0, 2, 0},
                                                      {2, 2, 3, 2}, {3, 1, 5, 0}, {3, 0, 4, 2}, {5, 1, 4, 1}};
            V3 tmp;
            Face* bestface = nullptr;
            int markid = 1;
            depth = -SIMD_INFINITY_;
            normal.set(0, 0, 0);
            root = nullptr;
            nfaces = 0;
            iterations = 0;
            failed = false;
            if (gjk->EncloseOrigin()) {
                const int (*pfidx)[3] = nullptr;
                int nfidx = 0;
                const int (*peidx)[4] = nullptr;
                int neidx = 0;
                Mkv* basemkv[5];
                Face* basefaces[6];
                switch (gjk->order) {
                case 3: pfidx = tetrahedron_fidx; nfidx = 4; peidx = tetrahedron_eidx; neidx = 6; break;
                case 4: pfidx = hexahedron_fidx; nfidx = 6; peidx = hexahedron_eidx; neidx = 9; break;
                }
                for (int i = 0; i <= gjk->order; ++i) {
                    basemkv[i] = allocMkv();
                    *basemkv[i] = gjk->simplex[i];
                }
                for (int i = 0; i < nfidx; ++i)
                    basefaces[i] = NewFace(basemkv[pfidx[i][0]], basemkv[pfidx[i][1]], basemkv[pfidx[i][2]]);
                for (int i = 0; i < neidx; ++i)
                    Link(basefaces[peidx[i][0]], peidx[i][1], basefaces[peidx[i][2]], peidx[i][3]);
            }
            if (0 == nfaces) return depth;
            for (; iterations < EPA_maxiterations; ++iterations) {
                Face* bf = FindBest();
                if (bf != nullptr) {
                    tmp.set(bf->n).scl(-1.0f);
                    Mkv* w = Support(tmp);
                    float d = bf->n.dot(w->w) + bf->d;
                    bestface = bf;
                    if (d < -accuracy) {
                        Face* cf = nullptr;
                        Face* ff = nullptr;
                        int nf = 0;
                        Detach(bf);
                        bf->mark = ++markid;
                        for (int i = 0; i < 3; ++i) nf += BuildHorizon(markid, w, bf->f[i], bf->e[i], &cf, &ff);
                        if (nf <= 2) break;
                        Link(cf, 1, ff, 2);
                    } else {
                        break;
                    }
                } else {
                    break;
                }
            }
            if (bestface != nullptr) {
                V3 b = GetCoordinates(bestface);
                normal.set(bestface->n);
                depth = jmaxf(0.0f, bestface->d);
                for (int i = 0; i < 2; ++i) {
                    float s = i != 0 ? -1.0f : 1.0f;
                    for (int j = 0; j < 3; ++j) {
                        tmp.set(bestface->v[j]->r).scl(s);
                        features[i][j] = gjk->LocalSupport(tmp, i);
                    }
                }
                V3 tmp1, tmp2, tmp3;
                for (int i = 0; i < 2; i++) {
                    tmp1.set(features[i][0]).scl(b.x);
                    tmp2.set(features[i][1]).scl(b.y);
                    tmp3.set(features[i][2]).scl(b.z);
                    nearest[i].set(tmp1.x + tmp2.x + tmp3.x, tmp1.y + tmp2.y + tmp3.y, tmp1.z + tmp2.z + tmp3.z);
                }
            } else {
                failed = true;
            }
            return depth;
        }
    };

    // :864-911
    static bool collide(const Shape* shape0, const Xf& wtrs0, const Shape* shape1, const Xf& wtrs1, float radialmargin,
                        EpaResults& results) {
        results.witnesses[0].set(0, 0, 0);
        results.witnesses[1].set(0, 0, 0);
        results.normal.set(0, 0, 0);
        results.depth = 0;
        results.status = 0;
        results.epa_iterations = 0;
        results.gjk_iterations = 0;
        GJK gjk;
        gjk.init(wtrs0.basis, wtrs0.origin, shape0, wtrs1.basis, wtrs1.origin, shape1, radialmargin + EPA_accuracy);
        bool col = gjk.SearchOrigin();
        results.gjk_iterations = gjk.iterations + 1;
        g_epaStats.calls++;
        g_epaStats.gjkIters += gjk.iterations + 1;
        if (gjk.iterations + 1 > g_epaStats.gjkItersMax) g_epaStats.gjkItersMax = gjk.iterations + 1;
        g_epaStats.curType = (shape0->type % 3) * 3 + (shape1->type % 3);
        if (col) {
            EPA epa;
            epa.gjk = &gjk;
            float pd = epa.EvaluatePD();
            results.epa_iterations = epa.iterations + 1;
            g_epaStats.epaIters += epa.iterations + 1;
            if (epa.iterations + 1 > g_epaStats.epaItersMax) g_epaStats.epaItersMax = epa.iterations + 1;
            g_epaStats.faces += (long)epa.facePool.size();
            if ((long)epa.facePool.size() > g_epaStats.facesMax) g_epaStats.facesMax = (long)epa.facePool.size();
            { int it = epa.iterations + 1; int b = it <= 4 ? 0 : (it <= 16 ? 1 : (it <= 64 ? 2 : 3)); g_epaStats.hist[g_epaStats.curType][b]++; }
            if (pd > 0) {
                results.status = 1;
                results.normal.set(epa.normal);
                results.depth = pd;
                results.witnesses[0].set(epa.nearest[0]);
                results.witnesses[1].set(epa.nearest[1]);
                return true;
            } else {
                if (epa.failed) results.status = 3;
            }
        } else {
            if (gjk.failed) results.status = 2;
        }
        return false;
    }
};

// np/GjkEpaPenetrationDepthSolver.java:41-63
static inline bool calcPenDepth(const Shape* a, const Shape* b, const Xf& ta, const Xf& tb, V3& wOnA, V3& wOnB,
                                int* epaStatus = nullptr) {
    float radialmargin = 0.0f;
    EpaResults results;
    bool ok = GjkEpaSolver::collide(a, ta, b, tb, radialmargin, results);
    if (epaStatus) *epaStatus = results.status;
    if (ok) {
        wOnA.set(results.witnesses[0]);
        wOnB.set(results.witnesses[1]);
        return true;
    }
    return false;
}

// ---------------------------------------------------------------------------------------
// np/GjkPairDetector.java:73-316 getClosestPoints
// ---------------------------------------------------------------------------------------
struct GjkOut {
    bool hasContact = false;
    V3 normalOnBInWorld, pointInWorld;
    float depth = 0;
    int lastUsedMethod = -1, curIter = 0, degenerateSimplex = 0;
    int deepPenetrationChecks = 0;  // BulletStats.gNumDeepPenetrationChecks contribution (:270)
};

// usePenetrationSolver = false: GjkPairDetector.init(..., penetrationDepthSolver = null) as np/GjkConvexCast.java:107 does
static inline void gjkGetClosestPoints(const Shape* minkowskiA, const Shape* minkowskiB, const Xf& transformA,
                                       const Xf& transformB, float maximumDistanceSquared, GjkOut& out,
                                       bool usePenetrationSolver = true) {
    static const float REL_ERROR2 = 1.0e-6f;  // :44
    V3 tmp;
    float distance = 0.0f;
    V3 normalInB(0, 0, 0);
    V3 pointOnA, pointOnB;
    Xf localTransA; localTransA.set(transformA);
    Xf localTransB; localTransB.set(transformB);
    V3 positionOffset;
    positionOffset.set(localTransA.origin).add(localTransB.origin);
    positionOffset.scl(0.5f);
    localTransA.origin.sub(positionOffset);
    localTransB.origin.sub(positionOffset);

    float marginA = minkowskiA->getMargin();
    float marginB = minkowskiB->getMargin();

    int curIter = 0;
    int gGjkMaxIter = 1000;
    V3 cachedSeparatingAxis(0, 1, 0);

    bool isValid = false;
    bool checkSimplex = false;
    bool checkPenetration = true;
    int degenerateSimplex = 0;
    int lastUsedMethod = -1;
    VoronoiSimplexSolver simplexSolver;
    {
        float squaredDistance = SIMD_INFINITY_;
        float delta = 0.0f;
        float margin = marginA + marginB;
        simplexSolver.reset();
        V3 sepAxisInA, sepAxisInB, pInA, qInB, pWorld, qWorld, w;
        V3 tmpPointOnA, tmpPointOnB, tmpNormalInB;
        for (;;) {
            sepAxisInA.set(cachedSeparatingAxis).scl(-1.0f);
            transposeTransform(sepAxisInA, sepAxisInA, transformA.basis);
            sepAxisInB.set(cachedSeparatingAxis);
            transposeTransform(sepAxisInB, sepAxisInB, transformB.basis);

            localGetSupportingVertexWithoutMargin(*minkowskiA, sepAxisInA, pInA);
            localGetSupportingVertexWithoutMargin(*minkowskiB, sepAxisInB, qInB);

            pWorld.set(pInA);
            localTransA.transform(pWorld);
            qWorld.set(qInB);
            localTransB.transform(qWorld);

            w.set(pWorld).sub(qWorld);
            delta = cachedSeparatingAxis.dot(w);

            if ((delta > 0.0f) && (delta * delta > squaredDistance * maximumDistanceSquared)) {
                checkPenetration = false;
                break;
            }
            if (simplexSolver.inSimplex(w)) {
                degenerateSimplex = 1;
                checkSimplex = true;
                break;
            }
            float f0 = squaredDistance - delta;
            float f1 = squaredDistance * REL_ERROR2;
            if (f0 <= f1) {
                if (f0 <= 0.0f) degenerateSimplex = 2;
                checkSimplex = true;
                break;
            }
            simplexSolver.addVertex(w, pWorld, qWorld);
            if (!simplexSolver.closest(cachedSeparatingAxis)) {
                degenerateSimplex = 3;
                checkSimplex = true;
                break;
            }
            if (cachedSeparatingAxis.len2() < REL_ERROR2) {
                degenerateSimplex = 6;
                checkSimplex = true;
                break;
            }
            float previousSquaredDistance = squaredDistance;
            squaredDistance = cachedSeparatingAxis.len2();
            if (previousSquaredDistance - squaredDistance <= FLT_EPSILON_ * previousSquaredDistance) {
                simplexSolver.backup_closest(cachedSeparatingAxis);
                checkSimplex = true;
                break;
            }
            if (curIter++ > gGjkMaxIter) break;
            bool check = !simplexSolver.fullSimplex();
            if (!check) {
                simplexSolver.backup_closest(cachedSeparatingAxis);
                break;
            }
        }

        if (checkSimplex) {
            simplexSolver.compute_points(pointOnA, pointOnB);
            normalInB.set(pointOnA).sub(pointOnB);
            float lenSqr = cachedSeparatingAxis.len2();
            if (lenSqr < 0.0001f) degenerateSimplex = 5;
            if (lenSqr > FLT_EPSILON_ * FLT_EPSILON_) {
                float rlen = 1.0f / jsqrt(lenSqr);
                normalInB.scl(rlen);
                float s = jsqrt(squaredDistance);
                tmp.set(cachedSeparatingAxis).scl(marginA / s);
                pointOnA.sub(tmp);
                tmp.set(cachedSeparatingAxis).scl(marginB / s);
                pointOnB.add(tmp);
                distance = ((1.0f / rlen) - margin);
                isValid = true;
                lastUsedMethod = 1;
            } else {
                lastUsedMethod = 2;
            }
        }

        bool catchDegeneratePenetrationCase = (degenerateSimplex != 0 && ((distance + margin) < 0.01f));
        if (usePenetrationSolver && checkPenetration && (!isValid || catchDegeneratePenetrationCase)) {  // :266-269
            out.deepPenetrationChecks++;
            bool isValid2 = calcPenDepth(minkowskiA, minkowskiB, localTransA, localTransB, tmpPointOnA, tmpPointOnB);
            if (isValid2) {
                tmpNormalInB.set(tmpPointOnB).sub(tmpPointOnA);
                float lenSqr = tmpNormalInB.len2();
                if (lenSqr > (FLT_EPSILON_ * FLT_EPSILON_)) {
                    tmpNormalInB.scl(1.0f / jsqrt(lenSqr));
                    tmp.set(tmpPointOnA).sub(tmpPointOnB);
                    float distance2 = -tmp.len();
                    if (!isValid || (distance2 < distance)) {
                        distance = distance2;
                        pointOnA.set(tmpPointOnA);
                        pointOnB.set(tmpPointOnB);
                        normalInB.set(tmpNormalInB);
                        isValid = true;
                        lastUsedMethod = 3;
                    }
                } else {
                    lastUsedMethod = 4;
                }
            } else {
                lastUsedMethod = 5;
            }
        }
    }
    out.lastUsedMethod = lastUsedMethod;
    out.curIter = curIter;
    out.degenerateSimplex = degenerateSimplex;
    if (isValid) {
        tmp.set(pointOnB).add(positionOffset);
        out.hasContact = true;
        out.normalOnBInWorld.set(normalInB);
        out.pointInWorld.set(tmp);
        out.depth = distance;
    }
}

}  // namespace orc
