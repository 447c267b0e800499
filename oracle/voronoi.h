// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// voronoi.h — restatement of np/VoronoiSimplexSolver.java (np/ = collision/narrowphase/).
#pragma once
#include "jmath.h"

namespace orc {

struct UsageBitfield {  // np/VoronoiSimplexSolver.java:635-647
    bool a = false, b = false, c = false, d = false;
    void reset() { a = b = c = d = false; }
};

struct SubSimplexClosestResult {  // np/VoronoiSimplexSolver.java:649-676
    V3 closestPointOnSimplex;
    UsageBitfield used;
    float bary[4] = {0, 0, 0, 0};
    bool degenerate = false;
    void reset() { degenerate = false; setBary(0, 0, 0, 0); used.reset(); }
    bool isValid() const { return bary[0] >= 0.0f && bary[1] >= 0.0f && bary[2] >= 0.0f && bary[3] >= 0.0f; }
    void setBary(float a, float b, float c, float d) { bary[0] = a; bary[1] = b; bary[2] = c; bary[3] = d; }
};

struct VoronoiSimplexSolver {
    int numVertices = 0;
    V3 W[5], P[5], Q[5];
    V3 cachedP1, cachedP2, cachedV, lastW;
    bool cachedValidClosest = false;
    SubSimplexClosestResult cachedBC;
    bool needsUpdate = true;

    void removeVertex(int index) {  // :79-85
        numVertices--;
        W[index].set(W[numVertices]);
        P[index].set(P[numVertices]);
        Q[index].set(Q[numVertices]);
    }
    void reduceVertices(const UsageBitfield& u) {  // :87-95
        if (numVertices >= 4 && !u.d) removeVertex(3);
        if (numVertices >= 3 && !u.c) removeVertex(2);
        if (numVertices >= 2 && !u.b) removeVertex(1);
        if (numVertices >= 1 && !u.a) removeVertex(0);
    }

    bool updateClosestVectorAndPoints() {  // :98-264
        if (needsUpdate) {
            cachedBC.reset();
            needsUpdate = false;
            switch (numVertices) {
            case 0:
                cachedValidClosest = false;
                break;
            case 1: {
                cachedP1.set(P[0]);
                cachedP2.set(Q[0]);
                cachedV.set(cachedP1).sub(cachedP2);
                cachedBC.reset();
                cachedBC.setBary(1, 0, 0, 0);
                cachedValidClosest = cachedBC.isValid();
                break;
            }
            case 2: {
                V3 tmp;
                const V3& from = W[0];
                const V3& to = W[1];
                V3 nearest;
                V3 p(0, 0, 0);
                V3 diff; diff.set(p).sub(from);
                V3 v; v.set(to).sub(from);
                float t = v.dot(diff);
                if (t > 0) {
                    float dotVV = v.dot(v);
                    if (t < dotVV) {
                        t /= dotVV;
                        tmp.set(v).scl(t);
                        diff.sub(tmp);
                        cachedBC.used.a = true;
                        cachedBC.used.b = true;
                    } else {
                        t = 1;
                        diff.sub(v);
                        cachedBC.used.b = true;
                    }
                } else {
                    t = 0;
                    cachedBC.used.a = true;
                }
                cachedBC.setBary(1.0f - t, t, 0, 0);
                tmp.set(v).scl(t);
                nearest.set(from).add(tmp);

                tmp.set(P[1]).sub(P[0]);
                tmp.scl(t);
                cachedP1.set(P[0]).add(tmp);

                tmp.set(Q[1]).sub(Q[0]);
                tmp.scl(t);
                cachedP2.set(Q[0]).add(tmp);

                cachedV.set(cachedP1).sub(cachedP2);
                reduceVertices(cachedBC.used);
                cachedValidClosest = cachedBC.isValid();
                break;
            }
            case 3: {
                V3 tmp1, tmp2, tmp3;
                V3 p(0, 0, 0);
                V3 a = W[0], b = W[1], c = W[2];
                closestPtPointTriangle(p, a, b, c, cachedBC);

                tmp1.set(P[0]).scl(cachedBC.bary[0]);
                tmp2.set(P[1]).scl(cachedBC.bary[1]);
                tmp3.set(P[2]).scl(cachedBC.bary[2]);
                cachedP1.set(tmp1.x + tmp2.x + tmp3.x, tmp1.y + tmp2.y + tmp3.y, tmp1.z + tmp2.z + tmp3.z);

                tmp1.set(Q[0]).scl(cachedBC.bary[0]);
                tmp2.set(Q[1]).scl(cachedBC.bary[1]);
                tmp3.set(Q[2]).scl(cachedBC.bary[2]);
                cachedP2.set(tmp1.x + tmp2.x + tmp3.x, tmp1.y + tmp2.y + tmp3.y, tmp1.z + tmp2.z + tmp3.z);

                cachedV.set(cachedP1).sub(cachedP2);
                reduceVertices(cachedBC.used);
                cachedValidClosest = cachedBC.isValid();
                break;
            }
            case 4: {
                V3 tmp1, tmp2, tmp3, tmp4;
                V3 p(0, 0, 0);
                V3 a = W[0], b = W[1], c = W[2], d = W[3];
                bool hasSeparation = closestPtPointTetrahedron(p, a, b, c, d, cachedBC);
                if (hasSeparation) {
                    tmp1.set(P[0]).scl(cachedBC.bary[0]);
                    tmp2.set(P[1]).scl(cachedBC.bary[1]);
                    tmp3.set(P[2]).scl(cachedBC.bary[2]);
                    tmp4.set(P[3]).scl(cachedBC.bary[3]);
                    cachedP1.set(tmp1.x + tmp2.x + tmp3.x + tmp4.x, tmp1.y + tmp2.y + tmp3.y + tmp4.y,
                                 tmp1.z + tmp2.z + tmp3.z + tmp4.z);
                    tmp1.set(Q[0]).scl(cachedBC.bary[0]);
                    tmp2.set(Q[1]).scl(cachedBC.bary[1]);
                    tmp3.set(Q[2]).scl(cachedBC.bary[2]);
                    tmp4.set(Q[3]).scl(cachedBC.bary[3]);
                    cachedP2.set(tmp1.x + tmp2.x + tmp3.x + tmp4.x, tmp1.y + tmp2.y + tmp3.y + tmp4.y,
                                 tmp1.z + tmp2.z + tmp3.z + tmp4.z);
                    cachedV.set(cachedP1).sub(cachedP2);
                    reduceVertices(cachedBC.used);
                } else {
                    if (cachedBC.degenerate) {
                        cachedValidClosest = false;
                    } else {
                        cachedValidClosest = true;
                        cachedV.set(0, 0, 0);
                    }
                    break;
                }
                cachedValidClosest = cachedBC.isValid();
                break;
            }
            default:
                cachedValidClosest = false;
            }
        }
        return cachedValidClosest;
    }

    static bool closestPtPointTriangle(const V3& p, const V3& a, const V3& b, const V3& c,
                                       SubSimplexClosestResult& result) {  // :267-389
        result.used.reset();
        V3 ab; ab.set(b).sub(a);
        V3 ac; ac.set(c).sub(a);
        V3 ap; ap.set(p).sub(a);
        float d1 = ab.dot(ap);
        float d2 = ac.dot(ap);
        if (d1 <= 0.0f && d2 <= 0.0f) {
            result.closestPointOnSimplex.set(a);
            result.used.a = true;
            result.setBary(1, 0, 0, 0);
            return true;
        }
        V3 bp; bp.set(p).sub(b);
        float d3 = ab.dot(bp);
        float d4 = ac.dot(bp);
        if (d3 >= 0.0f && d4 <= d3) {
            result.closestPointOnSimplex.set(b);
            result.used.b = true;
            result.setBary(0, 1, 0, 0);
            return true;
        }
        float vc = d1 * d4 - d3 * d2;
        if (vc <= 0.0f && d1 >= 0.0f && d3 <= 0.0f) {
            float v = d1 / (d1 - d3);
            result.closestPointOnSimplex.x = v * ab.x + a.x;
            result.closestPointOnSimplex.y = v * ab.y + a.y;
            result.closestPointOnSimplex.z = v * ab.z + a.z;
            result.used.a = true;
            result.used.b = true;
            result.setBary(1.0f - v, v, 0, 0);
            return true;
        }
        V3 cp; cp.set(p).sub(c);
        float d5 = ab.dot(cp);
        float d6 = ac.dot(cp);
        if (d6 >= 0.0f && d5 <= d6) {
            result.closestPointOnSimplex.set(c);
            result.used.c = true;
            result.setBary(0, 0, 1, 0);
            return true;
        }
        float vb = d5 * d2 - d1 * d6;
        if (vb <= 0.0f && d2 >= 0.0f && d6 <= 0.0f) {
            float w = d2 / (d2 - d6);
            result.closestPointOnSimplex.x = w * ac.x + a.x;
            result.closestPointOnSimplex.y = w * ac.y + a.y;
            result.closestPointOnSimplex.z = w * ac.z + a.z;
            result.used.a = true;
            result.used.c = true;
            result.setBary(1.0f - w, 0, w, 0);
            return true;
        }
        float va = d3 * d6 - d5 * d4;
        if (va <= 0.0f && (d4 - d3) >= 0.0f && (d5 - d6) >= 0.0f) {
            float w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
            V3 tmp; tmp.set(c).sub(b);
            result.closestPointOnSimplex.x = w * tmp.x + b.x;
            result.closestPointOnSimplex.y = w * tmp.y + b.y;
            result.closestPointOnSimplex.z = w * tmp.z + b.z;
            result.used.b = true;
            result.used.c = true;
            result.setBary(0, 1.0f - w, w, 0);
            return true;
        }
        float denom = 1.0f / (va + vb + vc);
        float v = vb * denom;
        float w = vc * denom;
        V3 tmp1, tmp2;
        tmp1.set(ab).scl(v);
        tmp2.set(ac).scl(w);
        result.closestPointOnSimplex.set(a.x + tmp1.x + tmp2.x, a.y + tmp1.y + tmp2.y, a.z + tmp1.z + tmp2.z);
        result.used.a = true;
        result.used.b = true;
        result.used.c = true;
        result.setBary(1.0f - v - w, v, w, 0);
        return true;
    }

    static int pointOutsideOfPlane(const V3& p, const V3& a, const V3& b, const V3& c, const V3& d) {  // :393-425
        V3 tmp, normal;
        normal.set(b).sub(a);
        tmp.set(c).sub(a);
        normal.crs(tmp);
        tmp.set(p).sub(a);
        float signp = tmp.dot(normal);
        tmp.set(d).sub(a);
        float signd = tmp.dot(normal);
        if (signd * signd < ((1e-4f) * (1e-4f))) return -1;
        return (signp * signd < 0.0f) ? 1 : 0;
    }

    static bool closestPtPointTetrahedron(const V3& p, const V3& a, const V3& b, const V3& c, const V3& d,
                                          SubSimplexClosestResult& fin) {  // :428-561
        SubSimplexClosestResult temp;
        temp.reset();
        V3 tmp, q;
        fin.closestPointOnSimplex.set(p);
        fin.used.reset();
        fin.used.a = fin.used.b = fin.used.c = fin.used.d = true;

        int oABC = pointOutsideOfPlane(p, a, b, c, d);
        int oACD = pointOutsideOfPlane(p, a, c, d, b);
        int oADB = pointOutsideOfPlane(p, a, d, b, c);
        int oBDC = pointOutsideOfPlane(p, b, d, c, a);

        if (oABC < 0 || oACD < 0 || oADB < 0 || oBDC < 0) {
            fin.degenerate = true;
            return false;
        }
        if (oABC == 0 && oACD == 0 && oADB == 0 && oBDC == 0) return false;

        float bestSqDist = 3.4028234663852886e38f;
        if (oABC != 0) {
            closestPtPointTriangle(p, a, b, c, temp);
            q.set(temp.closestPointOnSimplex);
            tmp.set(q).sub(p);
            float sqDist = tmp.dot(tmp);
            if (sqDist < bestSqDist) {
                bestSqDist = sqDist;
                fin.closestPointOnSimplex.set(q);
                fin.used.reset();
                fin.used.a = temp.used.a;
                fin.used.b = temp.used.b;
                fin.used.c = temp.used.c;
                fin.setBary(temp.bary[0], temp.bary[1], temp.bary[2], 0);
            }
        }
        if (oACD != 0) {
            closestPtPointTriangle(p, a, c, d, temp);
            q.set(temp.closestPointOnSimplex);
            tmp.set(q).sub(p);
            float sqDist = tmp.dot(tmp);
            if (sqDist < bestSqDist) {
                bestSqDist = sqDist;
                fin.closestPointOnSimplex.set(q);
                fin.used.reset();
                fin.used.a = temp.used.a;
                fin.used.c = temp.used.b;
                fin.used.d = temp.used.c;
                fin.setBary(temp.bary[0], 0, temp.bary[1], temp.bary[2]);
            }
        }
        if (oADB != 0) {
            closestPtPointTriangle(p, a, d, b, temp);
            q.set(temp.closestPointOnSimplex);
            tmp.set(q).sub(p);
            float sqDist = tmp.dot(tmp);
            if (sqDist < bestSqDist) {
                bestSqDist = sqDist;
                fin.closestPointOnSimplex.set(q);
                fin.used.reset();
                fin.used.a = temp.used.a;
                fin.used.b = temp.used.c;
                fin.used.d = temp.used.b;
                fin.setBary(temp.bary[0], temp.bary[2], 0, temp.bary[1]);
            }
        }
        if (oBDC != 0) {
            closestPtPointTriangle(p, b, d, c, temp);
            q.set(temp.closestPointOnSimplex);
            tmp.set(q).sub(p);
            float sqDist = tmp.dot(tmp);
            if (sqDist < bestSqDist) {
                bestSqDist = sqDist;
                fin.closestPointOnSimplex.set(q);
                fin.used.reset();
                fin.used.b = temp.used.a;
                fin.used.c = temp.used.c;
                fin.used.d = temp.used.b;
                fin.setBary(0, temp.bary[0], temp.bary[2], temp.bary[1]);
            }
        }
        return true;
    }

    void reset() {  // :564-570
        cachedValidClosest = false;
        numVertices = 0;
        needsUpdate = true;
        lastW.set(1e30f, 1e30f, 1e30f);
        cachedBC.reset();
    }
    void addVertex(const V3& w, const V3& p, const V3& q) {  // :572-581
        lastW.set(w);
        needsUpdate = true;
        W[numVertices].set(w);
        P[numVertices].set(p);
        Q[numVertices].set(q);
        numVertices++;
    }
    bool closest(V3& v) {  // :584-588
        bool ok = updateClosestVectorAndPoints();
        v.set(cachedV);
        return ok;
    }
    bool fullSimplex() const { return numVertices == 4; }  // :602-604
    bool inSimplex(const V3& w) const {  // :615-633
        bool found = false;
        for (int i = 0; i < numVertices; i++)
            if (W[i].equals(w)) found = true;
        if (w.equals(lastW)) return true;
        return found;
    }
    void backup_closest(V3& v) const { v.set(cachedV); }  // :635-637
    void compute_points(V3& p1, V3& p2) {  // :643-647
        updateClosestVectorAndPoints();
        p1.set(cachedP1);
        p2.set(cachedP2);
    }
};

}  // namespace orc
