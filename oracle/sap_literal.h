// ORACLE (test infrastructure only; never linked into the product) — "parity unpinned": the reference ships no tests.
//
// sap_literal.h — the reference's AxisSweep3, two restatements:
//   * SapQuantizer: bp/AxisSweep3Internal.java:201-216 quantize (+ :87-105 the quantisation scale), both widths
//     (bp/AxisSweep3.java:52 16 bit, bp/AxisSweep3_32.java:49 31 bit);
//   * LSap: a literal transliteration of the incremental algorithm — edge arrays with boundary sentinels (:122-147),
//     addHandle (:450-494), removeHandle (:496-535), updateHandle (:537-574), sortMinDown/Up, sortMaxDown/Up
//     (:219-374), testOverlap on edge INDICES (:174-195), the pair cache reduced to a set with the group/mask filter
//     and uid ordering (bp/HashedOverlappingPairCache.java:67-75,179-188,291-296).
// Minimum edges are even and maximum edges odd (quantize: "& handleMask | isMax"), so a min and a max never tie and the
// index order the literal algorithm tests equals the order of the quantised values: after every single edge swap the
// pair set is exactly {filter(a,b) and the quantised intervals overlap on the three axes}.  tests/
// test_oracle_differential.py checks that claim by running both on moving scenes, with removals.
#pragma once
#include <set>
#include <utility>
#include <vector>
#include "jmath.h"

namespace orc {

struct SapQuantizer {
    V3 worldMin, worldMax, quant;
    int handleMask = 0xfffe, sentinel = 0xffff;
    unsigned mask = 0xffffu;
    void init(const V3& mn, const V3& mx, bool wide) {
        worldMin = mn; worldMax = mx;
        handleMask = wide ? (int)0xfffffffe : 0xfffe;
        sentinel = wide ? 0x7fffffff : 0xffff;
        mask = wide ? 0xffffffffu : 0xffffu;
        V3 size; size.set(worldMax).sub(worldMin);
        const float maxInt = (float)sentinel;  // int / float promotes the int
        quant.set(maxInt / size.x, maxInt / size.y, maxInt / size.z);
    }
    static int j2i(float f) {  // Java (int) cast: toward zero, saturating, NaN -> 0
        if (f != f) return 0;
        if (f >= 2147483648.0f) return 0x7fffffff;
        if (f <= -2147483648.0f) return (int)0x80000000;
        return (int)f;
    }
    void quantize(unsigned out[3], const V3& p, int isMax) const {
        V3 c; c.set(p);
        c.x = jmaxf(c.x, worldMin.x); c.y = jmaxf(c.y, worldMin.y); c.z = jmaxf(c.z, worldMin.z);
        c.x = jminf(c.x, worldMax.x); c.y = jminf(c.y, worldMax.y); c.z = jminf(c.z, worldMax.z);
        V3 v; v.set(c).sub(worldMin);
        v.x = v.x * quant.x; v.y = v.y * quant.y; v.z = v.z * quant.z;
        out[0] = (unsigned)((j2i(v.x) & handleMask) | isMax) & mask;
        out[1] = (unsigned)((j2i(v.y) & handleMask) | isMax) & mask;
        out[2] = (unsigned)((j2i(v.z) & handleMask) | isMax) & mask;
    }
    // monotone map back to floats (what the device uses as the box it grids and sweeps; the pair predicate is on the ints)
    float dequant(unsigned q, int axis) const { return (float)q / quant.get(axis) + worldMin.get(axis); }
};

struct LSap {
    struct Handle {
        int minEdges[3], maxEdges[3];
        int group = 1, mask = -1, world = 0;
        int nextFree = 0;
        bool used = false;
    };
    struct Edge { unsigned pos; int handle; };  // isMax = pos & 1
    SapQuantizer Q;
    std::vector<Handle> H;
    std::vector<Edge> E[3];
    int firstFree = 1, numHandles = 0;
    std::set<std::pair<int, int>> pairs;

    void init(const V3& mn, const V3& mx, bool wide, int maxHandlesUser) {
        Q.init(mn, mx, wide);
        const int maxHandles = maxHandlesUser + 1;
        H.assign(maxHandles, Handle());
        for (int i = 1; i < maxHandles; i++) H[i].nextFree = i + 1;
        H[maxHandles - 1].nextFree = 0;
        firstFree = 1; numHandles = 0;
        for (int a = 0; a < 3; a++) {
            E[a].assign(maxHandles * 2, Edge{0u, 0});
            H[0].minEdges[a] = 0; H[0].maxEdges[a] = 1;
            E[a][0] = Edge{0u, 0};
            E[a][1] = Edge{(unsigned)Q.sentinel, 0};
        }
        pairs.clear();
    }
    bool filter(const Handle& a, const Handle& b) const {
        if (a.world != b.world) return false;
        return (a.group & b.mask) != 0 && (b.group & a.mask) != 0;
    }
    void addPair(int a, int b) {
        if (!filter(H[a], H[b])) return;
        pairs.insert(std::make_pair(a < b ? a : b, a < b ? b : a));
    }
    void removePair(int a, int b) { pairs.erase(std::make_pair(a < b ? a : b, a < b ? b : a)); }
    bool testOverlap(int ignoreAxis, const Handle& A, const Handle& B) const {  // :174-195, on edge indices
        for (int axis = 0; axis < 3; axis++)
            if (axis != ignoreAxis)
                if (A.maxEdges[axis] < B.minEdges[axis] || B.maxEdges[axis] < A.minEdges[axis]) return false;
        return true;
    }
    static bool isMax(const Edge& e) { return (e.pos & 1u) != 0; }
    void sortMinDown(int axis, int edge, bool upd) {  // :219-256
        std::vector<Edge>& A = E[axis];
        int e = edge, p = edge - 1;
        const int he = A[e].handle;
        while (A[e].pos < A[p].pos) {
            const int hp = A[p].handle;
            if (isMax(A[p])) {
                if (upd && testOverlap(axis, H[he], H[hp])) addPair(he, hp);
                H[hp].maxEdges[axis]++;
            } else {
                H[hp].minEdges[axis]++;
            }
            H[he].minEdges[axis]--;
            std::swap(A[e], A[p]);
            e--; p--;
        }
    }
    void sortMinUp(int axis, int edge, bool upd) {  // :258-294
        std::vector<Edge>& A = E[axis];
        int e = edge, n = edge + 1;
        const int he = A[e].handle;
        while (A[n].handle != 0 && A[e].pos >= A[n].pos) {
            const int hn = A[n].handle;
            if (isMax(A[n])) {
                if (upd) removePair(A[e].handle, A[n].handle);
                H[hn].maxEdges[axis]--;
            } else {
                H[hn].minEdges[axis]--;
            }
            H[he].minEdges[axis]++;
            std::swap(A[e], A[n]);
            e++; n++;
        }
    }
    void sortMaxDown(int axis, int edge, bool upd) {  // :296-335
        std::vector<Edge>& A = E[axis];
        int e = edge, p = edge - 1;
        const int he = A[e].handle;
        while (A[e].pos < A[p].pos) {
            const int hp = A[p].handle;
            if (!isMax(A[p])) {
                if (upd) removePair(A[e].handle, A[p].handle);
                H[hp].minEdges[axis]++;
            } else {
                H[hp].maxEdges[axis]++;
            }
            H[he].maxEdges[axis]--;
            std::swap(A[e], A[p]);
            e--; p--;
        }
    }
    void sortMaxUp(int axis, int edge, bool upd) {  // :337-374
        std::vector<Edge>& A = E[axis];
        int e = edge, n = edge + 1;
        const int he = A[e].handle;
        while (A[n].handle != 0 && A[e].pos >= A[n].pos) {
            const int hn = A[n].handle;
            if (!isMax(A[n])) {
                if (upd && testOverlap(axis, H[he], H[hn])) addPair(A[e].handle, A[n].handle);
                H[hn].minEdges[axis]--;
            } else {
                H[hn].maxEdges[axis]--;
            }
            H[he].maxEdges[axis]++;
            std::swap(A[e], A[n]);
            e++; n++;
        }
    }
    int addHandle(const V3& mn, const V3& mx, int group, int mask, int world) {  // :450-494
        unsigned qmin[3], qmax[3];
        Q.quantize(qmin, mn, 0);
        Q.quantize(qmax, mx, 1);
        const int handle = firstFree;
        firstFree = H[handle].nextFree;
        numHandles++;
        Handle& h = H[handle];
        h.group = group; h.mask = mask; h.world = world; h.used = true;
        const int limit = numHandles * 2;
        for (int axis = 0; axis < 3; axis++) {
            H[0].maxEdges[axis] += 2;
            E[axis][limit + 1] = E[axis][limit - 1];
            E[axis][limit - 1] = Edge{qmin[axis], handle};
            E[axis][limit] = Edge{qmax[axis], handle};
            h.minEdges[axis] = limit - 1;
            h.maxEdges[axis] = limit;
        }
        sortMinDown(0, h.minEdges[0], false);
        sortMaxDown(0, h.maxEdges[0], false);
        sortMinDown(1, h.minEdges[1], false);
        sortMaxDown(1, h.maxEdges[1], false);
        sortMinDown(2, h.minEdges[2], true);
        sortMaxDown(2, h.maxEdges[2], true);
        return handle;
    }
    void removeHandle(int handle) {  // :496-535
        for (auto it = pairs.begin(); it != pairs.end();) {  // removeOverlappingPairsContainingProxy
            if (it->first == handle || it->second == handle) it = pairs.erase(it);
            else ++it;
        }
        const int limit = numHandles * 2;
        for (int axis = 0; axis < 3; axis++) H[0].maxEdges[axis] -= 2;
        for (int axis = 0; axis < 3; axis++) {
            int mx = H[handle].maxEdges[axis];
            E[axis][mx].pos = (unsigned)Q.sentinel;
            sortMaxUp(axis, mx, false);
            int i = H[handle].minEdges[axis];
            E[axis][i].pos = (unsigned)Q.sentinel;
            sortMinUp(axis, i, false);
            E[axis][limit - 1].handle = 0;
            E[axis][limit - 1].pos = (unsigned)Q.sentinel;
        }
        H[handle].nextFree = firstFree;
        H[handle].used = false;
        firstFree = handle;
        numHandles--;
    }
    void updateHandle(int handle, const V3& mn, const V3& mx) {  // :537-574
        unsigned qmin[3], qmax[3];
        Q.quantize(qmin, mn, 0);
        Q.quantize(qmax, mx, 1);
        Handle& h = H[handle];
        for (int axis = 0; axis < 3; axis++) {
            const int emin = h.minEdges[axis], emax = h.maxEdges[axis];
            const long long dmin = (long long)(int)qmin[axis] - (long long)(int)E[axis][emin].pos;
            const long long dmax = (long long)(int)qmax[axis] - (long long)(int)E[axis][emax].pos;
            E[axis][emin].pos = qmin[axis];
            E[axis][emax].pos = qmax[axis];
            if (dmin < 0) sortMinDown(axis, emin, true);
            if (dmax > 0) sortMaxUp(axis, emax, true);
            if (dmin > 0) sortMinUp(axis, emin, true);
            if (dmax < 0) sortMaxDown(axis, emax, true);
        }
    }
};

}  // namespace orc
