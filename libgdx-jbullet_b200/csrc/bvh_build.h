// bvh_build.h — host-side construction of the quantized BVH for BvhTriangleMeshShape.
//
// Produces the node array sh/OptimizedBvh.java:283-342 (build) produces for
// useQuantizedAabbCompression=true: same quantisation grid (:132-144), same leaf boxes (:240-280, flat
// boxes padded to 0.002), same top-down split (:480-561: variance axis :676-707, mean split with the
// middle-third rebalance :622-674), same 16-byte node encoding (sh/QuantizedBvhNodes.java:34-48).  The
// tree is built once at registration on the host (it is a sequential partition) and uploaded; queries
// run on the device.  Host float code is compiled with -ffp-contract=off.
//
// Implementation notes (ours): leaves are kept as an index permutation over precomputed float centres
// and quantised boxes instead of swapping 16-byte nodes, and the recursion is an explicit stack.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace b2c {

struct HostBvh {
    std::vector<int32_t> nodes;  // 4 ints per node
    float qmin[3], qmax[3], quant[3];
    float localMin[3], localMax[3];  // vertex-coordinate extremes (TriangleMeshShape.recalcLocalAabb)
};

namespace bvhdetail {
static inline float jmax(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a >= b ? a : b)); }
static inline float jmin(float a, float b) { return (a != a) ? a : ((b != b) ? b : (a <= b ? a : b)); }
static inline int f2i(float f) {
    if (f != f) return 0;
    if (f >= 2147483648.0f) return 2147483647;
    if (f <= -2147483648.0f) return (int)0x80000000;
    return (int)f;
}
struct Q3 { uint16_t v[3]; };
}  // namespace bvhdetail

// verts: scaled vertex positions xyz; idx: 3 per triangle (all sub-parts back to back, in part order — the order
// sh/StridingMeshInterface.java:40-58 visits them); leafWord[t] = partId << 21 | index inside the part (:278), or null for a
// one-part mesh.
inline void buildQuantizedBvh(const float* verts, const int32_t* idx, int numTris, HostBvh& out, const int32_t* leafWord = nullptr) {
    using namespace bvhdetail;
    // mesh bounds over referenced vertices (sh/StridingMeshInterface.java calculateAabbBruteForce)
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int t = 0; t < numTris; t++)
        for (int k = 0; k < 3; k++) {
            const float* p = verts + 3 * (size_t)idx[3 * (size_t)t + k];
            for (int c = 0; c < 3; c++) { mn[c] = jmin(mn[c], p[c]); mx[c] = jmax(mx[c], p[c]); }
        }
    for (int c = 0; c < 3; c++) {
        out.localMin[c] = mn[c];
        out.localMax[c] = mx[c];
        out.qmin[c] = mn[c] - 1.0f;  // quantizationMargin 1 (:128-144)
        out.qmax[c] = mx[c] + 1.0f;
        out.quant[c] = 65535.0f / (out.qmax[c] - out.qmin[c]);
    }
    auto quantize = [&](const float p[3], uint16_t q[3]) {  // :1038-1056
        for (int c = 0; c < 3; c++) {
            float v = jmin(jmax(p[c], out.qmin[c]), out.qmax[c]);
            v = (v - out.qmin[c]) * out.quant[c];
            q[c] = (uint16_t)(f2i(v + 0.5f) & 0xFFFF);
        }
    };
    auto unquantize = [&](const uint16_t q[3], float p[3]) {  // :1058-1068
        for (int c = 0; c < 3; c++) p[c] = (float)(int)q[c] / out.quant[c] + out.qmin[c];
    };
    // leaves: quantised box, and the float centre of the UN-quantised box the reference sorts by (:103-113)
    std::vector<Q3> lmin(numTris), lmax(numTris);
    std::vector<float> centre(3 * (size_t)numTris), uqmin(3 * (size_t)numTris), uqmax(3 * (size_t)numTris);
    for (int t = 0; t < numTris; t++) {
        float a[3] = {1e30f, 1e30f, 1e30f}, b[3] = {-1e30f, -1e30f, -1e30f};
        for (int k = 0; k < 3; k++) {
            const float* p = verts + 3 * (size_t)idx[3 * (size_t)t + k];
            for (int c = 0; c < 3; c++) { a[c] = jmin(a[c], p[c]); b[c] = jmax(b[c], p[c]); }
        }
        for (int c = 0; c < 3; c++)
            if (b[c] - a[c] < 0.002f) { b[c] = b[c] + 0.001f; a[c] = a[c] - 0.001f; }
        quantize(a, lmin[t].v);
        quantize(b, lmax[t].v);
        float ua[3], ub[3];
        unquantize(lmin[t].v, ua);
        unquantize(lmax[t].v, ub);
        for (int c = 0; c < 3; c++) {
            uqmin[3 * (size_t)t + c] = ua[c];
            uqmax[3 * (size_t)t + c] = ub[c];
            centre[3 * (size_t)t + c] = (ub[c] + ua[c]) * 0.5f;
        }
    }
    std::vector<int> perm(numTris);
    for (int t = 0; t < numTris; t++) perm[t] = t;
    out.nodes.assign(4 * 2 * (size_t)(numTris > 0 ? numTris : 1), 0);
    int cur = 0;
    auto putNode = [&](int node, const uint16_t a[3], const uint16_t b[3], int32_t w) {
        int32_t* n = &out.nodes[4 * (size_t)node];
        n[0] = (int32_t)((uint32_t)a[0] | ((uint32_t)a[1] << 16));
        n[1] = (int32_t)((uint32_t)a[2] | ((uint32_t)b[0] << 16));
        n[2] = (int32_t)((uint32_t)b[1] | ((uint32_t)b[2] << 16));
        n[3] = w;
    };
    struct Frame { int start, end, node, stage, split; };
    std::vector<Frame> stack;
    if (numTris > 0) stack.push_back({0, numTris, -1, 0, 0});
    while (!stack.empty()) {
        Frame& fr = stack.back();
        if (fr.stage == 0) {
            int start = fr.start, end = fr.end, num = end - start;
            if (num == 1) {
                int t = perm[start];
                putNode(cur, lmin[t].v, lmax[t].v, leafWord ? leafWord[t] : t);  // (partId << 21) | triangleIndex
                cur++;
                stack.pop_back();
                continue;
            }
            // calcSplittingAxis (:676-707)
            float means[3] = {0, 0, 0}, var[3] = {0, 0, 0};
            for (int i = start; i < end; i++)
                for (int c = 0; c < 3; c++) means[c] = means[c] + centre[3 * (size_t)perm[i] + c];
            float inv = 1.0f / (float)num;
            for (int c = 0; c < 3; c++) means[c] = means[c] * inv;
            for (int i = start; i < end; i++)
                for (int c = 0; c < 3; c++) {
                    float d = centre[3 * (size_t)perm[i] + c] - means[c];
                    var[c] = var[c] + d * d;
                }
            float invm1 = 1.0f / ((float)num - 1);
            for (int c = 0; c < 3; c++) var[c] = var[c] * invm1;
            int axis = -1;
            float best = -1e30f;
            for (int c = 0; c < 3; c++)
                if (var[c] > best) { axis = c; best = var[c]; }
            if (axis < 0) axis = 0;
            // sortAndCalcSplittingIndex (:622-674): the mean is recomputed identically
            float splitValue = means[axis];
            int split = start;
            for (int i = start; i < end; i++) {
                if (centre[3 * (size_t)perm[i] + axis] > splitValue) {
                    int tmp = perm[i]; perm[i] = perm[split]; perm[split] = tmp;
                    split++;
                }
            }
            int third = num / 3;
            if ((split <= (start + third)) || (split >= (end - 1 - third))) split = start + (num >> 1);
            // internal node box: merge of re-quantised un-quantised leaf boxes (:521-527, :160-182)
            uint16_t a[3] = {65535, 65535, 65535}, b[3] = {0, 0, 0};
            {
                float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
                quantize(hi, b);
                quantize(lo, a);
            }
            for (int i = start; i < end; i++) {
                uint16_t qa[3], qb[3];
                quantize(&uqmin[3 * (size_t)perm[i]], qa);
                quantize(&uqmax[3 * (size_t)perm[i]], qb);
                for (int c = 0; c < 3; c++) {
                    if (a[c] > qa[c]) a[c] = qa[c];
                    if (b[c] < qb[c]) b[c] = qb[c];
                }
            }
            fr.node = cur;
            putNode(cur, a, b, 0);
            cur++;
            fr.split = split;
            fr.stage = 1;
            stack.push_back({start, split, -1, 0, 0});
        } else if (fr.stage == 1) {
            fr.stage = 2;
            int s = fr.split, e = fr.end;
            stack.push_back({s, e, -1, 0, 0});
        } else {
            out.nodes[4 * (size_t)fr.node + 3] = -(cur - fr.node);  // escape index (:540,:559)
            stack.pop_back();
        }
    }
    out.nodes.resize(4 * (size_t)cur);
}

}  // namespace b2c
