// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// bvh.h — restatement of sh/OptimizedBvh.java (quantized build + stackless AABB query),
// sh/QuantizedBvhNodes.java (16-byte node layout) and the mesh access of
// sh/StridingMeshInterface.java / sh/VertexData.java.  sh/ = collision/shapes/.
#pragma once
#include <cstdint>
#include <vector>
#include "jmath.h"

namespace orc {

struct MeshPart {  // one sub-part: sh/IndexedMesh.java:35-47 (indices already widened to int, sh/ByteBufferVertexData.java:75-84)
    std::vector<float> verts;    // xyz
    std::vector<int32_t> idx;    // 3 per triangle
    int numTriangles() const { return (int)idx.size() / 3; }
};
struct MeshData {  // sh/TriangleIndexVertexArray.java:45-100: a list of sub-parts, partId = position in the list
    std::vector<MeshPart> parts;
    V3 scaling{1, 1, 1};         // sh/StridingMeshInterface.java scaling
    int numTriangles() const {
        int n = 0;
        for (const MeshPart& p : parts) n += p.numTriangles();
        return n;
    }
    // sh/VertexData.java:50-55 getTriangle (vertex * scaling)
    void getTriangle(int part, int tri, V3 out[3]) const {
        const MeshPart& mp = parts[part];
        for (int i = 0; i < 3; i++) {
            int vi = mp.idx[tri * 3 + i];
            out[i].set(mp.verts[3 * vi] * scaling.x, mp.verts[3 * vi + 1] * scaling.y, mp.verts[3 * vi + 2] * scaling.z);
        }
    }
};

struct QNode {  // sh/QuantizedBvhNodes.java:34-48: u16 min[3], u16 max[3], i32 escapeOrTri
    uint16_t mn[3], mx[3];
    int32_t escapeIndexOrTriangleIndex;
    bool isLeaf() const { return escapeIndexOrTriangleIndex >= 0; }
};

struct Bvh {
    static constexpr int MAX_NUM_PARTS_IN_BITS = 10;  // sh/OptimizedBvh.java:65
    std::vector<QNode> leafNodes, nodes;
    int curNodeIndex = 0;
    V3 bvhAabbMin, bvhAabbMax, bvhQuantization;

    void setQuantizationValues(const V3& aabbMin, const V3& aabbMax, float quantizationMargin = 1.0f) {  // :132-144
        V3 clampValue(quantizationMargin, quantizationMargin, quantizationMargin);
        bvhAabbMin.set(aabbMin).sub(clampValue);
        bvhAabbMax.set(aabbMax).add(clampValue);
        V3 aabbSize; aabbSize.set(bvhAabbMax).sub(bvhAabbMin);
        bvhQuantization.set(65535.0f / aabbSize.x, 65535.0f / aabbSize.y, 65535.0f / aabbSize.z);
    }
    void quantizeWithClamp(const V3& point, uint16_t out[3]) const {  // :1038-1056
        V3 c = point;
        c.x = jmaxf(c.x, bvhAabbMin.x); c.y = jmaxf(c.y, bvhAabbMin.y); c.z = jmaxf(c.z, bvhAabbMin.z);
        c.x = jminf(c.x, bvhAabbMax.x); c.y = jminf(c.y, bvhAabbMax.y); c.z = jminf(c.z, bvhAabbMax.z);
        V3 v; v.set(c).sub(bvhAabbMin);
        v.set(v.x * bvhQuantization.x, v.y * bvhQuantization.y, v.z * bvhQuantization.z);
        out[0] = (uint16_t)(jf2i(v.x + 0.5f) & 0xFFFF);
        out[1] = (uint16_t)(jf2i(v.y + 0.5f) & 0xFFFF);
        out[2] = (uint16_t)(jf2i(v.z + 0.5f) & 0xFFFF);
    }
    V3 unQuantize(const uint16_t in[3]) const {  // :1058-1068
        V3 o;
        o.x = (float)(int)in[0] / bvhQuantization.x;
        o.y = (float)(int)in[1] / bvhQuantization.y;
        o.z = (float)(int)in[2] / bvhQuantization.z;
        o.add(bvhAabbMin);
        return o;
    }
    V3 leafMin(int i) const { return unQuantize(leafNodes[i].mn); }  // getAabbMin :103-113
    V3 leafMax(int i) const { return unQuantize(leafNodes[i].mx); }

    // :283-342 build (quantized) with :240-280 QuantizedNodeTriangleCallback
    void build(const MeshData& mesh, const V3& aabbMin, const V3& aabbMax) {
        setQuantizationValues(aabbMin, aabbMax);
        int T = mesh.numTriangles();
        leafNodes.resize(T);
        int t = 0;
        // sh/StridingMeshInterface.java:40-58 internalProcessAllTriangles: part by part, triangle by triangle
        for (int part = 0; part < (int)mesh.parts.size(); part++)
        for (int i = 0; i < mesh.parts[part].numTriangles(); i++, t++) {
            V3 tri[3];
            mesh.getTriangle(part, i, tri);
            V3 mn(1e30f, 1e30f, 1e30f), mx(-1e30f, -1e30f, -1e30f);
            for (int k = 0; k < 3; k++) {
                mn.x = jminf(mn.x, tri[k].x); mn.y = jminf(mn.y, tri[k].y); mn.z = jminf(mn.z, tri[k].z);
                mx.x = jmaxf(mx.x, tri[k].x); mx.y = jmaxf(mx.y, tri[k].y); mx.z = jmaxf(mx.z, tri[k].z);
            }
            const float MIN_AABB_DIMENSION = 0.002f, MIN_AABB_HALF_DIMENSION = 0.001f;
            if (mx.x - mn.x < MIN_AABB_DIMENSION) { mx.x = mx.x + MIN_AABB_HALF_DIMENSION; mn.x = mn.x - MIN_AABB_HALF_DIMENSION; }
            if (mx.y - mn.y < MIN_AABB_DIMENSION) { mx.y = mx.y + MIN_AABB_HALF_DIMENSION; mn.y = mn.y - MIN_AABB_HALF_DIMENSION; }
            if (mx.z - mn.z < MIN_AABB_DIMENSION) { mx.z = mx.z + MIN_AABB_HALF_DIMENSION; mn.z = mn.z - MIN_AABB_HALF_DIMENSION; }
            quantizeWithClamp(mn, leafNodes[t].mn);
            quantizeWithClamp(mx, leafNodes[t].mx);
            leafNodes[t].escapeIndexOrTriangleIndex = (part << (31 - MAX_NUM_PARTS_IN_BITS)) | i;  // :278
        }
        nodes.assign(2 * (size_t)T, QNode());
        curNodeIndex = 0;
        if (T > 0) buildTree(0, T);
        nodes.resize(curNodeIndex);
        leafNodes.clear();
    }

    int calcSplittingAxis(int startIndex, int endIndex) const {  // :676-707
        V3 means(0, 0, 0), variance(0, 0, 0), center;
        int numIndices = endIndex - startIndex;
        for (int i = startIndex; i < endIndex; i++) {
            center.set(leafMax(i)).add(leafMin(i));
            center.scl(0.5f);
            means.add(center);
        }
        means.scl(1.0f / (float)numIndices);
        V3 diff2;
        for (int i = startIndex; i < endIndex; i++) {
            center.set(leafMax(i)).add(leafMin(i));
            center.scl(0.5f);
            diff2.set(center).sub(means);
            diff2.set(diff2.x * diff2.x, diff2.y * diff2.y, diff2.z * diff2.z);
            variance.add(diff2);
        }
        variance.scl(1.0f / ((float)numIndices - 1));
        // lm/VectorUtil.java:41-58 maxAxis
        int maxIndex = -1;
        float maxVal = -1e30f;
        if (variance.x > maxVal) { maxIndex = 0; maxVal = variance.x; }
        if (variance.y > maxVal) { maxIndex = 1; maxVal = variance.y; }
        if (variance.z > maxVal) { maxIndex = 2; maxVal = variance.z; }
        return maxIndex;
    }
    int sortAndCalcSplittingIndex(int startIndex, int endIndex, int splitAxis) {  // :622-674
        int splitIndex = startIndex;
        int numIndices = endIndex - startIndex;
        V3 means(0, 0, 0), center;
        for (int i = startIndex; i < endIndex; i++) {
            center.set(leafMax(i)).add(leafMin(i));
            center.scl(0.5f);
            means.add(center);
        }
        means.scl(1.0f / (float)numIndices);
        float splitValue = means.get(splitAxis);
        for (int i = startIndex; i < endIndex; i++) {
            center.set(leafMax(i)).add(leafMin(i));
            center.scl(0.5f);
            if (center.get(splitAxis) > splitValue) {
                QNode t = leafNodes[i]; leafNodes[i] = leafNodes[splitIndex]; leafNodes[splitIndex] = t;
                splitIndex++;
            }
        }
        int rangeBalancedIndices = numIndices / 3;
        bool unbalanced = ((splitIndex <= (startIndex + rangeBalancedIndices)) ||
                           (splitIndex >= (endIndex - 1 - rangeBalancedIndices)));
        if (unbalanced) splitIndex = startIndex + (numIndices >> 1);
        return splitIndex;
    }
    void buildTree(int startIndex, int endIndex) {  // :480-561
        int numIndices = endIndex - startIndex;
        int curIndex = curNodeIndex;
        if (numIndices == 1) {
            nodes[curNodeIndex] = leafNodes[startIndex];
            curNodeIndex++;
            return;
        }
        int splitAxis = calcSplittingAxis(startIndex, endIndex);
        int splitIndex = sortAndCalcSplittingIndex(startIndex, endIndex, splitAxis);
        int internalNodeIndex = curNodeIndex;
        V3 tmp1(-1e30f, -1e30f, -1e30f), tmp2(1e30f, 1e30f, 1e30f);
        quantizeWithClamp(tmp1, nodes[curNodeIndex].mx);
        quantizeWithClamp(tmp2, nodes[curNodeIndex].mn);
        for (int i = startIndex; i < endIndex; i++) {  // mergeInternalNodeAabb :160-182
            uint16_t qmin[3], qmax[3];
            quantizeWithClamp(leafMin(i), qmin);
            quantizeWithClamp(leafMax(i), qmax);
            for (int k = 0; k < 3; k++) {
                if (nodes[curNodeIndex].mn[k] > qmin[k]) nodes[curNodeIndex].mn[k] = qmin[k];
                if (nodes[curNodeIndex].mx[k] < qmax[k]) nodes[curNodeIndex].mx[k] = qmax[k];
            }
        }
        curNodeIndex++;
        buildTree(startIndex, splitIndex);
        buildTree(splitIndex, endIndex);
        int escapeIndex = curNodeIndex - curIndex;
        nodes[internalNodeIndex].escapeIndexOrTriangleIndex = -escapeIndex;
    }

    // :709-740 reportAabbOverlappingNodex + :940-997 walkStacklessQuantizedTree; calls cb(partId, triIndex)
    template <class F>
    void reportAabbOverlappingNodex(const V3& aabbMin, const V3& aabbMax, F cb, int* nodesVisited = nullptr) const {
        uint16_t qmin[3], qmax[3];
        quantizeWithClamp(aabbMin, qmin);
        quantizeWithClamp(aabbMax, qmax);
        int curIndex = 0, endNodeIndex = curNodeIndex, walk = 0;
        while (curIndex < endNodeIndex) {
            walk++;
            const QNode& n = nodes[curIndex];
            bool overlap = true;  // :563-585 testQuantizedAabbAgainstQuantizedAabb
            overlap = (qmin[0] > n.mx[0] || qmax[0] < n.mn[0]) ? false : overlap;
            overlap = (qmin[2] > n.mx[2] || qmax[2] < n.mn[2]) ? false : overlap;
            overlap = (qmin[1] > n.mx[1] || qmax[1] < n.mn[1]) ? false : overlap;
            bool isLeaf = n.isLeaf();
            if (isLeaf && overlap) {
                int v = n.escapeIndexOrTriangleIndex;
                int tri = v & ~((~0) << (31 - MAX_NUM_PARTS_IN_BITS));
                int part = (int)((uint32_t)v >> (31 - MAX_NUM_PARTS_IN_BITS));
                cb(part, tri);
            }
            if (overlap || isLeaf) curIndex++;
            else curIndex += -n.escapeIndexOrTriangleIndex;
        }
        if (nodesVisited) *nodesVisited = walk;
    }
};

}  // namespace orc
