// common.cuh — device data layout and strict-float helpers for the B200 collision path.
//
// Float discipline: every value that feeds a bit-exact comparison (AABBs -> pair set,
// BVH quantisation -> triangle set) and every GJK/EPA branch decision is computed in IEEE binary32,
// one rounding per operation, in the reference's operation order.  This translation unit is compiled
// with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false, so a*b+c below is two roundings.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace b2c {

struct f3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ f3 mk3(float x, float y, float z) { f3 r; r.x = x; r.y = y; r.z = z; return r; }
__host__ __device__ __forceinline__ f3 add3(f3 a, f3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ f3 sub3(f3 a, f3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ f3 scl3(f3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ f3 neg3(f3 a) { return mk3(a.x * -1.0f, a.y * -1.0f, a.z * -1.0f); }  // scl(-1)
__host__ __device__ __forceinline__ float dot3(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__host__ __device__ __forceinline__ float len2_3(f3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
// a x b in libgdx Vector3.crs order
__host__ __device__ __forceinline__ f3 crs3(f3 a, f3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float jsqrtf(float x) { return __fsqrt_rn(x); }  // (float)Math.sqrt(x)
__device__ __forceinline__ float len3(f3 a) { return jsqrtf(len2_3(a)); }
// libgdx Vector3.nor(): no-op when len2 is exactly 0 or 1
__device__ __forceinline__ f3 nor3(f3 a) {
    float l2 = len2_3(a);
    if (l2 == 0.0f || l2 == 1.0f) return a;
    return scl3(a, 1.0f / jsqrtf(l2));
}
// Float.floatToIntBits equality (canonical NaN) as Vector3.equals uses it
__device__ __forceinline__ uint32_t f2bits(float f) { return (f != f) ? 0x7fc00000u : __float_as_uint(f); }
__device__ __forceinline__ bool eq3bits(f3 a, f3 b) {
    return f2bits(a.x) == f2bits(b.x) && f2bits(a.y) == f2bits(b.y) && f2bits(a.z) == f2bits(b.z);
}
// Math.max / Math.min on floats (NaN-propagating, -0 < +0)
__device__ __forceinline__ float jmaxf(float a, float b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0f && b == 0.0f) return __uint_as_float(__float_as_uint(a) & __float_as_uint(b));
    return a >= b ? a : b;
}
__device__ __forceinline__ float jminf(float a, float b) {
    if (a != a) return a;
    if (b != b) return b;
    if (a == 0.0f && b == 0.0f) return __uint_as_float(__float_as_uint(a) | __float_as_uint(b));
    return a <= b ? a : b;
}

// World transform of one body: three float4 rows (m_r0 m_r1 m_r2 origin_r) -> 128-bit loads.
struct Xf {
    float m[3][3];
    f3 o;
};
__device__ __forceinline__ Xf loadXf(const float4* __restrict__ xf4, int body) {
    float4 r0 = __ldg(xf4 + 3 * (size_t)body), r1 = __ldg(xf4 + 3 * (size_t)body + 1), r2 = __ldg(xf4 + 3 * (size_t)body + 2);
    Xf t;
    t.m[0][0] = r0.x; t.m[0][1] = r0.y; t.m[0][2] = r0.z;
    t.m[1][0] = r1.x; t.m[1][1] = r1.y; t.m[1][2] = r1.z;
    t.m[2][0] = r2.x; t.m[2][1] = r2.y; t.m[2][2] = r2.z;
    t.o = mk3(r0.w, r1.w, r2.w);
    return t;
}
// Vector3.mul(Matrix3): M * v
__host__ __device__ __forceinline__ f3 mulMV(const float m[3][3], f3 v) {
    return mk3(v.x * m[0][0] + v.y * m[0][1] + v.z * m[0][2], v.x * m[1][0] + v.y * m[1][1] + v.z * m[1][2],
               v.x * m[2][0] + v.y * m[2][1] + v.z * m[2][2]);
}
// MatrixUtil.transposeTransform: M^T * v  (lm/MatrixUtil.java:297-316)
__host__ __device__ __forceinline__ f3 mulMtV(const float m[3][3], f3 v) {
    return mk3(m[0][0] * v.x + m[1][0] * v.y + m[2][0] * v.z, m[0][1] * v.x + m[1][1] * v.y + m[2][1] * v.z,
               m[0][2] * v.x + m[1][2] * v.y + m[2][2] * v.z);
}
// Transform.transform: M*v + o  (lm/Transform.java:91-94)
__host__ __device__ __forceinline__ f3 xfPoint(const Xf& t, f3 v) { return add3(mulMV(t.m, v), t.o); }
// Transform.invXform: M^T(v - o) evaluated as Vector3.mul(transposed) (lm/Transform.java:133-140)
__host__ __device__ __forceinline__ f3 invXfPoint(const Xf& t, f3 v) {
    f3 d = sub3(v, t.o);
    return mk3(d.x * t.m[0][0] + d.y * t.m[1][0] + d.z * t.m[2][0], d.x * t.m[0][1] + d.y * t.m[1][1] + d.z * t.m[2][1],
               d.x * t.m[0][2] + d.y * t.m[1][2] + d.z * t.m[2][2]);
}
// Matrix3.mul: A*B with libgdx's accumulation order
__host__ __device__ __forceinline__ void mulMM(const float a[3][3], const float b[3][3], float r[3][3]) {
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) r[i][j] = a[i][0] * b[0][j] + a[i][1] * b[1][j] + a[i][2] * b[2][j];
}
// Transform.inverse() then .mul(tr): returns inv(a) * b  (lm/Transform.java:101-120)
__host__ __device__ __forceinline__ Xf invMul(const Xf& a, const Xf& b) {
    Xf inv;
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) inv.m[i][j] = a.m[j][i];
    inv.o = mulMV(inv.m, neg3(a.o));
    Xf r;
    r.o = xfPoint(inv, b.o);
    mulMM(inv.m, b.m, r.m);
    return r;
}

enum { SH_BOX = 0, SH_SPHERE = 1, SH_HULL = 2, SH_TRIANGLE = 3, SH_PLANE = 4, SH_MESH = 5, SH_COMPOUND = 6 };

// One registered shape (64 B, read through the read-only path).
struct ShapeDev {
    int type;
    float margin;        // getMargin(): sphere -> radius (sh/SphereShape.java:93-97), else collisionMargin
    float dims[3];       // box: implicitShapeDimensions (half extents minus margin); sphere: dims[0] = radius
    float aabbMin[3];    // hull / mesh local AABB (sh/PolyhedralConvexShape.java:177-201, sh/TriangleMeshShape.java:79-93)
    float aabbMax[3];
    int pointOffset;     // hull: first vertex in the float4 hull-point pool; compound: first child in the child table
    int numPoints;       // hull: vertices; compound: children
    float plane[4];      // static plane: unit normal, constant
    int mesh;            // index into the mesh table
};

// One child of a CompoundShape (sh/CompoundShapeChild.java:34-39): local transform + child shape id.  64 B.
// A compound's entries in the child table are its LEAVES (box / sphere / hull) in the depth-first order in which
// CompoundCollisionAlgorithm visits them; a child that is itself a CompoundShape contributes its own leaves, each chained
// (parent1) to a FRAME entry that holds the nested compound's local transform, so that a leaf's world transform is composed
// exactly as the reference's recursion does: ((orgTrans * frame) * ...) * childTrans (disp/CompoundCollisionAlgorithm.java:107).
struct CompoundChildDev {
    float m[9];          // childTransform basis, row-major
    float o[3];          // childTransform origin
    int shape;           // index into the shape table (box, sphere or hull); -1: a frame entry (a nested compound's transform)
    int parent1;         // 1 + index of the frame entry this entry hangs under, 0 = directly under the compound
    int pad[2];
};
constexpr int COMPOUND_MAX_DEPTH = 4;   // frames above a leaf (nesting depth 5)
// Transform.mul(tr1, tr2) (lm/Transform.java:122-131): origin = tr1.transform(tr2.origin), basis = tr1.basis * tr2.basis
__host__ __device__ __forceinline__ Xf mulXf(const Xf& a, const Xf& b) {
    Xf r;
    r.o = xfPoint(a, b.o);
    mulMM(a.m, b.m, r.m);
    return r;
}
__host__ __device__ __forceinline__ Xf compoundChildLocal(const CompoundChildDev& ch) {
    Xf l;
    l.m[0][0] = ch.m[0]; l.m[0][1] = ch.m[1]; l.m[0][2] = ch.m[2];
    l.m[1][0] = ch.m[3]; l.m[1][1] = ch.m[4]; l.m[1][2] = ch.m[5];
    l.m[2][0] = ch.m[6]; l.m[2][1] = ch.m[7]; l.m[2][2] = ch.m[8];
    l.o = mk3(ch.o[0], ch.o[1], ch.o[2]);
    return l;
}
// world transform of a compound's leaf under the object's transform `org`: the frames from the outermost down, then the leaf
__host__ __device__ __forceinline__ Xf compoundChildWorld(const Xf& org, const CompoundChildDev* table, const CompoundChildDev& leaf) {
    if (leaf.parent1 == 0) return mulXf(org, compoundChildLocal(leaf));
    int chain[COMPOUND_MAX_DEPTH];
    int d = 0;
    for (int p = leaf.parent1; p > 0 && d < COMPOUND_MAX_DEPTH; p = table[p - 1].parent1) chain[d++] = p - 1;
    Xf w = org;
    for (int k = d - 1; k >= 0; k--) w = mulXf(w, compoundChildLocal(table[chain[k]]));
    return mulXf(w, compoundChildLocal(leaf));
}

// One registered triangle mesh with its quantized BVH (sh/OptimizedBvh.java, sh/QuantizedBvhNodes.java).
struct MeshDev {
    const int4* nodes;   // 16 B: (minx|miny<<16, minz|maxx<<16, maxy|maxz<<16, escapeOrTriangle)
    const float* verts;  // xyz, pre-multiplied by the mesh scaling (sh/VertexData.java:50-55)
    const int* idx;      // 3 per triangle, all sub-parts back to back, rebased onto the concatenated vertex array
    int numNodes;
    int numTris;
    float qmin[3], qmax[3], quant[3];  // bvhAabbMin, bvhAabbMax, bvhQuantization
    int numParts;        // sh/TriangleIndexVertexArray sub-parts; a leaf names its triangle as partId << 21 | index (sh/OptimizedBvh.java:65,278)
    const int* partStart;  // [numParts + 1] first triangle of every part in the concatenated arrays (null when numParts == 1)
};

// AxisSweep3 quantisation (bp/AxisSweep3Internal.java:87-105, 201-216)
struct SapParams {
    float wmin[3], wmax[3], quant[3];
    int handleMask;      // 0xfffe / 0xfffffffe
    uint32_t mask;       // 0xffff / 0xffffffff
    int enabled;
};

enum { BF_STATIC = 1, BF_ALIVE = 2, BF_ACTIVE = 4, BF_OVERFLOW = 8, BF_INFIXED = 16 };

// Grid over the two non-sweep axes, chosen on the device each step (no host sync).
struct GridParams {
    float y0, z0;          // grid origin
    float invCellY, invCellZ;
    int ny, nz;            // cells per world
    int rowsPerWorld;      // ny*nz
    int nrows;             // numWorlds*ny*nz ; rows nrows + w = large proxies of world w, row nrows + numWorlds = dead slots
    int numWorlds;
    float cellY, cellZ;
    float x0, invX;        // sweep-axis quantisation: qx = clamp(floor((min.x - x0) * invX), 0, xmaxf)
    int xbits;             // bits of qx in the sort key (key = row << xbits | qx): 12 when the rows need <= 12 bits, fewer when that
    uint32_t xmask;        //   keeps the key inside 24 bits (three 8-bit radix passes instead of four)
    float xmaxf;           // (1 << xbits) - 1 as a float
};

// Per-step device counters (one 128-byte block, cleared by the first kernel of the step).
struct StepCounters {
    uint32_t pairCount;        // emitted overlapping pairs (may exceed capacity -> overflow)
    uint32_t pairOverflow;
    uint32_t boundsTicket;     // last-block election in the bounds kernel
    uint32_t extYBits, extZBits;         // max dynamic extent (float bits, extents are >= 0)
    uint32_t minYKey, minZKey, maxYKey, maxZKey;  // ordered-uint keys of min-corner bounds of gridded proxies
    uint32_t contactsAdded, gjkChecks, deepChecks, epaFailed;
    uint32_t meshItems, meshOverflow, numManifolds;
    uint32_t binCount[16];
    uint32_t epaCount;
    uint32_t largeCount;
    uint32_t epaRetry;
    uint32_t minXKey, maxXKey;  // sweep-axis bounds of the gridded proxies (min kept complemented)
    uint32_t migrateOverflow;   // partitioned world: a migration slot was too small
    uint32_t aabbTicket;        // last-block election in k_aabb (fused bounds -> grid)
    uint32_t keysTicket;        // last-block election in k_keys (fused radix histograms -> digit offsets)
    uint32_t haloOverflow;      // partitioned world: a halo slot was too small
    uint32_t maxRowLen;         // longest grid row of this step (row-grouped ordering, pairfind.cuh)
    uint32_t epaBig;            // penetration items sent straight to the large-pool tier (their pair overflowed the small pools before)
    uint32_t pad[5];
};

// monotone float <-> uint key (total order matching float compare for non-NaN; -0 canonicalised to +0)
__host__ __device__ __forceinline__ uint32_t floatKey(float f) {
    f = f + 0.0f;  // -0 -> +0
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float keyFloat(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

}  // namespace b2c
