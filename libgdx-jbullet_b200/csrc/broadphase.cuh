// broadphase.cuh — stage 0 (AABB update) and stage 1 (overlap-pair finding) kernels.
//
// Replaces, for one step:
//   disp/CollisionWorld.java:231-245 updateAabbs / :195-229 updateSingleAabb          -> k_aabb
//   bp/DbvtBroadphase.java:196-228 setAabb (+ bp/Dbvt.java:157-169 in-place expand)     -> setAabbState
//   bp/DbvtBroadphase.java:89-150 collide / bp/SimpleBroadphase.java:81-110            -> k_bounds, k_keys,
//       radix sort, k_gather, k_sweep, k_large (pair set == all filtered overlaps of the effective AABBs)
//   bp/HashedOverlappingPairCache.java:179-188,291-296 (filter, uid ordering)           -> inside k_sweep/k_large
//
// Layout: 32-bit key = (row << 12) | qx, qx = min.x quantised to 12 bits over the gridded proxies' x range (any
// monotone function of min.x orders the sweep correctly; the overlap test itself uses the original floats).  row = world*ny*nz + cy*nz + cz is a coarse grid
// cell over the two non-sweep axes (cell >= the largest gridded extent, so overlapping proxies are in
// adjacent rows); proxies too large for the grid (static planes, meshes, big static boxes) share one
// extra row and are tested against everything.  After the sort the sweep visits, for every proxy, the
// x-window of its 9 neighbour rows.
#pragma once
#include "common.cuh"

namespace b2c {

struct BodyArrays {
    float4* xf4;          // [3*N] world transforms
    int* shape;           // [N]
    uint32_t* filt;       // [N] group | mask<<16
    uint8_t* flags;       // [N] BF_*
    int* world;           // [N]
    float4* effMin;       // [N] effective AABB min (w unused)
    float4* effMax;       // [N]
    float4* leafMin;      // [N] DbvtNode volume (dbvt mode)
    float4* leafMax;      // [N]
    int* lastSet;         // [N] step index of last setAabb
    float2* material;     // [N] friction, restitution
};

// ---- shape AABBs (bit-exact restatement of the reference's float sequences) -------------------------
// lm/AabbUtil2.java:133-163
__device__ __forceinline__ void aabbFromHalfExtents(f3 he, float margin, const Xf& t, f3& mn, f3& mx) {
    f3 h = mk3(he.x + margin, he.y + margin, he.z + margin);
    f3 ext = mk3(dot3(mk3(fabsf(t.m[0][0]), fabsf(t.m[0][1]), fabsf(t.m[0][2])), h),
                 dot3(mk3(fabsf(t.m[1][0]), fabsf(t.m[1][1]), fabsf(t.m[1][2])), h),
                 dot3(mk3(fabsf(t.m[2][0]), fabsf(t.m[2][1]), fabsf(t.m[2][2])), h));
    mn = sub3(t.o, ext);
    mx = add3(t.o, ext);
}
// lm/AabbUtil2.java:165-209
__device__ __forceinline__ void aabbFromLocalBox(f3 lmin, f3 lmax, float margin, const Xf& t, f3& mn, f3& mx) {
    f3 he = scl3(sub3(lmax, lmin), 0.5f);
    he = mk3(he.x + margin, he.y + margin, he.z + margin);
    f3 lc = scl3(add3(lmax, lmin), 0.5f);
    f3 c = xfPoint(t, lc);
    f3 ext = mk3(dot3(mk3(fabsf(t.m[0][0]), fabsf(t.m[0][1]), fabsf(t.m[0][2])), he),
                 dot3(mk3(fabsf(t.m[1][0]), fabsf(t.m[1][1]), fabsf(t.m[1][2])), he),
                 dot3(mk3(fabsf(t.m[2][0]), fabsf(t.m[2][1]), fabsf(t.m[2][2])), he));
    mn = sub3(c, ext);
    mx = add3(c, ext);
}
__device__ __forceinline__ void shapeAabb(const ShapeDev& s, const Xf& t, f3& mn, f3& mx) {
    switch (s.type) {
    case SH_BOX:  // sh/BoxShape.java:147-151
        aabbFromHalfExtents(mk3(s.dims[0], s.dims[1], s.dims[2]), s.margin, t, mn, mx);
        break;
    case SH_SPHERE: {  // sh/SphereShape.java:57-65
        f3 e = mk3(s.margin, s.margin, s.margin);
        mn = sub3(t.o, e);
        mx = add3(t.o, e);
        break;
    }
    case SH_COMPOUND:  // sh/CompoundShape.java:124-160: the float sequence of lm/AabbUtil2.java:165-209 over the children's box
    case SH_HULL:  // sh/PolyhedralConvexShape.java:169-171 (margin counted twice)
        aabbFromLocalBox(mk3(s.aabbMin[0], s.aabbMin[1], s.aabbMin[2]), mk3(s.aabbMax[0], s.aabbMax[1], s.aabbMax[2]),
                         s.margin, t, mn, mx);
        break;
    case SH_MESH: {  // sh/TriangleMeshShape.java:95-128 (margin added after the projection)
        f3 lmin = mk3(s.aabbMin[0], s.aabbMin[1], s.aabbMin[2]), lmax = mk3(s.aabbMax[0], s.aabbMax[1], s.aabbMax[2]);
        f3 he = scl3(sub3(lmax, lmin), 0.5f);
        f3 lc = scl3(add3(lmax, lmin), 0.5f);
        f3 c = xfPoint(t, lc);
        f3 ext = mk3(dot3(mk3(fabsf(t.m[0][0]), fabsf(t.m[0][1]), fabsf(t.m[0][2])), he),
                     dot3(mk3(fabsf(t.m[1][0]), fabsf(t.m[1][1]), fabsf(t.m[1][2])), he),
                     dot3(mk3(fabsf(t.m[2][0]), fabsf(t.m[2][1]), fabsf(t.m[2][2])), he));
        ext = add3(ext, mk3(s.margin, s.margin, s.margin));
        mn = sub3(c, ext);
        mx = add3(c, ext);
        break;
    }
    default:  // SH_PLANE: sh/StaticPlaneShape.java:125-128
        mn = mk3(-1e30f, -1e30f, -1e30f);
        mx = mk3(1e30f, 1e30f, 1e30f);
    }
}

__device__ __forceinline__ bool aabbIntersect(f3 amin, f3 amax, f3 bmin, f3 bmax) {  // bp/DbvtAabbMm.java:209-212
    return (amin.x <= bmax.x) && (amax.x >= bmin.x) && (amin.y <= bmax.y) && (amax.y >= bmin.y) && (amin.z <= bmax.z) &&
           (amax.z >= bmin.z);
}

// bp/AxisSweep3Internal.java:201-216 quantize (Java's (int) cast == cvt.rzi: toward zero, saturating, NaN -> 0)
__device__ __forceinline__ uint32_t sapQuantize1(float p, int axis, int isMax, const SapParams& sp) {
    float c = jminf(jmaxf(p, sp.wmin[axis]), sp.wmax[axis]);
    float v = (c - sp.wmin[axis]) * sp.quant[axis];
    return (uint32_t)((__float2int_rz(v) & sp.handleMask) | isMax) & sp.mask;
}
// monotone float image of a quantised coordinate: the box the grid and the sweep windows work on
__device__ __forceinline__ float sapDequant(uint32_t q, int axis, const SapParams& sp) {
    return __uint2float_rn(q) / sp.quant[axis] + sp.wmin[axis];
}
// SAP modes: leafMin/leafMax hold the quantised bounds (bit patterns), effMin/effMax their float image
__device__ __forceinline__ void sapSetAabb(const BodyArrays& B, int i, f3 mn, f3 mx, const SapParams& sp) {
    uint32_t q0 = sapQuantize1(mn.x, 0, 0, sp), q1 = sapQuantize1(mn.y, 1, 0, sp), q2 = sapQuantize1(mn.z, 2, 0, sp);
    uint32_t r0 = sapQuantize1(mx.x, 0, 1, sp), r1 = sapQuantize1(mx.y, 1, 1, sp), r2 = sapQuantize1(mx.z, 2, 1, sp);
    B.leafMin[i] = make_float4(__uint_as_float(q0), __uint_as_float(q1), __uint_as_float(q2), 0.f);
    B.leafMax[i] = make_float4(__uint_as_float(r0), __uint_as_float(r1), __uint_as_float(r2), 0.f);
    B.effMin[i] = make_float4(sapDequant(q0, 0, sp), sapDequant(q1, 1, sp), sapDequant(q2, 2, sp), 0.f);
    B.effMax[i] = make_float4(sapDequant(r0, 0, sp), sapDequant(r1, 1, sp), sapDequant(r2, 2, sp), 0.f);
}
// overlap of the quantised boxes (the index-order test of bp/AxisSweep3Internal.java:174-195 in value form)
__device__ __forceinline__ bool sapOverlap(float4 aqmin, float4 aqmax, float4 bqmin, float4 bqmax) {
    return !(__float_as_uint(aqmax.x) < __float_as_uint(bqmin.x) || __float_as_uint(bqmax.x) < __float_as_uint(aqmin.x) ||
             __float_as_uint(aqmax.y) < __float_as_uint(bqmin.y) || __float_as_uint(bqmax.y) < __float_as_uint(aqmin.y) ||
             __float_as_uint(aqmax.z) < __float_as_uint(bqmin.z) || __float_as_uint(bqmax.z) < __float_as_uint(aqmin.z));
}

// BroadphaseInterface.setAabb for one proxy.  mode 0: bp/SimpleBroadphase.java:112-116;
// mode 1: bp/DbvtBroadphase.java:196-228 with the in-place Expand/SignedExpand of bp/Dbvt.java:157-169.
__device__ __forceinline__ void setAabbState(const BodyArrays& B, int i, f3 mn, f3 mx, int mode, int step, float dbvtMargin,
                                             float predicted, uint8_t& flags, const SapParams& sp) {
    if (mode >= 2) {  // bp/AxisSweep3Internal.java:591-594 setAabb -> updateHandle: new quantised bounds
        sapSetAabb(B, i, mn, mx, sp);
        B.lastSet[i] = step;
        return;
    }
    if (mode == 1) {
        f3 lmin = mk3(B.leafMin[i].x, B.leafMin[i].y, B.leafMin[i].z);
        f3 lmax = mk3(B.leafMax[i].x, B.leafMax[i].y, B.leafMax[i].z);
        f3 amin = mn, amax = mx;
        if (flags & BF_INFIXED) {
            lmin = amin; lmax = amax;
            flags &= ~BF_INFIXED;
        } else if (aabbIntersect(lmin, lmax, amin, amax)) {
            f3 emin = mk3(B.effMin[i].x, B.effMin[i].y, B.effMin[i].z);
            f3 emax = mk3(B.effMax[i].x, B.effMax[i].y, B.effMax[i].z);
            f3 delta = scl3(add3(mn, mx), 0.5f);
            f3 center = scl3(add3(emin, emax), 0.5f);  // bp/DbvtAabbMm.java:66-70
            delta = scl3(sub3(delta, center), predicted);
            bool contain = (lmin.x <= amin.x) && (lmin.y <= amin.y) && (lmin.z <= amin.z) && (lmax.x >= amax.x) &&
                           (lmax.y >= amax.y) && (lmax.z >= amax.z);
            if (!contain) {
                f3 e = mk3(dbvtMargin, dbvtMargin, dbvtMargin);
                amin = sub3(amin, e);
                amax = add3(amax, e);
                if (delta.x > 0) amax.x += delta.x; else amin.x += delta.x;
                if (delta.y > 0) amax.y += delta.y; else amin.y += delta.y;
                if (delta.z > 0) amax.z += delta.z; else amin.z += delta.z;
                lmin = amin; lmax = amax;
            }
        } else {
            lmin = amin; lmax = amax;  // teleporting
        }
        B.leafMin[i] = make_float4(lmin.x, lmin.y, lmin.z, 0.f);
        B.leafMax[i] = make_float4(lmax.x, lmax.y, lmax.z, 0.f);
        mn = amin; mx = amax;  // proxy.aabb aliases the (possibly expanded) volume
    }
    B.effMin[i] = make_float4(mn.x, mn.y, mn.z, 0.f);
    B.effMax[i] = make_float4(mx.x, mx.y, mx.z, 0.f);
    B.lastSet[i] = step;
}

// ---- grid over the two non-sweep axes, derived on the device (no host sync) ------------------------------------------------
// Reduction state of one thread / block: the largest extent of a non-static proxy (= the cell size: gridded proxies are never
// larger than a cell) and the min-corner bounds of the non-static proxies (the grid's origin and span).  Static proxies that
// are small enough for the grid but lie outside those bounds are CLAMPED into the border cells by cellOf / quantX: both maps
// are monotone, so two proxies whose min corners are closer than a cell still land in the same or adjacent cells.
struct BoundsAcc {
    float extY, extZ;
    uint32_t kyMin, kzMin, kyMax, kzMax, kxMin, kxMax;
    __device__ __forceinline__ void init() {
        extY = extZ = 0.f;
        kyMin = kzMin = kxMin = 0xffffffffu;
        kyMax = kzMax = kxMax = 0u;
    }
    // a: effMin, b: effMax of an alive, non-static proxy
    __device__ __forceinline__ void add(float4 a, float4 b) {
        const float ey = b.y - a.y, ez = b.z - a.z;
        if (ey == ey && ey < 1e29f) extY = ey;
        if (ez == ez && ez < 1e29f) extZ = ez;
        if (ey < 1e29f && ez < 1e29f && fabsf(a.x) < 1e29f && fabsf(a.y) < 1e29f && fabsf(a.z) < 1e29f) {  // false for NaN
            kyMin = kyMax = floatKey(a.y);
            kzMin = kzMax = floatKey(a.z);
            kxMin = kxMax = floatKey(a.x);
        }
    }
};

// Grid parameters from the finished reductions (one thread).
__device__ __forceinline__ void finishGrid(StepCounters* ctr, GridParams* grid, int numWorlds, int maxRows) {
    // gridded proxies have extents <= the largest dynamic extent (limit); the cell is 5 % larger, which
    // absorbs the rounding of the row computation so overlapping proxies always sit in adjacent rows
    const float limitY = __uint_as_float(*(volatile uint32_t*)&ctr->extYBits), limitZ = __uint_as_float(*(volatile uint32_t*)&ctr->extZBits);
    const float cellY = limitY * 1.05f + 1e-6f, cellZ = limitZ * 1.05f + 1e-6f;
    GridParams g;
    uint32_t a = ~*(volatile uint32_t*)&ctr->minYKey, b = *(volatile uint32_t*)&ctr->maxYKey;
    uint32_t c = ~*(volatile uint32_t*)&ctr->minZKey, d = *(volatile uint32_t*)&ctr->maxZKey;
    float y0 = 0.f, y1 = 0.f, z0 = 0.f, z1 = 0.f;
    if (a != 0xffffffffu) { y0 = keyFloat(a); y1 = keyFloat(b); z0 = keyFloat(c); z1 = keyFloat(d); }
    if (!(y1 - y0 >= 0.f) || !(z1 - z0 >= 0.f)) { y0 = y1 = z0 = z1 = 0.f; }  // never loop on NaN/inf bounds
    // coarsen until numWorlds*ny*nz fits the row table (cells only need to be >= the gridded extents)
    float cy = cellY, cz = cellZ;
    int ny = 2, nz = 2;
    for (int guard = 0; guard < 512; guard++) {
        float fy = floorf((y1 - y0) / cy) + 2.f, fz = floorf((z1 - z0) / cz) + 2.f;
        ny = nz = 2;
        if (fy < 16384.f && fz < 16384.f) {
            ny = (int)fy; nz = (int)fz;
            if ((long long)ny * nz * numWorlds <= (long long)maxRows) break;
        }
        cy *= 1.5f; cz *= 1.5f;
    }
    g.y0 = y0; g.z0 = z0;
    g.cellY = cy; g.cellZ = cz;
    g.invCellY = 1.0f / cy; g.invCellZ = 1.0f / cz;
    g.ny = ny; g.nz = nz;
    g.rowsPerWorld = ny * nz;
    g.nrows = ny * nz * numWorlds;
    g.numWorlds = numWorlds;
    // sort key = row << xbits | qx.  12 bits of x when the rows fit 12 bits; fewer when that keeps the key inside 24 bits
    // (one radix pass less); 12 again when the rows alone need more than 18 bits (four passes either way).
    int rowBits = 1;
    while ((1u << rowBits) < (uint32_t)(g.nrows + numWorlds + 2)) rowBits++;
    g.xbits = rowBits <= 12 ? 12 : (rowBits <= 18 ? 24 - rowBits : 12);
    g.xmask = (1u << g.xbits) - 1u;
    g.xmaxf = (float)g.xmask;
    {
        uint32_t xa = ~*(volatile uint32_t*)&ctr->minXKey, xb = *(volatile uint32_t*)&ctr->maxXKey;
        float x0 = 0.f, x1 = 0.f;
        if (xa != 0xffffffffu) { x0 = keyFloat(xa); x1 = keyFloat(xb); }
        float span = x1 - x0;
        g.x0 = x0;
        g.invX = (span > 0.f && span < 1e30f) ? g.xmaxf / span : 0.f;
    }
    *grid = g;
}

// Block reduction of a BoundsAcc + one set of atomics per block; the last block to arrive derives the grid.  Every thread of
// a 256-thread block must call it.
__device__ __forceinline__ void reduceBoundsAndFinish(BoundsAcc acc, StepCounters* ctr, GridParams* grid, uint32_t* ticket,
                                                      int numWorlds, int maxRows) {
    for (int o = 16; o > 0; o >>= 1) {
        acc.extY = fmaxf(acc.extY, __shfl_xor_sync(0xffffffffu, acc.extY, o));
        acc.extZ = fmaxf(acc.extZ, __shfl_xor_sync(0xffffffffu, acc.extZ, o));
        acc.kyMin = min(acc.kyMin, __shfl_xor_sync(0xffffffffu, acc.kyMin, o));
        acc.kzMin = min(acc.kzMin, __shfl_xor_sync(0xffffffffu, acc.kzMin, o));
        acc.kxMin = min(acc.kxMin, __shfl_xor_sync(0xffffffffu, acc.kxMin, o));
        acc.kyMax = max(acc.kyMax, __shfl_xor_sync(0xffffffffu, acc.kyMax, o));
        acc.kzMax = max(acc.kzMax, __shfl_xor_sync(0xffffffffu, acc.kzMax, o));
        acc.kxMax = max(acc.kxMax, __shfl_xor_sync(0xffffffffu, acc.kxMax, o));
    }
    __shared__ float sf[2][8];
    __shared__ uint32_t su[6][8];
    __shared__ bool isLast;
    if ((threadIdx.x & 31) == 0) {
        const int w = threadIdx.x >> 5;
        sf[0][w] = acc.extY; sf[1][w] = acc.extZ;
        su[0][w] = acc.kyMin; su[1][w] = acc.kzMin; su[2][w] = acc.kxMin; su[3][w] = acc.kyMax; su[4][w] = acc.kzMax; su[5][w] = acc.kxMax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            acc.extY = fmaxf(acc.extY, sf[0][w]); acc.extZ = fmaxf(acc.extZ, sf[1][w]);
            acc.kyMin = min(acc.kyMin, su[0][w]); acc.kzMin = min(acc.kzMin, su[1][w]); acc.kxMin = min(acc.kxMin, su[2][w]);
            acc.kyMax = max(acc.kyMax, su[3][w]); acc.kzMax = max(acc.kzMax, su[4][w]); acc.kxMax = max(acc.kxMax, su[5][w]);
        }
        // extents are >= 0, so their float bits order as uints; counters are zero-initialised, so minima are kept as maxima of
        // the complemented key
        if (acc.extY > 0.f) atomicMax(&ctr->extYBits, __float_as_uint(acc.extY));
        if (acc.extZ > 0.f) atomicMax(&ctr->extZBits, __float_as_uint(acc.extZ));
        if (acc.kyMin != 0xffffffffu) {
            atomicMax(&ctr->minYKey, ~acc.kyMin); atomicMax(&ctr->minZKey, ~acc.kzMin); atomicMax(&ctr->minXKey, ~acc.kxMin);
            atomicMax(&ctr->maxYKey, acc.kyMax); atomicMax(&ctr->maxZKey, acc.kzMax); atomicMax(&ctr->maxXKey, acc.kxMax);
        }
        __threadfence();
        const uint32_t t = atomicAdd(ticket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (isLast && threadIdx.x == 0) {
        __threadfence();
        finishGrid(ctr, grid, numWorlds, maxRows);
    }
}

// k_aabb: one thread per proxy slot.
//   * optionally repacks freshly uploaded SoA transform planes into the float4 rows (staging != null),
//   * active proxies: shape AABB, +-threshold, overflow guard, setAabb state machine; per proxy a host-supplied AABB
//     (b2c_set_aabbs: 6 planes of n floats + mask) takes the place of the shape AABB,
//   * forPairs (the launch that opens a pair calculation): dbvt-mode proxies that were not updated during this step move to
//     the fixed set (bp/DbvtBroadphase.java:96-111: eff is kept, the leaf volume becomes eff),
//   * doBounds: the grid of this step is derived from the non-static proxies (reduceBoundsAndFinish) — no separate pass
//     over the AABBs,
//   * owner != null (one world partitioned over several GPUs): only the proxies this rank owns are touched.
__global__ void __launch_bounds__(256)
k_aabb(BodyArrays B, const ShapeDev* __restrict__ shapes, int n, const float* __restrict__ staging, int stagingStride,
       int stagingCount, const float* __restrict__ extAabb, const uint8_t* __restrict__ extMask, int extStride, int mode,
       const int* __restrict__ stepPtr, float threshold, float dbvtMargin, float predicted, int doUpdate, StepCounters* ctr,
       SapParams sp, int forPairs, int doBounds, GridParams* grid, int numWorlds, int maxRows, const uint8_t* __restrict__ owner,
       int myRank) {
    const int step = *stepPtr;  // device-resident step index: the same captured launch serves every step
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    BoundsAcc acc;
    acc.init();
    if (i < n && (!owner || owner[i] == (uint8_t)myRank)) {
        uint8_t flags = B.flags[i];
        if (flags & BF_ALIVE) {
            if (staging && i < stagingCount) {
                const float* p = staging + i;
                float4 r0 = make_float4(p[0], p[(size_t)stagingStride], p[2 * (size_t)stagingStride], p[9 * (size_t)stagingStride]);
                float4 r1 = make_float4(p[3 * (size_t)stagingStride], p[4 * (size_t)stagingStride], p[5 * (size_t)stagingStride],
                                        p[10 * (size_t)stagingStride]);
                float4 r2 = make_float4(p[6 * (size_t)stagingStride], p[7 * (size_t)stagingStride], p[8 * (size_t)stagingStride],
                                        p[11 * (size_t)stagingStride]);
                B.xf4[3 * (size_t)i] = r0;
                B.xf4[3 * (size_t)i + 1] = r1;
                B.xf4[3 * (size_t)i + 2] = r2;
            }
            bool updated = false;
            // per proxy: a host-supplied AABB (setAabb) wins for the proxies it names; every other active proxy is still
            // updated from its transform when updateAabbs was asked for (the two are independent in the reference)
            if (extAabb && extMask[i]) {
                f3 mn = mk3(extAabb[i], extAabb[i + (size_t)extStride], extAabb[i + 2 * (size_t)extStride]);
                f3 mx = mk3(extAabb[i + 3 * (size_t)extStride], extAabb[i + 4 * (size_t)extStride],
                            extAabb[i + 5 * (size_t)extStride]);
                setAabbState(B, i, mn, mx, mode, step, dbvtMargin, predicted, flags, sp);
                updated = true;
            } else if (doUpdate && (flags & BF_ACTIVE)) {
                Xf t;
                {
                    float4 r0 = B.xf4[3 * (size_t)i], r1 = B.xf4[3 * (size_t)i + 1], r2 = B.xf4[3 * (size_t)i + 2];
                    t.m[0][0] = r0.x; t.m[0][1] = r0.y; t.m[0][2] = r0.z;
                    t.m[1][0] = r1.x; t.m[1][1] = r1.y; t.m[1][2] = r1.z;
                    t.m[2][0] = r2.x; t.m[2][1] = r2.y; t.m[2][2] = r2.z;
                    t.o = mk3(r0.w, r1.w, r2.w);
                }
                ShapeDev s = shapes[B.shape[i]];
                f3 mn, mx;
                shapeAabb(s, t, mn, mx);
                f3 ct = mk3(threshold, threshold, threshold);  // disp/CollisionWorld.java:203-207
                mn = sub3(mn, ct);
                mx = add3(mx, ct);
                f3 d = sub3(mx, mn);
                if ((flags & BF_STATIC) || (len2_3(d) < 1e12f)) {  // disp/CollisionWorld.java:212-214
                    setAabbState(B, i, mn, mx, mode, step, dbvtMargin, predicted, flags, sp);
                    updated = true;
                } else {
                    flags = (uint8_t)((flags | BF_OVERFLOW) & ~BF_ACTIVE);  // reference: DISABLE_SIMULATION (:217)
                }
            }
            if (forPairs && mode == 1 && !updated && !(flags & BF_INFIXED) && B.lastSet[i] < step) {
                flags |= BF_INFIXED;
                B.leafMin[i] = B.effMin[i];
                B.leafMax[i] = B.effMax[i];
            }
            B.flags[i] = flags;
            if (doBounds && !(flags & BF_STATIC)) acc.add(B.effMin[i], B.effMax[i]);
        }
    }
    if (doBounds) reduceBoundsAndFinish(acc, ctr, grid, &ctr->aabbTicket, numWorlds, maxRows);
}

// k_bounds: the same reduction over an explicit list of proxies (one world partitioned over several GPUs: the proxies this
// rank sweeps are the ones it owns plus the halo it imported, which k_aabb has not seen).
__global__ void __launch_bounds__(256)
k_bounds(BodyArrays B, const uint32_t* __restrict__ list, const uint32_t* __restrict__ nPtr, int numWorlds, int maxRows,
         StepCounters* ctr, GridParams* grid) {
    const uint32_t n = *nPtr;
    BoundsAcc acc;
    acc.init();
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t i = list[t];
        const uint8_t flags = B.flags[i];
        if ((flags & BF_ALIVE) && !(flags & BF_STATIC)) {
            BoundsAcc one;
            one.init();
            one.add(B.effMin[i], B.effMax[i]);
            acc.extY = fmaxf(acc.extY, one.extY); acc.extZ = fmaxf(acc.extZ, one.extZ);
            acc.kyMin = min(acc.kyMin, one.kyMin); acc.kzMin = min(acc.kzMin, one.kzMin); acc.kxMin = min(acc.kxMin, one.kxMin);
            acc.kyMax = max(acc.kyMax, one.kyMax); acc.kzMax = max(acc.kzMax, one.kzMax); acc.kxMax = max(acc.kxMax, one.kxMax);
        }
    }
    reduceBoundsAndFinish(acc, ctr, grid, &ctr->boundsTicket, numWorlds, maxRows);
}

// cell coordinate of a min corner; monotone in v, and corners closer than one (unslackened) cell land in
// the same or adjacent cells.
__device__ __forceinline__ int cellOf(float v, float v0, float inv, int ncell) {
    int c = (int)floorf((v - v0) * inv);
    return c < 0 ? 0 : (c >= ncell ? ncell - 1 : c);
}

// sweep coordinate: monotone in x (float ops are monotone), clamped to [0, xmaxf]
__device__ __forceinline__ uint32_t quantX(float x, float x0, float invX, float xmaxf) {
    float q = floorf((x - x0) * invX);
    q = q < 0.f ? 0.f : (q > xmaxf ? xmaxf : q);
    return (q == q) ? (uint32_t)q : 0u;
}

__device__ __forceinline__ bool filterPass(uint32_t fa, uint32_t fb) {  // bp/HashedOverlappingPairCache.java:179-188
    return ((fa & 0xffffu) & (fb >> 16)) != 0 && ((fb & 0xffffu) & (fa >> 16)) != 0;
}

}  // namespace b2c
