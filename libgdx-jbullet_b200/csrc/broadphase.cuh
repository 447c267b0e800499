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

// k_aabb: one thread per proxy slot.
//   * optionally repacks freshly uploaded SoA transform planes into the float4 rows (staging != null),
//   * active proxies: shape AABB, +-threshold, overflow guard, setAabb state machine,
//   * dbvt mode: proxies that were not updated this step move to the fixed set,
//   * every alive non-static proxy contributes its y/z extent to the grid cell size.
// extAabb != null selects the "host supplied AABBs" path (b2c_set_aabbs): 6 planes of n floats + mask.
__global__ void __launch_bounds__(256)
k_aabb(BodyArrays B, const ShapeDev* __restrict__ shapes, int n, const float* __restrict__ staging, int stagingStride,
       int stagingCount, const float* __restrict__ extAabb, const uint8_t* __restrict__ extMask, int extStride, int mode,
       const int* __restrict__ stepPtr, float threshold, float dbvtMargin, float predicted, int doUpdate, StepCounters* ctr,
       SapParams sp) {
    const int step = *stepPtr;  // device-resident step index: the same captured launch serves every step
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    float extY = 0.f, extZ = 0.f;
    if (i < n) {
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
            // per proxy: a host-supplied AABB (setAabb) wins for the proxies it names; every other active proxy is still
            // updated from its transform when updateAabbs was asked for (the two are independent in the reference)
            if (extAabb && extMask[i]) {
                f3 mn = mk3(extAabb[i], extAabb[i + (size_t)extStride], extAabb[i + 2 * (size_t)extStride]);
                f3 mx = mk3(extAabb[i + 3 * (size_t)extStride], extAabb[i + 4 * (size_t)extStride],
                            extAabb[i + 5 * (size_t)extStride]);
                setAabbState(B, i, mn, mx, mode, step, dbvtMargin, predicted, flags, sp);
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
                } else {
                    flags = (uint8_t)((flags | BF_OVERFLOW) & ~BF_ACTIVE);  // reference: DISABLE_SIMULATION (:217)
                }
            }
            B.flags[i] = flags;
            if (!(flags & BF_STATIC)) {
                float4 a = B.effMin[i], b = B.effMax[i];
                float ey = b.y - a.y, ez = b.z - a.z;
                if (ey == ey && ey < 1e29f) extY = ey;
                if (ez == ez && ez < 1e29f) extZ = ez;
            }
        }
    }
    // block max of the dynamic extents -> one atomic per block (extents >= 0, so float bits order as uints)
    for (int o = 16; o > 0; o >>= 1) {
        extY = fmaxf(extY, __shfl_xor_sync(0xffffffffu, extY, o));
        extZ = fmaxf(extZ, __shfl_xor_sync(0xffffffffu, extZ, o));
    }
    __shared__ float sy[8], sz[8];
    if ((threadIdx.x & 31) == 0) { sy[threadIdx.x >> 5] = extY; sz[threadIdx.x >> 5] = extZ; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) { extY = fmaxf(extY, sy[w]); extZ = fmaxf(extZ, sz[w]); }
        if (extY > 0.f) atomicMax(&ctr->extYBits, __float_as_uint(extY));
        if (extZ > 0.f) atomicMax(&ctr->extZBits, __float_as_uint(extZ));
    }
}

// In dbvt mode, proxies that did not receive setAabb during this step (but did the step before) move to
// the fixed set: eff is kept and the leaf volume becomes eff (bp/DbvtBroadphase.java:96-111).  This only
// flips a flag and copies 32 B, so it rides in k_bounds.
//
// k_bounds: classify large proxies (static and larger than the dynamic cell, or non-finite), reduce the
// min-corner bounds of the gridded ones; the last block to finish derives the grid.
__global__ void __launch_bounds__(256)
k_bounds(BodyArrays B, int n, int mode, const int* __restrict__ stepPtr, int numWorlds, int maxRows, StepCounters* ctr,
         GridParams* grid) {
    const int step = *stepPtr;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    // gridded proxies have extents <= the largest dynamic extent (limit); the cell is 5 % larger, which
    // absorbs the rounding of the row computation so overlapping proxies always sit in adjacent rows
    const float limitY = __uint_as_float(ctr->extYBits), limitZ = __uint_as_float(ctr->extZBits);
    float cellY = limitY * 1.05f + 1e-6f;
    float cellZ = limitZ * 1.05f + 1e-6f;
    uint32_t kyMin = 0xffffffffu, kzMin = 0xffffffffu, kyMax = 0u, kzMax = 0u, kxMin = 0xffffffffu, kxMax = 0u;
    if (i < n) {
        uint8_t flags = B.flags[i];
        if (flags & BF_ALIVE) {
            if (mode == 1 && !(flags & BF_INFIXED) && B.lastSet[i] < step) {
                flags |= BF_INFIXED;
                B.flags[i] = flags;
                B.leafMin[i] = B.effMin[i];
                B.leafMax[i] = B.effMax[i];
            }
            float4 a = B.effMin[i], b = B.effMax[i];
            float ey = b.y - a.y, ez = b.z - a.z;
            bool large = !(ey <= limitY) || !(ez <= limitZ) || !(fabsf(a.y) < 1e29f) || !(fabsf(a.z) < 1e29f) || !(fabsf(a.x) < 1e29f);
            if (!large) {
                kyMin = kyMax = floatKey(a.y);
                kzMin = kzMax = floatKey(a.z);
                kxMin = kxMax = floatKey(a.x);
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        kyMin = min(kyMin, __shfl_xor_sync(0xffffffffu, kyMin, o));
        kzMin = min(kzMin, __shfl_xor_sync(0xffffffffu, kzMin, o));
        kyMax = max(kyMax, __shfl_xor_sync(0xffffffffu, kyMax, o));
        kzMax = max(kzMax, __shfl_xor_sync(0xffffffffu, kzMax, o));
        kxMin = min(kxMin, __shfl_xor_sync(0xffffffffu, kxMin, o));
        kxMax = max(kxMax, __shfl_xor_sync(0xffffffffu, kxMax, o));
    }
    __shared__ uint32_t s[6][8];
    __shared__ bool isLast;
    if ((threadIdx.x & 31) == 0) {
        int w = threadIdx.x >> 5;
        s[0][w] = kyMin; s[1][w] = kzMin; s[2][w] = kyMax; s[3][w] = kzMax; s[4][w] = kxMin; s[5][w] = kxMax;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; w++) {
            kyMin = min(kyMin, s[0][w]); kzMin = min(kzMin, s[1][w]);
            kyMax = max(kyMax, s[2][w]); kzMax = max(kzMax, s[3][w]);
            kxMin = min(kxMin, s[4][w]); kxMax = max(kxMax, s[5][w]);
        }
        if (kyMin != 0xffffffffu) {
            // counters are zero-initialised, so the minima are kept as maxima of the complemented key
            atomicMax(&ctr->minYKey, ~kyMin); atomicMax(&ctr->minZKey, ~kzMin);
            atomicMax(&ctr->maxYKey, kyMax); atomicMax(&ctr->maxZKey, kzMax);
            atomicMax(&ctr->minXKey, ~kxMin); atomicMax(&ctr->maxXKey, kxMax);
        }
        __threadfence();
        uint32_t t = atomicAdd(&ctr->boundsTicket, 1u);
        isLast = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (isLast && threadIdx.x == 0) {
        __threadfence();
        GridParams g;
        uint32_t a = ~*(volatile uint32_t*)&ctr->minYKey, b = *(volatile uint32_t*)&ctr->maxYKey;
        uint32_t c = ~*(volatile uint32_t*)&ctr->minZKey, d = *(volatile uint32_t*)&ctr->maxZKey;
        float y0 = 0.f, y1 = 0.f, z0 = 0.f, z1 = 0.f;
        if (a != 0xffffffffu) { y0 = keyFloat(a); y1 = keyFloat(b); z0 = keyFloat(c); z1 = keyFloat(d); }
        if (!(y1 - y0 >= 0.f) || !(z1 - z0 >= 0.f)) { y0 = y1 = z0 = z1 = 0.f; }  // never loop on NaN/inf bounds
        // coarsen until numWorlds*ny*nz fits the row table (cells only need to be >= the gridded extents)
        float cy = cellY, cz = cellZ;
        int ny, nz;
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
        {
            uint32_t xa = ~*(volatile uint32_t*)&ctr->minXKey, xb = *(volatile uint32_t*)&ctr->maxXKey;
            float x0 = 0.f, x1 = 0.f;
            if (xa != 0xffffffffu) { x0 = keyFloat(xa); x1 = keyFloat(xb); }
            float span = x1 - x0;
            g.x0 = x0;
            g.invX = (span > 0.f && span < 1e30f) ? 4095.0f / span : 0.f;
        }
        *grid = g;
    }
}

// cell coordinate of a min corner; monotone in v, and corners closer than one (unslackened) cell land in
// the same or adjacent cells.
__device__ __forceinline__ int cellOf(float v, float v0, float inv, int ncell) {
    int c = (int)floorf((v - v0) * inv);
    return c < 0 ? 0 : (c >= ncell ? ncell - 1 : c);
}

// 12-bit sweep coordinate: monotone in x (float ops are monotone), clamped
__device__ __forceinline__ uint32_t quantX(float x, float x0, float invX) {
    float q = floorf((x - x0) * invX);
    q = q < 0.f ? 0.f : (q > 4095.f ? 4095.f : q);
    return (q == q) ? (uint32_t)q : 0u;
}

// k_keys: 32-bit key and payload for every slot.
__global__ void __launch_bounds__(256)
k_keys(BodyArrays B, int n, const StepCounters* __restrict__ ctr, const GridParams* __restrict__ grid,
       uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int* stepPtr) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *stepPtr = *stepPtr + 1;  // k_aabb and k_bounds of this step have read it (stream order); next step sees +1
    if (i >= n) return;
    GridParams g = *grid;
    // same large criterion as k_bounds
    const float limitY = __uint_as_float(ctr->extYBits), limitZ = __uint_as_float(ctr->extZBits);
    uint8_t flags = B.flags[i];
    uint32_t row;
    uint32_t xk = 0;
    if (!(flags & BF_ALIVE)) {
        row = (uint32_t)(g.nrows + g.numWorlds);
    } else {
        float4 a = B.effMin[i], b = B.effMax[i];
        float ey = b.y - a.y, ez = b.z - a.z;
        bool large = !(ey <= limitY) || !(ez <= limitZ) || !(fabsf(a.y) < 1e29f) || !(fabsf(a.z) < 1e29f) || !(fabsf(a.x) < 1e29f);
        if (large) {
            row = (uint32_t)(g.nrows + B.world[i]);  // one row of large proxies per world
        } else {
            int cy = cellOf(a.y, g.y0, g.invCellY, g.ny), cz = cellOf(a.z, g.z0, g.invCellZ, g.nz);
            row = (uint32_t)(B.world[i] * g.rowsPerWorld + cy * g.nz + cz);
        }
        xk = quantX(a.x, g.x0, g.invX);
    }
    keys[i] = (row << 12) | xk;
    vals[i] = (uint32_t)i;
}

// k_gather: sorted AABB SoA for 128-bit loads in the sweep, and the row start table.
//   smin[j] = (min.x, min.y, min.z, bodyIndex) ; smax[j] = (max.x, max.y, max.z, filter)
// rowStart[r] = first sorted index whose row >= r (rows without proxies get the next row's start).
__global__ void __launch_bounds__(256)
k_gather(BodyArrays B, int n, const uint32_t* keysA, const uint32_t* keysB, const uint32_t* valsA, const uint32_t* valsB,
         const uint32_t* __restrict__ side, const GridParams* __restrict__ grid, float4* __restrict__ smin,
         float4* __restrict__ smax, uint32_t* __restrict__ srow, uint32_t* __restrict__ rowStart, uint32_t* __restrict__ scyz) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t* keys = *side ? keysB : keysA;
    const uint32_t* vals = *side ? valsB : valsA;
    uint32_t k = keys[j];
    uint32_t row = k >> 12;
    uint32_t body = vals[j];
    float4 a = B.effMin[body], b = B.effMax[body];
    a.w = __uint_as_float(body);
    b.w = __uint_as_float(B.filt[body]);
    smin[j] = a;
    smax[j] = b;
    srow[j] = k;  // the whole sorted key: row = k >> 12, qx = k & 4095
    {
        // cell coordinates of the row, once per proxy (the sweep's 9 threads per proxy would each redo the divisions);
        // 0xffffffff marks the rows that are not part of the grid (large proxies, dead slots)
        const uint32_t nrows = (uint32_t)grid->nrows, rpw = (uint32_t)grid->rowsPerWorld, nz = (uint32_t)grid->nz;
        uint32_t c = 0xffffffffu;
        if (row < nrows) {
            const uint32_t rem = row % rpw;
            c = ((rem / nz) << 16) | (rem % nz);
        }
        scyz[j] = c;
    }
    uint32_t prev = j ? (keys[j - 1] >> 12) : 0xffffffffu;
    const uint32_t lastRow = (uint32_t)(grid->nrows + grid->numWorlds) + 1u;
    if (j == 0) {
        for (uint32_t r = 0; r <= row; r++) rowStart[r] = 0;
    } else if (prev != row) {
        for (uint32_t r = prev + 1; r <= row; r++) rowStart[r] = (uint32_t)j;
    }
    if (j == n - 1) {
        for (uint32_t r = row + 1; r <= lastRow; r++) rowStart[r] = (uint32_t)n;
    }
}

__device__ __forceinline__ bool filterPass(uint32_t fa, uint32_t fb) {  // bp/HashedOverlappingPairCache.java:179-188
    return ((fa & 0xffffu) & (fb >> 16)) != 0 && ((fb & 0xffffu) & (fa >> 16)) != 0;
}

// Warp-staged pair append.  A single global counter cannot take one atomic per warp round (same-address
// atomics serialise in L2: ~300 k of them per step cost more than the sweep itself), so every warp stages its
// hits in a private shared-memory buffer (ballot + popc for the slot) and flushes 32+ pairs at a time with ONE
// global atomic and coalesced 8-byte stores.
constexpr int PAIR_STAGE = 96;  // per-warp staging capacity (flush when > 64 are waiting)

struct PairStager {
    uint64_t* buf;      // this warp's shared-memory slice
    uint32_t* rowCnt;   // pairs per uid0 so far this step: the atomicAdd's return value is the pair's slot in its row
    int count;          // warp-uniform
    int uidBits;
    __device__ __forceinline__ void init(uint64_t* warpBuf, uint32_t* rowCounters, int bits) {
        buf = warpBuf; rowCnt = rowCounters; count = 0; uidBits = bits;
    }
    __device__ __forceinline__ void flush(uint64_t* __restrict__ pairKeys, uint32_t maxPairs, StepCounters* ctr) {
        if (count == 0) return;
        const int lane = threadIdx.x & 31;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctr->pairCount, (uint32_t)count);
        base = __shfl_sync(0xffffffffu, base, 0);
        __syncwarp();
        for (int k = lane; k < count; k += 32) {
            uint32_t pos = base + k;
            if (pos < maxPairs) {
                uint64_t key = buf[k];
                uint32_t slot = atomicAdd(&rowCnt[(uint32_t)(key >> uidBits)], 1u);
                pairKeys[pos] = key | ((uint64_t)slot << (2 * uidBits));  // pair_rows.cuh: slot | uid0 | uid1
            } else {
                ctr->pairOverflow = 1;
            }
        }
        __syncwarp();
        count = 0;
    }
    // Final flush of a 256-thread block: the eight warps reserve their output range with ONE atomic on the shared pair
    // counter (the per-warp atomics of the plain flush were the top stall of the sweep: ~250 k same-address atomics per
    // step at 100 k bodies).  Every thread of the block must call it.
    __device__ __forceinline__ void flushBlock(uint64_t* __restrict__ pairKeys, uint32_t maxPairs, StepCounters* ctr,
                                               uint32_t* sCnt /*[8]*/, uint32_t* sBase) {
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        if (lane == 0) sCnt[warp] = (uint32_t)count;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t tot = 0;
            for (int w = 0; w < 8; w++) { uint32_t c = sCnt[w]; sCnt[w] = tot; tot += c; }
            *sBase = tot ? atomicAdd(&ctr->pairCount, tot) : 0u;
        }
        __syncthreads();
        const uint32_t base = *sBase + sCnt[warp];
        for (int k = lane; k < count; k += 32) {
            uint32_t pos = base + k;
            if (pos < maxPairs) {
                uint64_t key = buf[k];
                uint32_t slot = atomicAdd(&rowCnt[(uint32_t)(key >> uidBits)], 1u);
                pairKeys[pos] = key | ((uint64_t)slot << (2 * uidBits));
            } else {
                ctr->pairOverflow = 1;
            }
        }
        count = 0;
    }
    // all 32 lanes call this together
    __device__ __forceinline__ void push(bool hit, uint32_t bodyA, uint32_t bodyB, uint64_t* __restrict__ pairKeys,
                                         uint32_t maxPairs, StepCounters* ctr) {
        uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m == 0) return;
        if (hit) {
            int lane = threadIdx.x & 31;
            uint32_t ua = bodyA + 1u, ub = bodyB + 1u;  // uid = slot + 1 (bp/DbvtBroadphase.java:179)
            uint32_t lo = ua < ub ? ua : ub, hi = ua < ub ? ub : ua;  // bp/HashedOverlappingPairCache.java:292-296
            buf[count + __popc(m & ((1u << lane) - 1u))] = ((uint64_t)lo << uidBits) | hi;
        }
        count += __popc(m);
        if (count > PAIR_STAGE - 32) flush(pairKeys, maxPairs, ctr);
    }
};

// k_sweep: blockIdx.y selects one of the 9 neighbour rows (dy,dz); each lane owns one sorted proxy i and
// walks the x-window of that row: lower_bound(min.x_i) then forward while min.x_j <= max.x_i.  The warp
// advances in lock step and compacts hits with ballot/popc.  The pair is emitted by the member that comes
// first in (min.x, sorted position) order, so every overlapping pair is produced exactly once.
__global__ void __launch_bounds__(256)
k_sweep(int n, const float4* __restrict__ smin, const float4* __restrict__ smax, const uint32_t* __restrict__ srow,
        const uint32_t* __restrict__ rowStart, const GridParams* __restrict__ grid, int uidBits,
        uint64_t* __restrict__ pairKeys, uint32_t* rowCnt, uint32_t maxPairs, StepCounters* ctr, int partLo, int partHi,
        const float4* __restrict__ qmin, const float4* __restrict__ qmax /* SAP modes: quantised bounds per body, else null */,
        const uint32_t* __restrict__ scyz) {
    __shared__ uint64_t stage[8][PAIR_STAGE];
    PairStager st;
    st.init(stage[threadIdx.x >> 5], rowCnt, uidBits);
    const int i = blockIdx.x * blockDim.x + threadIdx.x + partLo;  // this rank's slice [partLo, partHi) of the sorted list
    n = n < partHi ? n : partHi;
    const int nb = blockIdx.y;  // 0..8
    const int dy = nb / 3 - 1, dz = nb % 3 - 1;
    const int ny = grid->ny, nz = grid->nz;
    uint32_t j = 0, end = 0;
    float4 amin = make_float4(0, 0, 0, 0), amax = amin;
    uint32_t xkI = 0, xkMax = 0;
    if (i < n) {
        const uint32_t keyI = srow[i];
        const uint32_t row = keyI >> 12;
        const uint32_t cc = __ldg(scyz + i);
        if (cc != 0xffffffffu) {
            int cy = (int)(cc >> 16) + dy, cz = (int)(cc & 0xffffu) + dz;
            if (cy >= 0 && cy < ny && cz >= 0 && cz < nz) {
                uint32_t r2 = (uint32_t)((int)row + dy * nz + dz);  // same world: cy stays inside [0, ny)
                uint32_t lo = rowStart[r2], hi = rowStart[r2 + 1];
                amin = __ldg(smin + i);
                amax = __ldg(smax + i);
                xkI = keyI & 4095u;
                xkMax = quantX(amax.x, grid->x0, grid->invX);
                if (nb == 4) {
                    lo = (uint32_t)i + 1u;  // same row: everything after i has min.x >= min.x_i (stable sort)
                } else {
                    uint32_t a = lo, b = hi;  // lower_bound of min.x_i inside the neighbour row
                    while (a < b) {
                        uint32_t mid = (a + b) >> 1;
                        if ((__ldg(srow + mid) & 4095u) < xkI) a = mid + 1; else b = mid;
                    }
                    lo = a;
                }
                j = lo; end = hi;
            }
        }
    }
    while (__any_sync(0xffffffffu, j < end)) {
        bool hit = false;
        uint32_t bodyB = 0;
        if (j < end) {
            uint32_t xkJ = __ldg(srow + j) & 4095u;
            if (xkJ > xkMax) {
                j = end;  // window closed
            } else {
                // ties in min.x across rows: only the earlier sorted position emits
                if (xkJ != xkI || j > (uint32_t)i) {
                    float4 bmin = __ldg(smin + j);
                    float4 bmax = __ldg(smax + j);
                    hit = (amin.x <= bmax.x) && (amax.x >= bmin.x) && (amin.y <= bmax.y) && (amax.y >= bmin.y) &&
                          (amin.z <= bmax.z) && (amax.z >= bmin.z) &&
                          filterPass(__float_as_uint(amax.w), __float_as_uint(bmax.w));
                    bodyB = __float_as_uint(bmin.w);
                    // AxisSweep3 modes: the float boxes are a monotone image of the quantised ones (necessary condition);
                    // the pair predicate itself is on the quantised values
                    if (hit && qmin) {
                        const uint32_t bodyA = __float_as_uint(amin.w);
                        hit = sapOverlap(__ldg(qmin + bodyA), __ldg(qmax + bodyA), __ldg(qmin + bodyB), __ldg(qmax + bodyB));
                    }
                }
                j++;
            }
        }
        st.push(hit, __float_as_uint(amin.w), bodyB, pairKeys, maxPairs, ctr);
    }
    __shared__ uint32_t sCnt[8];
    __shared__ uint32_t sBase;
    st.flushBlock(pairKeys, maxPairs, ctr, sCnt, &sBase);
}

// k_large: proxies that do not fit the grid (row == nrows) against every proxy of the same world, and
// against each other once.
__global__ void __launch_bounds__(256)
k_large(int n, const float4* __restrict__ smin, const float4* __restrict__ smax, const uint32_t* __restrict__ rowStart,
        const GridParams* __restrict__ grid, const int* __restrict__ world, int numWorlds, int uidBits,
        uint64_t* __restrict__ pairKeys, uint32_t* rowCnt, uint32_t maxPairs, StepCounters* ctr, int partLo, int partHi,
        int partRank, const float4* __restrict__ qmin, const float4* __restrict__ qmax) {
    __shared__ uint64_t stage[8][PAIR_STAGE];
    PairStager st;
    st.init(stage[threadIdx.x >> 5], rowCnt, uidBits);
    const int nrows = grid->nrows, rpw = grid->rowsPerWorld;
    const uint32_t l0 = rowStart[nrows], l1 = rowStart[nrows + numWorlds];
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) ctr->largeCount = l1 - l0;
    for (uint32_t l = l0 + blockIdx.y; l < l1; l += gridDim.y) {
        float4 amin = __ldg(smin + l), amax = __ldg(smax + l);
        uint32_t bodyA = __float_as_uint(amin.w);
        uint32_t lo = 0, hi = l0, lend = l1;
        if (numWorlds > 1) {
            int w = world[bodyA];
            lo = rowStart[w * rpw];
            hi = rowStart[(w + 1) * rpw];
            lend = rowStart[nrows + w + 1];
        }
        // gridded proxies of the same world, then the large ones of the same world after l
        uint32_t total = (hi - lo) + (lend - (l + 1));
        for (uint32_t t0 = blockIdx.x * blockDim.x; t0 < total; t0 += gridDim.x * blockDim.x) {
            uint32_t t = t0 + threadIdx.x;
            bool hit = false;
            uint32_t bodyB = 0;
            if (t < total) {
                uint32_t j = t < (hi - lo) ? lo + t : (l + 1) + (t - (hi - lo));
                float4 bmin = __ldg(smin + j), bmax = __ldg(smax + j);
                bodyB = __float_as_uint(bmin.w);
                hit = (amin.x <= bmax.x) && (amax.x >= bmin.x) && (amin.y <= bmax.y) && (amax.y >= bmin.y) &&
                      (amin.z <= bmax.z) && (amax.z >= bmin.z) && filterPass(__float_as_uint(amax.w), __float_as_uint(bmax.w));
                if (hit && qmin) hit = sapOverlap(__ldg(qmin + bodyA), __ldg(qmax + bodyA), __ldg(qmin + bodyB), __ldg(qmax + bodyB));
                if (hit && numWorlds > 1 && j >= l0) hit = world[bodyB] == world[bodyA];
                // partitioned world: a (large, gridded) pair belongs to the rank whose slice holds the gridded member;
                // (large, large) pairs belong to rank 0
                if (hit) hit = (j >= l0) ? (partRank == 0) : ((int)j >= partLo && (int)j < partHi);
            }
            st.push(hit, bodyA, bodyB, pairKeys, maxPairs, ctr);
        }
    }
    __shared__ uint32_t sCnt[8];
    __shared__ uint32_t sBase;
    st.flushBlock(pairKeys, maxPairs, ctr, sCnt, &sBase);
}

}  // namespace b2c
