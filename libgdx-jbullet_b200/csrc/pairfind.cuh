// pairfind.cuh — stage 1 of the path, overlap-pair finding over the effective AABBs (device code only).
//
// Replaces bp/DbvtBroadphase.java:89-150 collide / bp/SimpleBroadphase.java:81-110 / the AxisSweep3 edge sort: the pair
// set is "all filtered overlaps of the per-proxy effective AABBs" (DESIGN.md §1), rebuilt from scratch every step:
//   k_keys     32-bit key = row << xbits | qx per proxy (row = coarse (y, z) grid cell, qx = quantised min.x), the radix
//              histograms of all digits in the same pass, and (last block) their exclusive scans          [fused rs_hist / rs_scan]
//   rs_pass    one onesweep pass per 8-bit digit (radix_sort.cuh); constant digits are skipped on the device
//   k_gather   sorted AABB SoA (float4 min | proxy, float4 max | filter), sorted keys, row start table
//   k_sweep    each warp takes 32 consecutive sorted proxies; for each of the 9 neighbour rows the UNION of their x-windows
//              is a contiguous piece of the sorted arrays, staged once per warp in shared memory with 1-D bulk copies (TMA,
//              cp.async.bulk + mbarrier); every lane then finds and walks its own window in shared memory
//   k_large    proxies that do not fit the grid (planes, meshes, big statics) against everything in their world
// Pairs are emitted through PairStager (ballot/popc compaction, one global atomic per block) in arbitrary order and put into
// canonical (uid0, uid1) order by pair_rows.cuh.
#pragma once
#include "broadphase.cuh"
#include "radix_sort.cuh"

namespace b2c {

// One world partitioned over several GPUs by slabs along `axis`: planes[0..nplanes) ascending, region r is
// [planes[r-1], planes[r]).  A pair belongs to the region that holds max(min_a[axis], min_b[axis]) — a coordinate that lies
// inside BOTH boxes, so the owning rank holds both proxies (owned or halo).
struct SlabFilter {
    int enabled, axis, rank, nplanes;
    float planes[15];
    __device__ __forceinline__ int region(float v) const {
        int r = 0;
        for (int k = 0; k < nplanes; k++) r += (v >= planes[k]) ? 1 : 0;
        return r;
    }
    __device__ __forceinline__ bool owns(float4 amin, float4 bmin) const {
        if (!enabled) return true;
        const float a = axis == 0 ? amin.x : (axis == 1 ? amin.y : amin.z);
        const float b = axis == 0 ? bmin.x : (axis == 1 ? bmin.y : bmin.z);
        return region(fmaxf(a, b)) == rank;
    }
    // does the box [mn, mx] touch region r?
    __device__ __forceinline__ bool touches(float mn, float mx, int r) const {
        const int lo = region(mn), hi = region(mx);
        return r >= lo && r <= hi;
    }
};

// k_keys: key and payload of every proxy of the step (list == null: slots 0..n-1), with the histograms of the four radix digits.
// The last block to finish turns the histograms into exclusive digit offsets and the skip flags (a digit with a single
// non-empty bin is the identity), so no separate histogram / scan launch is needed.  `st` was cleared by a memset node.
__global__ void __launch_bounds__(256)
k_keys(BodyArrays B, int nConst, const uint32_t* __restrict__ nPtr, const uint32_t* __restrict__ list, StepCounters* ctr,
       const GridParams* __restrict__ grid, uint32_t* __restrict__ keys, uint32_t* __restrict__ vals, int* stepPtr, RadixState* st,
       int npass, uint32_t* __restrict__ rowCount, uint32_t* __restrict__ slots) {
    __shared__ uint32_t sh[4][256];
    __shared__ bool isLast;
    for (int k = threadIdx.x; k < 4 * 256; k += 256) (&sh[0][0])[k] = 0;
    __syncthreads();
    const uint32_t n = nPtr ? *nPtr : (uint32_t)nConst;
    if (blockIdx.x == 0 && threadIdx.x == 0) *stepPtr = *stepPtr + 1;  // k_aabb of this step has read it (stream order)
    const GridParams g = *grid;
    const float limitY = __uint_as_float(ctr->extYBits), limitZ = __uint_as_float(ctr->extZBits);
    const int lane = threadIdx.x & 31;
    for (uint32_t t0 = blockIdx.x * 256; t0 < n; t0 += gridDim.x * 256) {
        const uint32_t t = t0 + threadIdx.x;
        const bool valid = t < n;
        const uint32_t i = valid ? (list ? list[t] : t) : 0u;
        const uint8_t flags = valid ? B.flags[i] : (uint8_t)0;
        uint32_t row, xk = 0;
        if (!(flags & BF_ALIVE)) {
            row = (uint32_t)(g.nrows + g.numWorlds);  // dead slots sort behind everything
        } else {
            const float4 a = B.effMin[i], b = B.effMax[i];
            const float ey = b.y - a.y, ez = b.z - a.z;
            // too large for a cell (only statics can be: the cell is the largest non-static extent) or not finite
            const bool large = !(ey <= limitY) || !(ez <= limitZ) || !(fabsf(a.y) < 1e29f) || !(fabsf(a.z) < 1e29f) || !(fabsf(a.x) < 1e29f);
            if (large) {
                row = (uint32_t)(g.nrows + B.world[i]);  // one row of large proxies per world
            } else {
                const int cy = cellOf(a.y, g.y0, g.invCellY, g.ny), cz = cellOf(a.z, g.z0, g.invCellZ, g.nz);
                row = (uint32_t)(B.world[i] * g.rowsPerWorld + cy * g.nz + cz);
            }
            xk = quantX(a.x, g.x0, g.invX, g.xmaxf);
        }
        const uint32_t key = (row << g.xbits) | xk;
        if (rowCount) {
            // row-grouped ordering (k_row_order below): the proxy takes a slot in its row; neighbouring proxies mostly share
            // their row, so one atomic per distinct row of the warp
            const uint32_t peers = __match_any_sync(0xffffffffu, valid ? row : 0xffffffffu);
            uint32_t base = 0;
            const int leader = __ffs(peers) - 1;
            if (valid && lane == leader) base = atomicAdd(&rowCount[row], (uint32_t)__popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (valid) {
                keys[t] = key;
                vals[t] = i;
                slots[t] = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
            }
            continue;
        }
        if (valid) {
            keys[t] = key;
            vals[t] = i;
            atomicAdd(&sh[0][key & 255u], 1u);  // low bits of qx: spread over the bins
        }
        // the upper digits are the same for most lanes of a warp (neighbouring proxies share their row): one shared-memory
        // atomic per distinct digit instead of a 32-way conflict
#pragma unroll
        for (int p = 1; p < 4; p++) {
            if (p < npass) {
                const uint32_t d = valid ? ((key >> (8 * p)) & 255u) : 0x100u;
                const uint32_t peers = __match_any_sync(0xffffffffu, d);
                if (valid && lane == __ffs(peers) - 1) atomicAdd(&sh[p][d], (uint32_t)__popc(peers));
            }
        }
    }
    if (rowCount) {  // no radix passes follow: only the count is published
        if (blockIdx.x == 0 && threadIdx.x == 0) st->n = n;
        return;
    }
    __syncthreads();
    for (int k = threadIdx.x; k < npass * 256; k += 256) {
        const uint32_t v = (&sh[0][0])[k];
        if (v) atomicAdd(&(&st->hist[0][0])[k], v);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) isLast = atomicAdd(&ctr->keysTicket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    // exclusive scan of the 256 bins of every digit: warp p scans digit p (8 consecutive bins per lane, shuffles across lanes)
    if (threadIdx.x == 0) st->n = n;
    const int p = threadIdx.x >> 5;
    if (p < npass) {
        uint32_t v[8], sum = 0;
        bool full = false;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v[k] = *(volatile uint32_t*)&st->hist[p][lane * 8 + k];
            full |= (v[k] == n);  // covers n == 0 as well
            sum += v[k];
        }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t tv = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += tv;
        }
        uint32_t run = incl - sum;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            st->hist[p][lane * 8 + k] = run;
            run += v[k];
        }
        if (__any_sync(0xffffffffu, full) && lane == 0) st->skip[p] = 1u;
    }
}

// k_gather: sorted AABB SoA for 128-bit / bulk loads in the sweep, and the row start table.
//   smin[j] = (min.x, min.y, min.z, bodyIndex) ; smax[j] = (max.x, max.y, max.z, filter)
// rowStart[r] = first sorted index whose row >= r (rows without proxies get the next row's start).
__global__ void __launch_bounds__(256)
k_gather(BodyArrays B, const uint32_t* keysA, const uint32_t* keysB, const uint32_t* valsA, const uint32_t* valsB,
         const RadixState* __restrict__ st, int npass, const GridParams* __restrict__ grid, float4* __restrict__ smin,
         float4* __restrict__ smax, uint32_t* __restrict__ skey, uint32_t* __restrict__ rowStart, uint32_t* __restrict__ scyz,
         uint32_t* __restrict__ nSortedOut) {
    const uint32_t n = st->n;
    const int side = rs_side_before(st, npass);  // which ping-pong side the executed passes left the data in
    const uint32_t* keys = side ? keysB : keysA;
    const uint32_t* vals = side ? valsB : valsA;
    const int xbits = grid->xbits;
    const uint32_t nrows = (uint32_t)grid->nrows, rpw = (uint32_t)grid->rowsPerWorld, nz = (uint32_t)grid->nz;
    const uint32_t lastRow = (uint32_t)(grid->nrows + grid->numWorlds) + 1u;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        *nSortedOut = n;
        if (n == 0)
            for (uint32_t r = 0; r <= lastRow; r++) rowStart[r] = 0;
    }
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
        const uint32_t k = keys[j];
        const uint32_t row = k >> xbits;
        const uint32_t body = vals[j];
        float4 a = B.effMin[body], b = B.effMax[body];
        a.w = __uint_as_float(body);
        b.w = __uint_as_float(B.filt[body]);
        smin[j] = a;
        smax[j] = b;
        skey[j] = k;
        // cell coordinates of the row, once per proxy; 0xffffffff marks the rows outside the grid (large proxies, dead slots)
        uint32_t c = 0xffffffffu;
        if (row < nrows) {
            const uint32_t rem = row % rpw;
            c = ((rem / nz) << 16) | (rem % nz);
        }
        scyz[j] = c;
        const uint32_t prev = j ? (keys[j - 1] >> xbits) : 0xffffffffu;
        if (j == 0) {
            for (uint32_t r = 0; r <= row; r++) rowStart[r] = 0;
        } else if (prev != row) {
            for (uint32_t r = prev + 1; r <= row; r++) rowStart[r] = j;
        }
        if (j == n - 1)
            for (uint32_t r = row + 1; r <= lastRow; r++) rowStart[r] = n;
    }
}

// ---- row-grouped ordering: the sorted order without radix passes ------------------------------------------------------------
// The sweep needs the proxies grouped by row and ordered by qx inside a row; the order of equal keys is irrelevant (it only
// decides WHICH member of a pair emits it).  Rows are short (a few hundred proxies at most in a compact world), so instead of
// 3-4 stable onesweep passes over all keys:
//   k_keys          every proxy takes a slot in its row (one atomic per distinct row of a warp)
//   k_row_offsets   exclusive scan of the row counts = the row start table (single pass, decoupled look-back) + longest row
//   k_row_place     proxy -> rowStart[row] + slot (grouped by row, arbitrary order inside)
//   k_row_order     one thread per grouped entry: its place inside the row is the number of row members with a smaller
//                   (qx, proxy index); writes the sorted AABB SoA / keys / cell coordinates directly (the gather is fused)
// Work is sum(len^2) key compares on L1-resident rows — a few microseconds for rows of 100-300.  The host switches back to the
// radix passes when the longest row it has seen makes that quadratic term matter (ROW_ORDER_MAX_LEN).
constexpr uint32_t ROW_ORDER_MAX_LEN = 4096;
constexpr int ROFF_PER = 16, ROFF_TILE = 256 * ROFF_PER;
struct RowOffsetsMisc {
    uint32_t ticket, pad[3];
};

__global__ void __launch_bounds__(256)
k_row_offsets(const uint32_t* __restrict__ rowCount, const GridParams* __restrict__ grid, uint32_t* __restrict__ rowStart,
              uint32_t* status, RowOffsetsMisc* misc, StepCounters* ctr) {
    __shared__ uint32_t sTile, sExcl, warpSum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nR = (uint32_t)(grid->nrows + grid->numWorlds) + 2u;  // rowStart[0 .. lastRow] with lastRow = nrows + numWorlds + 1
    if (threadIdx.x == 0) sTile = atomicAdd(&misc->ticket, 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    if (tile * ROFF_TILE >= nR) return;  // tiles are taken in ticket order, so every tile before a needed one is needed too
    const uint32_t base = tile * ROFF_TILE + threadIdx.x * ROFF_PER;
    uint32_t c[ROFF_PER];
    uint32_t sum = 0, mx = 0;
    const uint32_t nGrid = (uint32_t)grid->nrows;  // only grid rows are ordered (and count for the longest row)
#pragma unroll
    for (int k = 0; k < ROFF_PER; k += 4) {  // rowCount is 16-byte aligned and padded
        const uint4 v = (base + k < nR) ? *reinterpret_cast<const uint4*>(rowCount + base + k) : make_uint4(0, 0, 0, 0);
        c[k] = v.x; c[k + 1] = (base + k + 1 < nR) ? v.y : 0u; c[k + 2] = (base + k + 2 < nR) ? v.z : 0u;
        c[k + 3] = (base + k + 3 < nR) ? v.w : 0u;
        sum += c[k] + c[k + 1] + c[k + 2] + c[k + 3];
#pragma unroll
        for (int q = 0; q < 4; q++)
            if (base + k + q < nGrid) mx = max(mx, c[k + q]);
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    for (int o = 16; o > 0; o >>= 1) mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx) atomicMax(&ctr->maxRowLen, mx);
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < 8 ? warpSum[lane] : 0u, vi = v;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < 8) warpSum[lane] = vi - v;
        const uint32_t total = __shfl_sync(0xffffffffu, vi, 7);
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) rs_store_release(status, total | RS_FLAG_INC);
        } else {
            if (lane == 0) rs_store_release(status + tile, total | RS_FLAG_AGG);
            int t = (int)tile - 1;  // decoupled look-back, 32 predecessors per round
            for (;;) {
                const uint32_t sv = (t - lane >= 0) ? rs_load_relaxed(status + (t - lane)) : RS_FLAG_INC;
                const uint32_t notReady = __ballot_sync(0xffffffffu, (sv & ~RS_VAL_MASK) == 0u);
                const uint32_t incMask = __ballot_sync(0xffffffffu, (sv & ~RS_VAL_MASK) == RS_FLAG_INC);
                const int firstInc = incMask ? __ffs(incMask) - 1 : 32;
                const uint32_t need = firstInc >= 31 ? 0xffffffffu : ((2u << firstInc) - 1u);
                if (notReady & need) continue;  // a predecessor has not published yet
                uint32_t v2 = (lane <= firstInc) ? (sv & RS_VAL_MASK) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                excl += v2;
                if (firstInc < 32) break;
                t -= 32;
            }
            if (lane == 0) rs_store_release(status + tile, (excl + total) | RS_FLAG_INC);
        }
        if (lane == 0) sExcl = excl;
    }
    __syncthreads();
    uint32_t run = sExcl + warpSum[warp] + (incl - sum);
#pragma unroll
    for (int k = 0; k < ROFF_PER; k++) {
        if (base + k < nR) rowStart[base + k] = run;
        run += c[k];
    }
}

__global__ void __launch_bounds__(256)
k_row_place(const RadixState* __restrict__ st, const GridParams* __restrict__ grid, const uint32_t* __restrict__ keys,
            const uint32_t* __restrict__ vals, const uint32_t* __restrict__ slots, const uint32_t* __restrict__ rowStart,
            uint32_t* __restrict__ gkey, uint32_t* __restrict__ gval) {
    const uint32_t n = st->n;
    const int xbits = grid->xbits;
    for (uint32_t t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        const uint32_t k = keys[t];
        const uint32_t pos = rowStart[k >> xbits] + slots[t];
        gkey[pos] = k;
        gval[pos] = vals[t];
    }
}

__global__ void __launch_bounds__(256)
k_row_order(BodyArrays B, const RadixState* __restrict__ st, const GridParams* __restrict__ grid, const uint32_t* __restrict__ gkey,
            const uint32_t* __restrict__ gval, const uint32_t* __restrict__ rowStart, float4* __restrict__ smin,
            float4* __restrict__ smax, uint32_t* __restrict__ skey, uint32_t* __restrict__ scyz, uint32_t* __restrict__ nSortedOut) {
    const uint32_t n = st->n;
    const int xbits = grid->xbits;
    const uint32_t nrows = (uint32_t)grid->nrows, rpw = (uint32_t)grid->rowsPerWorld, nz = (uint32_t)grid->nz;
    if (blockIdx.x == 0 && threadIdx.x == 0) *nSortedOut = n;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint32_t k = gkey[p];
        const uint32_t body = gval[p];
        const uint32_t row = k >> xbits;
        const uint32_t a = rowStart[row], b = rowStart[row + 1];
        uint32_t rank = p - a;  // rows outside the grid (large proxies, dead slots) need no order
        if (row < nrows) {
            // neighbouring threads walk the same row: broadcast loads, eight in flight per thread.  Ties in the key are
            // broken by the proxy index; they are looked up only when a tie shows (rare: 4096 qx values per row)
            uint32_t less = 0, ties = 0;
#pragma unroll 8
            for (uint32_t q = a; q < b; q++) {
                const uint32_t kq = __ldg(gkey + q);
                less += kq < k ? 1u : 0u;
                ties += kq == k ? 1u : 0u;
            }
            rank = less;
            if (ties > 1u)
                for (uint32_t q = a; q < b; q++)
                    rank += (__ldg(gkey + q) == k && __ldg(gval + q) < body) ? 1u : 0u;
        }
        const uint32_t j = a + rank;
        float4 mn = B.effMin[body], mx = B.effMax[body];
        mn.w = __uint_as_float(body);
        mx.w = __uint_as_float(B.filt[body]);
        smin[j] = mn;
        smax[j] = mx;
        skey[j] = k;
        uint32_t c = 0xffffffffu;  // rows outside the grid: large proxies, dead slots
        if (row < nrows) {
            const uint32_t rem = row % rpw;
            c = ((rem / nz) << 16) | (rem % nz);
        }
        scyz[j] = c;
    }
}

// Warp-staged pair append.  A single global counter cannot take one atomic per warp round (same-address
// atomics serialise in L2), so every warp stages its hits in a private shared-memory buffer (ballot + popc for the slot) and
// flushes 32+ pairs at a time with ONE global atomic and coalesced 8-byte stores.
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
    // counter.  Every thread of the block must call it.
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

// ---- 1-D bulk copies (TMA) into shared memory, completion on an mbarrier --------------------------------------------------
__device__ __forceinline__ uint32_t smemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbarExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbarWait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smemAddr(bar)), "r"(parity)
        : "memory");
}
// size: multiple of 16 bytes; src and dst 16-byte aligned
__device__ __forceinline__ void bulkLoad(void* dstSmem, const void* srcGlobal, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

constexpr int SW_CH = 1024;   // sorted entries staged per chunk and block
struct __align__(128) SweepStage {
    float4 mn[SW_CH];
    float4 mx[SW_CH];
    uint32_t key[SW_CH];
};

// Both half-warps search at once (each its own [a, b) and target): 16 probes per round and half.  Returns the lower bound of
// `target` in skey[a, b).
__device__ __forceinline__ uint32_t halfWarpLowerBound(const uint32_t* __restrict__ skey, uint32_t a, uint32_t b, uint32_t target,
                                                       int lane) {
    const int sub = lane & 15;
    const int sh = lane & 16;
    while (__any_sync(0xffffffffu, b - a > 16u)) {
        const bool big = b - a > 16u;
        const uint32_t stepw = big ? (b - a + 15u) / 16u : 1u;
        const uint32_t pos = a + (uint32_t)sub * stepw;
        const bool less = big && pos < b && __ldg(skey + pos) < target;
        const int c = __popc((__ballot_sync(0xffffffffu, less) >> sh) & 0xffffu);  // a prefix of the half (keys are sorted)
        if (big) {
            // the answer lies in (a + (c-1)*stepw, a + c*stepw]
            const uint32_t na = c ? a + (uint32_t)(c - 1) * stepw + 1u : a;
            const uint32_t nb = c ? min(b, a + (uint32_t)c * stepw) : a;
            a = na;
            b = nb;
        }
    }
    const uint32_t pos = a + (uint32_t)sub;
    const bool less = pos < b && __ldg(skey + pos) < target;
    return a + (uint32_t)__popc((__ballot_sync(0xffffffffu, less) >> sh) & 0xffffu);
}

// k_sweep: a block owns 256 consecutive SORTED proxies (thread = proxy) and one BAND of neighbour rows: dy = blockIdx.y - 1
// and dz = -1, 0, +1 — three rows that are consecutive in key order (row = ... + cy * nz + cz).  A proxy's candidates in a
// target row are the sorted entries whose qx lies in its x-window [qx(min.x), qx(max.x)].  Every unordered pair is emitted
// once, by the member that comes first in (qx, sorted position) order; because the list is sorted by row first, that rule
// folds into the window START: rows behind the proxy's own row take qx >= its qx, rows before it qx > its qx, its own row
// the entries behind the proxy itself.  So a window is a plain index range [s, e) of the sorted arrays and needs no
// per-candidate key test.
// The windows of consecutive proxies lie next to each other: the union [lo, hi) over the block and the band is one short
// contiguous range (about the block's own length plus two rows).  Warp 0 finds its two ends with two 16-ary searches that run
// side by side in its half-warps; the range is staged in shared memory with three 1-D bulk copies (TMA: min, max, key)
// completing on one mbarrier; every thread then finds the six ends of its three windows with interleaved branch-free binary
// searches in shared memory and walks the three windows back to back in ONE fixed-trip loop (trip count = the warp's longest
// concatenation, so three short windows per lane even out the length across the warp).  The loop body is the reference's
// closed-interval overlap predicate on the original floats (bp/DbvtAabbMm.java:209-212) + the group/mask filter + the
// ballot/popc compaction of hits; the keys only select candidates.
// The sorted arrays are padded by SW_CH entries: a chunk may read a few entries past n, never past the allocation, and no
// thread looks at staged entries beyond hi.
__global__ void __launch_bounds__(256, 5)
k_sweep(const uint32_t* __restrict__ nPtr, const float4* __restrict__ smin, const float4* __restrict__ smax,
        const uint32_t* __restrict__ skey, const uint32_t* __restrict__ rowStart, const GridParams* __restrict__ grid, int uidBits,
        uint64_t* __restrict__ pairKeys, uint32_t* rowCnt, uint32_t maxPairs, StepCounters* ctr, SlabFilter slab,
        const float4* __restrict__ qmin, const float4* __restrict__ qmax /* SAP modes: quantised bounds per body, else null */,
        const uint32_t* __restrict__ scyz) {
    __shared__ SweepStage S;
    __shared__ uint64_t bar;
    __shared__ uint64_t pstage[8][PAIR_STAGE];
    __shared__ uint32_t sRed[2][8];
    __shared__ uint32_t sLoHi[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        mbarInit(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    uint32_t parity = 0;
    PairStager st;
    st.init(pstage[warp], rowCnt, uidBits);
    const uint32_t n = *nPtr;
    const int ny = grid->ny, nz = grid->nz, xbits = grid->xbits;
    const uint32_t xmask = grid->xmask;
    const float gx0 = grid->x0, ginvX = grid->invX, gxmax = grid->xmaxf;
    const int dy = (int)blockIdx.y - 1;
    for (uint32_t bbase = blockIdx.x * 256u; bbase < n; bbase += gridDim.x * 256u) {
        const uint32_t i = bbase + threadIdx.x;
        uint32_t keyI = 0, cc = 0xffffffffu;
        if (i < n) {
            keyI = __ldg(skey + i);
            cc = __ldg(scyz + i);
        }
        const int cy = (int)(cc >> 16) + dy, cz0 = (int)(cc & 0xffffu);
        const bool act = cc != 0xffffffffu && cy >= 0 && cy < ny;
        float4 amin = make_float4(0, 0, 0, 0), amax = amin;
        if (act) {
            amin = __ldg(smin + i);
            amax = __ldg(smax + i);
        }
        const uint32_t row = keyI >> xbits, xkI = keyI & xmask;
        const uint32_t xkMax = act ? quantX(amax.x, gx0, ginvX, gxmax) : 0u;
        const uint32_t tRow0 = (uint32_t)((int)row + dy * nz);  // the dz = 0 row of the band; same world: cy stays inside [0, ny)
        const bool vm = act && cz0 > 0, vp = act && cz0 + 1 < nz;  // the dz = -1 / +1 rows exist
        // window w (dz = w - 1) = sorted entries with key in [a_w, b_w): rows before the proxy's own take qx > xkI (key + 1 runs
        // into the next row when qx is at its maximum: an empty window, as it must be), rows behind it qx >= xkI
        const uint32_t a0 = (((tRow0 - 1u) << xbits) | xkI) + (dy <= 0 ? 1u : 0u);
        const uint32_t a1 = ((tRow0 << xbits) | xkI) + (dy < 0 ? 1u : 0u);
        const uint32_t a2 = (((tRow0 + 1u) << xbits) | xkI) + (dy < 0 ? 1u : 0u);
        const uint32_t b0 = (((tRow0 - 1u) << xbits) | xkMax) + 1u, b1 = ((tRow0 << xbits) | xkMax) + 1u,
                       b2 = (((tRow0 + 1u) << xbits) | xkMax) + 1u;
        // union of the block's windows: [lower bound of the smallest start key, lower bound of the largest end key)
        uint32_t loKey = act ? (vm ? a0 : a1) : 0xffffffffu;
        uint32_t hiKey = act ? (vp ? b2 : b1) : 0u;
        for (int o = 16; o > 0; o >>= 1) {
            loKey = min(loKey, __shfl_xor_sync(0xffffffffu, loKey, o));
            hiKey = max(hiKey, __shfl_xor_sync(0xffffffffu, hiKey, o));
        }
        __syncthreads();  // the previous iteration's readers of sRed / sLoHi / S are through
        if (lane == 0) { sRed[0][warp] = loKey; sRed[1][warp] = hiKey; }
        __syncthreads();
        if (warp == 0) {
            loKey = sRed[0][lane & 7];
            hiKey = sRed[1][lane & 7];
            for (int o = 4; o > 0; o >>= 1) {
                loKey = min(loKey, __shfl_xor_sync(0xffffffffu, loKey, o));
                hiKey = max(hiKey, __shfl_xor_sync(0xffffffffu, hiKey, o));
            }
            uint32_t lo = 0, hi = 0;
            if (loKey != 0xffffffffu) {  // warp-uniform: some proxy of the block has this band
                const bool upper = lane >= 16;
                // a start key may have run into the next row (qx + 1 past the maximum): search the row the key names; an end
                // key's row is the row of its last admissible entry
                const uint32_t r = upper ? ((hiKey - 1u) >> xbits) : (loKey >> xbits);
                const uint32_t a = __ldg(rowStart + r), b = __ldg(rowStart + r + 1);
                const uint32_t pos = halfWarpLowerBound(skey, a, b, upper ? hiKey : loKey, lane);
                lo = __shfl_sync(0xffffffffu, pos, 0);
                hi = __shfl_sync(0xffffffffu, pos, 16);
            }
            if (lane == 0) { sLoHi[0] = lo; sLoHi[1] = hi; }
        }
        __syncthreads();
        const uint32_t lo = sLoHi[0], hi = sLoHi[1];
        const uint32_t fa = __float_as_uint(amax.w);
        for (uint32_t cs = lo & ~3u; cs < hi; cs += SW_CH) {
            const uint32_t cv = min((uint32_t)SW_CH, hi - cs);          // valid entries of the chunk
            const uint32_t cnt = (cv + 3u) & ~3u;                       // staged entries: 16-byte multiples
            if (cs != (lo & ~3u)) __syncthreads();  // every thread has finished reading the previous chunk
            if (threadIdx.x == 0) {
                mbarExpectTx(&bar, cnt * 36u);
                bulkLoad(S.mn, smin + cs, cnt * 16u, &bar);
                bulkLoad(S.mx, smax + cs, cnt * 16u, &bar);
                bulkLoad(S.key, skey + cs, cnt * 4u, &bar);
            }
            mbarWait(&bar, parity);
            parity ^= 1u;
            // the part of each window that lies in this chunk: six lower bounds, searched side by side
            uint32_t s0 = 0, s1 = 0, s2 = 0, e0 = 0, e1 = 0, e2 = 0;
#pragma unroll
            for (uint32_t sstep = SW_CH; sstep > 0; sstep >>= 1) {  // from SW_CH itself: a result may be cv == SW_CH
                uint32_t t;
                t = s0 + sstep; if (t <= cv && S.key[t - 1] < a0) s0 = t;
                t = s1 + sstep; if (t <= cv && S.key[t - 1] < a1) s1 = t;
                t = s2 + sstep; if (t <= cv && S.key[t - 1] < a2) s2 = t;
                t = e0 + sstep; if (t <= cv && S.key[t - 1] < b0) e0 = t;
                t = e1 + sstep; if (t <= cv && S.key[t - 1] < b1) e1 = t;
                t = e2 + sstep; if (t <= cv && S.key[t - 1] < b2) e2 = t;
            }
            if (dy == 0 && i + 1u > cs) s1 = max(s1, min(cv, i + 1u - cs));  // own row: only the entries behind the proxy itself
            const uint32_t L0 = (vm && e0 > s0) ? e0 - s0 : 0u;
            const uint32_t L1 = (act && e1 > s1) ? e1 - s1 : 0u;
            const uint32_t L2 = (vp && e2 > s2) ? e2 - s2 : 0u;
            const uint32_t L01 = L0 + L1, L = L01 + L2;
            uint32_t maxL = L;
            for (int o = 16; o > 0; o >>= 1) maxL = max(maxL, __shfl_xor_sync(0xffffffffu, maxL, o));
            for (uint32_t t = 0; t < maxL; t++) {
                bool hit = false;
                uint32_t bodyB = 0;
                if (t < L) {
                    const uint32_t k = t < L0 ? s0 + t : (t < L01 ? s1 + (t - L0) : s2 + (t - L01));
                    const float4 bmin = S.mn[k], bmax = S.mx[k];
                    hit = (amin.x <= bmax.x) && (amax.x >= bmin.x) && (amin.y <= bmax.y) && (amax.y >= bmin.y) &&
                          (amin.z <= bmax.z) && (amax.z >= bmin.z) && filterPass(fa, __float_as_uint(bmax.w));
                    bodyB = __float_as_uint(bmin.w);
                    // AxisSweep3 modes: the float boxes are a monotone image of the quantised ones (necessary condition); the
                    // pair predicate itself is on the quantised values
                    if (hit && qmin) {
                        const uint32_t bodyA = __float_as_uint(amin.w);
                        hit = sapOverlap(__ldg(qmin + bodyA), __ldg(qmax + bodyA), __ldg(qmin + bodyB), __ldg(qmax + bodyB));
                    }
                    if (hit) hit = slab.owns(amin, bmin);
                }
                st.push(hit, __float_as_uint(amin.w), bodyB, pairKeys, maxPairs, ctr);
            }
        }
    }
    __shared__ uint32_t sCnt[8];
    __shared__ uint32_t sBase;
    st.flushBlock(pairKeys, maxPairs, ctr, sCnt, &sBase);
}

// k_large: proxies that do not fit the grid (rows nrows + world) against every proxy of the same world, and
// against each other once.
__global__ void __launch_bounds__(256)
k_large(const uint32_t* __restrict__ nPtr, const float4* __restrict__ smin, const float4* __restrict__ smax,
        const uint32_t* __restrict__ rowStart, const GridParams* __restrict__ grid, const int* __restrict__ world, int numWorlds,
        int uidBits, uint64_t* __restrict__ pairKeys, uint32_t* rowCnt, uint32_t maxPairs, StepCounters* ctr, SlabFilter slab,
        const float4* __restrict__ qmin, const float4* __restrict__ qmax) {
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
                if (hit) hit = slab.owns(amin, bmin);  // partitioned world: the pair's region decides the rank
            }
            st.push(hit, bodyA, bodyB, pairKeys, maxPairs, ctr);
        }
    }
    __shared__ uint32_t sCnt[8];
    __shared__ uint32_t sBase;
    st.flushBlock(pairKeys, maxPairs, ctr, sCnt, &sBase);
}

}  // namespace b2c
