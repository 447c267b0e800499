// pair_rows.cuh — canonical order of the overlapping-pair list without a general sort.
//
// bp/HashedOverlappingPairCache.java:291-296 orders every pair as (uid0 < uid1); the ABI returns the list sorted
// lexicographically, and manifold persistence (k_carry) matches pairs by that order.  The sweep emits pairs in
// arbitrary order.  Instead of radix-sorting 64-bit keys (5 onesweep passes at 100 k bodies) the list is built
// as rows keyed by uid0:
//   emission   every pair takes its slot inside its row with one atomicAdd on rowCnt[uid0] (PairStager::flush)
//   k_row_scan     exclusive scan of rowCnt -> rowStart (single pass, decoupled look-back); rows longer than
//                  ROW_SMALL go on the big-row list; rowStart doubles as next step's "first pair of uid0" table
//   k_row_scatter  uid1 of every emitted pair -> csr[rowStart[uid0] + slot]
//   k_row_rank     one thread per pair of a short row (<= ROW_SMALL): its place is the count of smaller uid1 in the row
//   k_row_sort_big one block per long row (large statics, meshes): the row is a SET of uids, so it is sorted by
//                  setting bits in a shared-memory bitmap over [min uid1, max uid1] and enumerating them in order
// All of it is integer/byte work bound by L2/HBM traffic: ~8 P read + 4 P write + 4 P read + 16 P write bytes.
#pragma once
#include "common.cuh"

namespace b2c {

constexpr int ROW_SMALL = 48;
constexpr int RSCAN_PER = 16, RSCAN_TILE = 256 * RSCAN_PER;  // rows per thread, rows per tile
constexpr uint32_t RS2_AGG = 1u << 30, RS2_INC = 2u << 30, RS2_VAL = (1u << 30) - 1u;
constexpr int BIG_WORDS = 8192;            // shared-memory bitmap of k_row_sort_big: 262 144 uids per chunk

struct RowMisc {
    uint32_t ticket, bigCount, pad[2];
};

__device__ __forceinline__ uint32_t row_ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void row_st_relaxed(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// rowStart[u] = number of pairs whose uid0 < u, for u in 0..nRows (nRows = max uid + 2).
__global__ void __launch_bounds__(256)
k_row_scan(const uint32_t* __restrict__ rowCnt, uint32_t nRows, uint32_t* __restrict__ rowStart, uint32_t* status, RowMisc* misc,
           uint32_t* __restrict__ bigRows, uint32_t* __restrict__ numPairsOut) {
    __shared__ uint32_t sTile, sExcl, warpSum[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) sTile = atomicAdd(&misc->ticket, 1u);
    __syncthreads();
    const uint32_t tile = sTile;
    const uint32_t base = tile * RSCAN_TILE + threadIdx.x * RSCAN_PER;
    uint32_t c[RSCAN_PER];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < RSCAN_PER; k += 4) {  // rowCnt is 16-byte aligned and padded to a multiple of 4
        uint4 v = (base + k < nRows) ? *reinterpret_cast<const uint4*>(rowCnt + base + k) : make_uint4(0, 0, 0, 0);
        c[k] = v.x; c[k + 1] = (base + k + 1 < nRows) ? v.y : 0u; c[k + 2] = (base + k + 2 < nRows) ? v.z : 0u;
        c[k + 3] = (base + k + 3 < nRows) ? v.w : 0u;
        sum += c[k] + c[k + 1] + c[k + 2] + c[k + 3];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        uint32_t v = lane < 8 ? warpSum[lane] : 0u, vi = v;
#pragma unroll
        for (int o = 1; o < 8; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, vi, o);
            if (lane >= o) vi += t;
        }
        if (lane < 8) warpSum[lane] = vi - v;
        const uint32_t total = __shfl_sync(0xffffffffu, vi, 7);
        uint32_t excl = 0;
        if (tile == 0) {
            if (lane == 0) row_st_relaxed(status, total | RS2_INC);
        } else {
            if (lane == 0) row_st_relaxed(status + tile, total | RS2_AGG);
            // decoupled look-back, 32 predecessors per round
            int t = (int)tile - 1;
            for (;;) {
                uint32_t sv = (t - lane >= 0) ? row_ld_relaxed(status + (t - lane)) : RS2_INC;
                uint32_t notReady = __ballot_sync(0xffffffffu, (sv & ~RS2_VAL) == 0u);
                uint32_t incMask = __ballot_sync(0xffffffffu, (sv & ~RS2_VAL) == RS2_INC);
                int firstInc = incMask ? __ffs(incMask) - 1 : 32;
                uint32_t need = firstInc >= 31 ? 0xffffffffu : ((2u << firstInc) - 1u);
                if (notReady & need) continue;  // a predecessor has not published yet
                uint32_t v2 = (lane <= firstInc) ? (sv & RS2_VAL) : 0u;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v2 += __shfl_xor_sync(0xffffffffu, v2, o);
                excl += v2;
                if (firstInc < 32) break;
                t -= 32;
            }
            if (lane == 0) row_st_relaxed(status + tile, (excl + total) | RS2_INC);
        }
        if (lane == 0) {
            sExcl = excl;
            if ((tile + 1) * RSCAN_TILE >= nRows) {  // last tile: grand total
                rowStart[nRows] = excl + total;
                *numPairsOut = excl + total;
            }
        }
    }
    __syncthreads();
    uint32_t run = sExcl + warpSum[warp] + (incl - sum);
#pragma unroll
    for (int k = 0; k < RSCAN_PER; k++) {
        if (base + k < nRows) {
            rowStart[base + k] = run;
            if (c[k] > (uint32_t)ROW_SMALL) bigRows[atomicAdd(&misc->bigCount, 1u)] = base + k;
        }
        run += c[k];
    }
}

// emitted key = slot << 2*uidBits | uid0 << uidBits | uid1
__global__ void __launch_bounds__(256)
k_row_scatter(const uint64_t* __restrict__ pairKeys, const StepCounters* __restrict__ ctr, uint32_t maxPairs, int uidBits,
              const uint32_t* __restrict__ rowStart, uint32_t* __restrict__ csr) {
    const uint32_t n = ctr->pairCount < maxPairs ? ctr->pairCount : maxPairs;
    const uint64_t mask = (1ull << uidBits) - 1ull;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        uint64_t k = pairKeys[p];
        uint32_t hi = (uint32_t)(k & mask), lo = (uint32_t)((k >> uidBits) & mask), slot = (uint32_t)(k >> (2 * uidBits));
        csr[rowStart[lo] + slot] = hi;
    }
}

// One thread per EMITTED pair: its place inside its (short) row is the number of row members smaller than its uid1 —
// the row is a set, so ranks are distinct.  O(row length) independent L1-resident loads per pair, no local arrays.
__global__ void __launch_bounds__(256)
k_row_rank(const uint64_t* __restrict__ pairKeys, const StepCounters* __restrict__ ctr, uint32_t maxPairs, int uidBits,
           const uint32_t* __restrict__ rowStart, const uint32_t* __restrict__ csr, int2* __restrict__ pairs,
           uint64_t* __restrict__ sortedKeys) {
    const uint32_t n = ctr->pairCount < maxPairs ? ctr->pairCount : maxPairs;
    const uint64_t mask = (1ull << uidBits) - 1ull;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        const uint64_t k = pairKeys[p];
        const uint32_t hi = (uint32_t)(k & mask), lo = (uint32_t)((k >> uidBits) & mask);
        const uint32_t s = rowStart[lo], cnt = rowStart[lo + 1] - s;
        if (cnt > (uint32_t)ROW_SMALL) continue;  // long rows: k_row_sort_big
        uint32_t pos = 0;
        for (uint32_t e = 0; e < cnt; e++) pos += (__ldg(csr + s + e) < hi) ? 1u : 0u;
        pairs[s + pos] = make_int2((int)lo, (int)hi);
        sortedKeys[s + pos] = ((uint64_t)lo << uidBits) | hi;
    }
}

constexpr int BIG_THREADS = 1024;
__global__ void __launch_bounds__(BIG_THREADS)
k_row_sort_big(const uint32_t* __restrict__ rowStart, const uint32_t* __restrict__ bigRows, const RowMisc* __restrict__ misc,
               const uint32_t* __restrict__ csr, int uidBits, int2* __restrict__ pairs, uint64_t* __restrict__ sortedKeys) {
    __shared__ uint32_t bm[BIG_WORDS];
    __shared__ uint32_t red[2][32];
    __shared__ uint32_t wsum[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nBig = misc->bigCount;
    for (uint32_t b = blockIdx.x; b < nBig; b += gridDim.x) {
        const uint32_t r = bigRows[b], s = rowStart[r], cnt = rowStart[r + 1] - s;
        uint32_t mn = 0xffffffffu, mx = 0u;
        for (uint32_t e = threadIdx.x; e < cnt; e += blockDim.x) {
            uint32_t h = csr[s + e];
            mn = min(mn, h);
            mx = max(mx, h);
        }
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        __syncthreads();  // previous row's use of red/bm is over
        if (lane == 0) { red[0][warp] = mn; red[1][warp] = mx; }
        __syncthreads();
        for (int w = 0; w < BIG_THREADS / 32; w++) { mn = min(mn, red[0][w]); mx = max(mx, red[1][w]); }
        uint32_t outBase = s;
        for (uint32_t c0 = mn; c0 <= mx; c0 += (uint32_t)BIG_WORDS * 32u) {
            const uint32_t span = min(mx - c0 + 1u, (uint32_t)BIG_WORDS * 32u);
            const uint32_t words = (span + 31u) >> 5;
            const uint32_t wpt = (words + blockDim.x - 1) / blockDim.x;  // words per thread, contiguous
            for (uint32_t w = threadIdx.x; w < words; w += blockDim.x) bm[w] = 0u;
            __syncthreads();
            for (uint32_t e = threadIdx.x; e < cnt; e += blockDim.x) {
                uint32_t h = csr[s + e] - c0;  // wraps for h < c0 -> out of range
                if (h < span) atomicOr(&bm[h >> 5], 1u << (h & 31u));
            }
            __syncthreads();
            const uint32_t w0 = min(threadIdx.x * wpt, words), w1 = min(w0 + wpt, words);
            uint32_t mine = 0;
            for (uint32_t w = w0; w < w1; w++) mine += __popc(bm[w]);
            uint32_t incl = mine;
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += t;
            }
            if (lane == 31) wsum[warp] = incl;
            __syncthreads();
            uint32_t before = 0, total = 0;
            for (int w = 0; w < BIG_THREADS / 32; w++) { if (w < warp) before += wsum[w]; total += wsum[w]; }
            uint32_t pos = outBase + before + (incl - mine);
            for (uint32_t w = w0; w < w1; w++) {
                uint32_t bits = bm[w];
                while (bits) {
                    uint32_t h = c0 + (w << 5) + (uint32_t)(__ffs(bits) - 1);
                    bits &= bits - 1u;
                    pairs[pos] = make_int2((int)r, (int)h);
                    sortedKeys[pos] = ((uint64_t)r << uidBits) | h;
                    pos++;
                }
            }
            outBase += total;
            __syncthreads();
            if (c0 + (uint32_t)BIG_WORDS * 32u < c0) break;  // uint32 wrap guard
        }
    }
}

}  // namespace b2c
