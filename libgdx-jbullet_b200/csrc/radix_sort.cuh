// radix_sort.cuh — LSD radix sort for the broadphase (64-bit projected-AABB keys + body index) and
// for the canonical pair list (64-bit packed (uidA,uidB) keys).
//
// Design (sm_100a, HBM/L2-bound integer work, no tensor cores):
//   * the digit histograms and their exclusive scans come from the kernel that WRITES the keys (k_keys, pairfind.cuh):
//     no separate histogram / scan launch and no extra read of the keys;
//   * one kernel per 8-bit digit ("onesweep"): tiles are taken in ticket order, each tile ranks its
//     keys with warp-level match-any multisplit, publishes its per-digit counts and resolves its
//     exclusive prefix by decoupled look-back over the preceding tiles, then scatters through shared
//     memory so global stores are coalesced runs;
//   * a digit whose histogram has a single non-empty bin is skipped on the device (no host sync):
//     every pass kernel derives the current ping-pong side from the skip flags of the passes before it.
// The sort is stable, which the sweep relies on for its tie-break by sorted position.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

namespace b2c {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_MAX_PASSES = 8;
constexpr uint32_t RS_FLAG_AGG = 1u << 30;
constexpr uint32_t RS_FLAG_INC = 2u << 30;
constexpr uint32_t RS_VAL_MASK = (1u << 30) - 1;
constexpr int RS_LOOKBACK = 16;

struct RadixState {
    uint32_t hist[RS_MAX_PASSES][256];  // after rs_scan: exclusive digit offsets
    uint32_t skip[RS_MAX_PASSES];       // 1 = digit is constant over all keys, pass is the identity
    uint32_t ticket[RS_MAX_PASSES];     // dynamic tile ids
    uint32_t n;                         // number of keys (device-resident so it can come from a counter)
    uint32_t pad[7];
};

__device__ __forceinline__ uint32_t rs_load_acquire(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t rs_load_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void rs_store_release(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// Which ping-pong side holds the data before pass p.
__device__ __forceinline__ int rs_side_before(const RadixState* st, int p) {
    int side = 0;
    for (int q = 0; q < p; q++) side ^= (st->skip[q] ? 0 : 1);
    return side;
}

template <typename K, bool HAS_VAL, int ITEMS>
__global__ void __launch_bounds__(RS_THREADS)
rs_pass(K* keys0, K* keys1, uint32_t* vals0, uint32_t* vals1, RadixState* st, uint32_t* status /*[numTiles][256]*/,
        int pass) {
    constexpr int TILE = RS_THREADS * ITEMS;
    if (st->skip[pass]) return;
    const uint32_t n = st->n;
    const uint32_t numTiles = (n + TILE - 1) / TILE;
    const int side = rs_side_before(st, pass);
    const K* __restrict__ src = side ? keys1 : keys0;
    K* __restrict__ dst = side ? keys0 : keys1;
    const uint32_t* __restrict__ vsrc = side ? vals1 : vals0;
    uint32_t* __restrict__ vdst = side ? vals0 : vals1;

    __shared__ uint32_t warpCnt[RS_WARPS][256];
    __shared__ uint32_t digitBase[256];   // global destination of the first key of digit d in this tile
    __shared__ uint32_t digitLocal[256];  // exclusive scan of tile counts (position inside the sorted tile)
    __shared__ uint32_t sTile;
    extern __shared__ __align__(16) unsigned char dynsmem[];
    K* skeys = reinterpret_cast<K*>(dynsmem);
    uint32_t* svals = reinterpret_cast<uint32_t*>(dynsmem + sizeof(K) * TILE);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t ltmask = (1u << lane) - 1u;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) sTile = atomicAdd(&st->ticket[pass], 1u);
        for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_THREADS) (&warpCnt[0][0])[i] = 0;
        __syncthreads();
        const uint32_t tile = sTile;
        if (tile >= numTiles) return;
        const uint32_t tileStart = tile * TILE;

        K key[ITEMS];
        uint32_t val[ITEMS];
        uint32_t rank[ITEMS];
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            uint32_t idx = tileStart + warp * (ITEMS * 32) + k * 32 + lane;
            bool valid = idx < n;
            key[k] = valid ? src[idx] : (K)0;
            if (HAS_VAL) val[k] = valid ? vsrc[idx] : 0u;
        }
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            uint32_t idx = tileStart + warp * (ITEMS * 32) + k * 32 + lane;
            bool valid = idx < n;
            uint32_t d = (uint32_t)(key[k] >> (8 * pass)) & 255u;
            uint32_t m = __match_any_sync(0xffffffffu, valid ? d : (0x100u + lane));
            uint32_t lower = __popc(m & ltmask);
            uint32_t pre = 0;
            if (valid && lower == 0) {
                pre = warpCnt[warp][d];
                warpCnt[warp][d] = pre + __popc(m);
            }
            __syncwarp();
            pre = __shfl_sync(0xffffffffu, pre, __ffs(m) - 1);
            rank[k] = pre + lower;
        }
        __syncthreads();
        // per digit: exclusive scan over warps, tile count
        const int d = threadIdx.x;
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; w++) {
            uint32_t c = warpCnt[w][d];
            warpCnt[w][d] = run;
            run += c;
        }
        const uint32_t tileCount = run;
        // publish and look back
        uint32_t* myStatus = status + (size_t)tile * 256 + d;
        uint32_t excl = 0;
        if (tile == 0) {
            rs_store_release(myStatus, tileCount | RS_FLAG_INC);
        } else {
            rs_store_release(myStatus, tileCount | RS_FLAG_AGG);
            // Look back over the preceding tiles until one with an inclusive prefix is found.  The walk is the
            // critical path of a pass (early on only tile 0 is inclusive, so tile t sums t aggregates), so the
            // statuses are fetched RS_LOOKBACK at a time as independent loads instead of one dependent load per tile.
            int t = (int)tile - 1;
            bool done = false;
            while (!done && t >= 0) {
                uint32_t sv[RS_LOOKBACK];
#pragma unroll
                for (int k = 0; k < RS_LOOKBACK; k++)
                    sv[k] = (t - k >= 0) ? rs_load_relaxed(status + (size_t)(t - k) * 256 + d) : RS_FLAG_INC;
                int used = 0;
#pragma unroll
                for (int k = 0; k < RS_LOOKBACK; k++) {
                    if (!done && used == k) {
                        uint32_t f = sv[k] & ~RS_VAL_MASK;
                        if (f != 0) {  // published (an unpublished one stops the batch; it is re-read next round)
                            excl += sv[k] & RS_VAL_MASK;
                            used = k + 1;
                            if (f == RS_FLAG_INC) done = true;
                        }
                    }
                }
                t -= used;
            }
            rs_store_release(myStatus, (excl + tileCount) | RS_FLAG_INC);
        }
        // exclusive scan of tileCount over digits -> digitLocal
        digitLocal[d] = tileCount;
        __syncthreads();
        for (int off = 1; off < 256; off <<= 1) {
            uint32_t tv = d >= off ? digitLocal[d - off] : 0;
            __syncthreads();
            digitLocal[d] += tv;
            __syncthreads();
        }
        const uint32_t localExcl = digitLocal[d] - tileCount;
        __syncthreads();
        digitLocal[d] = localExcl;
        digitBase[d] = st->hist[pass][d] + excl - localExcl;
        __syncthreads();
        // place keys in tile-sorted order in shared memory
#pragma unroll
        for (int k = 0; k < ITEMS; k++) {
            uint32_t idx = tileStart + warp * (ITEMS * 32) + k * 32 + lane;
            if (idx < n) {
                uint32_t dk = (uint32_t)(key[k] >> (8 * pass)) & 255u;
                uint32_t lp = digitLocal[dk] + warpCnt[warp][dk] + rank[k];
                skeys[lp] = key[k];
                if (HAS_VAL) svals[lp] = val[k];
            }
        }
        __syncthreads();
        const uint32_t tileN = min((uint32_t)TILE, n - tileStart);
        for (uint32_t i = threadIdx.x; i < tileN; i += RS_THREADS) {
            K kk = skeys[i];
            uint32_t dk = (uint32_t)(kk >> (8 * pass)) & 255u;
            uint32_t pos = digitBase[dk] + i;
            dst[pos] = kk;
            if (HAS_VAL) vdst[pos] = svals[i];
        }
    }
}

struct RadixSorter {
    RadixState* st = nullptr;
    uint32_t* status = nullptr;
    size_t statusWordsPerPass = 0;   // sized for the smallest tile (most tiles)
    uint32_t capacity = 0;
    int launches = 0;

    static int pickItems(uint32_t n) {
        const char* e = getenv("B2C_RS_ITEMS");  // tuning knob for experiments
        if (e) { int v = atoi(e); if (v == 4 || v == 8 || v == 16) return v; }
        // measured (profiles/): 1 M keys 0.146 / 0.125 / 0.113 ms with 4 / 8 / 16 keys per thread; 262 k keys best at 8;
        // 100 k keys best at 4 (more tiles than SMs matters more than the length of the look-back chain)
        return n >= 600000u ? 16 : (n >= 200000u ? 8 : 4);
    }

    cudaError_t init(uint32_t cap) {
        capacity = cap;
        const uint32_t tile = RS_THREADS * 4;
        statusWordsPerPass = (size_t)((cap + tile - 1) / tile + 1) * 256;
        // dynamic shared memory of the wide tiles (> 48 KB needs the opt-in), set once so that passes() only launches
        cudaFuncSetAttribute(rs_pass<uint32_t, true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * RS_THREADS * 16);
        cudaFuncSetAttribute(rs_pass<uint32_t, true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * RS_THREADS * 8);
        cudaError_t e = cudaMalloc(&st, sizeof(RadixState));
        if (e != cudaSuccess) return e;
        return cudaMalloc(&status, statusWordsPerPass * RS_MAX_PASSES * sizeof(uint32_t));
    }
    void destroy() {
        cudaFree(st);
        cudaFree(status);
        st = nullptr;
        status = nullptr;
    }

    // Clear the histograms / tickets / skip flags and the look-back status of the tiles `nUpper` keys can occupy (memset
    // nodes of the step graph).  The kernel that writes the keys then fills st->hist, scans it and sets st->n / st->skip.
    // nUpper shapes the launches (tile size); nMax bounds the count the device may actually hold (>= nUpper).
    cudaError_t reset(uint32_t nUpper, uint32_t nMax, int npass, cudaStream_t s) {
        cudaError_t e = cudaMemsetAsync(st, 0, sizeof(RadixState), s);
        if (e != cudaSuccess) return e;
        const uint32_t tile = RS_THREADS * pickItems(nUpper);
        const size_t words = (size_t)((nMax + tile - 1) / tile + 1) * 256;
        for (int p = 0; p < npass && e == cudaSuccess; p++)
            e = cudaMemsetAsync(status + statusWordsPerPass * p, 0, words * sizeof(uint32_t), s);
        return e;
    }

    // The onesweep passes over at most nUpper keys (the exact count is st->n on the device).  The sorted data ends on the
    // side rs_side_before(st, npass) names; consumers read that on the device.
    template <typename K, bool HAS_VAL>
    void passes(K* keys0, K* keys1, uint32_t* vals0, uint32_t* vals1, uint32_t nUpper, int npass, cudaStream_t s) {
        const int items = pickItems(nUpper);
        const uint32_t tile = RS_THREADS * items;
        unsigned pgrid = (nUpper + tile - 1) / tile;
        if (pgrid < 1) pgrid = 1;
        if (pgrid > 148 * 4) pgrid = 148 * 4;
        const size_t dyn = (sizeof(K) + (HAS_VAL ? 4 : 0)) * tile;
        for (int p = 0; p < npass; p++) {
            uint32_t* stp = status + statusWordsPerPass * p;
            if (items == 16) {
                rs_pass<K, HAS_VAL, 16><<<pgrid, RS_THREADS, dyn, s>>>(keys0, keys1, vals0, vals1, st, stp, p);
            } else if (items == 8) {
                rs_pass<K, HAS_VAL, 8><<<pgrid, RS_THREADS, dyn, s>>>(keys0, keys1, vals0, vals1, st, stp, p);
            } else {
                rs_pass<K, HAS_VAL, 4><<<pgrid, RS_THREADS, dyn, s>>>(keys0, keys1, vals0, vals1, st, stp, p);
            }
        }
        launches += npass;
    }
};

}  // namespace b2c
