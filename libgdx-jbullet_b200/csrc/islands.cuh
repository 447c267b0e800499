// islands.cuh — the two consumers that sit directly behind the pair list in the reference:
//
//   * pair add / remove deltas: what bp/HashedOverlappingPairCache.java:323-325 (add) and :135-137 (remove) report to
//     the ghost pair callback (disp/GhostPairCallback.java:40-68).  The device list is rebuilt every step, so the
//     deltas are the set differences of this step's and last step's sorted key lists (row-confined binary search).
//   * simulation islands: disp/SimulationIslandManager.java:57-110 (findUnions over every BROADPHASE pair whose two
//     objects merge islands, then tag = find(i); -1 for static objects) with disp/UnionFind.java.  On the device:
//     lock-free union-find that always hooks the larger root under the smaller one (CAS on the root, path halving
//     on the way; parents only ever decrease, so racing/stale reads are still ancestors), then a flatten pass.
//     The tag of an island is therefore its smallest body index — the reference's tag depends on its (irreproducible,
//     SURVEY Q2) pair insertion order, and its consumers only compare tags for equality.
#pragma once
#include "common.cuh"

namespace b2c {

__device__ __forceinline__ int islandRep(int* par, int v) {
    int cur = par[v];
    if (cur != v) {
        int next, prev = v;
        while (cur > (next = par[cur])) {
            par[prev] = next;  // path halving
            prev = cur;
            cur = next;
        }
    }
    return cur;
}

__global__ void __launch_bounds__(256) k_island_init(int* __restrict__ par, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) par[i] = i;
}

__global__ void __launch_bounds__(256)
k_island_unite(const int2* __restrict__ pairs, const uint32_t* __restrict__ numPairs, const uint8_t* __restrict__ flags, int* par) {
    const uint32_t n = *numPairs;
    for (uint32_t p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
        int2 pr = pairs[p];
        int a = pr.x - 1, b = pr.y - 1;
        // CollisionObject.mergesSimulationIslands (disp/CollisionObject.java:100-103): not static / kinematic
        uint8_t fa = flags[a], fb = flags[b];
        if ((fa & BF_STATIC) || (fb & BF_STATIC) || !(fa & BF_ALIVE) || !(fb & BF_ALIVE)) continue;
        int ra = islandRep(par, a), rb = islandRep(par, b);
        bool repeat;
        do {
            repeat = false;
            if (ra != rb) {
                int ret;
                if (ra < rb) {
                    if ((ret = atomicCAS(&par[rb], rb, ra)) != rb) { rb = ret; repeat = true; }
                } else {
                    if ((ret = atomicCAS(&par[ra], ra, rb)) != ra) { ra = ret; repeat = true; }
                }
            }
        } while (repeat);
    }
}

__global__ void __launch_bounds__(256)
k_island_flatten(int* par, const uint8_t* __restrict__ flags, int n, int* __restrict__ tags, uint32_t* __restrict__ numIslands) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    bool root = false;
    if (i < n) {
        uint8_t f = flags[i];
        if ((f & BF_ALIVE) && !(f & BF_STATIC)) {
            int r = par[i];
            while (r != par[r]) r = par[r];
            tags[i] = r;
            root = (r == i);
        } else {
            tags[i] = -1;
        }
    }
    uint32_t m = __ballot_sync(0xffffffffu, root);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(numIslands, (uint32_t)__popc(m));
}

// which: 0 = pairs of `keys` that are not in `otherKeys` are appended to out (unordered; one atomic per warp round)
__global__ void __launch_bounds__(256)
k_pair_delta(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ num, const uint64_t* __restrict__ otherKeys,
             const uint32_t* __restrict__ otherNum, const uint32_t* __restrict__ otherFirst, int uidBits, int2* __restrict__ out,
             uint32_t cap, uint32_t* __restrict__ outCount) {
    const uint32_t n = *num, on = *otherNum;
    const int lane = threadIdx.x & 31;
    const uint64_t mask = (1ull << uidBits) - 1ull;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < n; base += gridDim.x * blockDim.x) {
        const uint32_t p = base + lane;
        bool missing = false;
        uint64_t k = 0;
        if (p < n) {
            k = keys[p];
            missing = true;
            if (on) {
                uint32_t uid0 = (uint32_t)(k >> uidBits);
                uint32_t a = otherFirst[uid0], b = otherFirst[uid0 + 1];
                while (a < b) {
                    uint32_t mid = (a + b) >> 1;
                    if (otherKeys[mid] < k) a = mid + 1; else b = mid;
                }
                missing = !(a < on && otherKeys[a] == k);
            }
        }
        uint32_t m = __ballot_sync(0xffffffffu, missing);
        if (m == 0) continue;
        uint32_t slot = 0;
        if (lane == 0) slot = atomicAdd(outCount, (uint32_t)__popc(m));
        slot = __shfl_sync(0xffffffffu, slot, 0) + __popc(m & ((1u << lane) - 1u));
        if (missing && slot < cap) out[slot] = make_int2((int)(k >> uidBits), (int)(k & mask));
    }
}

}  // namespace b2c
