// halo.cuh — one world partitioned over several GPUs by slabs of space (BASELINE config C5; device code only).
//
// Space is cut by nranks-1 planes along one axis into regions; rank r owns region r.  Two kinds of ownership:
//   * a PROXY is owned by the rank whose region held its origin when the partition was set (b2c_set_partition_slabs).  Only
//     the owner runs updateAabbs / the DbvtBroadphase.setAabb state machine for it, so that state never has to move.
//   * a PAIR is owned by the region that holds max(min_a, min_b) along the axis (SlabFilter::owns) — a coordinate inside both
//     boxes, so both proxies touch that region.
// Every step each rank therefore needs the AABBs of all proxies that TOUCH its region: its own, plus the HALO — proxies
// owned elsewhere whose box reaches into the region.  The owners publish exactly those (box not entirely inside the home
// region) as 80-byte records {min | proxy, max | flags, 3 transform rows} in a fixed-size slot; ONE all-gather over NVLink
// (ncclAllGather, enqueued by the host binding on the ctx stream) hands every rank all slots; each rank keeps the records
// that touch its region (k_halo_import) and sorts / sweeps its local list = owned-and-touching + imported.  The transform
// travels with the box because the narrowphase of a cross-boundary pair runs on the pair's owner.
#pragma once
#include "pairfind.cuh"

namespace b2c {

struct HaloRecord {
    float4 mn;      // effective AABB min | proxy index
    float4 mx;      // effective AABB max | flags
    float4 xf[3];   // world transform rows
};  // 80 bytes
constexpr int HALO_HEADER_BYTES = 16;  // { uint32 count, pad[3] }

__host__ __device__ inline size_t haloSlotBytes(uint32_t cap) { return (size_t)HALO_HEADER_BYTES + (size_t)cap * sizeof(HaloRecord); }

__device__ __forceinline__ float axisOf(float4 v, int axis) { return axis == 0 ? v.x : (axis == 1 ? v.y : v.z); }

// owner[i] = region of the proxy's origin (all ranks run this on identical transforms -> identical tables)
__global__ void __launch_bounds__(256)
k_assign_owner(BodyArrays B, int n, SlabFilter slab, uint8_t* __restrict__ owner) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 r = B.xf4[3 * (size_t)i + slab.axis];  // row `axis` of the transform: .w is the origin coordinate
    owner[i] = (uint8_t)slab.region(r.w);
}

// Owned proxies whose box is not entirely inside the home region go into this rank's slot.
__global__ void __launch_bounds__(256)
k_halo_export(BodyArrays B, int n, const uint8_t* __restrict__ owner, SlabFilter slab, unsigned char* __restrict__ slot, uint32_t cap,
              StepCounters* ctr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool send = false;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    uint8_t flags = 0;
    if (i < n && owner[i] == (uint8_t)slab.rank) {
        flags = B.flags[i];
        if (flags & BF_ALIVE) {
            a = B.effMin[i];
            b = B.effMax[i];
            const int lo = slab.region(axisOf(a, slab.axis)), hi = slab.region(axisOf(b, slab.axis));
            send = !(lo == slab.rank && hi == slab.rank);
        }
    }
    const uint32_t m = __ballot_sync(0xffffffffu, send);
    if (m == 0) return;
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(reinterpret_cast<uint32_t*>(slot), (uint32_t)__popc(m));
    base = __shfl_sync(0xffffffffu, base, 0);
    if (send) {
        const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
        if (pos < cap) {
            HaloRecord* r = reinterpret_cast<HaloRecord*>(slot + HALO_HEADER_BYTES) + pos;
            a.w = __uint_as_float((uint32_t)i);
            b.w = __uint_as_float((uint32_t)flags);
            r->mn = a;
            r->mx = b;
            r->xf[0] = B.xf4[3 * (size_t)i];
            r->xf[1] = B.xf4[3 * (size_t)i + 1];
            r->xf[2] = B.xf4[3 * (size_t)i + 2];
        } else {
            ctr->haloOverflow = 1;
        }
    }
}

// ---- the same exchange without a collective: stores into the peers' memory over NVLink -------------------------------
// Each rank owns an INBOX = 2 parities x nranks slots of the layout above; slot [parity][src] is written only by rank `src`.
// k_halo_export_p2p sends every boundary record straight into slot [epoch & 1][me] of exactly the ranks whose slab the box
// reaches (peer pointers from cudaIpcOpenMemHandle, or plain pointers when the "ranks" share a process), so a rank
// receives only what touches it and nothing is padded to a slot size.  When the last block is done it publishes
// {count, epoch} in each destination's slot header with one 64-bit release store at system scope; k_halo_wait on the
// receiving side spins (acquire, system scope) until every source's header carries this step's epoch.  Two parities are
// enough: a rank can run at most one step ahead of a neighbour, because its next export comes after its own import,
// which waits for that neighbour's export of the current step.
struct HaloPeers {
    unsigned char* inbox[16];   // base of rank r's inbox in THIS process's address space
};
__device__ __forceinline__ void st_release_sys_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

struct HaloP2pState {
    uint32_t sendCount[16];   // records sent to each destination this step
    uint32_t ticket;          // blocks finished (fused kernel)
    uint32_t pad[3];
    uint32_t pushTicket[16];  // blocks finished per destination (two-phase variant)
    uint32_t migrateTicket[16];
};

__global__ void __launch_bounds__(256)
k_halo_export_p2p(BodyArrays B, int n, const uint8_t* __restrict__ owner, SlabFilter slab, HaloPeers peers, int nranks, size_t slotBytes,
                  uint32_t cap, uint32_t epoch, HaloP2pState* st, StepCounters* ctr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int lo = 1, hi = 0;   // destination regions [lo, hi] (empty: nothing to send)
    float4 a = make_float4(0, 0, 0, 0), b = a;
    uint8_t flags = 0;
    if (i < n && owner[i] == (uint8_t)slab.rank) {
        flags = B.flags[i];
        if (flags & BF_ALIVE) {
            a = B.effMin[i];
            b = B.effMax[i];
            lo = slab.region(axisOf(a, slab.axis));
            hi = slab.region(axisOf(b, slab.axis));
        }
    }
    const size_t mySlot = ((size_t)(epoch & 1u) * (size_t)nranks + (size_t)slab.rank) * slotBytes;
    for (int d = 0; d < nranks; d++) {   // warp-uniform loop: one atomic per destination per warp
        const bool want = d != slab.rank && lo <= d && d <= hi;
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (m == 0) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&st->sendCount[d], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (want) {
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) {
                HaloRecord* r = reinterpret_cast<HaloRecord*>(peers.inbox[d] + mySlot + HALO_HEADER_BYTES) + pos;
                float4 a2 = a, b2 = b;
                a2.w = __uint_as_float((uint32_t)i);
                b2.w = __uint_as_float((uint32_t)flags);
                r->mn = a2;
                r->mx = b2;
                r->xf[0] = B.xf4[3 * (size_t)i];
                r->xf[1] = B.xf4[3 * (size_t)i + 1];
                r->xf[2] = B.xf4[3 * (size_t)i + 2];
            } else {
                ctr->haloOverflow = 1;
            }
        }
    }
    // publish: a block that stored into a peer fences at system scope before it takes a ticket; the last block writes the headers
    if (__syncthreads_or(lo <= hi && (lo != slab.rank || hi != slab.rank))) __threadfence_system();
    __syncthreads();
    __shared__ uint32_t sLast;
    if (threadIdx.x == 0) sLast = (atomicAdd(&st->ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
    __syncthreads();
    if (sLast && threadIdx.x < (unsigned)nranks && (int)threadIdx.x != slab.rank) {
        __threadfence_system();
        const int d = threadIdx.x;
        uint32_t c = *(volatile uint32_t*)&st->sendCount[d];
        if (c > cap) c = cap;
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(peers.inbox[d] + mySlot), (unsigned long long)c | ((unsigned long long)epoch << 32));
    }
}

// Two-phase variant of the same exchange: k_halo_stage compacts the boundary records per destination in LOCAL memory
// (stage[d] = cap records), k_halo_push copies each destination's run into the peer's inbox with coalesced 16-byte stores
// (grid.y = destination) and the last block of a destination publishes its header — scattered 16-byte peer stores from
// one thread per record become full NVLink write packets, and only the push blocks pay a system-scope fence.
__global__ void __launch_bounds__(256)
k_halo_stage(BodyArrays B, int n, const uint8_t* __restrict__ owner, SlabFilter slab, int nranks, HaloRecord* __restrict__ stage,
             uint32_t cap, HaloP2pState* st, StepCounters* ctr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int lo = 1, hi = 0;
    float4 a = make_float4(0, 0, 0, 0), b = a;
    uint8_t flags = 0;
    if (i < n && owner[i] == (uint8_t)slab.rank) {
        flags = B.flags[i];
        if (flags & BF_ALIVE) {
            a = B.effMin[i];
            b = B.effMax[i];
            lo = slab.region(axisOf(a, slab.axis));
            hi = slab.region(axisOf(b, slab.axis));
        }
    }
    for (int d = 0; d < nranks; d++) {
        const bool want = d != slab.rank && lo <= d && d <= hi;
        const uint32_t m = __ballot_sync(0xffffffffu, want);
        if (m == 0) continue;
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&st->sendCount[d], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (want) {
            const uint32_t pos = base + __popc(m & ((1u << lane) - 1u));
            if (pos < cap) {
                HaloRecord* r = stage + (size_t)d * cap + pos;
                float4 a2 = a, b2 = b;
                a2.w = __uint_as_float((uint32_t)i);
                b2.w = __uint_as_float((uint32_t)flags);
                r->mn = a2;
                r->mx = b2;
                r->xf[0] = B.xf4[3 * (size_t)i];
                r->xf[1] = B.xf4[3 * (size_t)i + 1];
                r->xf[2] = B.xf4[3 * (size_t)i + 2];
            } else {
                ctr->haloOverflow = 1;
            }
        }
    }
}
__global__ void __launch_bounds__(256)
k_halo_push(const HaloRecord* __restrict__ stage, HaloPeers peers, int nranks, int rank, size_t slotBytes, uint32_t cap, uint32_t epoch,
            HaloP2pState* st, uint32_t* __restrict__ pushTicket /*[16], zeroed*/) {
    const int d = blockIdx.y;
    if (d == rank) return;
    uint32_t c = st->sendCount[d];
    if (c > cap) c = cap;
    const size_t mySlot = ((size_t)(epoch & 1u) * (size_t)nranks + (size_t)rank) * slotBytes;
    const float4* src = reinterpret_cast<const float4*>(stage + (size_t)d * cap);
    float4* dst = reinterpret_cast<float4*>(peers.inbox[d] + mySlot + HALO_HEADER_BYTES);
    const size_t words = (size_t)c * (sizeof(HaloRecord) / 16);
    for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words; w += (size_t)gridDim.x * blockDim.x) dst[w] = src[w];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&pushTicket[d], 1u) == gridDim.x - 1) {
        __threadfence_system();
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(peers.inbox[d] + mySlot), (unsigned long long)c | ((unsigned long long)epoch << 32));
    }
}

// The manifolds of pairs that changed owner travel the same way: the sender's departed-manifold slot (k_export_departed:
// { count, keys[cap], headers[cap], points[4 cap] }, narrowphase.cuh) is pushed — only its `count` live records — into slot
// [parity][me] of EVERY other rank's migration inbox (the sender cannot tell who owns the pair now), then published.
__global__ void __launch_bounds__(256)
k_migrate_push(const unsigned char* __restrict__ local, HaloPeers peers, int nranks, int rank, size_t slotBytes, uint32_t cap, uint32_t epoch,
               uint32_t* __restrict__ pushTicket /*[16], zeroed*/, StepCounters* ctr) {
    const int d = blockIdx.y;
    if (d == rank) return;
    uint32_t c = *reinterpret_cast<const uint32_t*>(local);
    if (c > cap) { c = cap; if (blockIdx.x == 0 && threadIdx.x == 0) ctr->migrateOverflow = 1; }
    unsigned char* dst = peers.inbox[d] + ((size_t)(epoch & 1u) * (size_t)nranks + (size_t)rank) * slotBytes;
    // three runs of 8-byte words: keys (8 B each), headers (32 B), points (4 x 96 B per manifold)
    const size_t off[3] = {16, 16 + (size_t)cap * 8, 16 + (size_t)cap * (8 + 32)};
    const size_t words[3] = {(size_t)c, (size_t)c * 4, (size_t)c * 48};
    for (int k = 0; k < 3; k++) {
        const unsigned long long* s8 = reinterpret_cast<const unsigned long long*>(local + off[k]);
        unsigned long long* d8 = reinterpret_cast<unsigned long long*>(dst + off[k]);
        for (size_t w = (size_t)blockIdx.x * blockDim.x + threadIdx.x; w < words[k]; w += (size_t)gridDim.x * blockDim.x) d8[w] = s8[w];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0 && atomicAdd(&pushTicket[d], 1u) == gridDim.x - 1) {
        __threadfence_system();
        st_release_sys_u64(reinterpret_cast<unsigned long long*>(dst), (unsigned long long)c | ((unsigned long long)epoch << 32));
    }
}

// One warp: lane s waits until source s has published this step's epoch in its slot of this rank's inbox.
__global__ void __launch_bounds__(32)
k_halo_wait(const unsigned char* __restrict__ inboxParity, int nranks, int rank, size_t slotBytes, uint32_t epoch, StepCounters* ctr) {
    const int s = threadIdx.x;
    if (s >= nranks || s == rank) return;
    const unsigned long long* hdr = reinterpret_cast<const unsigned long long*>(inboxParity + (size_t)s * slotBytes);
    const long long t0 = clock64();
    while ((uint32_t)(ld_acquire_sys_u64(hdr) >> 32) != epoch) {
        if (clock64() - t0 > 20000000000ll) { ctr->haloOverflow = 2; break; }   // ~10 s: a peer died; reported by the host as an error
        __nanosleep(200);
    }
}

// Every record of the other ranks' slots that touches this rank's region becomes a local (halo) proxy: its box, transform
// and flags are written into the rank's proxy arrays and its index is appended to the local list.
__global__ void __launch_bounds__(256)
k_halo_import(BodyArrays B, const unsigned char* __restrict__ slots, uint32_t cap, SlabFilter slab, uint32_t* __restrict__ list,
              uint32_t* __restrict__ nLocal, StepCounters* ctr, size_t slotStride) {
    const int src = blockIdx.y;
    if (src == slab.rank) return;  // own slot: those proxies are listed by k_list_owned
    const unsigned char* slot = slots + (size_t)src * slotStride;
    uint32_t cnt = *reinterpret_cast<const uint32_t*>(slot);
    if (cnt > cap) { if (threadIdx.x == 0 && blockIdx.x == 0) ctr->haloOverflow = 1; cnt = cap; }
    const HaloRecord* recs = reinterpret_cast<const HaloRecord*>(slot + HALO_HEADER_BYTES);
    const int lane = threadIdx.x & 31;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < cnt; base += gridDim.x * blockDim.x) {
        const uint32_t k = base + lane;
        bool take = false;
        HaloRecord r;
        if (k < cnt) {
            r = recs[k];
            take = slab.touches(axisOf(r.mn, slab.axis), axisOf(r.mx, slab.axis), slab.rank);
        }
        const uint32_t m = __ballot_sync(0xffffffffu, take);
        if (m == 0) continue;
        uint32_t pos = 0;
        if (lane == 0) pos = atomicAdd(nLocal, (uint32_t)__popc(m));
        pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
        if (take) {
            const uint32_t i = __float_as_uint(r.mn.w);
            B.effMin[i] = make_float4(r.mn.x, r.mn.y, r.mn.z, 0.f);
            B.effMax[i] = make_float4(r.mx.x, r.mx.y, r.mx.z, 0.f);
            B.flags[i] = (uint8_t)__float_as_uint(r.mx.w);
            B.xf4[3 * (size_t)i] = r.xf[0];
            B.xf4[3 * (size_t)i + 1] = r.xf[1];
            B.xf4[3 * (size_t)i + 2] = r.xf[2];
            list[pos] = i;
        }
    }
}

// Owned proxies that touch the home region (all but the ones that have drifted out of it entirely).
__global__ void __launch_bounds__(256)
k_list_owned(BodyArrays B, int n, const uint8_t* __restrict__ owner, SlabFilter slab, uint32_t* __restrict__ list,
             uint32_t* __restrict__ nLocal) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    bool take = false;
    if (i < n && owner[i] == (uint8_t)slab.rank && (B.flags[i] & BF_ALIVE))
        take = slab.touches(axisOf(B.effMin[i], slab.axis), axisOf(B.effMax[i], slab.axis), slab.rank);
    const uint32_t m = __ballot_sync(0xffffffffu, take);
    if (m == 0) return;
    uint32_t pos = 0;
    if (lane == 0) pos = atomicAdd(nLocal, (uint32_t)__popc(m));
    pos = __shfl_sync(0xffffffffu, pos, 0) + __popc(m & ((1u << lane) - 1u));
    if (take) list[pos] = (uint32_t)i;
}

}  // namespace b2c
