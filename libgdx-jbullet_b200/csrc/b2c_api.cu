// b2c_api.cu — the C ABI (include/b2c.h) over the CUDA collision path.  Host orchestration only: every
// per-step computation is a kernel in broadphase.cuh / narrowphase.cuh / radix_sort.cuh; there is no CPU
// fallback (b2c_create fails without an sm_100 device).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/b2c.h"
#include "broadphase.cuh"
#include "bvh_build.h"
#include "common.cuh"
#include "compound.cuh"
#include "islands.cuh"
#include "narrowphase.cuh"
#include "pair_rows.cuh"
#include "halo.cuh"
#include "pairfind.cuh"
#include "radix_sort.cuh"
#include "raycast.cuh"
#include "compound_flatten.h"
#include "convexcast.cuh"

using namespace b2c;

namespace {

constexpr int EPA_GRID = 148 * 2, EPA_BLOCK = 32;      // tier 0: one warp per block, per-lane pools in shared memory (2 blocks/SM)
constexpr int EPA_GRID3 = 148 * 16, EPA_BLOCK3 = 64;   // tier 2: pools in local memory, throughput variant
constexpr int COMPOUND_MESH_GRID = 148, COMPOUND_MESH_BLOCK = 64;  // k_compound_mesh: one full-size EPA pool per thread in global memory
constexpr int EPA_GRID2 = 148 * 3, EPA_BLOCK2 = 64;    // tier 1: one item per warp, large pools in shared memory (2 x 34 KB per block)

struct HostMesh {
    int4* nodes = nullptr;
    float* verts = nullptr;
    int* idx = nullptr;
    int* partStart = nullptr;
    HostBvh bvh;
};

}  // namespace

struct b2c_ctx {
    b2c_config cfg;
    int device = 0;
    cudaStream_t stream = nullptr;
    // side streams of the narrowphase (forked from / joined to `stream` with events inside one step):
    cudaStream_t streamCopy = nullptr;    // D2H of the pair list while the narrowphase is still running on `stream`
    cudaEvent_t evPairsReady = nullptr;   // recorded on `stream` when the broadphase of the current step is enqueued
    cudaStream_t streamClosed = nullptr;  // sphere-sphere / convex-plane bins, beside the GJK kernels
    cudaStream_t streamEpa = nullptr;     // penetration bin (few long-latency lanes), beside k_manifold_cc; high priority
    cudaEvent_t evFork[4] = {nullptr, nullptr, nullptr, nullptr}, evJoin[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    bool overlap = true;
    int mccBlocks = 8;                    // k_manifold_cc blocks per SM (B2C_MCC_BLOCKS)
    int epaLpw = 8;                       // active lanes per warp in the shared-memory EPA tier (B2C_EPA_LPW: 32/16/8/4)
    int epaHint = -1;                     // -1 unknown, 0 small penetration bin (shared-memory tier), 1 large (local-memory tier)
    // CUDA graphs of the whole step (b2c_step_device), one per launch signature
    struct StepGraph {
        uint64_t sig[4];
        cudaGraphExec_t exec;
        int launches;
    };
    std::vector<StepGraph> graphs;
    bool useGraphs = true;                // B2C_GRAPH=0 disables
    bool capturing = false;
    bool timeline = false;                // B2C_TIMELINE=1: print where the side-stream kernels ran (debug)
    cudaEvent_t tl[6] = {};
    std::string err;

    // shapes
    std::vector<ShapeDev> hShapes;
    ShapeDev* dShapes = nullptr;
    float4* dHullPts = nullptr;
    int hullPtsUsed = 0;
    std::vector<HostMesh> meshes;
    std::vector<MeshDev> hMeshes;
    MeshDev* dMeshes = nullptr;
    bool shapesDirty = false;
    bool hasPlane = false, hasMesh = false;

    SapParams sap{};   // AxisSweep3 modes (broadphase_mode 2 / 3)

    // bodies
    int nBodies = 0;  // slots in use (uids 1..nBodies)
    BodyArrays B{};
    std::vector<uint8_t> hFlags;
    std::vector<int> hShapeOf;
    // destroyed slots: `freePending` until a pair calculation has dropped their pairs (so the manifold carry can never hand a
    // dead proxy's manifolds to its successor), then `freeSlots`, which b2c_proxy_create reuses once the table is full
    std::vector<int> freePending, freeSlots;
    float* dStaging = nullptr;   // 12 planes x max_bodies
    float* hStagingPinned = nullptr;
    int stagingCount = 0;        // >0: planes uploaded for bodies 1..stagingCount, to be repacked by k_aabb
    float* dExtAabb = nullptr;   // 6 planes x max_bodies (b2c_set_aabbs)
    uint8_t* dExtMask = nullptr;
    bool extPending = false;
    bool aabbPending = false;
    int step = 0;                // host mirror of *dStep (the kernels read the device copy)
    int* dStep = nullptr;

    // broadphase
    uint32_t* dKeys[2] = {nullptr, nullptr};
    uint32_t* dVals[2] = {nullptr, nullptr};
    uint32_t* dSide = nullptr;        // [2]: body sort side, pair sort side
    float4* dSmin = nullptr;
    float4* dSmax = nullptr;
    uint32_t* dSrow = nullptr;
    uint32_t* dScyz = nullptr;        // (cy << 16 | cz) of every sorted proxy's grid row
    uint32_t* dRowStart = nullptr;
    int maxRows = 0;
    GridParams* dGrid = nullptr;
    StepCounters* dCtr = nullptr;
    StepCounters* hCtrPinned = nullptr;
    uint64_t* dPairKeys = nullptr;    // emitted (slot | uid0 | uid1) keys, emission order
    uint32_t* dCsr = nullptr;         // uid1 of every pair, grouped by uid0 (pair_rows.cuh)
    uint32_t* dRowZero = nullptr;     // one block cleared per step: rowCnt[nRows] | scan status[tiles] | RowMisc
    uint32_t* dBigRows = nullptr;
    uint32_t nRows = 0, rowTiles = 0;
    RadixSorter sortBodies;
    int uidBits = 1;

    // pairs + manifolds (ping-pong across steps)
    int2* dPairs = nullptr;
    uint64_t* dSortedKeys[2] = {nullptr, nullptr};
    uint32_t* dNumPairs[2] = {nullptr, nullptr};
    ManifoldHdr* dMHdr[2] = {nullptr, nullptr};
    b2c_manifold_point* dMPts[2] = {nullptr, nullptr};
    uint32_t* dPairFirst[2] = {nullptr, nullptr};  // first pair index per uid0 (for the next step's carry)
    int cur = 0;  // index of this step's pair keys / manifolds
    bool pairsValid = false;

    // narrowphase
    b2c_raw_contact* dRaw = nullptr;
    int8_t* dRawFlag = nullptr;
    uint8_t* dHist = nullptr;          // per-pair GJK iteration count of the last step (k_carry)
    uint8_t* dBinOf = nullptr;
    uint32_t* dBinItems = nullptr;
    uint32_t* dBinStart = nullptr;   // [17]
    uint32_t* dBinZero = nullptr;    // hist[16] | cursor[16] of the bin partition
    uint32_t binTiles = 0;
    uint32_t* dCursors = nullptr;
    uint32_t* dSurvivors = nullptr;
    uint32_t* dSurvSorted = nullptr;   // survivors ordered by last step's iteration count
    uint8_t* dSurvKey = nullptr;
    uint32_t* dSurvZero = nullptr;     // hist[16] | cursor[16] of the survivor partition
    uint32_t* dSurvStart = nullptr;    // [17]
    EpaItem* dEpaItems = nullptr;
    uint32_t maxEpa = 0;
    uint32_t* dEpaRetry = nullptr;
    uint32_t* dEpaBig = nullptr;      // items routed straight to the large-pool tier (capacity maxEpaRetry)
    uint32_t maxEpaRetry = 0;
    uint32_t* dMeshPair = nullptr;
    int* dMeshTri = nullptr;
    b2c_raw_contact* dRawMesh = nullptr;
    uint32_t* dMeshStart = nullptr;
    uint32_t* dMeshCount = nullptr;

    // stats
    b2c_stats stats{};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    int launches = 0;
    int32_t lastPairs = 0, lastManifolds = 0, lastContacts = 0;
    // one world partitioned over several GPUs by slabs (halo.cuh)
    SlabFilter slab{};                 // enabled = 0: the whole world lives here
    int partRanks = 1;
    uint8_t* dOwner = nullptr;         // [N] owning rank of every proxy
    uint32_t* dLocalList = nullptr;    // [N] proxies this rank sorts and sweeps this step (owned and touching + imported halo)
    uint32_t* dNLocal = nullptr;       // device count of dLocalList
    uint32_t* dNSorted = nullptr;      // device count of the sorted arrays (== nBodies / *dNLocal), written by k_gather
    uint32_t localHint = 0;            // host's last reading of *dNLocal (launch shapes only; 0 = unknown)
    // row-grouped ordering (pairfind.cuh): zeroed block = RowOffsetsMisc | scan status[roTiles + 4] | rowCount[maxRows + 8]
    uint32_t* dRowOrdZero = nullptr;
    // everything a pair calculation needs zeroed lives in ONE allocation cleared by one memset node at the start of the step:
    // StepCounters | row-order block (dRowOrdZero) | pair-row block (dRowZero); the same for the dispatch: bin partition |
    // survivor partition | cursors
    unsigned char* dZeroBp = nullptr;
    size_t zeroBpBytes = 0;
    unsigned char* dZeroNp = nullptr;
    size_t zeroNpBytes = 0;
    uint32_t* dSlots = nullptr;        // [N] slot of every proxy inside its row
    uint32_t roTiles = 0;
    uint32_t rowLenHint = 0;           // longest row of the last step the host has read
    bool forceRadix = false;
    bool haloExported = false, haloImported = false;
    // peer-to-peer halo exchange (halo.cuh): this rank's inbox, the peers' inboxes as mapped here, per-step send state
    unsigned char* dHaloInbox = nullptr;
    size_t haloSlotBytesP2p = 0;
    uint32_t haloCapP2p = 0, haloEpoch = 0;
    HaloPeers haloPeers = {};
    void* haloIpcOpened[16] = {};
    bool haloConnected = false;
    HaloP2pState* dHaloP2p = nullptr;
    HaloRecord* dHaloStage = nullptr;   // two-phase variant: per-destination runs of records before the push
    unsigned char* dMigInbox = nullptr; // migration inbox (same allocation as the halo inbox, behind it)
    unsigned char* dMigLocal = nullptr; // this rank's departed-manifold slot before the push
    HaloPeers migPeers = {};
    size_t migSlotBytes = 0, haloInboxBytes = 0;
    uint32_t migCap = 0;
    bool migExported = false;
    bool haloTwoPhase = true;
    uint32_t* dExportCount = nullptr;
    bool prof = false;
    cudaEvent_t stageEv[B2C_NUM_STAGES + 1] = {};
    cudaEvent_t evGjk[2] = {};     // around k_gjk alone (profiling)
    bool wantRaw = false;          // b2c_set_raw_records: the inspection records of pairs nobody on the device reads
    bool stageValid = false;

    // compound shapes (compound.cuh) — buffers are allocated when the first one is registered
    bool hasCompound = false;
    std::vector<CompoundChildDev> hChildren;
    std::vector<std::vector<CompoundDirectChild>> compoundDirect;   // per shape id: the addChildShape calls of a compound (else empty)
    CompoundChildDev* dChildren = nullptr;
    size_t capChildren = 0;
    uint32_t maxCompoundItems = 0;
    CompoundCounters* dCompoundCtr = nullptr;
    uint32_t* dCItemPair = nullptr;
    uint32_t* dCItemCode = nullptr;
    int* dCItemPrev = nullptr;
    b2c_raw_contact* dCRaw = nullptr;
    uint32_t* dCMeshStart = nullptr;                   // child x mesh items: their per-triangle records in dRawMesh
    uint32_t* dCMeshCount = nullptr;
    EpaScratch* dCBigScratch = nullptr;                // full-size EPA pools for k_compound_mesh (allocated when a world has both
    uint32_t numCBigScratch = 0;                       //   compounds and meshes)
    ManifoldHdr* dCH[2] = {nullptr, nullptr};          // child manifolds, ping-pong per dispatch
    b2c_manifold_point* dCP[2] = {nullptr, nullptr};
    int ccur = 0;                                      // index of the LATEST child-manifold arrays

    uint64_t* dNoCollide = nullptr;   // sorted keys of never-dispatched body pairs (b2c_set_no_collide_pairs)
    uint32_t numNoCollide = 0, capNoCollide = 0;

    // ray tests (allocated on first use)
    float4* dRayMin = nullptr;
    float4* dRayMax = nullptr;
    float* dRayIn = nullptr;       // from | to, 6 floats per ray
    RayOut* dRayOut = nullptr;
    float4* dRayChunkMin = nullptr;   // boxes of RAY_CHUNK consecutive bodies in the last broadphase's sorted order
    float4* dRayChunkMax = nullptr;
    int nSortedBodies = 0;            // bodies covered by dSmin's sorted order (0: no broadphase has run)
    uint32_t* dRayOverflow = nullptr;
    int rayCap = 0;
    float* dSweepIn = nullptr;     // basis 9 | from 3 | to 3 | cast shape id, 16 words per sweep
    RayOut* dSweepOut = nullptr;
    int sweepCap = 0;

    // islands + pair deltas (allocated on first use)
    int* dIslandPar = nullptr;
    int* dIslandTags = nullptr;
    int2* dDelta[2] = {nullptr, nullptr};
    uint32_t* dDeltaCounts = nullptr;  // [3]: added, removed, islands
    bool deltaPrefetch = false;        // b2c_set_pair_delta_prefetch: the deltas are computed inside every pair calculation
    bool deltaReady = false;           // dDelta / dDeltaCounts belong to the current pair list
    uint32_t* hDeltaCountsPinned = nullptr;

    // compact contact stream
    b2c_contact_header* dContactHdr = nullptr;
    b2c_manifold_point* dContactPts = nullptr;
    uint32_t* dContactCounts = nullptr;
    uint32_t capContactHdr = 0, capContactPts = 0;
    int contactPrefetch = -1;          // b2c_set_contact_prefetch: -1 off, else the stream format compacted at the end of every dispatch
    int contactReady = -1;             // format of the stream sitting in dContactHdr / dContactPts for the last dispatch, or -1
    uint32_t contactCounts[2] = {0, 0};
    bool contactCountsValid = false;   // contactCounts came back with the step counters of the dispatch that compacted
    // two-phase packed stream (prefetch format 2): the manifolds that are final before the penetration bin ends are compacted
    // early and can be downloaded while it runs (b2c_begin_contact_download)
    cudaEvent_t evContactsEarly = nullptr;
    bool earlyRecorded = false;        // the last dispatch recorded evContactsEarly
    struct { bool active; void* hdr; void* pts; uint32_t nH, nP; } earlyDl = {false, nullptr, nullptr, 0, 0};
    uint32_t* hEarlyCountsPinned = nullptr;
};

namespace {

#define CK(call)                                                                                 \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess) {                                                                 \
            ctx->err = std::string(#call) + ": " + cudaGetErrorString(e_);                       \
            return B2C_ERR_CUDA;                                                                 \
        }                                                                                        \
    } while (0)

template <class T>
cudaError_t dalloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, n * sizeof(T));
    if (e == cudaSuccess) e = cudaMemset(*p, 0, n * sizeof(T));
    return e;
}

int bitsFor(uint32_t v) {
    int b = 1;
    while ((1ull << b) <= v) b++;
    return b;
}

unsigned gridFor(uint32_t n, int block, unsigned cap = 148 * 8) {
    unsigned g = (n + block - 1) / block;
    if (g < 1) g = 1;
    return g > cap ? cap : g;
}

int32_t uploadShapes(b2c_ctx* ctx) {
    if (!ctx->shapesDirty) return B2C_OK;
    if (!ctx->hShapes.empty())
        CK(cudaMemcpyAsync(ctx->dShapes, ctx->hShapes.data(), ctx->hShapes.size() * sizeof(ShapeDev), cudaMemcpyHostToDevice,
                           ctx->stream));
    if (!ctx->hMeshes.empty())
        CK(cudaMemcpyAsync(ctx->dMeshes, ctx->hMeshes.data(), ctx->hMeshes.size() * sizeof(MeshDev), cudaMemcpyHostToDevice,
                           ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->shapesDirty = false;
    return B2C_OK;
}

// host copy of the shape AABB for createProxy (disp/CollisionWorld.java:113-119): computed on the device
// by a one-thread launch of the same code path, so the bits are identical to the per-step kernel.
__global__ void k_initial_aabb(BodyArrays B, const ShapeDev* shapes, int i, SapParams sp) {
    Xf t;
    float4 r0 = B.xf4[3 * (size_t)i], r1 = B.xf4[3 * (size_t)i + 1], r2 = B.xf4[3 * (size_t)i + 2];
    t.m[0][0] = r0.x; t.m[0][1] = r0.y; t.m[0][2] = r0.z;
    t.m[1][0] = r1.x; t.m[1][1] = r1.y; t.m[1][2] = r1.z;
    t.m[2][0] = r2.x; t.m[2][1] = r2.y; t.m[2][2] = r2.z;
    t.o = mk3(r0.w, r1.w, r2.w);
    f3 mn, mx;
    shapeAabb(shapes[B.shape[i]], t, mn, mx);
    if (sp.enabled) { sapSetAabb(B, i, mn, mx, sp); return; }  // createProxy -> addHandle: quantised creation AABB
    B.effMin[i] = make_float4(mn.x, mn.y, mn.z, 0.f);
    B.effMax[i] = make_float4(mx.x, mx.y, mx.z, 0.f);
    B.leafMin[i] = B.effMin[i];
    B.leafMax[i] = B.effMax[i];
}

// batched variant used when many proxies are created before the first step
__global__ void k_initial_aabb_range(BodyArrays B, const ShapeDev* shapes, int first, int count, SapParams sp) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    int i = first + k;
    Xf t;
    float4 r0 = B.xf4[3 * (size_t)i], r1 = B.xf4[3 * (size_t)i + 1], r2 = B.xf4[3 * (size_t)i + 2];
    t.m[0][0] = r0.x; t.m[0][1] = r0.y; t.m[0][2] = r0.z;
    t.m[1][0] = r1.x; t.m[1][1] = r1.y; t.m[1][2] = r1.z;
    t.m[2][0] = r2.x; t.m[2][1] = r2.y; t.m[2][2] = r2.z;
    t.o = mk3(r0.w, r1.w, r2.w);
    f3 mn, mx;
    shapeAabb(shapes[B.shape[i]], t, mn, mx);
    if (sp.enabled) { sapSetAabb(B, i, mn, mx, sp); return; }  // createProxy -> addHandle: quantised creation AABB
    B.effMin[i] = make_float4(mn.x, mn.y, mn.z, 0.f);
    B.effMax[i] = make_float4(mx.x, mx.y, mx.z, 0.f);
    B.leafMin[i] = B.effMin[i];
    B.leafMax[i] = B.effMax[i];
}

__global__ void k_scatter_xf(BodyArrays B, int n, const int* __restrict__ uids, const float* __restrict__ planes) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int i = uids[k] - 1;
    const float* p = planes + k;
    size_t s = (size_t)n;
    B.xf4[3 * (size_t)i] = make_float4(p[0], p[s], p[2 * s], p[9 * s]);
    B.xf4[3 * (size_t)i + 1] = make_float4(p[3 * s], p[4 * s], p[5 * s], p[10 * s]);
    B.xf4[3 * (size_t)i + 2] = make_float4(p[6 * s], p[7 * s], p[8 * s], p[11 * s]);
}

__global__ void k_get_aabbs(BodyArrays B, int n, float* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 a = B.effMin[i], b = B.effMax[i];
    out[6 * (size_t)i + 0] = a.x; out[6 * (size_t)i + 1] = a.y; out[6 * (size_t)i + 2] = a.z;
    out[6 * (size_t)i + 3] = b.x; out[6 * (size_t)i + 4] = b.y; out[6 * (size_t)i + 5] = b.z;
}

__global__ void k_clear_np_counters(StepCounters* c) {
    if (threadIdx.x == 0) {
        c->contactsAdded = c->gjkChecks = c->deepChecks = c->epaFailed = 0;
        c->meshItems = c->meshOverflow = c->epaCount = c->epaRetry = c->epaBig = 0;
    }
    if (threadIdx.x < 16) c->binCount[threadIdx.x] = 0;
}

// Flush pending AABB work (update from transforms and/or host-supplied AABBs) with one k_aabb launch.
// forPairs: this launch opens a pair-finding step, so the counters are cleared first and the extent
// reduction it performs feeds the grid.
int32_t runAabbKernel(b2c_ctx* ctx, bool forPairs) {
    int n = ctx->nBodies;
    if (forPairs) CK(cudaMemsetAsync(ctx->dZeroBp, 0, ctx->zeroBpBytes, ctx->stream));  // counters, row counts, scan states of the whole pair calculation
    if (n == 0) return B2C_OK;
    int32_t rc = uploadShapes(ctx);
    if (rc) return rc;
    const float* staging = ctx->stagingCount > 0 ? ctx->dStaging : nullptr;
    // forPairs: the grid of this step comes out of the same launch (single world: fused bounds reduction; partitioned world: the
    // reduction runs over the local list once the halo is in, see enqueueBroadphase)
    const bool slab = ctx->slab.enabled != 0;
    k_aabb<<<(n + 255) / 256, 256, 0, ctx->stream>>>(
        ctx->B, ctx->dShapes, n, staging, ctx->cfg.max_bodies, ctx->stagingCount, ctx->extPending ? ctx->dExtAabb : nullptr,
        ctx->dExtMask, ctx->cfg.max_bodies, ctx->cfg.broadphase_mode, ctx->dStep, ctx->cfg.contact_breaking_threshold,
        ctx->cfg.dbvt_margin, ctx->cfg.dbvt_predicted_frames, ctx->aabbPending ? 1 : 0, ctx->dCtr, ctx->sap,
        forPairs ? 1 : 0, (forPairs && !slab) ? 1 : 0, ctx->dGrid, ctx->cfg.num_worlds, ctx->maxRows - ctx->cfg.num_worlds,
        slab ? ctx->dOwner : nullptr, ctx->slab.rank);
    ctx->launches++;
    ctx->stagingCount = 0;
    if (ctx->extPending) {
        CK(cudaMemsetAsync(ctx->dExtMask, 0, (size_t)ctx->cfg.max_bodies, ctx->stream));
        ctx->extPending = false;
    }
    ctx->aabbPending = false;
    CK(cudaGetLastError());
    return B2C_OK;
}

static inline void mark(b2c_ctx* ctx, int k);

// launch-shape bound of the proxies one pair calculation sorts (the exact count of a partitioned world lives on the device)
static uint32_t slabUpper(const b2c_ctx* ctx) {
    const uint32_t n = (uint32_t)ctx->nBodies;
    if (!ctx->slab.enabled || ctx->localHint == 0) return n;
    const uint64_t h = (uint64_t)ctx->localHint * 5 / 4 + 4096;
    return h < n ? (uint32_t)h : n;
}

// Which ordering pipeline the next pair calculation uses: rows grouped and ordered by counting (default), or the radix passes
// when the longest row seen so far would make the quadratic row ordering matter (B2C_SORT=radix forces them).
static bool useRowOrder(const b2c_ctx* ctx) { return !ctx->forceRadix && ctx->rowLenHint <= ROW_ORDER_MAX_LEN; }

int32_t enqueueBroadphase(b2c_ctx* ctx) {
    int n = ctx->nBodies;
    const bool slab = ctx->slab.enabled != 0;
    if (!slab) {
        mark(ctx, 0);
        int32_t rc = runAabbKernel(ctx, true);  // AABB update + (fused) the grid of this step
        if (rc) return rc;
    } else if (!ctx->haloImported) {
        ctx->err = "partitioned world: b2c_mgpu_update_export_halo and b2c_mgpu_import_halo must precede the pair calculation";
        return B2C_ERR_STATE;
    }
    ctx->haloImported = false;
    ctx->cur ^= 1;
    const int cur = ctx->cur;
    cudaStream_t s = ctx->stream;
    if (n == 0) {
        CK(cudaMemsetAsync(ctx->dNumPairs[cur], 0, sizeof(uint32_t), s));
        CK(cudaEventRecord(ctx->evPairsReady, s));
        ctx->step++;
        CK(cudaMemcpyAsync(ctx->dStep, &ctx->step, sizeof(int), cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
        ctx->pairsValid = true;
        return B2C_OK;
    }
    // proxies this pair calculation sorts: all slots, or (partitioned world) the local list whose length lives on the device
    const uint32_t nUpper = slabUpper(ctx);
    const uint32_t* nPtr = slab ? ctx->dNLocal : nullptr;
    const uint32_t* list = slab ? ctx->dLocalList : nullptr;
    const unsigned nb = gridFor(nUpper, 256, 148 * 16);
    mark(ctx, 1);
    if (slab) {  // the grid over the local list (the halo arrived after k_aabb)
        k_bounds<<<gridFor(nUpper, 256, 148 * 4), 256, 0, s>>>(ctx->B, list, nPtr, ctx->cfg.num_worlds, ctx->maxRows - ctx->cfg.num_worlds,
                                                              ctx->dCtr, ctx->dGrid);
        ctx->launches++;
    }
    int npass = (12 + bitsFor((uint32_t)ctx->maxRows + 2u) + 7) / 8;
    if (npass > 4) npass = 4;
    ctx->sortBodies.launches = 0;
    if (useRowOrder(ctx)) {
        // rows are short: group by row with one atomic per proxy, order inside the rows by counting (pairfind.cuh)
        RowOffsetsMisc* misc = reinterpret_cast<RowOffsetsMisc*>(ctx->dRowOrdZero);
        uint32_t* roStatus = ctx->dRowOrdZero + 4;
        uint32_t* rowCount = ctx->dRowOrdZero + 4 + ctx->roTiles + 4;
        k_keys<<<nb, 256, 0, s>>>(ctx->B, n, nPtr, list, ctx->dCtr, ctx->dGrid, ctx->dKeys[0], ctx->dVals[0], ctx->dStep, ctx->sortBodies.st,
                                  npass, rowCount, ctx->dSlots);
        mark(ctx, 2);
        k_row_offsets<<<ctx->roTiles, 256, 0, s>>>(rowCount, ctx->dGrid, ctx->dRowStart, roStatus, misc, ctx->dCtr);
        k_row_place<<<nb, 256, 0, s>>>(ctx->sortBodies.st, ctx->dGrid, ctx->dKeys[0], ctx->dVals[0], ctx->dSlots, ctx->dRowStart,
                                       ctx->dKeys[1], ctx->dVals[1]);
        mark(ctx, 3);
        k_row_order<<<nb, 256, 0, s>>>(ctx->B, ctx->sortBodies.st, ctx->dGrid, ctx->dKeys[1], ctx->dVals[1], ctx->dRowStart, ctx->dSmin,
                                       ctx->dSmax, ctx->dSrow, ctx->dScyz, ctx->dNSorted);
        ctx->launches += 2;  // + k_keys and k_row_order, counted below like k_keys / k_gather of the radix path
    } else {
        CK(ctx->sortBodies.reset(nUpper, (uint32_t)n, npass, s));
        k_keys<<<nb, 256, 0, s>>>(ctx->B, n, nPtr, list, ctx->dCtr, ctx->dGrid, ctx->dKeys[0], ctx->dVals[0], ctx->dStep, ctx->sortBodies.st,
                                  npass, nullptr, nullptr);
        mark(ctx, 2);
        ctx->sortBodies.passes<uint32_t, true>(ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], nUpper, npass, s);
        mark(ctx, 3);
        k_gather<<<nb, 256, 0, s>>>(ctx->B, ctx->dKeys[0], ctx->dKeys[1], ctx->dVals[0], ctx->dVals[1], ctx->sortBodies.st, npass, ctx->dGrid,
                                    ctx->dSmin, ctx->dSmax, ctx->dSrow, ctx->dRowStart, ctx->dScyz, ctx->dNSorted);
    }
    uint32_t* rowCnt = ctx->dRowZero;
    uint32_t* rowStatus = ctx->dRowZero + ctx->nRows;
    RowMisc* rowMisc = reinterpret_cast<RowMisc*>(ctx->dRowZero + ctx->nRows + ctx->rowTiles);
    mark(ctx, 4);
    // k_large (a few proxies against everything, latency-bound) runs beside k_sweep: both only append pairs
    cudaStream_t sl = s;
    if (ctx->overlap) {
        sl = ctx->streamClosed;
        CK(cudaEventRecord(ctx->evFork[2], s));
        CK(cudaStreamWaitEvent(sl, ctx->evFork[2], 0));
    }
    k_sweep<<<dim3(gridFor(nUpper, 256, 148 * 8), 3), 256, 0, s>>>(ctx->dNSorted, ctx->dSmin, ctx->dSmax, ctx->dSrow, ctx->dRowStart, ctx->dGrid,
                                                          ctx->uidBits, ctx->dPairKeys, rowCnt, (uint32_t)ctx->cfg.max_pairs, ctx->dCtr, ctx->slab,
                                                          ctx->sap.enabled ? ctx->B.leafMin : nullptr,
                                                          ctx->sap.enabled ? ctx->B.leafMax : nullptr, ctx->dScyz);
    // one block column per large proxy (static planes, meshes, big statics; one floor per world in batched scenes): the
    // host sizes the grid from the last count it has read, the kernel strides over whatever there is
    const unsigned perWorld = (unsigned)(n / ctx->cfg.num_worlds + 1);
    const int lhint = ctx->stats.large_proxies > 16 ? ctx->stats.large_proxies : 16;
    dim3 lg(gridFor(perWorld, 256, 64), (unsigned)(lhint < 8192 ? lhint : 8192));
    mark(ctx, 5);
    k_large<<<lg, 256, 0, sl>>>(ctx->dNSorted, ctx->dSmin, ctx->dSmax, ctx->dRowStart, ctx->dGrid, ctx->B.world, ctx->cfg.num_worlds,
                               ctx->uidBits, ctx->dPairKeys, rowCnt, (uint32_t)ctx->cfg.max_pairs, ctx->dCtr, ctx->slab,
                               ctx->sap.enabled ? ctx->B.leafMin : nullptr, ctx->sap.enabled ? ctx->B.leafMax : nullptr);
    if (ctx->overlap) {
        CK(cudaEventRecord(ctx->evJoin[2], sl));
        CK(cudaStreamWaitEvent(s, ctx->evJoin[2], 0));
    }
    mark(ctx, 6);
    // canonical (uid0, uid1) order: rows keyed by uid0 (pair_rows.cuh); rowStart is also the "first pair of uid0" table
    uint32_t* rowStart = ctx->dPairFirst[cur];
    k_row_scan<<<ctx->rowTiles, 256, 0, s>>>(rowCnt, ctx->nRows, rowStart, rowStatus, rowMisc, ctx->dBigRows, ctx->dNumPairs[cur]);
    k_row_scatter<<<gridFor((uint32_t)ctx->cfg.max_pairs, 256), 256, 0, s>>>(ctx->dPairKeys, ctx->dCtr, (uint32_t)ctx->cfg.max_pairs,
                                                                             ctx->uidBits, rowStart, ctx->dCsr);
    // the long rows (a handful of blocks with serial phases) are ordered beside the short ones
    cudaStream_t sb = s;
    if (ctx->overlap) {
        sb = ctx->streamClosed;
        CK(cudaEventRecord(ctx->evFork[3], s));
        CK(cudaStreamWaitEvent(sb, ctx->evFork[3], 0));
    }
    k_row_sort_big<<<148, BIG_THREADS, 0, sb>>>(rowStart, ctx->dBigRows, rowMisc, ctx->dCsr, ctx->uidBits, ctx->dPairs,
                                                ctx->dSortedKeys[cur]);
    k_row_rank<<<gridFor((uint32_t)ctx->cfg.max_pairs, 256), 256, 0, s>>>(ctx->dPairKeys, ctx->dCtr, (uint32_t)ctx->cfg.max_pairs,
                                                                          ctx->uidBits, rowStart, ctx->dCsr, ctx->dPairs,
                                                                          ctx->dSortedKeys[cur]);
    if (ctx->overlap) {
        CK(cudaEventRecord(ctx->evJoin[3], sb));
        CK(cudaStreamWaitEvent(s, ctx->evJoin[3], 0));
    }
    mark(ctx, 7);
    // manifolds follow their pair into the new list (done here so a step without dispatch keeps them too)
    k_carry<<<gridFor((uint32_t)ctx->cfg.max_pairs, 256), 256, 0, s>>>(ctx->dSortedKeys[cur], ctx->dNumPairs[cur],
                                                                      ctx->dSortedKeys[cur ^ 1], ctx->dNumPairs[cur ^ 1],
                                                                      ctx->dPairFirst[cur ^ 1], ctx->dMHdr[cur ^ 1], ctx->dMPts[cur ^ 1],
                                                                      ctx->dMHdr[cur], ctx->dMPts[cur], ctx->uidBits, ctx->dCtr, ctx->dHist, ctx->dPairFirst[cur]);
    ctx->launches += 9 + ctx->sortBodies.launches;  // k_keys, k_gather, k_sweep, k_large, 4 pair-row kernels, k_carry (k_aabb counts itself)
    if (ctx->deltaPrefetch) {
        // pair-cache events of this calculation (added / removed pairs), ready together with the pair list so that their
        // download overlaps the narrowphase
        const unsigned dg = gridFor((uint32_t)ctx->cfg.max_pairs, 256);
        const int prev = cur ^ 1;
        CK(cudaMemsetAsync(ctx->dDeltaCounts, 0, 2 * sizeof(uint32_t), s));
        k_pair_delta<<<dg, 256, 0, s>>>(ctx->dSortedKeys[cur], ctx->dNumPairs[cur], ctx->dSortedKeys[prev], ctx->dNumPairs[prev],
                                        ctx->dPairFirst[prev], ctx->uidBits, ctx->dDelta[0], (uint32_t)ctx->cfg.max_pairs, ctx->dDeltaCounts);
        k_pair_delta<<<dg, 256, 0, s>>>(ctx->dSortedKeys[prev], ctx->dNumPairs[prev], ctx->dSortedKeys[cur], ctx->dNumPairs[cur],
                                        ctx->dPairFirst[cur], ctx->uidBits, ctx->dDelta[1], (uint32_t)ctx->cfg.max_pairs, ctx->dDeltaCounts + 1);
        ctx->launches += 2;
    }
    ctx->nSortedBodies = n;  // dSmin.w = proxy index of every sorted position: the ray tests reuse the order
    CK(cudaGetLastError());
    CK(cudaEventRecordWithFlags(ctx->evPairsReady, s, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
    ctx->step++;
    ctx->pairsValid = true;
    return B2C_OK;
}

static inline void mark(b2c_ctx* ctx, int k) {
    if (ctx->prof) cudaEventRecord(ctx->stageEv[k], ctx->stream);
}

NpArgs makeNpArgs(b2c_ctx* ctx) {
    NpArgs a;
    a.pairs = ctx->dPairs;
    a.numPairs = ctx->dNumPairs[ctx->cur];
    a.xf4 = ctx->B.xf4;
    a.shape = ctx->B.shape;
    a.flags = ctx->B.flags;
    a.material = ctx->B.material;
    a.shapes = ctx->dShapes;
    a.hullPts = ctx->dHullPts;
    a.meshes = ctx->dMeshes;
    a.mhdr = ctx->dMHdr[ctx->cur];
    a.mpts = ctx->dMPts[ctx->cur];
    a.raw = ctx->dRaw;
    a.rawFlag = ctx->dRawFlag;
    a.wantRaw = ctx->wantRaw ? 1 : 0;
    a.hist = ctx->dHist;
    a.binOf = ctx->dBinOf;
    a.binItems = ctx->dBinItems;
    a.binStart = ctx->dBinStart;
    a.binZero = ctx->dBinZero;
    a.ctr = ctx->dCtr;
    a.threshold = ctx->cfg.contact_breaking_threshold;
    a.maxPairs = (uint32_t)ctx->cfg.max_pairs;
    a.noCollide = ctx->dNoCollide;
    a.numNoCollide = ctx->numNoCollide;
    a.uidBits = ctx->uidBits;
    a.hasCompound = ctx->hasCompound ? 1 : 0;
    return a;
}

CompoundArgs makeCompoundArgs(b2c_ctx* ctx) {
    CompoundArgs c;
    c.children = ctx->dChildren;
    c.cc = ctx->dCompoundCtr;
    c.itemPair = ctx->dCItemPair;
    c.itemCode = ctx->dCItemCode;
    c.itemPrev = ctx->dCItemPrev;
    c.raw = ctx->dCRaw;
    c.meshStart = ctx->dCMeshStart;
    c.meshCount = ctx->dCMeshCount;
    c.bigScratch = ctx->dCBigScratch;
    c.numBigScratch = ctx->numCBigScratch;
    c.H = ctx->dCH[ctx->ccur ^ 1];
    c.P = ctx->dCP[ctx->ccur ^ 1];
    c.prevH = ctx->dCH[ctx->ccur];
    c.prevP = ctx->dCP[ctx->ccur];
    c.maxItems = ctx->maxCompoundItems;
    return c;
}

// Compaction of the touching manifolds into dContactHdr / dContactPts (format `mode`, see getContactsImpl)
// phase 0: everything (counts cleared first); 1: early part of the pair manifolds (counts cleared first, then snapshot into
// counts[2..3]); 2: the rest of the pair manifolds + the compound child manifolds, appended behind the early part
int32_t enqueueContactCompaction(b2c_ctx* ctx, int mode, int phase = 0) {
    cudaStream_t s = ctx->stream;
    NpArgs a = makeNpArgs(ctx);
    if (phase != 2) CK(cudaMemsetAsync(ctx->dContactCounts, 0, 4 * sizeof(uint32_t), s));
    auto launch = [&](const NpArgs& na, unsigned grid, const uint32_t* itemPair) {
        if (mode == 3)
            k_compact_contacts<3><<<grid, 256, 0, s>>>(na, ctx->dContactHdr, ctx->dContactPts, ctx->capContactHdr, ctx->capContactPts,
                                                       ctx->dContactCounts, itemPair, itemPair ? 0 : phase);
        else if (mode == 2)
            k_compact_contacts<2><<<grid, 256, 0, s>>>(na, ctx->dContactHdr, ctx->dContactPts, ctx->capContactHdr, ctx->capContactPts,
                                                       ctx->dContactCounts, itemPair, itemPair ? 0 : phase);
        else if (mode == 1)
            k_compact_contacts<1><<<grid, 256, 0, s>>>(na, ctx->dContactHdr, ctx->dContactPts, ctx->capContactHdr, ctx->capContactPts,
                                                       ctx->dContactCounts, itemPair, itemPair ? 0 : phase);
        else
            k_compact_contacts<0><<<grid, 256, 0, s>>>(na, ctx->dContactHdr, ctx->dContactPts, ctx->capContactHdr, ctx->capContactPts,
                                                       ctx->dContactCounts, itemPair, itemPair ? 0 : phase);
    };
    launch(a, gridFor((uint32_t)ctx->cfg.max_pairs, 256), nullptr);
    if (phase == 1) {
        CK(cudaMemcpyAsync(ctx->dContactCounts + 2, ctx->dContactCounts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToDevice, s));
        CK(cudaGetLastError());
        return B2C_OK;
    }
    if (ctx->hasCompound) {  // the child manifolds of compound pairs: the same compaction over the latest item arrays
        NpArgs a2 = a;
        a2.mhdr = ctx->dCH[ctx->ccur];
        a2.mpts = ctx->dCP[ctx->ccur];
        a2.numPairs = &ctx->dCompoundCtr->numItems;
        launch(a2, gridFor(ctx->maxCompoundItems, 256), ctx->dCItemPair);
    }
    CK(cudaGetLastError());
    return B2C_OK;
}

int32_t enqueueNarrowphase(b2c_ctx* ctx) {
    if (!ctx->pairsValid) {
        ctx->err = "dispatch_all_pairs before calculate_overlapping_pairs";
        return B2C_ERR_STATE;
    }
    cudaStream_t s = ctx->stream;
    NpArgs a = makeNpArgs(ctx);
    GjkArgs g;
    g.epaItems = ctx->dEpaItems;
    g.maxEpa = ctx->maxEpa;
    g.epaRetry = ctx->dEpaRetry;
    g.epaBig = ctx->dEpaBig;
    g.maxEpaRetry = ctx->maxEpaRetry;
    g.meshPair = ctx->dMeshPair;
    g.meshTri = ctx->dMeshTri;
    g.rawMesh = ctx->dRawMesh;
    g.meshStart = ctx->dMeshStart;
    g.meshCount = ctx->dMeshCount;
    g.maxMeshItems = (uint32_t)ctx->cfg.max_mesh_items;
    g.comp = CompoundArgs{};
    if (ctx->hasCompound) g.comp = makeCompoundArgs(ctx);
    const unsigned pg = gridFor((uint32_t)ctx->cfg.max_pairs, 256);
    mark(ctx, 8);
    k_clear_np_counters<<<1, 32, 0, s>>>(ctx->dCtr);
    CK(cudaMemsetAsync(ctx->dZeroNp, 0, ctx->zeroNpBytes, s));  // bin partition, survivor partition, work cursors
    k_classify<<<pg, 256, 0, s>>>(a);
    // stable partition of the pair indices by bin, one pass (k_partition16)
    {
        unsigned bg = ctx->binTiles < 148u * 4u ? ctx->binTiles : 148u * 4u;
        k_partition16<<<bg ? bg : 1, 256, 0, s>>>(ctx->dBinOf, ctx->dNumPairs[ctx->cur], ctx->dBinZero, ctx->dBinStart, nullptr,
                                                  ctx->dBinItems);
    }
    ctx->launches += 3;  // k_clear_np_counters, k_classify, k_partition16 (kernels only; memsets are not counted)
    mark(ctx, 9);
    // the closed-form bins touch only their own pairs' records: they run beside the GJK kernels
    cudaStream_t sc = s;
    if (ctx->overlap) {
        sc = ctx->streamClosed;
        CK(cudaEventRecord(ctx->evFork[0], s));
        CK(cudaStreamWaitEvent(sc, ctx->evFork[0], 0));
    }
    k_sphere_sphere<<<148 * 4, 256, 0, sc>>>(a);
    ctx->launches += 1;  // k_sphere_sphere
    if (ctx->hasPlane) { k_convex_plane<<<148 * 2, 256, 0, sc>>>(a); ctx->launches++; }
    if (ctx->overlap) CK(cudaEventRecord(ctx->evJoin[0], sc));
    mark(ctx, 10);
    k_gjk_prefilter<<<148 * 8, 256, 0, s>>>(a, ctx->dSurvivors, ctx->dCursors + 2, ctx->dSurvKey, ctx->dSurvZero);
    {
        unsigned bg = ctx->binTiles < 148u * 2u ? ctx->binTiles : 148u * 2u;
        k_partition16<<<bg ? bg : 1, 256, 0, s>>>(ctx->dSurvKey, ctx->dCursors + 2, ctx->dSurvZero, ctx->dSurvStart, ctx->dSurvivors,
                                                  ctx->dSurvSorted);
    }
    if (ctx->prof) cudaEventRecord(ctx->evGjk[0], s);
    k_gjk<<<148 * GJK_MINB, 128, 0, s>>>(a, g, ctx->dCursors, ctx->dSurvSorted, ctx->dCursors + 2, ctx->dSurvStart);
    if (ctx->prof) cudaEventRecord(ctx->evGjk[1], s);
    ctx->launches += 3;
    if (ctx->hasMesh) {
        k_mesh_query<<<148 * 4, 128, 0, s>>>(a, g);
        k_gjk_tri<<<148 * 4, 128, 0, s>>>(a, g, ctx->dCursors + 1);
        ctx->launches += 2;
    }
    if (ctx->hasCompound) {
        // CompoundShape pairs: expand into child work items, run their detectors; items that need the penetration solver
        // join the bin below, the per-child manifolds follow at the end of the dispatch
        CK(cudaMemsetAsync(ctx->dCompoundCtr, 0, sizeof(CompoundCounters), s));
        k_compound_expand<<<148, 128, 0, s>>>(a, g.comp);
        k_compound_gjk<<<148 * 4, 128, 0, s>>>(a, g, ctx->dCursors + 3);
        ctx->launches += 2;
    }
    mark(ctx, 11);
    // the penetration bin is a handful of long, latency-bound lanes: it runs on its own (high-priority) stream while
    // k_manifold_cc streams through the manifolds of all the other pairs
    cudaStream_t se = s;
    if (ctx->overlap) {
        se = ctx->streamEpa;
        CK(cudaEventRecord(ctx->evFork[1], s));
        CK(cudaStreamWaitEvent(se, ctx->evFork[1], 0));
    }
    {
        static_assert(EPA_SMALL_STRIDE % 8 == 4, "lane chunks need an odd word stride");
        const int smem = EPA_BLOCK * EPA_SMALL_STRIDE;
        // blocks of two kernels share an SM only when both run with the same L1/shared split: k_manifold_cc has to ask for
        // the split the shared-memory EPA pools force, or it would wait for every EPA block to retire
        if (ctx->timeline) cudaEventRecord(ctx->tl[0], se);
        // which variant: the host knows the size of the previous step's bin (when it has read the counters); the kernels
        // handle any count either way, so a stale hint only costs time.  Without a hint both are launched and the device decides.
        const int hint = ctx->epaHint;
        if (hint <= 0) { k_epa<0><<<EPA_GRID, 32 * (32 / ctx->epaLpw), smem, se>>>(a, g, hint < 0 ? 0 : 1, ctx->epaLpw); ctx->launches++; }
        if (ctx->timeline) cudaEventRecord(ctx->tl[1], se);
        if (hint != 0) { k_epa<2><<<EPA_GRID3, EPA_BLOCK3, 0, se>>>(a, g, hint < 0 ? 0 : 1, 32); ctx->launches++; }
    }
    {
        const int smem1 = (EPA_BLOCK2 / 32) * (int)sizeof(EpaScratch);
        k_epa<1><<<EPA_GRID2, EPA_BLOCK2, smem1, se>>>(a, g, 0, 32);  // retry tier + the manifolds of the whole bin
        // the items whose pair overflowed the small pools before (epaRoute) run in the large-pool tier from the start, on a
        // third stream beside the two kernels above — at C3 this halves the penetration chain (tier 0, THEN tier 1 for ~20 items)
        cudaStream_t sp = ctx->overlap ? ctx->streamClosed : se;
        if (ctx->overlap) CK(cudaStreamWaitEvent(sp, ctx->evFork[1], 0));
        k_epa<1><<<EPA_GRID2, EPA_BLOCK2, smem1, sp>>>(a, g, 1, 32);
        if (ctx->overlap) CK(cudaEventRecord(ctx->evJoin[4], sp));
        ctx->launches++;
    }
    if (ctx->timeline) cudaEventRecord(ctx->tl[2], se);
    if (ctx->overlap) CK(cudaEventRecord(ctx->evJoin[1], se));
    if (ctx->timeline) cudaEventRecord(ctx->tl[3], s);
    k_manifold_cc<<<148 * ctx->mccBlocks, 256, 0, s>>>(a);
    if (ctx->timeline) cudaEventRecord(ctx->tl[4], s);
    ctx->launches += 2;
    ctx->earlyRecorded = false;
    if (ctx->contactPrefetch >= 2) {
        // every manifold outside the penetration bin and the mesh bin is final now: compact those and let the host start their
        // download (b2c_begin_contact_download) while the penetration bin is still running on its stream
        if (ctx->overlap) CK(cudaStreamWaitEvent(s, ctx->evJoin[0], 0));
        int32_t rce = enqueueContactCompaction(ctx, ctx->contactPrefetch, 1);
        if (rce) return rce;
        ctx->launches += 1;
        CK(cudaEventRecordWithFlags(ctx->evContactsEarly, s, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        ctx->earlyRecorded = true;
    }
    if (ctx->overlap) {
        CK(cudaStreamWaitEvent(s, ctx->evJoin[1], 0));
        CK(cudaStreamWaitEvent(s, ctx->evJoin[0], 0));
        CK(cudaStreamWaitEvent(s, ctx->evJoin[4], 0));
    }
    if (ctx->hasMesh) { k_mesh_manifold<<<148 * 4, 128, 0, s>>>(a, g); ctx->launches++; }
    if (ctx->hasCompound) {
        // per-child manifolds of the compound pairs, once every detector (and the penetration bin) has finished
        k_compound_manifold<<<gridFor(ctx->maxCompoundItems, 128, 148 * 8), 128, 0, s>>>(a, g.comp);
        ctx->launches += 1;
        if (ctx->hasMesh) {  // child x triangle mesh: BVH query + per-triangle detector + fold, one thread per child work item
            k_compound_mesh<<<COMPOUND_MESH_GRID, COMPOUND_MESH_BLOCK, 0, s>>>(a, g);
            ctx->launches += 1;
        }
        ctx->ccur ^= 1;
    }
    mark(ctx, 12);
    ctx->stageValid = ctx->prof;
    CK(cudaGetLastError());
    if (ctx->contactPrefetch >= 0) {  // the contact stream is compacted behind the dispatch: its getter only copies
        int32_t rc = enqueueContactCompaction(ctx, ctx->contactPrefetch, ctx->contactPrefetch >= 2 ? 2 : 0);
        if (rc) return rc;
        ctx->launches += ctx->hasCompound ? 2 : 1;
    }
    return B2C_OK;
}

int32_t readCounters(b2c_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->hCtrPinned, ctx->dCtr, sizeof(StepCounters), cudaMemcpyDeviceToHost, ctx->stream));
    if (ctx->contactReady >= 0)
        CK(cudaMemcpyAsync(ctx->hCtrPinned + 1, ctx->dContactCounts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->contactReady >= 0) { memcpy(ctx->contactCounts, ctx->hCtrPinned + 1, sizeof(ctx->contactCounts)); ctx->contactCountsValid = true; }
    const StepCounters& c = *ctx->hCtrPinned;
    uint32_t np = c.pairCount;
    ctx->stats.num_pairs = (int32_t)(np < (uint32_t)ctx->cfg.max_pairs ? np : (uint32_t)ctx->cfg.max_pairs);
    ctx->stats.num_manifolds = (int32_t)c.numManifolds;
    ctx->stats.num_contacts_added = (int32_t)c.contactsAdded;
    ctx->stats.gjk_checks = (int32_t)c.gjkChecks;
    ctx->stats.deep_penetration_checks = (int32_t)c.deepChecks;
    ctx->epaHint = c.epaCount > EPA_SMEM_LANES ? 1 : 0;
    ctx->stats.epa_failed = (int32_t)(c.epaFailed & 0x3fffffffu);
    ctx->stats.epa_retries = (int32_t)c.epaRetry;
    ctx->stats.mesh_items = (int32_t)c.meshItems;
    ctx->stats.large_proxies = (int32_t)c.largeCount;
    ctx->lastPairs = ctx->stats.num_pairs;
    ctx->lastManifolds = ctx->stats.num_manifolds;
    ctx->lastContacts = ctx->stats.num_contacts_added;
    if (c.pairOverflow || np > (uint32_t)ctx->cfg.max_pairs) {
        char buf[160];
        snprintf(buf, sizeof buf, "overlapping-pair capacity exceeded: need %u, max_pairs %d", np, ctx->cfg.max_pairs);
        ctx->err = buf;
        return B2C_ERR_CAPACITY;
    }
    if (c.meshOverflow) {
        char buf[160];
        snprintf(buf, sizeof buf, "mesh work-item capacity exceeded: need %u, max_mesh_items %d", c.meshItems, ctx->cfg.max_mesh_items);
        ctx->err = buf;
        return B2C_ERR_CAPACITY;
    }
    if (useRowOrder(ctx)) ctx->rowLenHint = c.maxRowLen;  // the radix path does not measure rows: once chosen it stays
    if (ctx->slab.enabled) {
        uint32_t nl = 0;
        CK(cudaMemcpyAsync(&nl, ctx->dNLocal, sizeof(nl), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        // re-shape the launches only when the local list has outgrown the bound or shrunk a lot (a new shape = a new graph)
        const uint32_t up = slabUpper(ctx);
        if (ctx->localHint == 0 || nl + 1024 > up || (uint64_t)nl * 2 < up) ctx->localHint = nl ? nl : 1;
    }
    if (c.haloOverflow == 2) {
        ctx->err = "partitioned world: a peer never published its boundary proxies for this step (peer-to-peer halo exchange timed out)";
        return B2C_ERR_STATE;
    }
    if (c.haloOverflow) {
        ctx->err = "partitioned world: halo slot too small";
        return B2C_ERR_CAPACITY;
    }
    if (c.migrateOverflow) {
        ctx->err = "partitioned world: manifold migration slot too small";
        return B2C_ERR_CAPACITY;
    }
    if (c.epaFailed >= 0x40000000u) {
        ctx->err = "penetration-solver work list capacity exceeded";
        return B2C_ERR_CAPACITY;
    }
    if (ctx->hasCompound) {
        CompoundCounters cc;
        CK(cudaMemcpyAsync(&cc, ctx->dCompoundCtr, sizeof(cc), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (cc.overflow) {
            char buf[160];
            snprintf(buf, sizeof buf, "compound work-item capacity exceeded: need %u, max_compound_items %u", cc.itemCount,
                     ctx->maxCompoundItems);
            ctx->err = buf;
            return B2C_ERR_CAPACITY;
        }
    }
    return B2C_OK;
}

}  // namespace

// Everything that decides WHAT enqueueBroadphase / enqueueNarrowphase launch (pointers that ping-pong, launch shapes that
// follow host hints, optional kernels).  Two calls with the same signature enqueue identical work, so the captured graph
// of the first serves the second.  kind: 0 = broadphase + narrowphase (one step), 1 = broadphase, 2 = narrowphase.
static void stepSignature(const b2c_ctx* ctx, int kind, uint64_t sig[4]) {
    const int lhint = ctx->stats.large_proxies > 16 ? ctx->stats.large_proxies : 16;
    sig[0] = ((uint64_t)(uint32_t)ctx->nBodies << 32) | (uint32_t)ctx->stagingCount;
    sig[1] = ((uint64_t)(uint32_t)(ctx->cur & 1)) | ((uint64_t)(ctx->extPending ? 1 : 0) << 1) | ((uint64_t)(ctx->hasPlane ? 1 : 0) << 2) |
             ((uint64_t)(ctx->hasMesh ? 1 : 0) << 3) | ((uint64_t)(ctx->overlap ? 1 : 0) << 4) | ((uint64_t)(ctx->aabbPending ? 1 : 0) << 5) |
             ((uint64_t)(uint32_t)kind << 6) | ((uint64_t)(uint32_t)(ctx->epaHint + 1) << 8) | ((uint64_t)(ctx->hasCompound ? 1 : 0) << 10) |
             ((uint64_t)(uint32_t)(ctx->ccur & 1) << 11) | ((uint64_t)(uint32_t)(ctx->contactPrefetch + 1) << 12) | ((uint64_t)(uint32_t)lhint << 16) |
             ((uint64_t)(uint32_t)ctx->slab.rank << 40) | ((uint64_t)(uint32_t)ctx->partRanks << 52);
    sig[2] = (uint64_t)(uintptr_t)ctx->dNoCollide ^ ((uint64_t)slabUpper(ctx) << 40);
    sig[3] = ((uint64_t)ctx->numNoCollide << 32) | (uint32_t)ctx->epaLpw | ((uint32_t)ctx->mccBlocks << 8) |
             ((uint32_t)(ctx->deltaPrefetch ? 1 : 0) << 16) | ((uint32_t)(useRowOrder(ctx) ? 1 : 0) << 17) |
             ((uint32_t)(ctx->wantRaw ? 1 : 0) << 18);
}

static void dropStepGraphs(b2c_ctx* ctx) {
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
}

// Enqueue the broadphase and/or the narrowphase on the ctx stream.  A step is ~28 short kernels plus memsets and
// side-stream joins; issued one by one the host falls behind the device in the broadphase (5-15 us kernels), so the
// sequence is captured once per launch signature into a CUDA graph and replayed with a single cudaGraphLaunch.
static int32_t enqueuePhasesImpl(b2c_ctx* ctx, int kind);
static int32_t enqueuePhases(b2c_ctx* ctx, int kind) {
    ctx->contactReady = -1;  // whatever sat in the contact buffers belongs to an older pair list / dispatch
    ctx->contactCountsValid = false;
    ctx->earlyDl.active = false;
    if (kind != 1) ctx->earlyRecorded = false;
    if (kind != 2) ctx->deltaReady = false;
    int32_t rc = enqueuePhasesImpl(ctx, kind);
    if (rc == B2C_OK && kind != 1) ctx->contactReady = ctx->contactPrefetch;
    if (rc == B2C_OK && kind != 2) ctx->deltaReady = ctx->deltaPrefetch;
    if (rc == B2C_OK && kind != 2 && !ctx->freePending.empty()) {
        // this pair calculation no longer sees the proxies destroyed before it: their slots may be handed out again
        ctx->freeSlots.insert(ctx->freeSlots.end(), ctx->freePending.begin(), ctx->freePending.end());
        ctx->freePending.clear();
    }
    return rc;
}
static int32_t enqueuePhasesImpl(b2c_ctx* ctx, int kind) {
    cudaStream_t s = ctx->stream;
    const bool broad = kind != 2, narrow = kind != 1;
    if (narrow && !broad && !ctx->pairsValid) {
        ctx->err = "dispatch_all_pairs before calculate_overlapping_pairs";
        return B2C_ERR_STATE;
    }
    if (broad && ctx->slab.enabled && !ctx->haloImported) {
        ctx->err = "partitioned world: b2c_mgpu_update_export_halo and b2c_mgpu_import_halo must precede the pair calculation";
        return B2C_ERR_STATE;
    }
    const bool graphable = ctx->useGraphs && !ctx->prof && !ctx->timeline && ctx->nBodies > 0;
    if (!graphable) {
        int32_t rc = B2C_OK;
        if (broad) rc = enqueueBroadphase(ctx);
        if (rc) return rc;
        if (kind == 0) CK(cudaEventRecord(ctx->ev[2], s));
        if (narrow) rc = enqueueNarrowphase(ctx);
        return rc;
    }
    int32_t rc = uploadShapes(ctx);  // not capturable (synchronous copy); a no-op unless shapes were registered since
    if (rc) return rc;
    uint64_t sig[4];
    stepSignature(ctx, kind, sig);
    b2c_ctx::StepGraph* hit = nullptr;
    for (auto& g : ctx->graphs)
        if (g.sig[0] == sig[0] && g.sig[1] == sig[1] && g.sig[2] == sig[2] && g.sig[3] == sig[3]) { hit = &g; break; }
    if (hit) {
        // the host-side state transitions the enqueue functions would have made
        if (broad) {
            ctx->stagingCount = 0;
            ctx->extPending = false;
            ctx->aabbPending = false;
            ctx->cur ^= 1;
            ctx->step++;
            ctx->pairsValid = true;
            ctx->nSortedBodies = ctx->nBodies;
            ctx->haloImported = false;
        }
        if (narrow && ctx->hasCompound) ctx->ccur ^= 1;
        if (narrow) ctx->earlyRecorded = ctx->contactPrefetch >= 2;
        ctx->stageValid = false;
        ctx->launches += hit->launches;
        CK(cudaGraphLaunch(hit->exec, s));
        return B2C_OK;
    }
    if (ctx->graphs.size() >= 24) dropStepGraphs(ctx);
    const int launchesBefore = ctx->launches;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true;
    if (broad) rc = enqueueBroadphase(ctx);
    if (rc == B2C_OK && kind == 0) cudaEventRecordWithFlags(ctx->ev[2], s, cudaEventRecordExternal);
    if (rc == B2C_OK && narrow) rc = enqueueNarrowphase(ctx);
    ctx->capturing = false;
    cudaError_t ce = cudaStreamEndCapture(s, &graph);
    if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (ce != cudaSuccess) { ctx->err = std::string("cudaStreamEndCapture: ") + cudaGetErrorString(ce); return B2C_ERR_CUDA; }
    b2c_ctx::StepGraph g;
    memcpy(g.sig, sig, sizeof(sig));
    g.launches = ctx->launches - launchesBefore;
    g.exec = nullptr;
    ce = cudaGraphInstantiate(&g.exec, graph, 0);
    cudaGraphDestroy(graph);
    if (ce != cudaSuccess) { ctx->err = std::string("cudaGraphInstantiate: ") + cudaGetErrorString(ce); return B2C_ERR_CUDA; }
    ctx->graphs.push_back(g);
    CK(cudaGraphLaunch(g.exec, s));
    return B2C_OK;
}

static void sapConfigure(b2c_ctx* ctx, const float mn[3], const float mx[3]) {
    const bool wide = ctx->cfg.broadphase_mode == B2C_BP_SAP32;
    const int sentinel = wide ? 0x7fffffff : 0xffff;  // bp/AxisSweep3_32.java:49, bp/AxisSweep3.java:52
    ctx->sap.handleMask = wide ? (int)0xfffffffe : 0xfffe;
    ctx->sap.mask = wide ? 0xffffffffu : 0xffffu;
    for (int c = 0; c < 3; c++) {
        ctx->sap.wmin[c] = mn[c];
        ctx->sap.wmax[c] = mx[c];
        const float size = mx[c] - mn[c];
        ctx->sap.quant[c] = (float)sentinel / size;  // bp/AxisSweep3Internal.java:103-105: maxInt / aabbSize
    }
    ctx->sap.enabled = 1;
}


extern "C" {

void b2c_default_config(b2c_config* cfg) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->device = 0;
    cfg->broadphase_mode = B2C_BP_DBVT;
    cfg->max_bodies = 131072;
    cfg->max_pairs = 2 * 1024 * 1024;
    cfg->max_shapes = 4096;
    cfg->max_hull_points = 1 << 20;
    cfg->max_mesh_items = 1 << 20;
    cfg->num_worlds = 1;
    cfg->contact_breaking_threshold = 0.02f;
    cfg->dbvt_margin = 0.05f;
    cfg->dbvt_predicted_frames = 2.0f;
}

int32_t b2c_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int i = 0; i < n; i++) {
        cudaDeviceProp p;
        if (cudaGetDeviceProperties(&p, i) == cudaSuccess && p.major == 10) ok++;
    }
    return ok;
}

int32_t b2c_create(const b2c_config* cfg, b2c_ctx** out) {
    if (!cfg || !out) return B2C_ERR_BAD_ARG;
    *out = nullptr;
    if (cfg->max_bodies < 1 || cfg->max_pairs < 1 || cfg->max_shapes < 1 || cfg->num_worlds < 1) return B2C_ERR_BAD_ARG;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || cfg->device < 0 || cfg->device >= ndev) {
        cudaGetLastError();
        return B2C_ERR_CUDA;  // no device: there is deliberately no CPU path
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess || prop.major != 10) return B2C_ERR_CUDA;
    if (cfg->broadphase_mode < B2C_BP_TIGHT || cfg->broadphase_mode > B2C_BP_SAP32) return B2C_ERR_BAD_ARG;
    b2c_ctx* ctx = new b2c_ctx();
    ctx->cfg = *cfg;
    ctx->device = cfg->device;
    if (cfg->broadphase_mode >= B2C_BP_SAP16) {
        const float mn[3] = {-1000.f, -1000.f, -1000.f}, mx[3] = {1000.f, 1000.f, 1000.f};
        sapConfigure(ctx, mn, mx);
    }
    auto fail = [&](int32_t rc) { b2c_destroy(ctx); return rc; };
#define CKC(call)                                  \
    do {                                           \
        if ((call) != cudaSuccess) {               \
            cudaGetLastError();                    \
            return fail(B2C_ERR_CUDA);             \
        }                                          \
    } while (0)
    CKC(cudaSetDevice(cfg->device));
    CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    {
        int prLo = 0, prHi = 0;
        CKC(cudaDeviceGetStreamPriorityRange(&prLo, &prHi));
        CKC(cudaStreamCreateWithFlags(&ctx->streamCopy, cudaStreamNonBlocking));
        CKC(cudaEventCreateWithFlags(&ctx->evPairsReady, cudaEventDisableTiming));
        CKC(cudaStreamCreateWithPriority(&ctx->streamClosed, cudaStreamNonBlocking, prLo));
        CKC(cudaStreamCreateWithPriority(&ctx->streamEpa, cudaStreamNonBlocking, prHi));
        for (int i = 0; i < 4; i++) CKC(cudaEventCreateWithFlags(&ctx->evFork[i], cudaEventDisableTiming));
        for (int i = 0; i < 5; i++) CKC(cudaEventCreateWithFlags(&ctx->evJoin[i], cudaEventDisableTiming));
        // one-time kernel attributes (kept out of the per-step path so that it can be captured into a graph)
        cudaFuncSetAttribute(k_epa<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, EPA_BLOCK * EPA_SMALL_STRIDE);
        cudaFuncSetAttribute(k_epa<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (EPA_BLOCK2 / 32) * (int)sizeof(EpaScratch));
        const char* gr = getenv("B2C_GRAPH");
        ctx->useGraphs = !(gr && gr[0] == '0');
        const char* e = getenv("B2C_OVERLAP");  // measurement knob: 0 = everything on one stream
        ctx->overlap = !(e && e[0] == '0');
        const char* mb = getenv("B2C_MCC_BLOCKS");
        if (mb) { int v = atoi(mb); if (v >= 1 && v <= 8) ctx->mccBlocks = v; }
        const char* l = getenv("B2C_EPA_LPW");
        if (l) { int v = atoi(l); if (v == 32 || v == 16 || v == 8 || v == 4) ctx->epaLpw = v; }
        const char* t = getenv("B2C_TIMELINE");
        ctx->timeline = t && t[0] == '1';
        for (int i = 0; i < 6; i++) CKC(cudaEventCreate(&ctx->tl[i]));
    }
    const size_t N = (size_t)cfg->max_bodies, P = (size_t)cfg->max_pairs;
    CKC(dalloc(&ctx->dShapes, (size_t)cfg->max_shapes));
    CKC(dalloc(&ctx->dHullPts, (size_t)(cfg->max_hull_points > 0 ? cfg->max_hull_points : 1) + 4));  // + 4: HullS::support reads in fours
    CKC(dalloc(&ctx->dMeshes, (size_t)cfg->max_shapes));
    CKC(dalloc(&ctx->B.xf4, 3 * N));
    CKC(dalloc(&ctx->B.shape, N));
    CKC(dalloc(&ctx->B.filt, N));
    CKC(dalloc(&ctx->B.flags, N));
    CKC(dalloc(&ctx->B.world, N));
    CKC(dalloc(&ctx->B.effMin, N));
    CKC(dalloc(&ctx->B.effMax, N));
    CKC(dalloc(&ctx->B.leafMin, N));
    CKC(dalloc(&ctx->B.leafMax, N));
    CKC(dalloc(&ctx->B.lastSet, N));
    CKC(dalloc(&ctx->B.material, N));
    CKC(dalloc(&ctx->dStep, (size_t)1));
    CKC(dalloc(&ctx->dStaging, 12 * N));
    CKC(cudaMallocHost((void**)&ctx->hStagingPinned, 12 * N * sizeof(float)));
    CKC(dalloc(&ctx->dExtAabb, 6 * N));
    CKC(dalloc(&ctx->dExtMask, N));
    for (int i = 0; i < 2; i++) {
        CKC(dalloc(&ctx->dKeys[i], N));
        CKC(dalloc(&ctx->dVals[i], N));
        CKC(dalloc(&ctx->dSortedKeys[i], P));
        CKC(dalloc(&ctx->dNumPairs[i], (size_t)1));
        CKC(dalloc(&ctx->dMHdr[i], P));
        CKC(dalloc(&ctx->dMPts[i], 4 * P));
        CKC(dalloc(&ctx->dPairFirst[i], N + 4));
    }
    if (N + 2 >= (1u << 21)) return fail(B2C_ERR_BAD_ARG);  // emitted pair key = slot | uid0 | uid1 in 3 x uidBits <= 63 bits
    CKC(dalloc(&ctx->dPairKeys, P));
    CKC(dalloc(&ctx->dCsr, P));
    ctx->nRows = (uint32_t)N + 2u;
    ctx->rowTiles = (ctx->nRows + RSCAN_TILE - 1) / RSCAN_TILE;
    CKC(dalloc(&ctx->dBigRows, (size_t)ctx->nRows));
    CKC(dalloc(&ctx->dSide, (size_t)4));
    CKC(dalloc(&ctx->dSmin, N + SW_CH));   // + SW_CH: the sweep stages whole chunks (pairfind.cuh)
    CKC(dalloc(&ctx->dSmax, N + SW_CH));
    CKC(dalloc(&ctx->dSrow, N + SW_CH));
    CKC(dalloc(&ctx->dScyz, N));
    CKC(dalloc(&ctx->dNSorted, (size_t)4));
    CKC(dalloc(&ctx->dNLocal, (size_t)4));
    ctx->maxRows = (int)(2 * N + 64 > (size_t)(64 * cfg->num_worlds) ? 2 * N + 64 : (size_t)(64 * cfg->num_worlds));
    if (ctx->maxRows > (1 << 20) - 4) ctx->maxRows = (1 << 20) - 4;  // the row shares a 32-bit key with 12 bits of x
    if ((long long)cfg->num_worlds * 5 > ctx->maxRows) return fail(B2C_ERR_BAD_ARG);
    CKC(dalloc(&ctx->dRowStart, (size_t)ctx->maxRows + 8));
    ctx->roTiles = (uint32_t)((ctx->maxRows + 8 + ROFF_TILE - 1) / ROFF_TILE);
    ctx->roTiles = (ctx->roTiles + 3u) & ~3u;  // keeps rowCount (behind misc and status) 16-byte aligned for its 128-bit loads
    CKC(dalloc(&ctx->dSlots, N));
    { const char* e = getenv("B2C_SORT"); ctx->forceRadix = e && e[0] == 'r'; }
    CKC(dalloc(&ctx->dGrid, (size_t)1));
    CKC(cudaMallocHost((void**)&ctx->hCtrPinned, 2 * sizeof(StepCounters)));  // [1]: the prefetched contact-stream counts
    CKC(ctx->sortBodies.init((uint32_t)N));
    ctx->uidBits = bitsFor((uint32_t)N + 1u);
    CKC(dalloc(&ctx->dPairs, P));
    CKC(dalloc(&ctx->dRaw, P));
    CKC(dalloc(&ctx->dRawFlag, P));
    if (P > (size_t)(1u << 24)) return fail(B2C_ERR_BAD_ARG);  // pair index must fit 24 bits next to the bin byte
    CKC(dalloc(&ctx->dBinOf, P));
    CKC(dalloc(&ctx->dHist, P));
    CKC(dalloc(&ctx->dBinItems, P));
    CKC(dalloc(&ctx->dBinStart, (size_t)32));
    ctx->binTiles = (uint32_t)((P + BIN_TILE - 1) / BIN_TILE);
    {
        auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
        const size_t ctrB = up(sizeof(StepCounters));
        const size_t ordB = up(((size_t)4 + ctx->roTiles + 4 + ctx->maxRows + 8 + 16) * sizeof(uint32_t));
        const size_t rowB = up(((size_t)ctx->nRows + ctx->rowTiles + 8) * sizeof(uint32_t) + sizeof(RowMisc));
        ctx->zeroBpBytes = ctrB + ordB + rowB;
        CKC(dalloc(&ctx->dZeroBp, ctx->zeroBpBytes));
        ctx->dCtr = reinterpret_cast<StepCounters*>(ctx->dZeroBp);
        ctx->dRowOrdZero = reinterpret_cast<uint32_t*>(ctx->dZeroBp + ctrB);
        ctx->dRowZero = reinterpret_cast<uint32_t*>(ctx->dZeroBp + ctrB + ordB);
        const size_t binB = up(32 * sizeof(uint32_t));
        ctx->zeroNpBytes = 2 * binB + 256;
        CKC(dalloc(&ctx->dZeroNp, ctx->zeroNpBytes));
        ctx->dBinZero = reinterpret_cast<uint32_t*>(ctx->dZeroNp);
        ctx->dSurvZero = reinterpret_cast<uint32_t*>(ctx->dZeroNp + binB);
        ctx->dCursors = reinterpret_cast<uint32_t*>(ctx->dZeroNp + 2 * binB);
    }
    CKC(dalloc(&ctx->dExportCount, (size_t)1));
    CKC(dalloc(&ctx->dSurvivors, P));
    CKC(dalloc(&ctx->dSurvSorted, P));
    CKC(dalloc(&ctx->dSurvKey, P));
    CKC(dalloc(&ctx->dSurvStart, (size_t)32));
    ctx->maxEpa = (uint32_t)(P / 4 + 1024);
    CKC(dalloc(&ctx->dEpaItems, (size_t)ctx->maxEpa));
    ctx->maxEpaRetry = ctx->maxEpa;
    CKC(dalloc(&ctx->dEpaRetry, (size_t)ctx->maxEpaRetry));
    CKC(dalloc(&ctx->dEpaBig, (size_t)ctx->maxEpaRetry));
    const size_t MI = (size_t)(cfg->max_mesh_items > 0 ? cfg->max_mesh_items : 1);
    CKC(dalloc(&ctx->dMeshPair, MI));
    CKC(dalloc(&ctx->dMeshTri, MI));
    CKC(dalloc(&ctx->dRawMesh, MI));
    CKC(dalloc(&ctx->dMeshStart, P));
    CKC(dalloc(&ctx->dMeshCount, P));
    for (int i = 0; i < 5; i++) CKC(cudaEventCreate(&ctx->ev[i]));
    for (int i = 0; i <= B2C_NUM_STAGES; i++) CKC(cudaEventCreate(&ctx->stageEv[i]));
    for (int i = 0; i < 2; i++) CKC(cudaEventCreate(&ctx->evGjk[i]));
    ctx->capContactHdr = (uint32_t)P;
    ctx->capContactPts = (uint32_t)std::min<size_t>(4 * P, 0xfffffff0u);  // a manifold holds up to 4 points (np/PersistentManifold.java:47)
    CKC(dalloc(&ctx->dContactHdr, (size_t)ctx->capContactHdr));
    CKC(dalloc(&ctx->dContactPts, (size_t)ctx->capContactPts));
    CKC(dalloc(&ctx->dContactCounts, (size_t)4));
    CKC(cudaEventCreateWithFlags(&ctx->evContactsEarly, cudaEventDisableTiming));
    CKC(cudaMallocHost((void**)&ctx->hEarlyCountsPinned, 2 * sizeof(uint32_t)));
#undef CKC
    *out = ctx;
    return B2C_OK;
}

void b2c_destroy(b2c_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
    ctx->graphs.clear();
    cudaFree(ctx->dShapes); cudaFree(ctx->dHullPts); cudaFree(ctx->dMeshes);
    for (auto& m : ctx->meshes) { cudaFree(m.nodes); cudaFree(m.verts); cudaFree(m.idx); cudaFree(m.partStart); }
    cudaFree(ctx->B.xf4); cudaFree(ctx->B.shape); cudaFree(ctx->B.filt); cudaFree(ctx->B.flags); cudaFree(ctx->B.world);
    cudaFree(ctx->B.effMin); cudaFree(ctx->B.effMax); cudaFree(ctx->B.leafMin); cudaFree(ctx->B.leafMax);
    cudaFree(ctx->B.lastSet); cudaFree(ctx->B.material);
    cudaFree(ctx->dStep);
    cudaFree(ctx->dStaging); cudaFreeHost(ctx->hStagingPinned); cudaFree(ctx->dExtAabb); cudaFree(ctx->dExtMask);
    for (int i = 0; i < 2; i++) {
        cudaFree(ctx->dKeys[i]); cudaFree(ctx->dVals[i]); cudaFree(ctx->dSortedKeys[i]);
        cudaFree(ctx->dNumPairs[i]); cudaFree(ctx->dMHdr[i]); cudaFree(ctx->dMPts[i]); cudaFree(ctx->dPairFirst[i]);
    }
    cudaFree(ctx->dZeroBp); cudaFree(ctx->dZeroNp); cudaFree(ctx->dSlots);
    cudaFree(ctx->dNSorted); cudaFree(ctx->dNLocal); cudaFree(ctx->dOwner); cudaFree(ctx->dLocalList);
    for (int r = 0; r < 16; r++) if (ctx->haloIpcOpened[r]) cudaIpcCloseMemHandle(ctx->haloIpcOpened[r]);
    cudaFree(ctx->dHaloInbox); cudaFree(ctx->dHaloP2p); cudaFree(ctx->dHaloStage); cudaFree(ctx->dMigLocal);
    cudaFree(ctx->dSide); cudaFree(ctx->dSmin); cudaFree(ctx->dSmax); cudaFree(ctx->dSrow); cudaFree(ctx->dScyz); cudaFree(ctx->dRowStart);
    cudaFree(ctx->dGrid); cudaFreeHost(ctx->hCtrPinned);
    ctx->sortBodies.destroy();
    cudaFree(ctx->dPairKeys); cudaFree(ctx->dCsr); cudaFree(ctx->dBigRows);
    cudaFree(ctx->dPairs); cudaFree(ctx->dRaw); cudaFree(ctx->dRawFlag); cudaFree(ctx->dBinOf); cudaFree(ctx->dHist); cudaFree(ctx->dBinItems); cudaFree(ctx->dBinStart); cudaFree(ctx->dExportCount); cudaFree(ctx->dSurvivors); cudaFree(ctx->dSurvSorted); cudaFree(ctx->dSurvKey); cudaFree(ctx->dSurvStart);
    cudaFree(ctx->dEpaItems); cudaFree(ctx->dEpaRetry); cudaFree(ctx->dEpaBig); cudaFree(ctx->dMeshPair);
    cudaFree(ctx->dMeshTri); cudaFree(ctx->dRawMesh); cudaFree(ctx->dMeshStart); cudaFree(ctx->dMeshCount);
    for (int i = 0; i < 5; i++) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
    for (int i = 0; i <= B2C_NUM_STAGES; i++) if (ctx->stageEv[i]) cudaEventDestroy(ctx->stageEv[i]);
    for (int i = 0; i < 2; i++) if (ctx->evGjk[i]) cudaEventDestroy(ctx->evGjk[i]);
    cudaFree(ctx->dContactHdr); cudaFree(ctx->dContactPts); cudaFree(ctx->dContactCounts);
    cudaFree(ctx->dRayChunkMin); cudaFree(ctx->dRayChunkMax); cudaFree(ctx->dRayMin); cudaFree(ctx->dRayMax); cudaFree(ctx->dRayIn); cudaFree(ctx->dRayOut); cudaFree(ctx->dRayOverflow);
    cudaFree(ctx->dSweepIn); cudaFree(ctx->dSweepOut);
    cudaFree(ctx->dNoCollide);
    cudaFree(ctx->dChildren); cudaFree(ctx->dCompoundCtr); cudaFree(ctx->dCItemPair); cudaFree(ctx->dCItemCode); cudaFree(ctx->dCItemPrev);
    cudaFree(ctx->dCRaw); cudaFree(ctx->dCMeshStart); cudaFree(ctx->dCMeshCount); cudaFree(ctx->dCBigScratch);
    for (int i = 0; i < 2; i++) { cudaFree(ctx->dCH[i]); cudaFree(ctx->dCP[i]); }
    cudaFree(ctx->dIslandPar); cudaFree(ctx->dIslandTags); cudaFree(ctx->dDelta[0]); cudaFree(ctx->dDelta[1]); cudaFree(ctx->dDeltaCounts);
    for (int i = 0; i < 4; i++) if (ctx->evFork[i]) cudaEventDestroy(ctx->evFork[i]);
    for (int i = 0; i < 5; i++) if (ctx->evJoin[i]) cudaEventDestroy(ctx->evJoin[i]);
    if (ctx->streamCopy) { cudaStreamSynchronize(ctx->streamCopy); cudaStreamDestroy(ctx->streamCopy); }
    if (ctx->evPairsReady) cudaEventDestroy(ctx->evPairsReady);
    if (ctx->evContactsEarly) cudaEventDestroy(ctx->evContactsEarly);
    cudaFreeHost(ctx->hEarlyCountsPinned);
    cudaFreeHost(ctx->hDeltaCountsPinned);
    if (ctx->streamClosed) cudaStreamDestroy(ctx->streamClosed);
    if (ctx->streamEpa) cudaStreamDestroy(ctx->streamEpa);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* b2c_last_error_string(const b2c_ctx* ctx) { return ctx ? ctx->err.c_str() : "null ctx"; }

int32_t b2c_set_world_aabb(b2c_ctx* ctx, const float mn[3], const float mx[3]) {
    if (!ctx || !mn || !mx) return B2C_ERR_BAD_ARG;
    for (int c = 0; c < 3; c++)
        if (!(mx[c] > mn[c])) return B2C_ERR_BAD_ARG;
    if (ctx->cfg.broadphase_mode != B2C_BP_SAP16 && ctx->cfg.broadphase_mode != B2C_BP_SAP32) return B2C_OK;  // only the SAP modes quantise
    if (ctx->nBodies > 0) { ctx->err = "b2c_set_world_aabb after proxies were created"; return B2C_ERR_STATE; }
    sapConfigure(ctx, mn, mx);
    return B2C_OK;
}

// ---- shapes ------------------------------------------------------------------------------------------
// A world with both compounds and triangle meshes: k_compound_mesh finishes deep (child, triangle) penetrations in the thread,
// its rare large polytopes in one full-size pool per thread (global memory)
static int32_t ensureCompoundMeshScratch(b2c_ctx* ctx) {
    if (!ctx->hasCompound || !ctx->hasMesh || ctx->dCBigScratch) return B2C_OK;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);  // graphs captured a null scratch pointer
    ctx->graphs.clear();
    const uint32_t n = (uint32_t)(COMPOUND_MESH_GRID * COMPOUND_MESH_BLOCK);
    CK(cudaMalloc((void**)&ctx->dCBigScratch, (size_t)n * sizeof(EpaScratch)));
    ctx->numCBigScratch = n;
    return B2C_OK;
}

static int32_t addShape(b2c_ctx* ctx, const ShapeDev& s, int32_t* out) {
    if ((int)ctx->hShapes.size() >= ctx->cfg.max_shapes) { ctx->err = "shape table full"; return B2C_ERR_CAPACITY; }
    ctx->hShapes.push_back(s);
    ctx->shapesDirty = true;
    if (out) *out = (int32_t)ctx->hShapes.size() - 1;
    return B2C_OK;
}

int32_t b2c_shape_register_box(b2c_ctx* ctx, const float he[3], float margin, int32_t* out) {
    if (!ctx || !he) return B2C_ERR_BAD_ARG;
    ShapeDev s{};
    s.type = SH_BOX;
    s.margin = margin >= 0.f ? margin : 0.04f;  // BulletGlobals.CONVEX_DISTANCE_MARGIN
    // sh/BoxShape.java:46-50: implicitShapeDimensions = halfExtents * localScaling(1) - margin
    for (int c = 0; c < 3; c++) s.dims[c] = he[c] * 1.0f - s.margin;
    return addShape(ctx, s, out);
}
int32_t b2c_shape_register_sphere(b2c_ctx* ctx, float radius, int32_t* out) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    ShapeDev s{};
    s.type = SH_SPHERE;
    s.dims[0] = radius;
    s.margin = radius * 1.0f;  // getMargin() = getRadius() = implicit.x * localScaling.x (sh/SphereShape.java:83-97)
    return addShape(ctx, s, out);
}
int32_t b2c_shape_register_hull(b2c_ctx* ctx, const float* pts, int32_t n, float margin, int32_t* out) {
    if (!ctx || !pts || n < 1) return B2C_ERR_BAD_ARG;
    if (ctx->hullPtsUsed + n > ctx->cfg.max_hull_points) { ctx->err = "hull point pool full"; return B2C_ERR_CAPACITY; }
    ShapeDev s{};
    s.type = SH_HULL;
    s.margin = margin >= 0.f ? margin : 0.04f;
    s.pointOffset = ctx->hullPtsUsed;
    s.numPoints = n;
    std::vector<float4> packed((size_t)n);
    // sh/PolyhedralConvexShape.java:177-201 recalcLocalAabb: extreme coordinate of point*scaling(1) +- margin
    float mx[3] = {0, 0, 0}, mn[3] = {0, 0, 0};
    float wmx[3] = {-1e30f, -1e30f, -1e30f}, wmn[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = 0; i < n; i++) {
        float v[3] = {pts[3 * i] * 1.0f, pts[3 * i + 1] * 1.0f, pts[3 * i + 2] * 1.0f};
        packed[i] = make_float4(v[0], v[1], v[2], 0.f);
        for (int c = 0; c < 3; c++) {
            float dpos = v[c], dneg = -v[c];  // dot with +-unit axis reduces to +-coordinate (other terms are *0)
            dpos = (c == 0 ? 1.0f * v[0] + 0.0f * v[1] + 0.0f * v[2] : (c == 1 ? 0.0f * v[0] + 1.0f * v[1] + 0.0f * v[2] : 0.0f * v[0] + 0.0f * v[1] + 1.0f * v[2]));
            dneg = (c == 0 ? -1.0f * v[0] + 0.0f * v[1] + 0.0f * v[2] : (c == 1 ? 0.0f * v[0] + -1.0f * v[1] + 0.0f * v[2] : 0.0f * v[0] + 0.0f * v[1] + -1.0f * v[2]));
            if (dpos > wmx[c]) { wmx[c] = dpos; mx[c] = v[c]; }
            if (dneg > wmn[c]) { wmn[c] = dneg; mn[c] = v[c]; }
        }
    }
    for (int c = 0; c < 3; c++) { s.aabbMax[c] = mx[c] + s.margin; s.aabbMin[c] = mn[c] - s.margin; }
    cudaSetDevice(ctx->device);
    CK(cudaMemcpy(ctx->dHullPts + ctx->hullPtsUsed, packed.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice));
    ctx->hullPtsUsed += n;
    return addShape(ctx, s, out);
}
int32_t b2c_shape_register_plane(b2c_ctx* ctx, const float nrm[3], float c, int32_t* out) {
    if (!ctx || !nrm) return B2C_ERR_BAD_ARG;
    ShapeDev s{};
    s.type = SH_PLANE;
    s.margin = 0.f;  // sh/ConcaveShape.java:35
    // sh/StaticPlaneShape.java:45-48: planeNormal.set(n).nor()  (libgdx nor(): untouched if len2 is 0 or 1)
    float l2 = nrm[0] * nrm[0] + nrm[1] * nrm[1] + nrm[2] * nrm[2];
    float k = 1.0f;
    bool scale = !(l2 == 0.f || l2 == 1.f);
    if (scale) k = 1.0f / (float)std::sqrt((double)l2);
    for (int i = 0; i < 3; i++) s.plane[i] = scale ? nrm[i] * k : nrm[i];
    s.plane[3] = c;
    ctx->hasPlane = true;
    return addShape(ctx, s, out);
}
int32_t b2c_shape_register_mesh(b2c_ctx* ctx, const void* vbase, int32_t nv, int32_t vstride, const void* ibase, int32_t nt,
                                int32_t istride, const float scaling[3], int32_t* out) {
    b2c_indexed_mesh part{vbase, nv, vstride, ibase, nt, istride, B2C_INDEX_INT32};
    return b2c_shape_register_mesh_parts(ctx, &part, 1, scaling, out);
}

// sh/TriangleIndexVertexArray.java:72-100 (a list of IndexedMesh parts, each with its own index type) behind
// sh/BvhTriangleMeshShape.java:68-90.  The parts are concatenated on the host (vertices pre-multiplied by the scaling,
// indices rebased), the quantized BVH is built over the triangles in part order with leaf words partId << 21 | index.
int32_t b2c_shape_register_mesh_parts(b2c_ctx* ctx, const b2c_indexed_mesh* parts, int32_t nparts, const float scaling[3],
                                      int32_t* out) {
    if (!ctx || !parts || nparts < 1) return B2C_ERR_BAD_ARG;
    if (nparts > (1 << 10)) { ctx->err = "more than 1024 mesh parts (sh/OptimizedBvh.java:65 MAX_NUM_PARTS_IN_BITS)"; return B2C_ERR_BAD_ARG; }
    float sc[3] = {1.f, 1.f, 1.f};
    if (scaling) { sc[0] = scaling[0]; sc[1] = scaling[1]; sc[2] = scaling[2]; }
    size_t totalV = 0, totalT = 0;
    for (int p = 0; p < nparts; p++) {
        const b2c_indexed_mesh& m = parts[p];
        if (!m.vertex_base || !m.index_base || m.num_vertices < 1 || m.num_triangles < 1 || m.vertex_stride < 12) return B2C_ERR_BAD_ARG;
        if (m.index_type != B2C_INDEX_INT16 && m.index_type != B2C_INDEX_INT32) { ctx->err = "mesh index type must be SHORT or INTEGER"; return B2C_ERR_BAD_ARG; }
        // sh/TriangleIndexVertexArray.java:94: the per-index stride is triangleIndexStride / 3
        if (m.index_stride / 3 < m.index_type) { ctx->err = "triangle index stride smaller than three indices"; return B2C_ERR_BAD_ARG; }
        if (m.num_triangles >= (1 << 21)) { ctx->err = "mesh part exceeds 2^21 triangles (sh/OptimizedBvh.java:65)"; return B2C_ERR_BAD_ARG; }
        totalV += (size_t)m.num_vertices;
        totalT += (size_t)m.num_triangles;
    }
    if (totalT > 0x7fffffffu / 3 || totalV > 0x7fffffffu / 3) return B2C_ERR_BAD_ARG;
    std::vector<float> verts(3 * totalV);
    std::vector<int32_t> idx(3 * totalT), leafWord(totalT), partStart((size_t)nparts + 1);
    size_t v0 = 0, t0 = 0;
    for (int p = 0; p < nparts; p++) {
        const b2c_indexed_mesh& m = parts[p];
        for (int i = 0; i < m.num_vertices; i++) {
            const float* q = (const float*)((const char*)m.vertex_base + (size_t)i * m.vertex_stride);
            for (int c = 0; c < 3; c++) verts[3 * (v0 + i) + c] = q[c] * sc[c];  // sh/VertexData.java:50-55
        }
        const size_t istep = (size_t)(m.index_stride / 3);
        partStart[p] = (int32_t)t0;
        for (int t = 0; t < m.num_triangles; t++) {
            for (int c = 0; c < 3; c++) {
                const char* q = (const char*)m.index_base + ((size_t)3 * t + c) * istep;  // sh/ByteBufferVertexData.java:75-84
                int32_t v;
                if (m.index_type == B2C_INDEX_INT16) { uint16_t h; memcpy(&h, q, 2); v = (int32_t)h; }  // getShort & 0xFFFF
                else memcpy(&v, q, 4);
                if (v < 0 || v >= m.num_vertices) { ctx->err = "mesh index out of range"; return B2C_ERR_BAD_ARG; }
                idx[3 * (t0 + t) + c] = v + (int32_t)v0;
            }
            leafWord[t0 + t] = (int32_t)(((uint32_t)p << 21) | (uint32_t)t);
        }
        v0 += (size_t)m.num_vertices;
        t0 += (size_t)m.num_triangles;
    }
    partStart[nparts] = (int32_t)t0;
    HostMesh hm;
    buildQuantizedBvh(verts.data(), idx.data(), (int)totalT, hm.bvh, nparts > 1 ? leafWord.data() : nullptr);
    cudaSetDevice(ctx->device);
    size_t nn = hm.bvh.nodes.size() / 4;
    CK(cudaMalloc((void**)&hm.nodes, nn * sizeof(int4)));
    CK(cudaMalloc((void**)&hm.verts, verts.size() * sizeof(float)));
    CK(cudaMalloc((void**)&hm.idx, idx.size() * sizeof(int)));
    CK(cudaMemcpy(hm.nodes, hm.bvh.nodes.data(), nn * sizeof(int4), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(hm.verts, verts.data(), verts.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(hm.idx, idx.data(), idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    MeshDev md{};
    md.nodes = hm.nodes; md.verts = hm.verts; md.idx = hm.idx;
    md.numNodes = (int)nn; md.numTris = (int)totalT;
    md.numParts = nparts;
    md.partStart = nullptr;
    if (nparts > 1) {
        CK(cudaMalloc((void**)&hm.partStart, partStart.size() * sizeof(int)));
        CK(cudaMemcpy(hm.partStart, partStart.data(), partStart.size() * sizeof(int), cudaMemcpyHostToDevice));
        md.partStart = hm.partStart;
    }
    for (int c = 0; c < 3; c++) { md.qmin[c] = hm.bvh.qmin[c]; md.qmax[c] = hm.bvh.qmax[c]; md.quant[c] = hm.bvh.quant[c]; }
    ShapeDev s{};
    s.type = SH_MESH;
    s.margin = 0.f;  // sh/ConcaveShape.java:35
    for (int c = 0; c < 3; c++) { s.aabbMax[c] = hm.bvh.localMax[c] + s.margin; s.aabbMin[c] = hm.bvh.localMin[c] - s.margin; }
    s.mesh = (int)ctx->hMeshes.size();
    ctx->hMeshes.push_back(md);
    ctx->meshes.push_back(std::move(hm));
    ctx->hasMesh = true;
    int32_t rcs = ensureCompoundMeshScratch(ctx);
    if (rcs) return rcs;
    return addShape(ctx, s, out);
}
// sh/CompoundShape.java:50-82: new CompoundShape() then addChildShape(localTransform_i, child_i) for i = 0..n-1
int32_t b2c_shape_register_compound(b2c_ctx* ctx, int32_t n, const int32_t* childShapes, const float* childXf12, int32_t* out) {
    if (!ctx || !childShapes || !childXf12 || n < 1 || n > 32767) return B2C_ERR_BAD_ARG;
    for (int i = 0; i < n; i++) {
        if (childShapes[i] < 0 || childShapes[i] >= (int)ctx->hShapes.size()) return B2C_ERR_BAD_HANDLE;
        const int t = ctx->hShapes[childShapes[i]].type;
        if (t != SH_BOX && t != SH_SPHERE && t != SH_HULL && t != SH_COMPOUND) {
            ctx->err = "compound children must be box, sphere, convex hull or compound shapes";
            return B2C_ERR_BAD_ARG;
        }
    }
    if ((int)ctx->hShapes.size() >= ctx->cfg.max_shapes) { ctx->err = "shape table full"; return B2C_ERR_CAPACITY; }
    if (ctx->partRanks > 1) { ctx->err = "compound shapes are not supported in a partitioned world"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    if (!ctx->dCompoundCtr) {  // first compound: the child work-item arrays
        const int want = ctx->cfg.max_compound_items;
        size_t M = want > 0 ? (size_t)want : std::max<size_t>(65536, (size_t)ctx->cfg.max_pairs);
        if (M > ((size_t)1 << 26)) M = (size_t)1 << 26;
        ctx->maxCompoundItems = (uint32_t)M;
        CK(dalloc(&ctx->dCompoundCtr, (size_t)1));
        CK(dalloc(&ctx->dCItemPair, M));
        CK(dalloc(&ctx->dCItemCode, M));
        CK(dalloc(&ctx->dCItemPrev, M));
        CK(dalloc(&ctx->dCRaw, M));
        CK(dalloc(&ctx->dCMeshStart, M));
        CK(dalloc(&ctx->dCMeshCount, M));
        for (int i = 0; i < 2; i++) {
            CK(dalloc(&ctx->dCH[i], M));
            CK(dalloc(&ctx->dCP[i], 4 * M));
        }
    }
    // the table gets this compound's frames and leaves (compound_flatten.h); behind them, for the AABB kernel below only, its
    // DIRECT children (a nested compound counts with its own local AABB, sh/CompoundShape.java:60-80)
    std::vector<CompoundDirectChild> direct((size_t)n);
    for (int i = 0; i < n; i++) {
        direct[(size_t)i].shape = childShapes[i];
        for (int k = 0; k < 12; k++) direct[(size_t)i].xf12[k] = childXf12[12 * i + k];
    }
    int first = 0, numLeaves = 0;
    const size_t tableBefore = ctx->hChildren.size();
    if (!flattenCompound(ctx->hChildren, direct,
                         [&](int sid) -> const std::vector<CompoundDirectChild>* {
                             return (sid < (int)ctx->compoundDirect.size() && !ctx->compoundDirect[(size_t)sid].empty()) ? &ctx->compoundDirect[(size_t)sid] : nullptr;
                         },
                         first, numLeaves)) {
        ctx->hChildren.resize(tableBefore);
        ctx->err = "compound shapes nested deeper than 4 levels";
        return B2C_ERR_CAPACITY;
    }
    if (numLeaves < 1 || numLeaves > 32767) {
        ctx->hChildren.resize(tableBefore);
        ctx->err = "a compound must have between 1 and 32767 leaf shapes";
        return B2C_ERR_BAD_ARG;
    }
    const size_t tableAfter = ctx->hChildren.size();
    const int firstDirect = (int)tableAfter;
    for (int i = 0; i < n; i++) {
        CompoundChildDev ch{};
        for (int k = 0; k < 9; k++) ch.m[k] = childXf12[12 * i + k];
        for (int k = 0; k < 3; k++) ch.o[k] = childXf12[12 * i + 9 + k];
        ch.shape = childShapes[i];
        ctx->hChildren.push_back(ch);
    }
    if (ctx->hChildren.size() > ctx->capChildren) {  // the table moves: graphs that captured the old pointer are stale
        CK(cudaStreamSynchronize(ctx->stream));
        for (auto& g : ctx->graphs) cudaGraphExecDestroy(g.exec);
        ctx->graphs.clear();
        cudaFree(ctx->dChildren);
        ctx->dChildren = nullptr;
        ctx->capChildren = std::max<size_t>(1024, 2 * ctx->hChildren.size());
        CK(dalloc(&ctx->dChildren, ctx->capChildren));
    }
    CK(cudaMemcpy(ctx->dChildren, ctx->hChildren.data(), ctx->hChildren.size() * sizeof(CompoundChildDev), cudaMemcpyHostToDevice));
    int32_t rc = uploadShapes(ctx);  // the children's records must be on the device for the AABB kernel below
    if (rc) return rc;
    float* d6 = nullptr;
    float h6[6];
    CK(cudaMalloc((void**)&d6, 6 * sizeof(float)));
    k_compound_local_aabb<<<1, 1, 0, ctx->stream>>>(ctx->dShapes, ctx->dChildren, firstDirect, n, d6);
    cudaError_t ce = cudaMemcpyAsync(h6, d6, sizeof(h6), cudaMemcpyDeviceToHost, ctx->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(ctx->stream);
    cudaFree(d6);
    ctx->hChildren.resize(tableAfter);   // the direct-children scratch entries go again
    CK(ce);
    ShapeDev s{};
    s.type = SH_COMPOUND;
    s.margin = 0.f;  // sh/CompoundShape.java:49 collisionMargin
    for (int c = 0; c < 3; c++) { s.aabbMin[c] = h6[c]; s.aabbMax[c] = h6[3 + c]; }
    s.pointOffset = first;
    s.numPoints = numLeaves;
    ctx->hasCompound = true;
    int32_t rcs = ensureCompoundMeshScratch(ctx);
    if (rcs) return rcs;
    int32_t sid = -1;
    int32_t rca = addShape(ctx, s, &sid);
    if (rca) return rca;
    if ((int)ctx->compoundDirect.size() <= sid) ctx->compoundDirect.resize((size_t)sid + 1);
    ctx->compoundDirect[(size_t)sid] = direct;
    if (out) *out = sid;
    return B2C_OK;
}
int32_t b2c_mesh_get_bvh(b2c_ctx* ctx, int32_t shape, void* nodesOut, int32_t cap, int32_t* numNodes, float quant9[9]) {
    if (!ctx || shape < 0 || shape >= (int)ctx->hShapes.size() || ctx->hShapes[shape].type != SH_MESH) return B2C_ERR_BAD_HANDLE;
    const HostMesh& hm = ctx->meshes[ctx->hShapes[shape].mesh];
    int nn = (int)(hm.bvh.nodes.size() / 4);
    if (numNodes) *numNodes = nn;
    if (quant9)
        for (int c = 0; c < 3; c++) { quant9[c] = hm.bvh.qmin[c]; quant9[3 + c] = hm.bvh.qmax[c]; quant9[6 + c] = hm.bvh.quant[c]; }
    if (nodesOut) {
        if (cap < nn) return B2C_ERR_CAPACITY;
        memcpy(nodesOut, hm.bvh.nodes.data(), (size_t)nn * 16);
    }
    return B2C_OK;
}

// ---- proxies -----------------------------------------------------------------------------------------
int32_t b2c_proxy_create(b2c_ctx* ctx, int32_t shape, const float t[12], int16_t group, int16_t mask, int32_t flags, int32_t world,
                         int32_t* uidOut) {
    if (!ctx || !t) return B2C_ERR_BAD_ARG;
    if (shape < 0 || shape >= (int)ctx->hShapes.size()) return B2C_ERR_BAD_HANDLE;
    if (world < 0 || world >= ctx->cfg.num_worlds) return B2C_ERR_BAD_ARG;
    // max_bodies caps the LIVE proxies: when every slot has been handed out, the slot (and uid) of a proxy destroyed before the
    // last pair calculation is reused, lowest first (AxisSweep3 reuses handles the same way, bp/AxisSweep3Internal.java:380-395)
    const bool reuse = ctx->nBodies >= ctx->cfg.max_bodies;
    if (reuse && ctx->freeSlots.empty()) {
        ctx->err = ctx->freePending.empty() ? "proxy capacity exceeded"
                                            : "proxy capacity exceeded (destroyed slots become reusable after the next pair calculation)";
        return B2C_ERR_CAPACITY;
    }
    cudaSetDevice(ctx->device);
    int i = ctx->nBodies;
    if (reuse) {
        auto it = std::min_element(ctx->freeSlots.begin(), ctx->freeSlots.end());
        i = *it;
        ctx->freeSlots.erase(it);
    }
    float4 rows[3] = {make_float4(t[0], t[1], t[2], t[9]), make_float4(t[3], t[4], t[5], t[10]), make_float4(t[6], t[7], t[8], t[11])};
    uint32_t filt = ((uint32_t)(uint16_t)group) | ((uint32_t)(uint16_t)mask << 16);
    uint8_t fl = (uint8_t)(BF_ALIVE | BF_ACTIVE | ((flags & 1) ? BF_STATIC : 0));
    float2 mat = make_float2(0.5f, 0.0f);  // disp/CollisionObject.java:95 friction, :71 restitution
    int32_t rc = uploadShapes(ctx);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->B.xf4 + 3 * (size_t)i, rows, sizeof(rows), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.shape + i, &shape, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.filt + i, &filt, sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.flags + i, &fl, 1, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.world + i, &world, sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.material + i, &mat, sizeof(float2), cudaMemcpyHostToDevice, s));
    int ls = ctx->step;  // createProxy counts as a setAabb in the current window (bp/DbvtBroadphase.java:177-180)
    CK(cudaMemcpyAsync(ctx->B.lastSet + i, &ls, sizeof(int), cudaMemcpyHostToDevice, s));
    k_initial_aabb<<<1, 1, 0, s>>>(ctx->B, ctx->dShapes, i, ctx->sap);
    CK(cudaStreamSynchronize(s));  // host temporaries above go out of scope
    if (reuse) {
        ctx->hFlags[i] = fl;
        ctx->hShapeOf[i] = shape;
    } else {
        ctx->hFlags.push_back(fl);
        ctx->hShapeOf.push_back(shape);
        ctx->nBodies++;
    }
    if (uidOut) *uidOut = i + 1;  // ++gid (a recycled slot keeps its uid)
    return B2C_OK;
}

int32_t b2c_proxy_create_batch(b2c_ctx* ctx, int32_t n, const int32_t* shapes, const float* planes, const int16_t* groups,
                               const int16_t* masks, const int32_t* flags, const int32_t* worlds, int32_t* firstUid) {
    if (!ctx || !shapes || !planes || !groups || !masks || n < 0) return B2C_ERR_BAD_ARG;
    if (ctx->nBodies + n > ctx->cfg.max_bodies) { ctx->err = "proxy capacity exceeded"; return B2C_ERR_CAPACITY; }
    if (firstUid) *firstUid = ctx->nBodies + 1;
    if (n == 0) return B2C_OK;
    for (int k = 0; k < n; k++) {
        if (shapes[k] < 0 || shapes[k] >= (int)ctx->hShapes.size()) return B2C_ERR_BAD_HANDLE;
        if (worlds && (worlds[k] < 0 || worlds[k] >= ctx->cfg.num_worlds)) return B2C_ERR_BAD_ARG;
    }
    cudaSetDevice(ctx->device);
    int32_t rc = uploadShapes(ctx);
    if (rc) return rc;
    const int first = ctx->nBodies;
    std::vector<float4> rows(3 * (size_t)n);
    std::vector<uint32_t> filt((size_t)n);
    std::vector<uint8_t> fl((size_t)n);
    std::vector<int> wl((size_t)n, 0), ls((size_t)n, ctx->step);
    std::vector<float2> mat((size_t)n, make_float2(0.5f, 0.0f));
    const size_t S = (size_t)n;
    for (int k = 0; k < n; k++) {
        const float* p = planes + k;
        rows[3 * (size_t)k] = make_float4(p[0], p[S], p[2 * S], p[9 * S]);
        rows[3 * (size_t)k + 1] = make_float4(p[3 * S], p[4 * S], p[5 * S], p[10 * S]);
        rows[3 * (size_t)k + 2] = make_float4(p[6 * S], p[7 * S], p[8 * S], p[11 * S]);
        filt[k] = ((uint32_t)(uint16_t)groups[k]) | ((uint32_t)(uint16_t)masks[k] << 16);
        fl[k] = (uint8_t)(BF_ALIVE | BF_ACTIVE | ((flags && (flags[k] & 1)) ? BF_STATIC : 0));
        if (worlds) wl[k] = worlds[k];
    }
    cudaStream_t s = ctx->stream;
    CK(cudaMemcpyAsync(ctx->B.xf4 + 3 * (size_t)first, rows.data(), rows.size() * sizeof(float4), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.shape + first, shapes, S * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.filt + first, filt.data(), S * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.flags + first, fl.data(), S, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.world + first, wl.data(), S * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.material + first, mat.data(), S * sizeof(float2), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->B.lastSet + first, ls.data(), S * sizeof(int), cudaMemcpyHostToDevice, s));
    k_initial_aabb_range<<<(n + 255) / 256, 256, 0, s>>>(ctx->B, ctx->dShapes, first, n, ctx->sap);
    CK(cudaStreamSynchronize(s));
    for (int k = 0; k < n; k++) { ctx->hFlags.push_back(fl[k]); ctx->hShapeOf.push_back(shapes[k]); }
    ctx->nBodies += n;
    return B2C_OK;
}

int32_t b2c_proxy_destroy(b2c_ctx* ctx, int32_t uid) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (uid < 1 || uid > ctx->nBodies || !(ctx->hFlags[uid - 1] & BF_ALIVE)) return B2C_ERR_BAD_HANDLE;
    cudaSetDevice(ctx->device);
    ctx->hFlags[uid - 1] = 0;
    uint8_t z = 0;
    CK(cudaMemcpyAsync(ctx->B.flags + (uid - 1), &z, 1, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->freePending.push_back(uid - 1);
    return B2C_OK;
}

int32_t b2c_proxy_set_material(b2c_ctx* ctx, int32_t uid, float friction, float restitution) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (uid < 1 || uid > ctx->nBodies) return B2C_ERR_BAD_HANDLE;
    cudaSetDevice(ctx->device);
    float2 m = make_float2(friction, restitution);
    CK(cudaMemcpyAsync(ctx->B.material + (uid - 1), &m, sizeof(m), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return B2C_OK;
}

// ---- per-step inputs ---------------------------------------------------------------------------------
int32_t b2c_set_transforms(b2c_ctx* ctx, int32_t n, const int32_t* uids, const float* planes) {
    if (!ctx || !planes || n < 0) return B2C_ERR_BAD_ARG;
    if (n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    if (!uids) {
        // 12 planes of n floats -> staging with plane stride max_bodies; repacked into float4 rows by k_aabb
        CK(cudaMemcpy2DAsync(ctx->dStaging, (size_t)ctx->cfg.max_bodies * sizeof(float), planes, (size_t)n * sizeof(float),
                             (size_t)n * sizeof(float), 12, cudaMemcpyHostToDevice, s));
        ctx->stagingCount = n;
        return B2C_OK;
    }
    for (int i = 0; i < n; i++)
        if (uids[i] < 1 || uids[i] > ctx->nBodies) return B2C_ERR_BAD_HANDLE;
    if (ctx->stagingCount > 0) {  // an unflushed full upload precedes this partial one: repack it first
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    int* dU = nullptr;
    float* dP = nullptr;
    CK(cudaMalloc((void**)&dU, (size_t)n * sizeof(int)));
    CK(cudaMalloc((void**)&dP, (size_t)n * 12 * sizeof(float)));
    CK(cudaMemcpyAsync(dU, uids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(dP, planes, (size_t)n * 12 * sizeof(float), cudaMemcpyHostToDevice, s));
    k_scatter_xf<<<(n + 255) / 256, 256, 0, s>>>(ctx->B, n, dU, dP);
    CK(cudaStreamSynchronize(s));
    cudaFree(dU);
    cudaFree(dP);
    return B2C_OK;
}

int32_t b2c_set_activation(b2c_ctx* ctx, int32_t n, const int32_t* uids, const uint8_t* active) {
    if (!ctx || !active || n < 0 || n > ctx->cfg.max_bodies) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    // flags live on the host too (alive/static are host decisions); flush any device-side change first
    if (ctx->aabbPending || ctx->extPending || ctx->stagingCount) {
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    std::vector<uint8_t> dev((size_t)ctx->nBodies);
    if (ctx->nBodies) {
        CK(cudaMemcpyAsync(dev.data(), ctx->B.flags, (size_t)ctx->nBodies, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (int k = 0; k < n; k++) {
        int uid = uids ? uids[k] : k + 1;
        if (uid < 1 || uid > ctx->nBodies) return B2C_ERR_BAD_HANDLE;
        uint8_t f = dev[uid - 1];
        if (!(f & BF_ALIVE)) continue;
        f = active[k] ? (uint8_t)(f | BF_ACTIVE) : (uint8_t)(f & ~BF_ACTIVE);
        dev[uid - 1] = f;
    }
    if (ctx->nBodies) {
        CK(cudaMemcpyAsync(ctx->B.flags, dev.data(), (size_t)ctx->nBodies, cudaMemcpyHostToDevice, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return B2C_OK;
}

int32_t b2c_set_no_collide_pairs(b2c_ctx* ctx, int32_t n, const int32_t* uidPairs) {
    if (!ctx || n < 0 || (n > 0 && !uidPairs)) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    std::vector<uint64_t> keys((size_t)n);
    for (int i = 0; i < n; i++) {
        int a = uidPairs[2 * i], b = uidPairs[2 * i + 1];
        if (a < 1 || b < 1 || a > ctx->cfg.max_bodies || b > ctx->cfg.max_bodies || a == b) return B2C_ERR_BAD_HANDLE;
        uint32_t lo = (uint32_t)(a < b ? a : b), hi = (uint32_t)(a < b ? b : a);
        keys[i] = ((uint64_t)lo << ctx->uidBits) | hi;
    }
    std::sort(keys.begin(), keys.end());
    keys.erase(std::unique(keys.begin(), keys.end()), keys.end());
    CK(cudaStreamSynchronize(ctx->stream));  // a step in flight may still be reading the old list
    if (keys.size() > ctx->capNoCollide) {
        cudaFree(ctx->dNoCollide);
        ctx->dNoCollide = nullptr;
        ctx->capNoCollide = 0;
        CK(cudaMalloc((void**)&ctx->dNoCollide, keys.size() * sizeof(uint64_t)));
        ctx->capNoCollide = (uint32_t)keys.size();
    }
    if (!keys.empty()) CK(cudaMemcpy(ctx->dNoCollide, keys.data(), keys.size() * sizeof(uint64_t), cudaMemcpyHostToDevice));
    ctx->numNoCollide = (uint32_t)keys.size();
    return B2C_OK;
}

int32_t b2c_set_aabbs(b2c_ctx* ctx, int32_t n, const int32_t* uids, const float* mm) {
    if (!ctx || !mm || n < 0 || n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t N = (size_t)ctx->cfg.max_bodies;
    if (!uids) {
        CK(cudaMemcpy2DAsync(ctx->dExtAabb, N * sizeof(float), mm, (size_t)n * sizeof(float), (size_t)n * sizeof(float), 6,
                             cudaMemcpyHostToDevice, s));
        CK(cudaMemsetAsync(ctx->dExtMask, 1, (size_t)n, s));
    } else {
        std::vector<float> tmp(6 * N, 0.f);
        std::vector<uint8_t> mask(N, 0);
        if (ctx->extPending) {
            CK(cudaMemcpyAsync(tmp.data(), ctx->dExtAabb, 6 * N * sizeof(float), cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(mask.data(), ctx->dExtMask, N, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
        }
        for (int k = 0; k < n; k++) {
            int uid = uids[k];
            if (uid < 1 || uid > ctx->nBodies) return B2C_ERR_BAD_HANDLE;
            for (int c = 0; c < 6; c++) tmp[c * N + (uid - 1)] = mm[(size_t)c * n + k];
            mask[uid - 1] = 1;
        }
        CK(cudaMemcpyAsync(ctx->dExtAabb, tmp.data(), 6 * N * sizeof(float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(ctx->dExtMask, mask.data(), N, cudaMemcpyHostToDevice, s));
        CK(cudaStreamSynchronize(s));
    }
    ctx->extPending = true;
    return B2C_OK;
}

// ---- the path ----------------------------------------------------------------------------------------
int32_t b2c_update_aabbs(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    ctx->aabbPending = true;  // executed by the k_aabb launch that opens calculate_overlapping_pairs (or by a getter)
    return B2C_OK;
}

int32_t b2c_calculate_overlapping_pairs(b2c_ctx* ctx, int32_t* numPairs) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    ctx->launches = 0;
    int32_t rc = enqueuePhases(ctx, 1);
    if (rc) return rc;
    rc = readCounters(ctx);
    if (numPairs) *numPairs = ctx->lastPairs;
    ctx->stats.kernel_launches = ctx->launches;
    return rc;
}

int32_t b2c_get_pairs(b2c_ctx* ctx, int32_t* out, int32_t cap, int32_t* numOut) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    // The pair list is final once the broadphase has run: copy it on a second stream that only waits for that point,
    // so after b2c_step_device the download overlaps the narrowphase still running on the ctx stream.
    cudaStream_t cs = ctx->streamCopy;
    CK(cudaStreamWaitEvent(cs, ctx->evPairsReady, 0));
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, ctx->dNumPairs[ctx->cur], sizeof(n), cudaMemcpyDeviceToHost, cs));
    CK(cudaStreamSynchronize(cs));
    if (numOut) *numOut = (int32_t)n;
    if (!out) return B2C_OK;
    if ((uint32_t)cap < n) { ctx->err = "pair output buffer too small"; return B2C_ERR_CAPACITY; }
    if (n) {
        CK(cudaMemcpyAsync(out, ctx->dPairs, (size_t)n * sizeof(int2), cudaMemcpyDeviceToHost, cs));
        CK(cudaStreamSynchronize(cs));
    }
    return B2C_OK;
}

int32_t b2c_dispatch_all_pairs(b2c_ctx* ctx, int32_t* numManifolds, int32_t* numContacts) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    int before = ctx->launches;
    int32_t rc = enqueuePhases(ctx, 2);
    if (rc) return rc;
    rc = readCounters(ctx);
    if (numManifolds) *numManifolds = ctx->lastManifolds;
    if (numContacts) *numContacts = ctx->lastContacts;
    ctx->stats.kernel_launches = ctx->launches;
    (void)before;
    return rc;
}

int32_t b2c_set_transforms_device(b2c_ctx* ctx, int32_t n, const float* dPlanes) {
    if (!ctx || !dPlanes || n < 0 || n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    cudaSetDevice(ctx->device);
    CK(cudaMemcpy2DAsync(ctx->dStaging, (size_t)ctx->cfg.max_bodies * sizeof(float), dPlanes, (size_t)n * sizeof(float),
                         (size_t)n * sizeof(float), 12, cudaMemcpyDeviceToDevice, ctx->stream));
    ctx->stagingCount = n;
    return B2C_OK;
}

int32_t b2c_transforms_written(b2c_ctx* ctx, int32_t n) {
    if (!ctx || n < 0 || n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    ctx->stagingCount = n;
    return B2C_OK;
}

int32_t b2c_step_device(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    ctx->launches = 0;
    ctx->aabbPending = true;
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], s));
    int32_t rc = enqueuePhases(ctx, 0);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[3], s));
    ctx->stats.kernel_launches = ctx->launches;
    return B2C_OK;
}

int32_t b2c_sync_counts(b2c_ctx* ctx, int32_t* numPairs, int32_t* numManifolds, int32_t* numContacts) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    int32_t rc = readCounters(ctx);
    if (numPairs) *numPairs = ctx->lastPairs;
    if (numManifolds) *numManifolds = ctx->lastManifolds;
    if (numContacts) *numContacts = ctx->lastContacts;
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[2]) == cudaSuccess) ctx->stats.ms_broadphase = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]) == cudaSuccess) ctx->stats.ms_narrowphase = ms;
    if (cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]) == cudaSuccess) ctx->stats.ms_total = ms;
    cudaGetLastError();
    return rc;
}

int32_t b2c_step(b2c_ctx* ctx, int32_t n, const float* planes, int32_t* numPairs, int32_t* numManifolds, int32_t* numContacts) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (planes) {
        int32_t rc = b2c_set_transforms(ctx, n, nullptr, planes);
        if (rc) return rc;
    }
    int32_t rc = b2c_step_device(ctx);
    if (rc) return rc;
    return b2c_sync_counts(ctx, numPairs, numManifolds, numContacts);
}

// ---- results -----------------------------------------------------------------------------------------
// number of child algorithms of a compound pair: children(body0) x children(body1), a non-compound side counting as one
static uint32_t hostCompoundCount(const b2c_ctx* ctx, int uid0, int uid1) {
    const ShapeDev& S0 = ctx->hShapes[ctx->hShapeOf[uid0 - 1]];
    const ShapeDev& S1 = ctx->hShapes[ctx->hShapeOf[uid1 - 1]];
    const uint32_t n0 = S0.type == SH_COMPOUND ? (uint32_t)S0.numPoints : 1u;
    const uint32_t n1 = S1.type == SH_COMPOUND ? (uint32_t)S1.numPoints : 1u;
    return n0 * n1;
}
static int32_t compoundNumItems(b2c_ctx* ctx, uint32_t* nOut) {
    *nOut = 0;
    if (!ctx->hasCompound) return B2C_OK;
    CompoundCounters cc;
    CK(cudaMemcpyAsync(&cc, ctx->dCompoundCtr, sizeof(cc), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    *nOut = cc.numItems <= ctx->maxCompoundItems ? cc.numItems : 0;
    return B2C_OK;
}

int32_t b2c_get_manifolds(b2c_ctx* ctx, b2c_manifold* out, int32_t cap, int32_t onlyTouching, int32_t* numOut) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, ctx->dNumPairs[ctx->cur], sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<ManifoldHdr> hdr(n);
    std::vector<b2c_manifold_point> pts(4 * (size_t)n);
    if (n) {
        CK(cudaMemcpyAsync(hdr.data(), ctx->dMHdr[ctx->cur], (size_t)n * sizeof(ManifoldHdr), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(pts.data(), ctx->dMPts[ctx->cur], 4 * (size_t)n * sizeof(b2c_manifold_point), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    // child manifolds of the compound pairs (compound.cuh): the latest arrays, indexed through the pair's header
    uint32_t nItems = 0;
    int32_t rcc = compoundNumItems(ctx, &nItems);
    if (rcc) return rcc;
    std::vector<ManifoldHdr> chdr(nItems);
    std::vector<b2c_manifold_point> cpts(4 * (size_t)nItems);
    if (nItems) {
        CK(cudaMemcpyAsync(chdr.data(), ctx->dCH[ctx->ccur], (size_t)nItems * sizeof(ManifoldHdr), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(cpts.data(), ctx->dCP[ctx->ccur], 4 * (size_t)nItems * sizeof(b2c_manifold_point), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    int32_t k = 0;
    auto emit = [&](const ManifoldHdr& h, const b2c_manifold_point* pp, int child0, int child1) {
        if (onlyTouching && h.num_contacts == 0) return;
        if (out && k < cap) {
            b2c_manifold& m = out[k];
            memset(&m, 0, sizeof(m));
            m.pair_uid0 = h.pair_uid0; m.pair_uid1 = h.pair_uid1; m.body0 = h.body0; m.body1 = h.body1;
            m.num_contacts = h.num_contacts; m.algorithm = h.algorithm;
            m.child0 = child0; m.child1 = child1;
            for (int q = 0; q < h.num_contacts && q < 4; q++) m.points[q] = pp[q];
        }
        k++;
    };
    for (uint32_t p = 0; p < n; p++) {
        const ManifoldHdr& h = hdr[p];
        if (h.algorithm == 0) continue;
        if (h.algorithm == 5) {  // compound pair: one manifold per child algorithm, in the order the reference runs them
            const uint32_t start = (uint32_t)h.pad1, count = hostCompoundCount(ctx, h.pair_uid0, h.pair_uid1);
            for (uint32_t q = start; q < start + count && q < nItems; q++)
                if (chdr[q].algorithm != 0) emit(chdr[q], &cpts[4 * (size_t)q], chdr[q].pad0, chdr[q].pad1);
            continue;
        }
        emit(h, &pts[4 * (size_t)p], -1, -1);
    }
    if (numOut) *numOut = k;
    if (out && k > cap) { ctx->err = "manifold output buffer too small"; return B2C_ERR_CAPACITY; }
    return B2C_OK;
}

int32_t b2c_set_raw_records(b2c_ctx* ctx, int32_t on) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    ctx->wantRaw = on != 0;
    return B2C_OK;
}

int32_t b2c_get_raw_contacts(b2c_ctx* ctx, b2c_raw_contact* out, int32_t cap, int32_t* numOut) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    if (!ctx->wantRaw) {
        ctx->err = "raw detector records are an inspection channel and off by default: call b2c_set_raw_records(ctx, 1) before the dispatch";
        return B2C_ERR_STATE;
    }
    cudaSetDevice(ctx->device);
    uint32_t n = 0;
    CK(cudaMemcpyAsync(&n, ctx->dNumPairs[ctx->cur], sizeof(n), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    std::vector<b2c_raw_contact> raw(n);
    std::vector<uint32_t> ms(n), mc(n);
    std::vector<uint8_t> binOf(n);
    uint32_t nCItems = 0;
    int32_t rcc = compoundNumItems(ctx, &nCItems);
    if (rcc) return rcc;
    std::vector<ManifoldHdr> hdr(ctx->hasCompound ? n : 0);
    std::vector<b2c_raw_contact> craw(nCItems);
    std::vector<uint32_t> cms(nCItems), cmc(nCItems);
    if (ctx->hasCompound && n) {
        CK(cudaMemcpyAsync(hdr.data(), ctx->dMHdr[ctx->cur], (size_t)n * sizeof(ManifoldHdr), cudaMemcpyDeviceToHost, ctx->stream));
        if (nCItems) CK(cudaMemcpyAsync(craw.data(), ctx->dCRaw, (size_t)nCItems * sizeof(b2c_raw_contact), cudaMemcpyDeviceToHost, ctx->stream));
        if (nCItems && ctx->hasMesh) {
            CK(cudaMemcpyAsync(cms.data(), ctx->dCMeshStart, (size_t)nCItems * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(cmc.data(), ctx->dCMeshCount, (size_t)nCItems * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    if (n) {
        CK(cudaMemcpyAsync(raw.data(), ctx->dRaw, (size_t)n * sizeof(b2c_raw_contact), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaMemcpyAsync(binOf.data(), ctx->dBinOf, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        if (ctx->hasMesh) {
            CK(cudaMemcpyAsync(ms.data(), ctx->dMeshStart, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaMemcpyAsync(mc.data(), ctx->dMeshCount, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        CK(cudaStreamSynchronize(ctx->stream));
    }
    uint32_t nItems = 0;
    std::vector<b2c_raw_contact> rawMesh;
    if (ctx->hasMesh) {
        // read the count of THIS dispatch from the device (the pinned mirror is only as fresh as the last readCounters)
        CK(cudaMemcpyAsync(&nItems, &ctx->dCtr->meshItems, sizeof(nItems), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        if (nItems > (uint32_t)ctx->cfg.max_mesh_items) nItems = 0;
        rawMesh.resize(nItems);
        if (nItems) {
            CK(cudaMemcpyAsync(rawMesh.data(), ctx->dRawMesh, (size_t)nItems * sizeof(b2c_raw_contact), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
        }
    }
    int32_t k = 0;
    for (uint32_t p = 0; p < n; p++) {
        if (binOf[p] == BIN_SKIP || binOf[p] == BIN_COMPOUND_KEEP) continue;  // not dispatched (both bodies inactive, or no algorithm for the type pair)
        if (binOf[p] == BIN_COMPOUND) {  // one record per child algorithm (tri = -2 - k)
            if (hdr[p].algorithm != 5) continue;
            const uint32_t start = (uint32_t)hdr[p].pad1, count = hostCompoundCount(ctx, hdr[p].pair_uid0, hdr[p].pair_uid1);
            for (uint32_t q = start; q < start + count && q < nCItems; q++) {
                if (craw[q].has_contact == -3) {  // child x mesh: one record per (child, triangle) in the mesh item array
                    for (uint32_t t = cms[q]; t < cms[q] + cmc[q] && t < nItems; t++) {
                        if (out && k < cap) out[k] = rawMesh[t];
                        k++;
                    }
                    continue;
                }
                if (out && k < cap) out[k] = craw[q];
                k++;
            }
            continue;
        }
        if (raw[p].has_contact == -3) {
            for (uint32_t q = ms[p]; q < ms[p] + mc[p] && q < nItems; q++) {
                if (out && k < cap) out[k] = rawMesh[q];
                k++;
            }
            continue;
        }
        if (out && k < cap) out[k] = raw[p];
        k++;
    }
    if (numOut) *numOut = k;
    if (out && k > cap) { ctx->err = "raw contact output buffer too small"; return B2C_ERR_CAPACITY; }
    return B2C_OK;
}

int32_t b2c_get_aabbs(b2c_ctx* ctx, float* out, int32_t n) {
    if (!ctx || !out || n < 0 || n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    cudaSetDevice(ctx->device);
    if (ctx->aabbPending || ctx->extPending || ctx->stagingCount) {
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    float* d = nullptr;
    CK(cudaMalloc((void**)&d, (size_t)n * 6 * sizeof(float)));
    k_get_aabbs<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->B, n, d);
    CK(cudaMemcpyAsync(out, d, (size_t)n * 6 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d);
    return B2C_OK;
}

int32_t b2c_get_broadphase_aabb(b2c_ctx* ctx, float mn[3], float mx[3]) {
    if (!ctx || !mn || !mx) return B2C_ERR_BAD_ARG;
    // bp/SimpleBroadphase.java:118-121 and bp/DbvtBroadphase.java getBroadphaseAabb: unbounded;
    // bp/AxisSweep3Internal.java:622-627: the world box
    for (int c = 0; c < 3; c++) {
        mn[c] = ctx->sap.enabled ? ctx->sap.wmin[c] : -1e30f;
        mx[c] = ctx->sap.enabled ? ctx->sap.wmax[c] : 1e30f;
    }
    return B2C_OK;
}

int32_t b2c_get_stats(b2c_ctx* ctx, b2c_stats* out) {
    if (!ctx || !out) return B2C_ERR_BAD_ARG;
    *out = ctx->stats;
    return B2C_OK;
}

static int32_t getContactsImpl(b2c_ctx* ctx, int mode, void* hOut, int32_t capH, void* pOut, int32_t capP, int32_t* nH,
                               int32_t* nP) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t ptSize = mode >= 2 ? sizeof(b2c_packed_point) : (mode == 1 ? sizeof(b2c_solver_point) : sizeof(b2c_manifold_point));
    const size_t hdSize = mode >= 2 ? sizeof(b2c_packed_header) : sizeof(b2c_contact_header);
    uint32_t counts[2] = {0, 0};
    if (ctx->contactReady == mode && ctx->contactCountsValid) {
        // compacted behind the dispatch (b2c_set_contact_prefetch) and the counts came back with the step counters
        // (b2c_sync_counts / b2c_dispatch_all_pairs): nothing to launch, nothing to wait for
        counts[0] = ctx->contactCounts[0];
        counts[1] = ctx->contactCounts[1];
    } else if (ctx->contactReady == mode) {
        CK(cudaMemcpyAsync(counts, ctx->dContactCounts, sizeof(counts), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ctx->contactCounts[0] = counts[0]; ctx->contactCounts[1] = counts[1];
        ctx->contactCountsValid = true;
    } else {
        int32_t rc = enqueueContactCompaction(ctx, mode);
        if (rc) return rc;
        ctx->contactReady = mode;
        CK(cudaMemcpyAsync(counts, ctx->dContactCounts, sizeof(counts), cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        ctx->contactCounts[0] = counts[0]; ctx->contactCounts[1] = counts[1];
        ctx->contactCountsValid = true;
    }
    if (nH) *nH = (int32_t)counts[0];
    if (nP) *nP = (int32_t)counts[1];
    if (counts[0] > ctx->capContactHdr || counts[1] > ctx->capContactPts) { ctx->err = "contact stream capacity exceeded"; return B2C_ERR_CAPACITY; }
    if (!hOut && !pOut) return B2C_OK;
    if ((uint32_t)capH < counts[0] || (uint32_t)capP < counts[1]) { ctx->err = "contact output buffers too small"; return B2C_ERR_CAPACITY; }
    uint32_t h0 = 0, p0 = 0;
    if (mode >= 2 && ctx->earlyDl.active && ctx->earlyDl.hdr == hOut && ctx->earlyDl.pts == pOut && ctx->earlyDl.nH <= counts[0] &&
        ctx->earlyDl.nP <= counts[1]) {
        h0 = ctx->earlyDl.nH;  // the first part is already on its way (b2c_begin_contact_download): copy only what the
        p0 = ctx->earlyDl.nP;  // end of the dispatch appended behind it
    }
    if (hOut && counts[0] > h0)
        CK(cudaMemcpyAsync((char*)hOut + (size_t)h0 * hdSize, (const char*)ctx->dContactHdr + (size_t)h0 * hdSize,
                           (size_t)(counts[0] - h0) * hdSize, cudaMemcpyDeviceToHost, s));
    if (pOut && counts[1] > p0)
        CK(cudaMemcpyAsync((char*)pOut + (size_t)p0 * ptSize, (const char*)ctx->dContactPts + (size_t)p0 * ptSize,
                           (size_t)(counts[1] - p0) * ptSize, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (h0 || p0) CK(cudaStreamSynchronize(ctx->streamCopy));
    ctx->earlyDl.active = false;
    return B2C_OK;
}

int32_t b2c_get_contacts(b2c_ctx* ctx, b2c_contact_header* hOut, int32_t capH, b2c_manifold_point* pOut, int32_t capP,
                         int32_t* nH, int32_t* nP) {
    return getContactsImpl(ctx, 0, hOut, capH, pOut, capP, nH, nP);
}

int32_t b2c_get_solver_contacts(b2c_ctx* ctx, b2c_contact_header* hOut, int32_t capH, b2c_solver_point* pOut, int32_t capP,
                                int32_t* nH, int32_t* nP) {
    return getContactsImpl(ctx, 1, hOut, capH, pOut, capP, nH, nP);
}

int32_t b2c_get_packed_contacts(b2c_ctx* ctx, b2c_packed_header* hOut, int32_t capH, b2c_packed_point* pOut, int32_t capP,
                                int32_t* nH, int32_t* nP) {
    return getContactsImpl(ctx, 2, hOut, capH, pOut, capP, nH, nP);
}

int32_t b2c_get_packed_contacts_uid(b2c_ctx* ctx, b2c_packed_uid_header* hOut, int32_t capH, b2c_packed_point* pOut, int32_t capP,
                                    int32_t* nH, int32_t* nP) {
    return getContactsImpl(ctx, 3, hOut, capH, pOut, capP, nH, nP);
}

int32_t b2c_begin_contact_download(b2c_ctx* ctx, void* hOut, int32_t capH, b2c_packed_point* pOut, int32_t capP) {
    if (!ctx || !hOut || !pOut || capH < 0 || capP < 0) return B2C_ERR_BAD_ARG;
    if (ctx->contactPrefetch < 2 || !ctx->earlyRecorded) {
        ctx->err = "b2c_begin_contact_download needs b2c_set_contact_prefetch(ctx, 2 or 3) and an enqueued dispatch";
        return B2C_ERR_STATE;
    }
    cudaSetDevice(ctx->device);
    cudaStream_t cs = ctx->streamCopy;
    CK(cudaStreamWaitEvent(cs, ctx->evContactsEarly, 0));
    CK(cudaMemcpyAsync(ctx->hEarlyCountsPinned, ctx->dContactCounts + 2, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
    CK(cudaStreamSynchronize(cs));  // returns when k_manifold_cc has finished; the penetration bin may still be running
    const uint32_t nH = ctx->hEarlyCountsPinned[0], nP = ctx->hEarlyCountsPinned[1];
    if (nH > ctx->capContactHdr || nP > ctx->capContactPts) { ctx->err = "contact stream capacity exceeded"; return B2C_ERR_CAPACITY; }
    if ((uint32_t)capH < nH || (uint32_t)capP < nP) { ctx->err = "contact output buffers too small"; return B2C_ERR_CAPACITY; }
    if (nH) CK(cudaMemcpyAsync(hOut, ctx->dContactHdr, (size_t)nH * sizeof(b2c_packed_header), cudaMemcpyDeviceToHost, cs));
    if (nP) CK(cudaMemcpyAsync(pOut, ctx->dContactPts, (size_t)nP * sizeof(b2c_packed_point), cudaMemcpyDeviceToHost, cs));
    ctx->earlyDl.active = true;
    ctx->earlyDl.hdr = hOut;
    ctx->earlyDl.pts = pOut;
    ctx->earlyDl.nH = nH;
    ctx->earlyDl.nP = nP;
    return B2C_OK;
}

int32_t b2c_set_contact_prefetch(b2c_ctx* ctx, int32_t format) {
    if (!ctx || format < -1 || format > 3) return B2C_ERR_BAD_ARG;
    ctx->contactPrefetch = format;
    return B2C_OK;
}

int32_t b2c_get_pair_deltas(b2c_ctx* ctx, int32_t* addedOut, int32_t capA, int32_t* removedOut, int32_t capR, int32_t* nA, int32_t* nR) {
    if (!ctx || capA < 0 || capR < 0) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t P = (size_t)ctx->cfg.max_pairs;
    if (!ctx->dDeltaCounts) CK(dalloc(&ctx->dDeltaCounts, (size_t)4));
    for (int i = 0; i < 2; i++)
        if (!ctx->dDelta[i]) CK(dalloc(&ctx->dDelta[i], P));
    if (ctx->deltaReady) {
        // computed inside the pair calculation: copy on the side stream, which only waits for the broadphase
        cudaStream_t cs = ctx->streamCopy;
        CK(cudaStreamWaitEvent(cs, ctx->evPairsReady, 0));
        CK(cudaMemcpyAsync(ctx->hDeltaCountsPinned, ctx->dDeltaCounts, 2 * sizeof(uint32_t), cudaMemcpyDeviceToHost, cs));
        CK(cudaStreamSynchronize(cs));
        const uint32_t c0 = ctx->hDeltaCountsPinned[0], c1 = ctx->hDeltaCountsPinned[1];
        if (nA) *nA = (int32_t)c0;
        if (nR) *nR = (int32_t)c1;
        if ((addedOut && (uint32_t)capA < c0) || (removedOut && (uint32_t)capR < c1)) {
            ctx->err = "pair delta output buffer too small";
            return B2C_ERR_CAPACITY;
        }
        if (addedOut && c0) CK(cudaMemcpyAsync(addedOut, ctx->dDelta[0], (size_t)c0 * sizeof(int2), cudaMemcpyDeviceToHost, cs));
        if (removedOut && c1) CK(cudaMemcpyAsync(removedOut, ctx->dDelta[1], (size_t)c1 * sizeof(int2), cudaMemcpyDeviceToHost, cs));
        CK(cudaStreamSynchronize(cs));
        return B2C_OK;
    }
    const int cur = ctx->cur, prev = cur ^ 1;
    CK(cudaMemsetAsync(ctx->dDeltaCounts, 0, 2 * sizeof(uint32_t), s));
    const unsigned g = gridFor((uint32_t)P, 256);
    k_pair_delta<<<g, 256, 0, s>>>(ctx->dSortedKeys[cur], ctx->dNumPairs[cur], ctx->dSortedKeys[prev], ctx->dNumPairs[prev],
                                   ctx->dPairFirst[prev], ctx->uidBits, ctx->dDelta[0], (uint32_t)P, ctx->dDeltaCounts);
    k_pair_delta<<<g, 256, 0, s>>>(ctx->dSortedKeys[prev], ctx->dNumPairs[prev], ctx->dSortedKeys[cur], ctx->dNumPairs[cur],
                                   ctx->dPairFirst[cur], ctx->uidBits, ctx->dDelta[1], (uint32_t)P, ctx->dDeltaCounts + 1);
    uint32_t c[2] = {0, 0};
    CK(cudaMemcpyAsync(c, ctx->dDeltaCounts, sizeof(c), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (nA) *nA = (int32_t)c[0];
    if (nR) *nR = (int32_t)c[1];
    if ((addedOut && (uint32_t)capA < c[0]) || (removedOut && (uint32_t)capR < c[1])) {
        ctx->err = "pair delta output buffer too small";
        return B2C_ERR_CAPACITY;
    }
    if (addedOut && c[0]) CK(cudaMemcpyAsync(addedOut, ctx->dDelta[0], (size_t)c[0] * sizeof(int2), cudaMemcpyDeviceToHost, s));
    if (removedOut && c[1]) CK(cudaMemcpyAsync(removedOut, ctx->dDelta[1], (size_t)c[1] * sizeof(int2), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2C_OK;
}

int32_t b2c_set_pair_delta_prefetch(b2c_ctx* ctx, int32_t on) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    if (on) {
        const size_t P = (size_t)ctx->cfg.max_pairs;
        if (!ctx->dDeltaCounts) CK(dalloc(&ctx->dDeltaCounts, (size_t)4));
        for (int i = 0; i < 2; i++)
            if (!ctx->dDelta[i]) CK(dalloc(&ctx->dDelta[i], P));
        if (!ctx->hDeltaCountsPinned) CK(cudaMallocHost((void**)&ctx->hDeltaCountsPinned, 2 * sizeof(uint32_t)));
    }
    ctx->deltaPrefetch = on != 0;
    ctx->deltaReady = false;
    return B2C_OK;
}

int32_t b2c_compute_islands(b2c_ctx* ctx, int32_t* tagsOut, int32_t n, int32_t* numIslands) {
    if (!ctx || n < 0 || n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const size_t N = (size_t)ctx->cfg.max_bodies;
    if (!ctx->dDeltaCounts) CK(dalloc(&ctx->dDeltaCounts, (size_t)4));
    if (!ctx->dIslandPar) CK(dalloc(&ctx->dIslandPar, N));
    if (!ctx->dIslandTags) CK(dalloc(&ctx->dIslandTags, N));
    const int nb = ctx->nBodies;
    if (nb == 0) { if (numIslands) *numIslands = 0; return B2C_OK; }
    CK(cudaMemsetAsync(ctx->dDeltaCounts + 2, 0, sizeof(uint32_t), s));
    k_island_init<<<(nb + 255) / 256, 256, 0, s>>>(ctx->dIslandPar, nb);
    k_island_unite<<<gridFor((uint32_t)ctx->cfg.max_pairs, 256), 256, 0, s>>>(ctx->dPairs, ctx->dNumPairs[ctx->cur], ctx->B.flags,
                                                                            ctx->dIslandPar);
    k_island_flatten<<<(nb + 255) / 256, 256, 0, s>>>(ctx->dIslandPar, ctx->B.flags, nb, ctx->dIslandTags, ctx->dDeltaCounts + 2);
    uint32_t c = 0;
    CK(cudaMemcpyAsync(&c, ctx->dDeltaCounts + 2, sizeof(c), cudaMemcpyDeviceToHost, s));
    if (tagsOut && n) CK(cudaMemcpyAsync(tagsOut, ctx->dIslandTags, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    if (numIslands) *numIslands = (int32_t)c;
    return B2C_OK;
}

int32_t b2c_ray_test_closest(b2c_ctx* ctx, int32_t n, const float* from, const float* to, int16_t group, int16_t mask, int32_t* uidOut,
                             float* fracOut, float* nrmOut, float* ptOut) {
    if (!ctx || n < 0 || (n > 0 && (!from || !to))) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    if (ctx->stagingCount || ctx->extPending) {  // transforms uploaded but not yet repacked: flush them (no AABB update)
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    int32_t rc = uploadShapes(ctx);
    if (rc) return rc;
    const size_t N = (size_t)ctx->cfg.max_bodies;
    if (!ctx->dRayMin) CK(dalloc(&ctx->dRayMin, N));
    if (!ctx->dRayMax) CK(dalloc(&ctx->dRayMax, N));
    if (!ctx->dRayOverflow) CK(dalloc(&ctx->dRayOverflow, (size_t)1));
    if (!ctx->dRayChunkMin) CK(dalloc(&ctx->dRayChunkMin, N / RAY_CHUNK + 2));
    if (!ctx->dRayChunkMax) CK(dalloc(&ctx->dRayChunkMax, N / RAY_CHUNK + 2));
    if (n > ctx->rayCap) {
        cudaFree(ctx->dRayIn); cudaFree(ctx->dRayOut);
        ctx->dRayIn = nullptr; ctx->dRayOut = nullptr; ctx->rayCap = 0;
        CK(cudaMalloc((void**)&ctx->dRayIn, (size_t)n * 6 * sizeof(float)));
        CK(cudaMalloc((void**)&ctx->dRayOut, (size_t)n * sizeof(RayOut)));
        ctx->rayCap = n;
    }
    const int nb = ctx->nBodies;
    CK(cudaMemcpyAsync(ctx->dRayIn, from, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->dRayIn + 3 * (size_t)n, to, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->dRayOverflow, 0, sizeof(uint32_t), s));
    const int nSorted = ctx->nSortedBodies < nb ? ctx->nSortedBodies : nb;
    if (nb > 0) {
        k_ray_aabbs<<<(nb + 255) / 256, 256, 0, s>>>(ctx->B, ctx->dShapes, nb, ctx->dRayMin, ctx->dRayMax);
        const int nChunks = (nb + RAY_CHUNK - 1) / RAY_CHUNK;
        k_ray_chunks<<<(nChunks + 127) / 128, 128, 0, s>>>(ctx->dRayMin, ctx->dRayMax, ctx->dSmin, nSorted, nb, ctx->dRayChunkMin,
                                                          ctx->dRayChunkMax);
    }
    const uint32_t cbFilter = ((uint32_t)(uint16_t)group) | ((uint32_t)(uint16_t)mask << 16);
    const unsigned grid = (unsigned)(n < 148 * 16 ? n : 148 * 16);
    k_ray_test<<<grid, RAY_THREADS, 0, s>>>(ctx->B, ctx->dShapes, ctx->dHullPts, ctx->dMeshes, ctx->dChildren, ctx->dSmin, nSorted, ctx->dRayChunkMin, ctx->dRayChunkMax, nb, ctx->dRayMin, ctx->dRayMax, ctx->dRayIn,
                                            ctx->dRayIn + 3 * (size_t)n, n, cbFilter, ctx->dRayOut, ctx->dRayOverflow);
    std::vector<RayOut> host((size_t)n);
    uint32_t ov = 0;
    CK(cudaMemcpyAsync(host.data(), ctx->dRayOut, (size_t)n * sizeof(RayOut), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(&ov, ctx->dRayOverflow, sizeof(ov), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    (void)ov;  // informational: the most boxes a ray met when it had to take the index-ordered tile path (raycast.cuh)
    for (int i = 0; i < n; i++) {
        if (uidOut) uidOut[i] = host[i].uid;
        if (fracOut) fracOut[i] = host[i].fraction;
        if (nrmOut) { nrmOut[3 * i] = host[i].normal[0]; nrmOut[3 * i + 1] = host[i].normal[1]; nrmOut[3 * i + 2] = host[i].normal[2]; }
        if (ptOut) { ptOut[3 * i] = host[i].point[0]; ptOut[3 * i + 1] = host[i].point[1]; ptOut[3 * i + 2] = host[i].point[2]; }
    }
    return B2C_OK;
}

// Shared body of the two sweep entry points.  me / radius non-null: the CCD motion-clamping sweeps (ClosestNotMe callback).
static int32_t runConvexSweeps(b2c_ctx* ctx, int32_t n, const int32_t* castShapes, const float* basis9, const float* from, const float* to,
                               uint32_t cbFilter, float allowedPenetration, const int32_t* meBodies, const float* radius, int32_t* uidOut,
                               float* fracOut, float* nrmOut, float* ptOut) {
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    if (ctx->stagingCount || ctx->extPending) {
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    int32_t rc = uploadShapes(ctx);
    if (rc) return rc;
    const size_t N = (size_t)ctx->cfg.max_bodies;
    if (!ctx->dRayMin) CK(dalloc(&ctx->dRayMin, N));
    if (!ctx->dRayMax) CK(dalloc(&ctx->dRayMax, N));
    if (!ctx->dRayOverflow) CK(dalloc(&ctx->dRayOverflow, (size_t)1));
    if (!ctx->dRayChunkMin) CK(dalloc(&ctx->dRayChunkMin, N / RAY_CHUNK + 2));
    if (!ctx->dRayChunkMax) CK(dalloc(&ctx->dRayChunkMax, N / RAY_CHUNK + 2));
    if (n > ctx->sweepCap) {
        cudaFree(ctx->dSweepIn); cudaFree(ctx->dSweepOut);
        ctx->dSweepIn = nullptr; ctx->dSweepOut = nullptr; ctx->sweepCap = 0;
        CK(cudaMalloc((void**)&ctx->dSweepIn, (size_t)n * 16 * sizeof(float)));  // basis 9 | from 3 | to 3 | cast shape id (or: me, radius)
        CK(cudaMalloc((void**)&ctx->dSweepOut, (size_t)n * sizeof(RayOut)));
        ctx->sweepCap = n;
    }
    const int nb = ctx->nBodies;
    float* dBasis = ctx->dSweepIn;
    float* dFrom = dBasis + 9 * (size_t)n;
    float* dTo = dFrom + 3 * (size_t)n;
    int* dShape = reinterpret_cast<int*>(dTo + 3 * (size_t)n);
    SweepNotMe nm{};
    CK(cudaMemcpyAsync(dTo, to, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
    if (meBodies) {
        // the basis area holds the radii, the shape-id area the 0-based body indices
        CK(cudaMemcpyAsync(dBasis, radius, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dShape, meBodies, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
        nm.me = dShape;
        nm.radius = dBasis;
        if (ctx->pairsValid) {
            const int cur = ctx->cur;
            nm.keys = ctx->dSortedKeys[cur];
            nm.numPairs = ctx->dNumPairs[cur];
            nm.first = ctx->dPairFirst[cur];
            nm.uidBits = ctx->uidBits;
            nm.mhdr = ctx->dMHdr[cur];
            nm.compH = ctx->hasCompound ? ctx->dCH[ctx->ccur] : nullptr;
        }
    } else {
        CK(cudaMemcpyAsync(dBasis, basis9, (size_t)n * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dFrom, from, (size_t)n * 3 * sizeof(float), cudaMemcpyHostToDevice, s));
        CK(cudaMemcpyAsync(dShape, castShapes, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, s));
    }
    CK(cudaMemsetAsync(ctx->dRayOverflow, 0, sizeof(uint32_t), s));
    const int nSorted = ctx->nSortedBodies < nb ? ctx->nSortedBodies : nb;
    if (nb > 0) {
        k_ray_aabbs<<<(nb + 255) / 256, 256, 0, s>>>(ctx->B, ctx->dShapes, nb, ctx->dRayMin, ctx->dRayMax);
        const int nChunks = (nb + RAY_CHUNK - 1) / RAY_CHUNK;
        k_ray_chunks<<<(nChunks + 127) / 128, 128, 0, s>>>(ctx->dRayMin, ctx->dRayMax, ctx->dSmin, nSorted, nb, ctx->dRayChunkMin,
                                                          ctx->dRayChunkMax);
    }
    const unsigned grid = (unsigned)(n < 148 * 8 ? n : 148 * 8);
    k_convex_sweep<<<grid, SWEEP_THREADS, 0, s>>>(ctx->B, ctx->dShapes, ctx->dHullPts, ctx->dMeshes, ctx->dChildren, ctx->dSmin, nSorted,
                                                  ctx->dRayChunkMin, ctx->dRayChunkMax, nb, ctx->dRayMin, ctx->dRayMax, dShape, dBasis,
                                                  dFrom, dTo, n, cbFilter, allowedPenetration, ctx->dSweepOut, ctx->dRayOverflow, nm);
    std::vector<RayOut> host((size_t)n);
    CK(cudaMemcpyAsync(host.data(), ctx->dSweepOut, (size_t)n * sizeof(RayOut), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int i = 0; i < n; i++) {
        if (uidOut) uidOut[i] = host[i].uid;
        if (fracOut) fracOut[i] = host[i].fraction;
        if (nrmOut) { nrmOut[3 * i] = host[i].normal[0]; nrmOut[3 * i + 1] = host[i].normal[1]; nrmOut[3 * i + 2] = host[i].normal[2]; }
        if (ptOut) { ptOut[3 * i] = host[i].point[0]; ptOut[3 * i + 1] = host[i].point[1]; ptOut[3 * i + 2] = host[i].point[2]; }
    }
    return B2C_OK;
}

int32_t b2c_convex_sweep_closest(b2c_ctx* ctx, int32_t n, const int32_t* castShapes, const float* basis9, const float* from, const float* to,
                                 int16_t group, int16_t mask, float allowedPenetration, int32_t* uidOut, float* fracOut, float* nrmOut,
                                 float* ptOut) {
    if (!ctx || n < 0 || (n > 0 && (!castShapes || !basis9 || !from || !to))) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    for (int i = 0; i < n; i++) {
        const int sid = castShapes[i];
        if (sid < 0 || sid >= (int)ctx->hShapes.size()) { ctx->err = "convex sweep: unknown cast shape id"; return B2C_ERR_BAD_HANDLE; }
        const int ty = ctx->hShapes[(size_t)sid].type;
        if (ty != SH_BOX && ty != SH_SPHERE && ty != SH_HULL) {
            ctx->err = "convex sweep: the cast shape must be convex (box, sphere or hull), as ConvexShape castShape is";
            return B2C_ERR_BAD_ARG;
        }
    }
    const uint32_t cbFilter = ((uint32_t)(uint16_t)group) | ((uint32_t)(uint16_t)mask << 16);
    return runConvexSweeps(ctx, n, castShapes, basis9, from, to, cbFilter, allowedPenetration, nullptr, nullptr, uidOut, fracOut, nrmOut, ptOut);
}

int32_t b2c_ccd_sweep_not_me(b2c_ctx* ctx, int32_t n, const int32_t* bodyUids, const float* radius, const float* predicted, float allowedPenetration,
                             int32_t* uidOut, float* fracOut, float* nrmOut, float* ptOut) {
    if (!ctx || n < 0 || (n > 0 && (!bodyUids || !radius || !predicted))) return B2C_ERR_BAD_ARG;
    if (n == 0) return B2C_OK;
    std::vector<int32_t> me((size_t)n);
    for (int i = 0; i < n; i++) {
        const int u = bodyUids[i];
        if (u < 1 || u > ctx->nBodies || !(ctx->hFlags[(size_t)u - 1] & BF_ALIVE)) { ctx->err = "ccd sweep: unknown body uid"; return B2C_ERR_BAD_HANDLE; }
        if (!(radius[i] >= 0.f)) { ctx->err = "ccd sweep: negative swept-sphere radius"; return B2C_ERR_BAD_ARG; }
        me[(size_t)i] = u - 1;
    }
    return runConvexSweeps(ctx, n, nullptr, nullptr, nullptr, predicted, 0u, allowedPenetration, me.data(), radius, uidOut, fracOut, nrmOut, ptOut);
}

int32_t b2c_set_profiling(b2c_ctx* ctx, int32_t on) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    ctx->prof = on != 0;
    ctx->stageValid = false;
    return B2C_OK;
}

int32_t b2c_get_stage_times(b2c_ctx* ctx, float ms[B2C_NUM_STAGES]) {
    if (!ctx || !ms) return B2C_ERR_BAD_ARG;
    if (!ctx->stageValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->timeline) {
        float v[5] = {0, 0, 0, 0, 0};
        for (int k = 0; k < 5; k++) cudaEventElapsedTime(&v[k], ctx->stageEv[11], ctx->tl[k]);
        fprintf(stderr, "[b2c timeline, ms after stage 11 start] epa0 %.3f..%.3f  epa-chain end %.3f | manifold_cc %.3f..%.3f\n", v[0], v[1],
                v[2], v[3], v[4]);
        cudaGetLastError();
    }
    for (int k = 0; k < B2C_NUM_STAGES; k++) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, ctx->stageEv[k], ctx->stageEv[k + 1]) != cudaSuccess) { cudaGetLastError(); t = 0.f; }
        ms[k] = t;
    }
    return B2C_OK;
}

int32_t b2c_get_gjk_kernel_time(b2c_ctx* ctx, float* ms) {
    if (!ctx || !ms) return B2C_ERR_BAD_ARG;
    if (!ctx->stageValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    float t = 0.f;
    if (cudaEventElapsedTime(&t, ctx->evGjk[0], ctx->evGjk[1]) != cudaSuccess) { cudaGetLastError(); t = 0.f; }
    *ms = t;
    return B2C_OK;
}

const char* b2c_stage_name(int32_t k) {
    static const char* names[B2C_NUM_STAGES] = {"aabb", "bounds_keys", "sort_proxies", "gather", "sweep", "large", "sort_pairs",
                                                "unpack_carry", "classify_bin", "closed_form", "gjk_mesh", "epa_fold_count"};
    return (k >= 0 && k < B2C_NUM_STAGES) ? names[k] : "";
}

// Slab partition along `axis` with explicit planes (ascending, nranks - 1 of them).  Call it on every rank after the proxies
// exist and while all ranks hold the same transforms: the owner table is derived from the resident origins.
int32_t b2c_set_partition_slabs(b2c_ctx* ctx, int32_t rank, int32_t nranks, int32_t axis, const float* planes) {
    if (!ctx || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks || axis < 0 || axis > 2) return B2C_ERR_BAD_ARG;
    if (nranks > 1 && !planes) return B2C_ERR_BAD_ARG;
    if (nranks > 1 && ctx->hasCompound) {  // child manifolds do not migrate between ranks (compound.cuh)
        ctx->err = "compound shapes are not supported in a partitioned world";
        return B2C_ERR_STATE;
    }
    for (int k = 1; k < nranks - 1; k++)
        if (!(planes[k] >= planes[k - 1])) { ctx->err = "partition planes must ascend"; return B2C_ERR_BAD_ARG; }
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    dropStepGraphs(ctx);
    ctx->slab = SlabFilter{};
    ctx->partRanks = nranks;
    ctx->localHint = 0;
    ctx->haloExported = ctx->haloImported = false;
    if (nranks == 1) return B2C_OK;
    if (ctx->stagingCount || ctx->extPending) {  // transforms uploaded but not yet repacked: the owner table reads the rows
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    const size_t N = (size_t)ctx->cfg.max_bodies;
    if (!ctx->dOwner) CK(dalloc(&ctx->dOwner, N));
    if (!ctx->dLocalList) CK(dalloc(&ctx->dLocalList, N));
    SlabFilter f{};
    f.enabled = 1; f.axis = axis; f.rank = rank; f.nplanes = nranks - 1;
    for (int k = 0; k < nranks - 1; k++) f.planes[k] = planes[k];
    if (ctx->nBodies > 0) {
        k_assign_owner<<<(ctx->nBodies + 255) / 256, 256, 0, ctx->stream>>>(ctx->B, ctx->nBodies, f, ctx->dOwner);
        CK(cudaStreamSynchronize(ctx->stream));
    }
    ctx->slab = f;
    return B2C_OK;
}

// The same with planes chosen here: slabs along the axis the proxy origins spread most over, cut at equal-count quantiles
// of the origins (identical on every rank, because all ranks hold the same transforms when they call it).
int32_t b2c_set_partition(b2c_ctx* ctx, int32_t rank, int32_t nranks) {
    if (!ctx || nranks < 1 || nranks > 16 || rank < 0 || rank >= nranks) return B2C_ERR_BAD_ARG;
    if (nranks == 1) return b2c_set_partition_slabs(ctx, 0, 1, 0, nullptr);
    cudaSetDevice(ctx->device);
    if (ctx->stagingCount || ctx->extPending) {
        int32_t rc = runAabbKernel(ctx, false);
        if (rc) return rc;
    }
    const int n = ctx->nBodies;
    std::vector<float4> rows(3 * (size_t)(n > 0 ? n : 1));
    if (n > 0) {
        CK(cudaMemcpyAsync(rows.data(), ctx->B.xf4, 3 * (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    int axis = 0;
    std::vector<float> best;
    double bestSpan = -1.0;
    for (int a = 0; a < 3; a++) {
        std::vector<float> v;
        v.reserve((size_t)n);
        for (int i = 0; i < n; i++)
            if ((ctx->hFlags[i] & BF_ALIVE) && !(ctx->hFlags[i] & BF_STATIC)) {
                const float o = rows[3 * (size_t)i + a].w;
                if (o == o && std::fabs(o) < 1e29f) v.push_back(o);
            }
        if (v.empty()) continue;
        std::sort(v.begin(), v.end());
        const double span = (double)v.back() - (double)v.front();
        if (span > bestSpan) { bestSpan = span; axis = a; best.swap(v); }
    }
    float planes[15];
    for (int k = 0; k < nranks - 1; k++) {
        if (best.empty()) { planes[k] = (float)k; continue; }
        const size_t idx = (size_t)(((unsigned long long)best.size() * (unsigned)(k + 1)) / (unsigned)nranks);
        planes[k] = best[idx < best.size() ? idx : best.size() - 1];
    }
    return b2c_set_partition_slabs(ctx, rank, nranks, axis, planes);
}

int32_t b2c_get_partition(b2c_ctx* ctx, int32_t* axis_out, float* planes_out, uint8_t* owner_out, int32_t n) {
    if (!ctx || n < 0 || n > ctx->nBodies) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled) { ctx->err = "the world is not partitioned"; return B2C_ERR_STATE; }
    if (axis_out) *axis_out = ctx->slab.axis;
    if (planes_out) for (int k = 0; k < ctx->slab.nplanes; k++) planes_out[k] = ctx->slab.planes[k];
    if (owner_out && n) {
        cudaSetDevice(ctx->device);
        CK(cudaMemcpyAsync(owner_out, ctx->dOwner, (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return B2C_OK;
}

int64_t b2c_mgpu_halo_slot_bytes(int32_t cap) { return cap < 0 ? 0 : (int64_t)haloSlotBytes((uint32_t)cap); }

// updateAabbs for the proxies this rank owns, then the boundary ones (box not entirely inside the home slab) into `slot`.
int32_t b2c_mgpu_update_export_halo(b2c_ctx* ctx, void* slot, int32_t cap) {
    if (!ctx || !slot || cap < 1) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled) { ctx->err = "the world is not partitioned"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    ctx->launches = 0;
    ctx->aabbPending = true;
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], s));
    mark(ctx, 0);
    int32_t rc = runAabbKernel(ctx, true);  // clears the step counters, k_aabb over the owned proxies
    if (rc) return rc;
    CK(cudaMemsetAsync(slot, 0, HALO_HEADER_BYTES, s));
    if (ctx->nBodies > 0) {
        k_halo_export<<<(ctx->nBodies + 255) / 256, 256, 0, s>>>(ctx->B, ctx->nBodies, ctx->dOwner, ctx->slab, (unsigned char*)slot,
                                                                (uint32_t)cap, ctx->dCtr);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    ctx->haloExported = true;
    return B2C_OK;
}

// After the all-gather of the slots: adopt the records that touch this rank's slab and build the step's local list.
int32_t b2c_mgpu_import_halo(b2c_ctx* ctx, const void* slots, int32_t nslots, int32_t cap) {
    if (!ctx || !slots || cap < 1 || nslots < 1) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled || nslots != ctx->partRanks) { ctx->err = "halo import: one slot per rank of the partition"; return B2C_ERR_STATE; }
    if (!ctx->haloExported) { ctx->err = "b2c_mgpu_import_halo before b2c_mgpu_update_export_halo"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    CK(cudaMemsetAsync(ctx->dNLocal, 0, sizeof(uint32_t), s));
    dim3 grid(gridFor((uint32_t)cap, 256, 148), (unsigned)nslots);
    k_halo_import<<<grid, 256, 0, s>>>(ctx->B, (const unsigned char*)slots, (uint32_t)cap, ctx->slab, ctx->dLocalList, ctx->dNLocal, ctx->dCtr,
                                       haloSlotBytes((uint32_t)cap));
    if (ctx->nBodies > 0)
        k_list_owned<<<(ctx->nBodies + 255) / 256, 256, 0, s>>>(ctx->B, ctx->nBodies, ctx->dOwner, ctx->slab, ctx->dLocalList, ctx->dNLocal);
    ctx->launches += 2;
    CK(cudaGetLastError());
    ctx->haloExported = false;
    ctx->haloImported = true;
    return B2C_OK;
}

// ---- the halo exchange as peer-to-peer stores (no collective) -------------------------------------------------------
int32_t b2c_mgpu_p2p_init(b2c_ctx* ctx, int32_t cap, int32_t migrateCap, void* ipcHandleOut, void** inboxOut) {
    if (!ctx || cap < 1 || migrateCap < 1) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled) { ctx->err = "the world is not partitioned"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    CK(cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < 16; r++) if (ctx->haloIpcOpened[r]) { cudaIpcCloseMemHandle(ctx->haloIpcOpened[r]); ctx->haloIpcOpened[r] = nullptr; }
    cudaFree(ctx->dHaloInbox); ctx->dHaloInbox = nullptr;
    cudaFree(ctx->dHaloStage); ctx->dHaloStage = nullptr;
    cudaFree(ctx->dMigLocal); ctx->dMigLocal = nullptr;
    { const char* e = getenv("B2C_HALO_PUSH"); ctx->haloTwoPhase = !(e && e[0] == '0'); }   // 0 = the single fused kernel (A/B)
    ctx->haloConnected = false;
    ctx->haloCapP2p = (uint32_t)cap;
    ctx->haloSlotBytesP2p = (haloSlotBytes((uint32_t)cap) + 255) & ~(size_t)255;
    ctx->migCap = (uint32_t)migrateCap;
    ctx->migSlotBytes = mgpuSlotBytes((uint32_t)migrateCap);   // the stride k_import_arrival_slots assumes
    ctx->haloInboxBytes = 2 * (size_t)ctx->partRanks * ctx->haloSlotBytesP2p;
    const size_t migBytes = 2 * (size_t)ctx->partRanks * ctx->migSlotBytes;
    // ONE plain cudaMalloc for both inboxes: the allocation is exported with a single cudaIpcGetMemHandle
    CK(cudaMalloc((void**)&ctx->dHaloInbox, ctx->haloInboxBytes + migBytes));
    CK(cudaMemset(ctx->dHaloInbox, 0, ctx->haloInboxBytes + migBytes));
    ctx->dMigInbox = ctx->dHaloInbox + ctx->haloInboxBytes;
    CK(cudaMalloc((void**)&ctx->dMigLocal, ctx->migSlotBytes));
    if (!ctx->dHaloP2p) CK(cudaMalloc((void**)&ctx->dHaloP2p, sizeof(HaloP2pState)));
    ctx->haloEpoch = 0;
    if (ipcHandleOut) {
        cudaIpcMemHandle_t h;
        CK(cudaIpcGetMemHandle(&h, ctx->dHaloInbox));
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "b2c.h documents a 64-byte handle");
        memcpy(ipcHandleOut, &h, sizeof h);
    }
    if (inboxOut) *inboxOut = ctx->dHaloInbox;
    return B2C_OK;
}

int32_t b2c_mgpu_p2p_connect(b2c_ctx* ctx, const void* ipcHandles, void* const* inboxPtrs) {
    if (!ctx || (!ipcHandles && !inboxPtrs)) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled || !ctx->dHaloInbox) { ctx->err = "b2c_mgpu_p2p_connect before b2c_mgpu_p2p_init"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    const int R = ctx->partRanks, me = ctx->slab.rank;
    for (int r = 0; r < R; r++) {
        if (r == me) { ctx->haloPeers.inbox[r] = ctx->dHaloInbox; continue; }
        if (inboxPtrs) { ctx->haloPeers.inbox[r] = (unsigned char*)inboxPtrs[r]; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const unsigned char*)ipcHandles + 64 * (size_t)r, sizeof h);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            ctx->err = std::string("cudaIpcOpenMemHandle of a peer's halo inbox failed: ") + cudaGetErrorString(e);
            cudaGetLastError();
            return B2C_ERR_CUDA;
        }
        ctx->haloIpcOpened[r] = p;
        ctx->haloPeers.inbox[r] = (unsigned char*)p;
    }
    // every rank uses the same capacities, so the migration inbox sits at the same offset in every allocation
    for (int r = 0; r < R; r++) ctx->migPeers.inbox[r] = ctx->haloPeers.inbox[r] + ctx->haloInboxBytes;
    ctx->haloConnected = true;
    return B2C_OK;
}

int32_t b2c_mgpu_p2p_export_halo(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled || !ctx->haloConnected) { ctx->err = "peer-to-peer halo exchange is not connected"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    ctx->launches = 0;
    ctx->aabbPending = true;
    cudaStream_t s = ctx->stream;
    CK(cudaEventRecord(ctx->ev[0], s));
    mark(ctx, 0);
    int32_t rc = runAabbKernel(ctx, true);  // clears the step counters, k_aabb over the owned proxies
    if (rc) return rc;
    ctx->haloEpoch++;
    CK(cudaMemsetAsync(ctx->dHaloP2p, 0, sizeof(HaloP2pState), s));
    const int nb = ctx->nBodies > 0 ? ctx->nBodies : 1;   // at least one block: the headers must be published
    if (ctx->haloTwoPhase) {
        if (!ctx->dHaloStage) CK(cudaMalloc((void**)&ctx->dHaloStage, (size_t)ctx->partRanks * ctx->haloCapP2p * sizeof(HaloRecord)));
        k_halo_stage<<<(nb + 255) / 256, 256, 0, s>>>(ctx->B, ctx->nBodies, ctx->dOwner, ctx->slab, ctx->partRanks, ctx->dHaloStage,
                                                     ctx->haloCapP2p, ctx->dHaloP2p, ctx->dCtr);
        k_halo_push<<<dim3(16, (unsigned)ctx->partRanks), 256, 0, s>>>(ctx->dHaloStage, ctx->haloPeers, ctx->partRanks, ctx->slab.rank,
                                                                       ctx->haloSlotBytesP2p, ctx->haloCapP2p, ctx->haloEpoch, ctx->dHaloP2p,
                                                                       reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(ctx->dHaloP2p) + offsetof(HaloP2pState, pushTicket)));
        ctx->launches += 2;
    } else {
        k_halo_export_p2p<<<(nb + 255) / 256, 256, 0, s>>>(ctx->B, ctx->nBodies, ctx->dOwner, ctx->slab, ctx->haloPeers, ctx->partRanks,
                                                          ctx->haloSlotBytesP2p, ctx->haloCapP2p, ctx->haloEpoch, ctx->dHaloP2p, ctx->dCtr);
        ctx->launches++;
    }
    CK(cudaGetLastError());
    ctx->haloExported = true;
    return B2C_OK;
}

int32_t b2c_mgpu_p2p_import_halo(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled || !ctx->haloConnected) { ctx->err = "peer-to-peer halo exchange is not connected"; return B2C_ERR_STATE; }
    if (!ctx->haloExported) { ctx->err = "b2c_mgpu_p2p_import_halo before b2c_mgpu_p2p_export_halo"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const unsigned char* slots = ctx->dHaloInbox + (size_t)(ctx->haloEpoch & 1u) * (size_t)ctx->partRanks * ctx->haloSlotBytesP2p;
    CK(cudaMemsetAsync(ctx->dNLocal, 0, sizeof(uint32_t), s));
    k_halo_wait<<<1, 32, 0, s>>>(slots, ctx->partRanks, ctx->slab.rank, ctx->haloSlotBytesP2p, ctx->haloEpoch, ctx->dCtr);
    dim3 grid(gridFor(ctx->haloCapP2p, 256, 148), (unsigned)ctx->partRanks);
    k_halo_import<<<grid, 256, 0, s>>>(ctx->B, slots, ctx->haloCapP2p, ctx->slab, ctx->dLocalList, ctx->dNLocal, ctx->dCtr, ctx->haloSlotBytesP2p);
    if (ctx->nBodies > 0)
        k_list_owned<<<(ctx->nBodies + 255) / 256, 256, 0, s>>>(ctx->B, ctx->nBodies, ctx->dOwner, ctx->slab, ctx->dLocalList, ctx->dNLocal);
    ctx->launches += 3;
    CK(cudaGetLastError());
    ctx->haloExported = false;
    ctx->haloImported = true;
    return B2C_OK;
}

// Manifolds of pairs this rank no longer owns -> every other rank's migration inbox (replaces export_departed_slot + the
// all-gather); then wait + adoption (replaces import_arrival_slots).
int32_t b2c_mgpu_p2p_export_departed(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled || !ctx->haloConnected) { ctx->err = "peer-to-peer exchange is not connected"; return B2C_ERR_STATE; }
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    int32_t rc = b2c_mgpu_export_departed_slot(ctx, ctx->dMigLocal, (int32_t)ctx->migCap);
    if (rc) return rc;
    cudaStream_t s = ctx->stream;
    uint32_t* tickets = reinterpret_cast<uint32_t*>(reinterpret_cast<unsigned char*>(ctx->dHaloP2p) + offsetof(HaloP2pState, migrateTicket));
    k_migrate_push<<<dim3(4, (unsigned)ctx->partRanks), 256, 0, s>>>(ctx->dMigLocal, ctx->migPeers, ctx->partRanks, ctx->slab.rank,
                                                                     ctx->migSlotBytes, ctx->migCap, ctx->haloEpoch, tickets, ctx->dCtr);
    ctx->launches++;
    CK(cudaGetLastError());
    ctx->migExported = true;
    return B2C_OK;
}

int32_t b2c_mgpu_p2p_import_arrivals(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    if (!ctx->slab.enabled || !ctx->haloConnected) { ctx->err = "peer-to-peer exchange is not connected"; return B2C_ERR_STATE; }
    if (!ctx->migExported) { ctx->err = "b2c_mgpu_p2p_import_arrivals before b2c_mgpu_p2p_export_departed"; return B2C_ERR_STATE; }
    cudaSetDevice(ctx->device);
    const unsigned char* slots = ctx->dMigInbox + (size_t)(ctx->haloEpoch & 1u) * (size_t)ctx->partRanks * ctx->migSlotBytes;
    k_halo_wait<<<1, 32, 0, ctx->stream>>>(slots, ctx->partRanks, ctx->slab.rank, ctx->migSlotBytes, ctx->haloEpoch, ctx->dCtr);
    ctx->launches++;
    ctx->migExported = false;
    return b2c_mgpu_import_arrival_slots(ctx, slots, ctx->partRanks, (int32_t)ctx->migCap);
}

int32_t b2c_mgpu_broadphase(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    if (!ctx->slab.enabled) {  // a partition of one: the plain pair calculation
        ctx->launches = 0;
        ctx->aabbPending = true;
        CK(cudaEventRecord(ctx->ev[0], ctx->stream));
    }
    int32_t rc = enqueuePhases(ctx, 1);
    CK(cudaEventRecord(ctx->ev[2], ctx->stream));
    return rc;
}

int32_t b2c_mgpu_export_departed(b2c_ctx* ctx, uint64_t* keys, void* hdrs, b2c_manifold_point* pts, int32_t cap, int32_t* countOut) {
    if (!ctx || !keys || !hdrs || !pts || cap < 0 || !countOut) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const int cur = ctx->cur, prev = cur ^ 1;
    CK(cudaMemsetAsync(ctx->dExportCount, 0, sizeof(uint32_t), s));
    k_export_departed<<<gridFor((uint32_t)ctx->cfg.max_pairs, 256), 256, 0, s>>>(
        ctx->dSortedKeys[prev], ctx->dNumPairs[prev], ctx->dMHdr[prev], ctx->dMPts[prev], ctx->dSortedKeys[cur], ctx->dNumPairs[cur],
        ctx->dPairFirst[cur], ctx->uidBits, keys, (ManifoldHdr*)hdrs, pts, (uint32_t)cap, ctx->dExportCount);
    ctx->launches++;
    uint32_t c = 0;
    CK(cudaMemcpyAsync(&c, ctx->dExportCount, sizeof(c), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *countOut = (int32_t)c;
    if (c > (uint32_t)cap) { ctx->err = "departed-manifold export buffer too small"; return B2C_ERR_CAPACITY; }
    return B2C_OK;
}

int32_t b2c_mgpu_import_arrivals(b2c_ctx* ctx, const uint64_t* keys, const void* hdrs, const b2c_manifold_point* pts, int32_t count) {
    if (!ctx || count < 0 || (count > 0 && (!keys || !hdrs || !pts))) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    if (count == 0) return B2C_OK;
    cudaSetDevice(ctx->device);
    const int cur = ctx->cur;
    k_import_arrivals<<<gridFor((uint32_t)count, 256), 256, 0, ctx->stream>>>(
        keys, (const ManifoldHdr*)hdrs, pts, (uint32_t)count, ctx->dSortedKeys[cur], ctx->dNumPairs[cur], ctx->dPairFirst[cur],
        ctx->uidBits, ctx->dMHdr[cur], ctx->dMPts[cur], ctx->dCtr);
    ctx->launches++;
    CK(cudaGetLastError());
    return B2C_OK;
}

int64_t b2c_mgpu_slot_bytes(int32_t cap) { return cap < 0 ? 0 : (int64_t)mgpuSlotBytes((uint32_t)cap); }

int32_t b2c_mgpu_export_departed_slot(b2c_ctx* ctx, void* slot, int32_t cap) {
    if (!ctx || !slot || cap < 1) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    cudaStream_t s = ctx->stream;
    const int cur = ctx->cur, prev = cur ^ 1;
    unsigned char* b = (unsigned char*)slot;
    CK(cudaMemsetAsync(b, 0, 16, s));
    k_export_departed<<<gridFor((uint32_t)ctx->cfg.max_pairs, 256), 256, 0, s>>>(
        ctx->dSortedKeys[prev], ctx->dNumPairs[prev], ctx->dMHdr[prev], ctx->dMPts[prev], ctx->dSortedKeys[cur], ctx->dNumPairs[cur],
        ctx->dPairFirst[cur], ctx->uidBits, (uint64_t*)(b + 16), (ManifoldHdr*)(b + 16 + (size_t)cap * 8),
        (b2c_manifold_point*)(b + 16 + (size_t)cap * (8 + sizeof(ManifoldHdr))), (uint32_t)cap, (uint32_t*)b);
    ctx->launches++;
    CK(cudaGetLastError());
    return B2C_OK;
}

int32_t b2c_mgpu_import_arrival_slots(b2c_ctx* ctx, const void* slots, int32_t nslots, int32_t cap) {
    if (!ctx || !slots || nslots < 1 || cap < 1) return B2C_ERR_BAD_ARG;
    if (!ctx->pairsValid) return B2C_ERR_STATE;
    cudaSetDevice(ctx->device);
    const int cur = ctx->cur;
    dim3 grid(gridFor((uint32_t)cap, 256, 64), (unsigned)nslots);
    k_import_arrival_slots<<<grid, 256, 0, ctx->stream>>>((const unsigned char*)slots, (uint32_t)nslots, (uint32_t)cap,
                                                           ctx->dSortedKeys[cur], ctx->dNumPairs[cur], ctx->dPairFirst[cur], ctx->uidBits,
                                                           ctx->dMHdr[cur], ctx->dMPts[cur], ctx->dCtr);
    ctx->launches++;
    CK(cudaGetLastError());
    return B2C_OK;
}

int32_t b2c_mgpu_narrowphase(b2c_ctx* ctx) {
    if (!ctx) return B2C_ERR_BAD_ARG;
    cudaSetDevice(ctx->device);
    int32_t rc = enqueuePhases(ctx, 2);
    if (rc) return rc;
    CK(cudaEventRecord(ctx->ev[3], ctx->stream));
    ctx->stats.kernel_launches = ctx->launches;
    return B2C_OK;
}

void* b2c_stream(b2c_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
float* b2c_device_transforms(b2c_ctx* ctx) { return ctx ? ctx->dStaging : nullptr; }

}  // extern "C"
