// TEST INFRASTRUCTURE ONLY — never built into, loaded by or shipped with the product (the package fails without a GPU).
// Host emulation of the compound kernels (compound.cuh + the shared EPA bin of narrowphase.cuh): the device headers are compiled
// for the CPU with tests/emu/cuda_runtime.h (a shim of the CUDA built-ins, one emulated lane per warp) and driven kernel by
// kernel against the oracle on a seeded scene; every record must be BIT-identical.  It checks the kernels' logic — item
// expansion and persistence, detector / manifold semantics, traversal orders — where no GPU is available; scheduling, atomics
// across threads and launch plumbing are only exercised by the -m gpu tests.  Run by tests/test_emu_kernels.py.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>
#include "cuda_runtime.h"
#include "../../oracle/world.h"
#include "../../libgdx-jbullet_b200/csrc/compound.cuh"
#include "../../libgdx-jbullet_b200/csrc/compound_flatten.h"

namespace b2c { alignas(16) unsigned char epaSmem[2 * sizeof(EpaScratch)]; }

using namespace b2c;

static std::mt19937 rng(12345);
static float uf(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }

struct Sc {
    orc::World W;
    std::vector<ShapeDev> shapes;
    std::vector<float4> hull;
    std::vector<CompoundChildDev> children;
    int addBox(float x, float y, float z) {
        W.shapes.emplace_back(); W.meshes.emplace_back(nullptr);
        orc::initBox(W.shapes.back(), orc::V3(x, y, z));
        ShapeDev s{}; s.type = SH_BOX; s.margin = 0.04f; s.dims[0] = x * 1.0f - s.margin; s.dims[1] = y * 1.0f - s.margin; s.dims[2] = z * 1.0f - s.margin;
        shapes.push_back(s); return (int)shapes.size() - 1;
    }
    int addSphere(float r) {
        W.shapes.emplace_back(); W.meshes.emplace_back(nullptr);
        orc::initSphere(W.shapes.back(), r);
        ShapeDev s{}; s.type = SH_SPHERE; s.dims[0] = r; s.margin = r * 1.0f;
        shapes.push_back(s); return (int)shapes.size() - 1;
    }
    int addHull(int n, float rad) {
        std::vector<float> pts(3 * n);
        for (int i = 0; i < n; i++) {
            float v[3] = {uf(-1, 1), uf(-1, 1), uf(-1, 1)};
            float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-3f;
            for (int c = 0; c < 3; c++) pts[3 * i + c] = v[c] / l * rad * uf(0.8f, 1.0f);
        }
        W.shapes.emplace_back(); W.meshes.emplace_back(nullptr);
        orc::initHull(W.shapes.back(), pts.data(), n);
        ShapeDev s{}; s.type = SH_HULL; s.margin = 0.04f; s.pointOffset = (int)hull.size(); s.numPoints = n;
        float mx[3] = {0, 0, 0}, mn[3] = {0, 0, 0}, wmx[3] = {-1e30f, -1e30f, -1e30f}, wmn[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = 0; i < n; i++) {
            float v[3] = {pts[3 * i], pts[3 * i + 1], pts[3 * i + 2]};
            hull.push_back(make_float4(v[0], v[1], v[2], 0.f));
            for (int c = 0; c < 3; c++) {
                if (v[c] > wmx[c]) { wmx[c] = v[c]; mx[c] = v[c]; }
                if (-v[c] > wmn[c]) { wmn[c] = -v[c]; mn[c] = v[c]; }
            }
        }
        for (int c = 0; c < 3; c++) { s.aabbMax[c] = mx[c] + s.margin; s.aabbMin[c] = mn[c] - s.margin; }
        shapes.push_back(s); return (int)shapes.size() - 1;
    }
    int addPlane(float nx, float ny, float nz, float c) {
        W.shapes.emplace_back(); W.meshes.emplace_back(nullptr);
        orc::initPlane(W.shapes.back(), orc::V3(nx, ny, nz), c);
        ShapeDev s{}; s.type = SH_PLANE; s.margin = 0;
        s.plane[0] = W.shapes.back().planeNormal.x; s.plane[1] = W.shapes.back().planeNormal.y; s.plane[2] = W.shapes.back().planeNormal.z; s.plane[3] = c;
        shapes.push_back(s); return (int)shapes.size() - 1;
    }
    std::vector<std::vector<CompoundDirectChild>> directOfShape;   // per shape id, as b2c_api.cu keeps it
    int addCompound(const std::vector<int>& kids, const std::vector<float>& xf12) {
        int sid = W.addCompound((int)kids.size(), kids.data(), xf12.data());
        // the same table b2c_shape_register_compound builds: frames + leaves (compound_flatten.h), then the direct children as
        // scratch entries for the local-AABB kernel
        std::vector<CompoundDirectChild> direct(kids.size());
        for (size_t i = 0; i < kids.size(); i++) { direct[i].shape = kids[i]; for (int k = 0; k < 12; k++) direct[i].xf12[k] = xf12[12 * i + k]; }
        int firstLeaf = 0, numLeaves = 0;
        if (!flattenCompound(children, direct, [&](int sh) -> const std::vector<CompoundDirectChild>* {
                return (sh < (int)directOfShape.size() && !directOfShape[sh].empty()) ? &directOfShape[sh] : nullptr; }, firstLeaf, numLeaves)) {
            printf("compound nesting too deep\n"); exit(1);
        }
        const size_t keep = children.size();
        int first = (int)children.size();
        for (size_t i = 0; i < kids.size(); i++) {
            CompoundChildDev ch{};
            for (int k = 0; k < 9; k++) ch.m[k] = xf12[12 * i + k];
            for (int k = 0; k < 3; k++) ch.o[k] = xf12[12 * i + 9 + k];
            ch.shape = kids[i];
            children.push_back(ch);
        }
        float out6[6];
        blockIdx = {0, 0, 0}; threadIdx = {0, 0, 0}; blockDim = {1, 1, 1}; gridDim = {1, 1, 1};
        k_compound_local_aabb(shapes.data(), children.data(), first, (int)kids.size(), out6);
        ShapeDev s{}; s.type = SH_COMPOUND; s.margin = 0;
        for (int c = 0; c < 3; c++) { s.aabbMin[c] = out6[c]; s.aabbMax[c] = out6[3 + c]; }
        children.resize(keep);
        s.pointOffset = firstLeaf; s.numPoints = numLeaves;
        shapes.push_back(s);
        if (directOfShape.size() < shapes.size()) directOfShape.resize(shapes.size());
        directOfShape[shapes.size() - 1] = direct;
        const orc::Shape& os = W.shapes[sid];
        float o6[6] = {os.localAabbMin.x, os.localAabbMin.y, os.localAabbMin.z, os.localAabbMax.x, os.localAabbMax.y, os.localAabbMax.z};
        if (memcmp(o6, out6, 24)) { printf("compound local AABB differs\n"); exit(1); }
        return (int)shapes.size() - 1;
    }
};

static void randRot(float m[9]) {
    float q[4] = {uf(-1, 1), uf(-1, 1), uf(-1, 1), uf(-1, 1)};
    float l = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-6f;
    float x = q[0] / l, y = q[1] / l, z = q[2] / l, w = q[3] / l;
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
static std::vector<float> idXf(float x, float y, float z) { return {1, 0, 0, 0, 1, 0, 0, 0, 1, x, y, z}; }
static void app(std::vector<float>& a, const std::vector<float>& b) { a.insert(a.end(), b.begin(), b.end()); }

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 150;
    const int STEPS = 8; const float SP = argc > 2 ? atof(argv[2]) : 1.0f; const float YOFF = argc > 4 ? atof(argv[4]) : 0.f;
    Sc sc;
    sc.W.mode = orc::BP_DBVT;
    const bool useMesh = argc > 3;
    int plane = sc.addPlane(0, 1, 0, useMesh ? -50.f : 0.f);
    // heightfield mesh under the lattice
    const int C = 24; static std::vector<float> verts; static std::vector<int> idx; static std::vector<int4> nodes; std::vector<MeshDev> meshes;
    int meshShape = -1;
    if (useMesh) {
        for (int i = 0; i <= C; i++) for (int j = 0; j <= C; j++) { verts.push_back(i * 0.5f - 2.f); verts.push_back(0.25f * sinf(i * 0.9f) * cosf(j * 0.7f)); verts.push_back(j * 0.5f - 2.f); }
        for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) { int v00 = i * (C + 1) + j, v10 = (i + 1) * (C + 1) + j, v01 = v00 + 1, v11 = v10 + 1; idx.insert(idx.end(), {v00, v01, v10, v10, v01, v11}); }
        meshShape = sc.W.addMesh(verts.data(), (int)verts.size() / 3, idx.data(), (int)idx.size() / 3);
        nodes.resize(sc.W.meshes[meshShape]->bvh.nodes.size());
        memcpy(nodes.data(), sc.W.meshes[meshShape]->bvh.nodes.data(), nodes.size() * 16);
        MeshDev md{}; md.nodes = nodes.data(); md.verts = verts.data(); md.idx = idx.data(); md.numNodes = (int)nodes.size(); md.numTris = (int)idx.size() / 3;
        const orc::Bvh& b = sc.W.meshes[meshShape]->bvh;
        md.qmin[0] = b.bvhAabbMin.x; md.qmin[1] = b.bvhAabbMin.y; md.qmin[2] = b.bvhAabbMin.z; md.qmax[0] = b.bvhAabbMax.x; md.qmax[1] = b.bvhAabbMax.y; md.qmax[2] = b.bvhAabbMax.z;
        md.quant[0] = b.bvhQuantization.x; md.quant[1] = b.bvhQuantization.y; md.quant[2] = b.bvhQuantization.z; meshes.push_back(md);
        ShapeDev s{}; s.type = SH_MESH; s.mesh = 0; const orc::Shape& os = sc.W.shapes[meshShape];
        s.aabbMin[0] = os.localAabbMin.x; s.aabbMin[1] = os.localAabbMin.y; s.aabbMin[2] = os.localAabbMin.z; s.aabbMax[0] = os.localAabbMax.x; s.aabbMax[1] = os.localAabbMax.y; s.aabbMax[2] = os.localAabbMax.z;
        sc.shapes.push_back(s);
        if ((int)sc.shapes.size() - 1 != meshShape) { printf("shape index mismatch\n"); return 1; }
    }
    int sSmall = sc.addSphere(0.3f), sBig = sc.addSphere(0.4f);
    int bar = sc.addBox(0.45f, 0.12f, 0.12f), slab = sc.addBox(0.4f, 0.15f, 0.3f), post = sc.addBox(0.15f, 0.4f, 0.15f);
    int hl = sc.addHull(16, 0.35f), hl2 = sc.addHull(16, 0.4f);
    int bx = sc.addBox(0.3f, 0.35f, 0.4f);
    std::vector<int> comps;
    { std::vector<float> x; app(x, idXf(-0.5f, 0, 0)); app(x, idXf(0, 0, 0)); app(x, idXf(0.5f, 0, 0)); comps.push_back(sc.addCompound({sSmall, bar, sBig}, x)); }
    { std::vector<float> x; app(x, idXf(0, -0.25f, 0)); app(x, idXf(0.25f, 0.3f, 0)); comps.push_back(sc.addCompound({slab, post}, x)); }
    { std::vector<float> x(24); randRot(x.data()); x[9] = -0.2f; x[10] = 0; x[11] = 0.1f; randRot(x.data() + 12); x[21] = 0.3f; x[22] = 0.1f; x[23] = -0.1f; comps.push_back(sc.addCompound({hl, sSmall}, x)); }
    { std::vector<float> x; app(x, idXf(0, 0, 0)); app(x, {0, -1, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0}); app(x, {0, 0, 1, 0, 1, 0, -1, 0, 0, 0, 0, 0}); comps.push_back(sc.addCompound({bar, bar, bar}, x)); }
    { std::vector<float> x; app(x, idXf(0, 0.2f, 0)); comps.push_back(sc.addCompound({sBig}, x)); }
    // nested: children that are CompoundShapes themselves (one and two levels), with rotated frames
    { std::vector<float> x(24); randRot(x.data()); x[9] = 0.1f; x[10] = 0.35f; x[11] = 0; randRot(x.data() + 12); x[21] = -0.1f; x[22] = -0.3f; x[23] = 0.05f; comps.push_back(sc.addCompound({comps[0], post}, x)); }
    { std::vector<float> x(24); randRot(x.data()); x[9] = 0; x[10] = 0.2f; x[11] = 0.1f; randRot(x.data() + 12); x[21] = 0.2f; x[22] = -0.4f; x[23] = 0; comps.push_back(sc.addCompound({comps[5], sSmall}, x)); }
    { std::vector<float> x; app(x, idXf(0, 0, 0)); comps.push_back(sc.addCompound({comps[2]}, x)); }   // identity frame around compound 2
    std::vector<int> plain = {sSmall, sBig, hl, hl2, bx, bar};
    // bodies
    std::vector<int> bodyShape;
    std::vector<float> base;  // 12 per body
    std::vector<float> vel;
    auto addBody = [&](int shape, const float* xf12, bool isStatic) {
        orc::Xf x;
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) x.basis.m[r][c] = xf12[3 * r + c];
        x.origin.set(xf12[9], xf12[10], xf12[11]);
        sc.W.addBody(shape, x, isStatic ? 2 : 1, isStatic ? (-1 ^ 2) : -1, isStatic, 0);
        bodyShape.push_back(shape);
        base.insert(base.end(), xf12, xf12 + 12);
        for (int c = 0; c < 3; c++) vel.push_back(isStatic ? 0.f : uf(-0.01f, 0.01f));
    };
    { auto g = idXf(0, 0, 0); addBody(plane, g.data(), true); }
    if (useMesh) { auto g = idXf(0, 0.1f, 0); addBody(meshShape, g.data(), true); }
    int m = (int)ceil(cbrt((double)N));
    for (int i = 0; i < N; i++) {
        float xf[12];
        randRot(xf);
        xf[9] = (i % m) * SP + uf(-0.12f, 0.12f);
        xf[10] = (i / (m * m)) * SP * 0.8f + (useMesh ? 0.45f : 0.5f) + YOFF + uf(-0.12f, 0.12f);
        xf[11] = ((i / m) % m) * SP + uf(-0.12f, 0.12f);
        int shape = (uf(0, 1) < 0.5f) ? comps[rng() % comps.size()] : plain[rng() % plain.size()];
        addBody(shape, xf, false);
    }
    const int NB = (int)bodyShape.size();
    sc.hull.resize(sc.hull.size() + 8);
    // device-side arrays
    std::vector<float4> xf4(3 * NB);
    std::vector<float2> material(NB, make_float2(0.5f, 0.0f));
    const uint32_t MAXI = 1 << 18;
    std::vector<uint32_t> itemPair(MAXI), itemCode(MAXI);
    std::vector<EpaItem> epaItems(MAXI); std::vector<uint32_t> epaRetry(MAXI), epaBig(MAXI);
    std::vector<int> itemPrev(MAXI);
    std::vector<b2c_raw_contact> craw(MAXI), rawMesh(MAXI);
    std::vector<uint32_t> cMeshStart(MAXI), cMeshCount(MAXI);
    static std::vector<EpaScratch> big(1);
    std::vector<ManifoldHdr> CH[2] = {std::vector<ManifoldHdr>(MAXI), std::vector<ManifoldHdr>(MAXI)};
    std::vector<b2c_manifold_point> CP[2] = {std::vector<b2c_manifold_point>(4 * MAXI), std::vector<b2c_manifold_point>(4 * MAXI)};
    int ccur = 0;
    std::map<std::pair<int, int>, ManifoldHdr> prevHdr;
    StepCounters ctr{};
    CompoundCounters cc{};
    long totalItems = 0, totalTouch = 0, totalRetry = 0, totalDeep = 0, keepItems = 0, bigUsed = 0;
    for (int step = 0; step < STEPS; step++) {
        // transforms + activity
        for (int b = 0; b < NB; b++) {
            orc::Body& B = sc.W.bodies[b];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B.xf.basis.m[r][c] = base[12 * b + 3 * r + c];
            B.xf.origin.set(base[12 * b + 9] + vel[3 * b] * step, base[12 * b + 10] + vel[3 * b + 1] * step, base[12 * b + 11] + vel[3 * b + 2] * step);
            B.active = (step < 3 || step == 7) ? true : (uf(0, 1) > 0.4f);
            for (int r = 0; r < 3; r++)
                xf4[3 * b + r] = make_float4(B.xf.basis.m[r][0], B.xf.basis.m[r][1], B.xf.basis.m[r][2], r == 0 ? B.xf.origin.x : (r == 1 ? B.xf.origin.y : B.xf.origin.z));
        }
        sc.W.updateAabbs();
        sc.W.calculateOverlappingPairs();
        long deepBefore = sc.W.deepPenetrationChecks;
        sc.W.dispatchAllPairs();
        // ---- emulated device side
        const uint32_t P = (uint32_t)sc.W.pairs.size();
        std::vector<int2> pairs(P);
        std::vector<ManifoldHdr> mhdr(P);
        std::vector<uint32_t> binItems;
        std::vector<uint32_t> keepList;
        for (uint32_t p = 0; p < P; p++) {
            auto pr = sc.W.pairs[p];
            pairs[p] = make_int2(pr.first, pr.second);
            auto it = prevHdr.find(pr);
            ManifoldHdr h{};
            if (it != prevHdr.end()) { h = it->second; if (h.algorithm == 5) h.pad0 = 0; }   // k_carry
            else { h.pair_uid0 = pr.first; h.pair_uid1 = pr.second; }
            mhdr[p] = h;
            const orc::Body& b0 = sc.W.bodies[pr.first - 1];
            const orc::Body& b1 = sc.W.bodies[pr.second - 1];
            int t0 = sc.shapes[bodyShape[pr.first - 1]].type, t1 = sc.shapes[bodyShape[pr.second - 1]].type;
            bool dispatch = b0.active || b1.active;
            if (compoundPairSupported(t0, t1)) { if (dispatch) binItems.push_back(p); else keepList.push_back(p); }
        }
        uint32_t binStart[17] = {0};
        binStart[BIN_COMPOUND] = 0; binStart[BIN_COMPOUND_KEEP] = (uint32_t)binItems.size();
        binItems.insert(binItems.end(), keepList.begin(), keepList.end());
        binStart[BIN_COMPOUND_KEEP + 1] = (uint32_t)binItems.size();
        uint32_t numPairs = P;
        NpArgs a{};
        a.pairs = pairs.data(); a.numPairs = &numPairs; a.xf4 = xf4.data(); a.shape = bodyShape.data(); a.material = material.data();
        a.shapes = sc.shapes.data(); a.hullPts = sc.hull.data(); a.mhdr = mhdr.data(); a.binItems = binItems.data(); a.binStart = binStart;
        a.ctr = &ctr; a.threshold = 0.02f; a.maxPairs = P; a.hasCompound = 1;
        ctr = StepCounters{};
        cc = CompoundCounters{};
        CompoundArgs c{};
        c.children = sc.children.data(); c.cc = &cc; c.itemPair = itemPair.data(); c.itemCode = itemCode.data(); c.itemPrev = itemPrev.data();
        c.raw = craw.data(); c.meshStart = cMeshStart.data(); c.meshCount = cMeshCount.data(); c.bigScratch = big.data(); c.numBigScratch = 1;
        a.meshes = meshes.data();
        c.H = CH[ccur ^ 1].data(); c.P = CP[ccur ^ 1].data(); c.prevH = CH[ccur].data(); c.prevP = CP[ccur].data(); c.maxItems = MAXI;
        blockIdx = {0, 0, 0}; threadIdx = {0, 0, 0}; blockDim = {1, 1, 1}; gridDim = {1, 1, 1};
        GjkArgs g{};
        g.rawMesh = rawMesh.data(); g.maxMeshItems = MAXI;
        g.epaItems = epaItems.data(); g.maxEpa = MAXI; g.epaRetry = epaRetry.data(); g.maxEpaRetry = MAXI; g.epaBig = epaBig.data(); g.comp = c;
        uint32_t cursor = 0;
        k_compound_expand(a, c);
        k_compound_gjk(a, g, &cursor);
        k_epa<2>(a, g, 1, 32);
        blockDim = {32, 1, 1};
        for (unsigned t = 0; t < 32; t++) { threadIdx.x = t; k_epa<1>(a, g, 1, 32); }   // items routed straight to the large pools
        for (unsigned t = 0; t < 32; t++) { threadIdx.x = t; k_epa<1>(a, g, 0, 32); }
        threadIdx.x = 0;
        blockDim = {1, 1, 1};
        k_compound_manifold(a, c);
        if (useMesh) {
            memset((void*)big.data(), 0xAB, sizeof(EpaScratch));
            g.comp = c;
            k_compound_mesh(a, g);
            const unsigned char* bytes = reinterpret_cast<const unsigned char*>(big.data());
            for (size_t q = 0; q < sizeof(EpaScratch); q++) if (bytes[q] != 0xAB) { bigUsed++; break; }
        }
        ccur ^= 1;
        prevHdr.clear();
        for (uint32_t p = 0; p < P; p++) prevHdr[sc.W.pairs[p]] = mhdr[p];
        totalItems += cc.numItems; totalRetry += ctr.epaRetry; totalDeep += ctr.deepChecks;
        // ---- compare
        long oracleKidManifolds = 0;
        std::map<std::pair<int, int>, uint32_t> pairIndex;
        for (uint32_t p = 0; p < P; p++) pairIndex[sc.W.pairs[p]] = p;
        // raw records of the oracle, keyed (uid0, uid1, tri)
        std::map<std::tuple<int, int, int>, const orc::RawContact*> oraw;
        for (auto& r : sc.W.raw) if (r.tri <= -2) oraw[std::make_tuple(r.uid0, r.uid1, r.tri)] = &r;
        size_t meshRawSeen = 0;
        size_t rawSeen = 0;
        for (auto& kv : sc.W.pairState) {
            const orc::PairState& ps = kv.second;
            if (ps.kids.empty()) continue;
            uint32_t p = pairIndex.at(kv.first);
            const ManifoldHdr& h = mhdr[p];
            if (h.algorithm != 5) { printf("step %d pair (%d,%d): header algorithm %d\n", step, kv.first.first, kv.first.second, h.algorithm); return 1; }
            const orc::Body& b0 = sc.W.bodies[kv.first.first - 1];
            const orc::Body& b1 = sc.W.bodies[kv.first.second - 1];
            bool dispatched = b0.active || b1.active;
            for (size_t k = 0; k < ps.kids.size(); k++) {
                const orc::PersistentManifold& om = ps.kids[k].manifold;
                uint32_t item = (uint32_t)h.pad1 + (uint32_t)k;
                const ManifoldHdr& gh = CH[ccur][item];
                const b2c_manifold_point* gp = &CP[ccur][4 * (size_t)item];
                oracleKidManifolds += ps.kids[k].hasManifold ? 1 : 0;
                if (!dispatched) keepItems++;
                if (gh.body0 != om.body0 || gh.body1 != om.body1 || gh.num_contacts != om.cachedPoints || gh.pad0 != ps.kidChild[k].first ||
                    gh.pad1 != ps.kidChild[k].second || gh.pair_uid0 != kv.first.first || gh.pair_uid1 != kv.first.second) {
                    printf("step %d pair (%d,%d) kid %zu header differs: gpu b(%d,%d) n %d ch(%d,%d) alg %d | oracle b(%d,%d) n %d ch(%d,%d)\n", step, kv.first.first,
                           kv.first.second, k, gh.body0, gh.body1, gh.num_contacts, gh.pad0, gh.pad1, gh.algorithm, om.body0, om.body1, om.cachedPoints,
                           ps.kidChild[k].first, ps.kidChild[k].second);
                    return 1;
                }
                for (int q = 0; q < om.cachedPoints; q++) {
                    const orc::ManifoldPoint& op = om.pointCache[q];
                    float of[16] = {op.localPointA.x, op.localPointA.y, op.localPointA.z, op.localPointB.x, op.localPointB.y, op.localPointB.z,
                                    op.positionWorldOnA.x, op.positionWorldOnA.y, op.positionWorldOnA.z, op.positionWorldOnB.x, op.positionWorldOnB.y,
                                    op.positionWorldOnB.z, op.normalWorldOnB.x, op.normalWorldOnB.y, op.normalWorldOnB.z, op.distance1};
                    if (memcmp(of, &gp[q], 64) || gp[q].life_time != op.lifeTime || gp[q].src_slot != op.srcSlot ||
                        gp[q].combined_friction != op.combinedFriction || gp[q].combined_restitution != op.combinedRestitution || gp[q].index1 != op.index1) {
                        printf("step %d pair (%d,%d) kid %zu point %d differs (life %d/%d src %d/%d dist %g/%g)\n", step, kv.first.first, kv.first.second, k, q,
                               gp[q].life_time, op.lifeTime, gp[q].src_slot, op.srcSlot, gp[q].distance, op.distance1);
                        return 1;
                    }
                    totalTouch++;
                }
                if (dispatched && craw[item].has_contact == -3) {
                    for (uint32_t t = cMeshStart[item]; t < cMeshStart[item] + cMeshCount[item]; t++) {
                        const b2c_raw_contact& gr = rawMesh[t];
                        auto itr = oraw.find(std::make_tuple(gr.uid0, gr.uid1, gr.tri));
                        if (itr == oraw.end()) { printf("step %d: oracle mesh raw record missing (tri %d)\n", step, gr.tri); return 1; }
                        const orc::RawContact& orr = *itr->second;
                        rawSeen++; meshRawSeen++;
                        if (gr.has_contact != orr.hasContact || gr.method != orr.method || gr.iters != orr.iters || memcmp(gr.normal, orr.normal, 12) ||
                            memcmp(gr.point, orr.point, 12) || memcmp(&gr.depth, &orr.depth, 4)) {
                            printf("step %d pair (%d,%d) kid %zu mesh raw differs: has %d/%d method %d/%d\n", step, kv.first.first, kv.first.second, k, gr.has_contact, orr.hasContact, gr.method, orr.method);
                            return 1;
                        }
                    }
                } else if (dispatched) {
                    auto itr = oraw.find(std::make_tuple(kv.first.first, kv.first.second, -2 - (int)k));
                    if (itr == oraw.end()) { printf("step %d: oracle raw record missing\n", step); return 1; }
                    const orc::RawContact& orr = *itr->second;
                    const b2c_raw_contact& gr = craw[item];
                    rawSeen++;
                    if (gr.uid0 != orr.uid0 || gr.uid1 != orr.uid1 || gr.tri != orr.tri || gr.has_contact != orr.hasContact || gr.method != orr.method ||
                        gr.iters != orr.iters || memcmp(gr.normal, orr.normal, 12) || memcmp(gr.point, orr.point, 12) || memcmp(&gr.depth, &orr.depth, 4)) {
                        printf("step %d pair (%d,%d) kid %zu raw differs: has %d/%d method %d/%d iters %d/%d depth %g/%g\n", step, kv.first.first, kv.first.second,
                               k, gr.has_contact, orr.hasContact, gr.method, orr.method, gr.iters, orr.iters, gr.depth, orr.depth);
                        return 1;
                    }
                }
            }
        }
        if (rawSeen != oraw.size()) { printf("step %d: raw count %zu vs oracle %zu\n", step, rawSeen, oraw.size()); return 1; }
        if ((long)ctr.numManifolds != oracleKidManifolds) { printf("step %d: numManifolds %u vs oracle kid manifolds %ld\n", step, ctr.numManifolds, oracleKidManifolds); return 1; }
        if ((long)ctr.deepChecks > sc.W.deepPenetrationChecks - deepBefore) { printf("deep count\n"); return 1; }
        if (useMesh) printf("   mesh raw records %zu index1 check\n", meshRawSeen);
        printf("step %d ok: pairs %u compound items %u retry %u deep %u kid manifolds %ld contactsAdded %u\n", step, P, cc.numItems, ctr.epaRetry, ctr.deepChecks,
               oracleKidManifolds, ctr.contactsAdded);
    }
    printf("ALL OK items %ld touching points %ld retries %ld deep %ld keepItems %ld steps-with-large-pool-in-mesh-kernel %ld\n", totalItems, totalTouch, totalRetry, totalDeep,
           keepItems, bigUsed);
    return 0;
}
