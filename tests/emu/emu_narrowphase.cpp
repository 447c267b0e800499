// TEST INFRASTRUCTURE ONLY — never built into, loaded by or shipped with the product (the package fails without a GPU).
// Host emulation of the narrowphase kernels of narrowphase.cuh (k_carry, k_sphere_sphere, k_convex_plane, k_gjk, k_mesh_query,
// k_gjk_tri, k_epa<2>/<1>, k_manifold_cc, k_mesh_manifold): compiled for the CPU with tests/emu/cuda_runtime.h (one emulated
// lane per warp) and driven against the oracle on a seeded scene of boxes, spheres and hulls over a static plane and a
// triangle mesh, with bodies going to sleep and waking up.  Raw detector records (method and iteration count included),
// manifold headers and points must be BIT-identical.  The classification / binning kernels need a full thread block and are
// restated on the host here; the prefilter is skipped (k_gjk runs every convex-convex pair from scratch, which is what it
// does for survivors anyway).  Run by tests/test_emu_kernels.py.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>
#include "cuda_runtime.h"
#include "../../oracle/world.h"
#include "../../libgdx-jbullet_b200/csrc/compound.cuh"  // pulls in broadphase.cuh + narrowphase.cuh

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
    int addCompound(const std::vector<int>& kids, const std::vector<float>& xf12) {
        int sid = W.addCompound((int)kids.size(), kids.data(), xf12.data());
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
        s.pointOffset = first; s.numPoints = (int)kids.size();
        shapes.push_back(s);
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
    const int N = argc > 1 ? atoi(argv[1]) : 200;
    const float SP = argc > 2 ? atof(argv[2]) : 0.8f;
    const int STEPS = 8;
    Sc sc;
    sc.W.mode = orc::BP_DBVT;
    int plane = sc.addPlane(0, 1, 0, -0.3f);
    // heightfield mesh under the lattice
    const int C = 24; static std::vector<float> verts; static std::vector<int> idx; static std::vector<int4> nodes; std::vector<MeshDev> meshes;
    for (int i = 0; i <= C; i++) for (int j = 0; j <= C; j++) { verts.push_back(i * 0.5f - 2.f); verts.push_back(0.25f * sinf(i * 0.9f) * cosf(j * 0.7f)); verts.push_back(j * 0.5f - 2.f); }
    for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) { int v00 = i * (C + 1) + j, v10 = (i + 1) * (C + 1) + j, v01 = v00 + 1, v11 = v10 + 1; idx.insert(idx.end(), {v00, v01, v10, v10, v01, v11}); }
    int meshShape = sc.W.addMesh(verts.data(), (int)verts.size() / 3, idx.data(), (int)idx.size() / 3);
    nodes.resize(sc.W.meshes[meshShape]->bvh.nodes.size());
    memcpy(nodes.data(), sc.W.meshes[meshShape]->bvh.nodes.data(), nodes.size() * 16);
    {
        MeshDev md{}; md.nodes = nodes.data(); md.verts = verts.data(); md.idx = idx.data(); md.numNodes = (int)nodes.size(); md.numTris = (int)idx.size() / 3;
        const orc::Bvh& b = sc.W.meshes[meshShape]->bvh;
        md.qmin[0] = b.bvhAabbMin.x; md.qmin[1] = b.bvhAabbMin.y; md.qmin[2] = b.bvhAabbMin.z; md.qmax[0] = b.bvhAabbMax.x; md.qmax[1] = b.bvhAabbMax.y; md.qmax[2] = b.bvhAabbMax.z;
        md.quant[0] = b.bvhQuantization.x; md.quant[1] = b.bvhQuantization.y; md.quant[2] = b.bvhQuantization.z; meshes.push_back(md);
        ShapeDev s{}; s.type = SH_MESH; s.mesh = 0; const orc::Shape& os = sc.W.shapes[meshShape];
        s.aabbMin[0] = os.localAabbMin.x; s.aabbMin[1] = os.localAabbMin.y; s.aabbMin[2] = os.localAabbMin.z; s.aabbMax[0] = os.localAabbMax.x; s.aabbMax[1] = os.localAabbMax.y; s.aabbMax[2] = os.localAabbMax.z;
        sc.shapes.push_back(s);
        if ((int)sc.shapes.size() - 1 != meshShape) { printf("shape index mismatch\n"); return 1; }
    }
    std::vector<int> plain = {sc.addSphere(0.3f), sc.addSphere(0.4f), sc.addHull(16, 0.35f), sc.addHull(12, 0.4f), sc.addBox(0.3f, 0.35f, 0.4f), sc.addBox(0.45f, 0.15f, 0.2f)};
    std::vector<int> bodyShape;
    std::vector<float> base, vel;
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
    { auto g = idXf(0, 0.1f, 0); addBody(meshShape, g.data(), true); }
    int m = (int)ceil(cbrt((double)N));
    for (int i = 0; i < N; i++) {
        float xf[12];
        randRot(xf);
        xf[9] = (i % m) * SP + uf(-0.12f, 0.12f);
        xf[10] = (i / (m * m)) * SP * 0.8f + 0.45f + uf(-0.12f, 0.12f);
        xf[11] = ((i / m) % m) * SP + uf(-0.12f, 0.12f);
        addBody(plain[rng() % plain.size()], xf, false);
    }
    const int NB = (int)bodyShape.size();
    sc.hull.resize(sc.hull.size() + 8);
    const int uidBits = 12;
    std::vector<float4> xf4(3 * NB);
    std::vector<float2> material(NB, make_float2(0.5f, 0.0f));
    std::vector<uint8_t> flags(NB);
    const uint32_t MAXP = 1 << 16, MAXI = 1 << 17;
    std::vector<ManifoldHdr> H[2] = {std::vector<ManifoldHdr>(MAXP), std::vector<ManifoldHdr>(MAXP)};
    std::vector<b2c_manifold_point> P[2] = {std::vector<b2c_manifold_point>(4 * MAXP), std::vector<b2c_manifold_point>(4 * MAXP)};
    std::vector<uint64_t> keys[2] = {std::vector<uint64_t>(MAXP), std::vector<uint64_t>(MAXP)};
    std::vector<uint32_t> first[2] = {std::vector<uint32_t>(NB + 4), std::vector<uint32_t>(NB + 4)};
    uint32_t numPairs[2] = {0, 0};
    int cur = 0;
    std::vector<b2c_raw_contact> raw(MAXP), rawMesh(MAXI);
    std::vector<int8_t> rawFlag(MAXP);
    std::vector<uint8_t> hist(MAXP), binOf(MAXP);
    std::vector<uint32_t> binItems(MAXP), meshPair(MAXI), meshStart(MAXP), meshCount(MAXP), epaRetry(MAXI), epaBig(MAXI);
    std::vector<int> meshTri(MAXI);
    std::vector<EpaItem> epaItems(MAXI);
    StepCounters ctr{};
    long totRaw = 0, totPts = 0, totDeep = 0, totMesh = 0, totRetry = 0, totBig = 0;
    for (int step = 0; step < STEPS; step++) {
        for (int b = 0; b < NB; b++) {
            orc::Body& B = sc.W.bodies[b];
            for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) B.xf.basis.m[r][c] = base[12 * b + 3 * r + c];
            B.xf.origin.set(base[12 * b + 9] + vel[3 * b] * step, base[12 * b + 10] + vel[3 * b + 1] * step, base[12 * b + 11] + vel[3 * b + 2] * step);
            B.active = (step < 3 || step == 7) ? true : (uf(0, 1) > 0.4f);
            flags[b] = (uint8_t)(BF_ALIVE | (B.active ? BF_ACTIVE : 0) | (B.isStatic ? BF_STATIC : 0));
            for (int r = 0; r < 3; r++)
                xf4[3 * b + r] = make_float4(B.xf.basis.m[r][0], B.xf.basis.m[r][1], B.xf.basis.m[r][2], r == 0 ? B.xf.origin.x : (r == 1 ? B.xf.origin.y : B.xf.origin.z));
        }
        sc.W.updateAabbs();
        sc.W.calculateOverlappingPairs();
        sc.W.dispatchAllPairs();
        // ---- emulated device side: pair list + row table of this step
        cur ^= 1;
        const uint32_t Pn = (uint32_t)sc.W.pairs.size();
        if (Pn > MAXP) { printf("too many pairs\n"); return 1; }
        std::vector<int2> pairs(Pn);
        for (uint32_t p = 0; p < Pn; p++) {
            pairs[p] = make_int2(sc.W.pairs[p].first, sc.W.pairs[p].second);
            keys[cur][p] = ((uint64_t)(uint32_t)pairs[p].x << uidBits) | (uint32_t)pairs[p].y;
        }
        { uint32_t p = 0; for (int u = 0; u < NB + 4; u++) { while (p < Pn && pairs[p].x < u) p++; first[cur][u] = p; } }
        numPairs[cur] = Pn;
        ctr = StepCounters{};
        blockIdx = {0, 0, 0}; threadIdx = {0, 0, 0}; blockDim = {1, 1, 1}; gridDim = {1, 1, 1};
        k_carry(keys[cur].data(), &numPairs[cur], keys[cur ^ 1].data(), &numPairs[cur ^ 1], first[cur ^ 1].data(), H[cur ^ 1].data(), P[cur ^ 1].data(),
                H[cur].data(), P[cur].data(), uidBits, &ctr, hist.data(), first[cur].data());
        // classification and stable binning, restated from k_classify / k_partition16
        std::vector<std::vector<uint32_t>> bins(16);
        for (uint32_t p = 0; p < Pn; p++) {
            const int b0 = pairs[p].x - 1, b1 = pairs[p].y - 1;
            int bin = BIN_SKIP;
            if ((flags[b0] & BF_ACTIVE) || (flags[b1] & BF_ACTIVE)) {
                const int t0 = sc.shapes[bodyShape[b0]].type, t1 = sc.shapes[bodyShape[b1]].type;
                if (t0 == SH_SPHERE && t1 == SH_SPHERE) bin = BIN_SS;
                else if ((isConvexType(t0) && t1 == SH_PLANE) || (isConvexType(t1) && t0 == SH_PLANE)) bin = BIN_CP;
                else if (isConvexType(t0) && isConvexType(t1)) bin = (hist[p] >= 2 ? BIN_PS0 : BIN_GJK0) + (t0 == SH_HULL ? 2 : 0) + (t1 == SH_HULL ? 1 : 0);
                else if ((isConvexType(t0) && t1 == SH_MESH) || (isConvexType(t1) && t0 == SH_MESH)) bin = BIN_MESH;
            }
            binOf[p] = (uint8_t)bin;
            bins[bin].push_back(p);
        }
        uint32_t binStart[17];
        { uint32_t o = 0; for (int b = 0; b < 16; b++) { binStart[b] = o; for (uint32_t p : bins[b]) binItems[o++] = p; } binStart[16] = o; }
        NpArgs a{};
        uint32_t nP = Pn;
        a.pairs = pairs.data(); a.numPairs = &nP; a.xf4 = xf4.data(); a.shape = bodyShape.data(); a.flags = flags.data(); a.material = material.data();
        a.shapes = sc.shapes.data(); a.hullPts = sc.hull.data(); a.meshes = meshes.data(); a.mhdr = H[cur].data(); a.mpts = P[cur].data();
        a.raw = raw.data(); a.rawFlag = rawFlag.data(); a.wantRaw = 1; /* the harness compares every raw detector record */ a.hist = hist.data(); a.binOf = binOf.data(); a.binItems = binItems.data(); a.binStart = binStart;
        a.ctr = &ctr; a.threshold = 0.02f; a.maxPairs = MAXP; a.uidBits = uidBits;
        GjkArgs g{};
        g.epaItems = epaItems.data(); g.maxEpa = MAXI; g.epaRetry = epaRetry.data(); g.maxEpaRetry = MAXI; g.epaBig = epaBig.data();
        g.meshPair = meshPair.data(); g.meshTri = meshTri.data(); g.rawMesh = rawMesh.data(); g.meshStart = meshStart.data(); g.meshCount = meshCount.data();
        g.maxMeshItems = MAXI;
        k_sphere_sphere(a);
        k_convex_plane(a);
        // every convex-convex pair goes to k_gjk as a survivor without history (the prefilter only decides early, never differently)
        std::vector<uint32_t> survivors(binItems.begin() + binStart[BIN_GJK0], binItems.begin() + binStart[BIN_COUNT]);
        uint32_t survCount = (uint32_t)survivors.size(), cursor = 0, cursorTri = 0;
        uint32_t survStart[17] = {0};
        for (int b = 15; b <= 16; b++) survStart[b] = b == 16 ? survCount : 0;
        k_gjk(a, g, &cursor, survivors.data(), &survCount, survStart);
        k_mesh_query(a, g);
        k_gjk_tri(a, g, &cursorTri);
        k_epa<2>(a, g, 1, 32);
        blockDim = {32, 1, 1};
        for (unsigned t = 0; t < 32; t++) { threadIdx.x = t; k_epa<1>(a, g, 1, 32); }   // items routed straight to the large pools
        for (unsigned t = 0; t < 32; t++) { threadIdx.x = t; k_epa<1>(a, g, 0, 32); }
        threadIdx.x = 0;
        blockDim = {1, 1, 1};
        k_manifold_cc(a);
        k_mesh_manifold(a, g);
        totDeep += ctr.deepChecks; totMesh += ctr.meshItems; totRetry += ctr.epaRetry; totBig += ctr.epaBig;
        // ---- compare raw records
        std::map<std::tuple<int, int, int>, const orc::RawContact*> oraw;
        for (auto& r : sc.W.raw) oraw[std::make_tuple(r.uid0, r.uid1, r.tri)] = &r;
        size_t seen = 0;
        auto cmpRaw = [&](const b2c_raw_contact& gr) -> bool {
            auto it = oraw.find(std::make_tuple(gr.uid0, gr.uid1, gr.tri));
            if (it == oraw.end()) { printf("step %d: raw (%d,%d,%d) missing in the oracle\n", step, gr.uid0, gr.uid1, gr.tri); return false; }
            const orc::RawContact& o = *it->second;
            seen++;
            if (gr.has_contact != o.hasContact || gr.method != o.method || gr.iters != o.iters || memcmp(gr.normal, o.normal, 12) || memcmp(gr.point, o.point, 12) ||
                memcmp(&gr.depth, &o.depth, 4)) {
                printf("step %d raw (%d,%d,%d) differs: has %d/%d method %d/%d iters %d/%d depth %.9g/%.9g\n", step, gr.uid0, gr.uid1, gr.tri, gr.has_contact, o.hasContact,
                       gr.method, o.method, gr.iters, o.iters, gr.depth, o.depth);
                return false;
            }
            return true;
        };
        for (uint32_t p = 0; p < Pn; p++) {
            if (binOf[p] == BIN_SKIP) continue;
            if (binOf[p] == BIN_MESH) { for (uint32_t k = meshStart[p]; k < meshStart[p] + meshCount[p]; k++) if (!cmpRaw(rawMesh[k])) return 1; }
            else if (!cmpRaw(raw[p])) return 1;
        }
        if (seen != oraw.size()) { printf("step %d: raw count %zu vs oracle %zu\n", step, seen, oraw.size()); return 1; }
        totRaw += (long)seen;
        // ---- compare manifolds
        long oracleManifolds = 0;
        for (uint32_t p = 0; p < Pn; p++) {
            auto it = sc.W.pairState.find(sc.W.pairs[p]);
            const bool oHas = it != sc.W.pairState.end() && it->second.hasManifold;
            const ManifoldHdr& gh = H[cur][p];
            if ((gh.algorithm != 0) != oHas) { printf("step %d pair (%d,%d): manifold existence %d/%d\n", step, pairs[p].x, pairs[p].y, gh.algorithm, (int)oHas); return 1; }
            if (!oHas) continue;
            oracleManifolds++;
            const orc::PersistentManifold& om = it->second.manifold;
            if (gh.body0 != om.body0 || gh.body1 != om.body1 || gh.num_contacts != om.cachedPoints) {
                printf("step %d pair (%d,%d) header differs: b(%d,%d) n %d | b(%d,%d) n %d\n", step, pairs[p].x, pairs[p].y, gh.body0, gh.body1, gh.num_contacts, om.body0,
                       om.body1, om.cachedPoints);
                return 1;
            }
            for (int q = 0; q < om.cachedPoints; q++) {
                const orc::ManifoldPoint& op = om.pointCache[q];
                const b2c_manifold_point& gp = P[cur][4 * (size_t)p + q];
                float of[16] = {op.localPointA.x, op.localPointA.y, op.localPointA.z, op.localPointB.x, op.localPointB.y, op.localPointB.z,
                                op.positionWorldOnA.x, op.positionWorldOnA.y, op.positionWorldOnA.z, op.positionWorldOnB.x, op.positionWorldOnB.y,
                                op.positionWorldOnB.z, op.normalWorldOnB.x, op.normalWorldOnB.y, op.normalWorldOnB.z, op.distance1};
                if (memcmp(of, &gp, 64) || gp.life_time != op.lifeTime || gp.src_slot != op.srcSlot || gp.index1 != op.index1 ||
                    gp.combined_friction != op.combinedFriction || gp.combined_restitution != op.combinedRestitution) {
                    printf("step %d pair (%d,%d) point %d differs (life %d/%d src %d/%d idx %d/%d dist %.9g/%.9g)\n", step, pairs[p].x, pairs[p].y, q, gp.life_time,
                           op.lifeTime, gp.src_slot, op.srcSlot, gp.index1, op.index1, gp.distance, op.distance1);
                    return 1;
                }
                totPts++;
            }
        }
        if ((long)ctr.numManifolds != oracleManifolds) { printf("step %d: numManifolds %u vs oracle %ld\n", step, ctr.numManifolds, oracleManifolds); return 1; }
        printf("step %d ok: pairs %u raw %zu manifolds %ld deep %u mesh items %u contactsAdded %u\n", step, Pn, seen, oracleManifolds, ctr.deepChecks, ctr.meshItems, ctr.contactsAdded);
    }
    printf("ALL OK raw %ld points %ld deep %ld mesh items %ld retries %ld routed-to-large-pools %ld\n", totRaw, totPts, totDeep, totMesh, totRetry, totBig);
    return 0;
}
