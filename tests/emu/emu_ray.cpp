// TEST INFRASTRUCTURE ONLY — never built into, loaded by or shipped with the product (the package fails without a GPU).
// Host emulation of the ray kernels (raycast.cuh): the device headers are compiled
// for the CPU with tests/emu/cuda_runtime.h (a shim of the CUDA built-ins, one emulated lane per warp) and driven kernel by
// kernel against the oracle on a seeded scene; every record must be BIT-identical.  It checks the kernels' logic — item
// expansion and persistence, detector / manifold semantics, traversal orders — where no GPU is available; scheduling, atomics
// across threads and launch plumbing are only exercised by the -m gpu tests.  Run by tests/test_emu_kernels.py.
#include <cstdio>
#include <cstdlib>
#include <random>
#include <algorithm>
#include <vector>
#define B2C_RAY_THREADS 1
#define B2C_RAY_MAX_CAND 12    // small: rays that meet more boxes exercise the index-ordered tile path of k_ray_test
#define B2C_SWEEP_MAX_CAND 128   // small lists: the rounds of k_convex_sweep (several flushes per sweep) are exercised
#define B2C_SWEEP_HIT_CAP 3
#include "cuda_runtime.h"
#include "../../oracle/world.h"
#include "../../libgdx-jbullet_b200/csrc/compound.cuh"
#include "../../libgdx-jbullet_b200/csrc/raycast.cuh"
#include "../../libgdx-jbullet_b200/csrc/convexcast.cuh"
#include "../../libgdx-jbullet_b200/csrc/compound_flatten.h"
namespace b2c { alignas(16) unsigned char epaSmem[2 * sizeof(EpaScratch)]; }
using namespace b2c;
static std::mt19937 rng(777);
static float uf(float a, float b) { return std::uniform_real_distribution<float>(a, b)(rng); }
static void randRot(float m[9]) {
    float q[4] = {uf(-1, 1), uf(-1, 1), uf(-1, 1), uf(-1, 1)};
    float l = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) + 1e-6f;
    float x = q[0] / l, y = q[1] / l, z = q[2] / l, w = q[3] / l;
    m[0] = 1 - 2 * (y * y + z * z); m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
    m[3] = 2 * (x * y + w * z); m[4] = 1 - 2 * (x * x + z * z); m[5] = 2 * (y * z - w * x);
    m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = 1 - 2 * (x * x + y * y);
}
int main(int argc, char** argv) {
    const int NB = argc > 1 ? atoi(argv[1]) : 200, NR = argc > 2 ? atoi(argv[2]) : 3000;
    const bool tiltPlane = argc > 3;
    orc::World W; W.mode = orc::BP_TIGHT;
    std::vector<ShapeDev> shapes; std::vector<float4> hull; std::vector<CompoundChildDev> children; std::vector<MeshDev> meshes;
    auto box = [&](float x, float y, float z) { W.shapes.emplace_back(); W.meshes.emplace_back(nullptr); orc::initBox(W.shapes.back(), orc::V3(x, y, z));
        ShapeDev s{}; s.type = SH_BOX; s.margin = 0.04f; s.dims[0] = x - s.margin; s.dims[1] = y - s.margin; s.dims[2] = z - s.margin; shapes.push_back(s); return (int)shapes.size() - 1; };
    auto sphere = [&](float r) { W.shapes.emplace_back(); W.meshes.emplace_back(nullptr); orc::initSphere(W.shapes.back(), r);
        ShapeDev s{}; s.type = SH_SPHERE; s.dims[0] = r; s.margin = r; shapes.push_back(s); return (int)shapes.size() - 1; };
    auto hullS = [&](int n, float rad) {
        std::vector<float> pts(3 * n);
        for (int i = 0; i < n; i++) { float v[3] = {uf(-1, 1), uf(-1, 1), uf(-1, 1)}; float l = sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]) + 1e-3f; for (int c = 0; c < 3; c++) pts[3 * i + c] = v[c] / l * rad; }
        W.shapes.emplace_back(); W.meshes.emplace_back(nullptr); orc::initHull(W.shapes.back(), pts.data(), n);
        ShapeDev s{}; s.type = SH_HULL; s.margin = 0.04f; s.pointOffset = (int)hull.size(); s.numPoints = n;
        const orc::Shape& os = W.shapes.back();
        s.aabbMin[0] = os.localAabbMin.x; s.aabbMin[1] = os.localAabbMin.y; s.aabbMin[2] = os.localAabbMin.z;
        s.aabbMax[0] = os.localAabbMax.x; s.aabbMax[1] = os.localAabbMax.y; s.aabbMax[2] = os.localAabbMax.z;
        for (int i = 0; i < n; i++) hull.push_back(make_float4(pts[3 * i], pts[3 * i + 1], pts[3 * i + 2], 0));
        shapes.push_back(s); return (int)shapes.size() - 1; };
    auto plane = [&](float nx, float ny, float nz, float c) { W.shapes.emplace_back(); W.meshes.emplace_back(nullptr); orc::initPlane(W.shapes.back(), orc::V3(nx, ny, nz), c);
        ShapeDev s{}; s.type = SH_PLANE; s.plane[0] = W.shapes.back().planeNormal.x; s.plane[1] = W.shapes.back().planeNormal.y; s.plane[2] = W.shapes.back().planeNormal.z; s.plane[3] = c;
        shapes.push_back(s); return (int)shapes.size() - 1; };
    std::vector<std::vector<CompoundDirectChild>> directOfShape;
    auto compound = [&](const std::vector<int>& kids, const std::vector<float>& xf) { int sid = W.addCompound((int)kids.size(), kids.data(), xf.data());
        std::vector<CompoundDirectChild> direct(kids.size());
        for (size_t i = 0; i < kids.size(); i++) { direct[i].shape = kids[i]; for (int k = 0; k < 12; k++) direct[i].xf12[k] = xf[12 * i + k]; }
        int firstLeaf = 0, numLeaves = 0;
        flattenCompound(children, direct, [&](int sh) -> const std::vector<CompoundDirectChild>* { return (sh < (int)directOfShape.size() && !directOfShape[sh].empty()) ? &directOfShape[sh] : nullptr; }, firstLeaf, numLeaves);
        ShapeDev s{}; s.type = SH_COMPOUND; s.pointOffset = firstLeaf; s.numPoints = numLeaves;
        const orc::Shape& os = W.shapes[sid];
        s.aabbMin[0] = os.localAabbMin.x; s.aabbMin[1] = os.localAabbMin.y; s.aabbMin[2] = os.localAabbMin.z;
        s.aabbMax[0] = os.localAabbMax.x; s.aabbMax[1] = os.localAabbMax.y; s.aabbMax[2] = os.localAabbMax.z;
        shapes.push_back(s);
        if (directOfShape.size() < shapes.size()) directOfShape.resize(shapes.size());
        directOfShape[shapes.size() - 1] = direct;
        return (int)shapes.size() - 1; };
    // heightfield mesh 24x24 cells over [0,12]^2
    const int C = 24; std::vector<float> verts; std::vector<int> idx;
    for (int i = 0; i <= C; i++) for (int j = 0; j <= C; j++) { verts.push_back(i * 0.5f); verts.push_back(0.6f * sinf(i * 0.7f) * cosf(j * 0.5f) + uf(-0.05f, 0.05f)); verts.push_back(j * 0.5f); }
    for (int i = 0; i < C; i++) for (int j = 0; j < C; j++) { int v00 = i * (C + 1) + j, v10 = (i + 1) * (C + 1) + j, v01 = v00 + 1, v11 = v10 + 1; idx.insert(idx.end(), {v00, v01, v10, v10, v01, v11}); }
    int meshShape = W.addMesh(verts.data(), (int)verts.size() / 3, idx.data(), (int)idx.size() / 3);
    std::vector<int4> nodes(W.meshes[meshShape]->bvh.nodes.size());
    memcpy(nodes.data(), W.meshes[meshShape]->bvh.nodes.data(), nodes.size() * 16);
    { MeshDev md{}; md.nodes = nodes.data(); md.verts = verts.data(); md.idx = idx.data(); md.numNodes = (int)nodes.size(); md.numTris = (int)idx.size() / 3;
      const orc::Bvh& b = W.meshes[meshShape]->bvh;
      md.qmin[0] = b.bvhAabbMin.x; md.qmin[1] = b.bvhAabbMin.y; md.qmin[2] = b.bvhAabbMin.z; md.qmax[0] = b.bvhAabbMax.x; md.qmax[1] = b.bvhAabbMax.y; md.qmax[2] = b.bvhAabbMax.z;
      md.quant[0] = b.bvhQuantization.x; md.quant[1] = b.bvhQuantization.y; md.quant[2] = b.bvhQuantization.z; meshes.push_back(md);
      ShapeDev s{}; s.type = SH_MESH; s.mesh = 0; const orc::Shape& os = W.shapes[meshShape];
      s.aabbMin[0] = os.localAabbMin.x; s.aabbMin[1] = os.localAabbMin.y; s.aabbMin[2] = os.localAabbMin.z; s.aabbMax[0] = os.localAabbMax.x; s.aabbMax[1] = os.localAabbMax.y; s.aabbMax[2] = os.localAabbMax.z;
      shapes.push_back(s); }
    int pl = tiltPlane ? plane(0.2f, 1.0f, -0.1f, -1.5f) : plane(0, 1, 0, -1.0f);
    int sS = sphere(0.3f), sB = sphere(0.45f), bx = box(0.4f, 0.3f, 0.35f), bar = box(0.5f, 0.1f, 0.1f), hl = hullS(14, 0.4f);
    std::vector<float> x3 = {1, 0, 0, 0, 1, 0, 0, 0, 1, -0.5f, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1, 0.5f, 0, 0};
    int c1 = compound({sS, bar, sB}, x3);
    std::vector<float> x2(24); randRot(x2.data()); x2[9] = -0.2f; x2[10] = 0.1f; x2[11] = 0; randRot(x2.data() + 12); x2[21] = 0.3f; x2[22] = 0; x2[23] = 0.2f;
    int c2 = compound({hl, bx}, x2);
    // a compound whose children are compounds (rotated frames, two levels): rays and sweeps visit its leaves depth first
    std::vector<float> x4(24); randRot(x4.data()); x4[9] = 0.2f; x4[10] = 0.3f; x4[11] = 0; randRot(x4.data() + 12); x4[21] = -0.3f; x4[22] = -0.2f; x4[23] = 0.1f;
    int c3 = compound({c1, c2}, x4);
    std::vector<float> x5(24); randRot(x5.data()); x5[9] = 0; x5[10] = 0.4f; x5[11] = 0; randRot(x5.data() + 12); x5[21] = 0.1f; x5[22] = -0.5f; x5[23] = 0;
    int c4 = compound({c3, bar}, x5);
    std::vector<int> kinds = {sS, sB, bx, bar, hl, c1, c2, c3, c4};
    std::vector<int> bodyShape; std::vector<float4> xf4; std::vector<uint32_t> filt; std::vector<uint8_t> flags;
    auto addBody = [&](int shape, const float* t, int group, int mask) {
        orc::Xf x; for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) x.basis.m[r][c] = t[3 * r + c]; x.origin.set(t[9], t[10], t[11]);
        W.addBody(shape, x, group, mask, false, 0); bodyShape.push_back(shape);
        for (int r = 0; r < 3; r++) xf4.push_back(make_float4(t[3 * r], t[3 * r + 1], t[3 * r + 2], t[9 + r]));
        filt.push_back(((uint32_t)(uint16_t)group) | ((uint32_t)(uint16_t)mask << 16)); flags.push_back(BF_ALIVE | BF_ACTIVE); };
    { float t[12] = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0}; addBody(pl, t, 2, -1 ^ 2); }
    { float t[12]; randRot(t); if (!tiltPlane) { t[0] = 1; t[1] = 0; t[2] = 0; t[3] = 0; t[4] = 1; t[5] = 0; t[6] = 0; t[7] = 0; t[8] = 1; } t[9] = 0.25f; t[10] = 0.5f; t[11] = -0.5f; addBody(meshShape, t, 8, -1 ^ 2); }
    for (int i = 0; i < NB; i++) { float t[12]; randRot(t); t[9] = uf(0, 12); t[10] = uf(0.5f, 5); t[11] = uf(0, 12); addBody(kinds[rng() % kinds.size()], t, (i % 7 == 0) ? 4 : 1, -1); }
    hull.resize(hull.size() + 8);
    const int N = (int)bodyShape.size();
    BodyArrays B{}; B.xf4 = xf4.data(); B.shape = bodyShape.data(); B.filt = filt.data(); B.flags = flags.data();
    std::vector<float4> rmin(N), rmax(N);
    blockDim = {1, 1, 1}; gridDim = {1, 1, 1}; blockIdx = {0, 0, 0};
    for (int i = 0; i < N; i++) { blockIdx.x = i; threadIdx.x = 0; k_ray_aabbs(B, shapes.data(), N, rmin.data(), rmax.data()); }
    blockIdx.x = 0;
    // a random permutation of the first nSorted bodies stands in for the broadphase's sorted order
    const int nSorted = N - 37;
    std::vector<float4> sortedMin(N);
    { std::vector<int> perm(nSorted); for (int i = 0; i < nSorted; i++) perm[i] = i; std::shuffle(perm.begin(), perm.end(), rng);
      for (int i = 0; i < nSorted; i++) { sortedMin[i] = make_float4(0, 0, 0, 0); sortedMin[i].w = __uint_as_float((uint32_t)perm[i]); } }
    const int nChunks = (N + RAY_CHUNK - 1) / RAY_CHUNK;
    std::vector<float4> cmin(nChunks), cmax(nChunks);
    blockDim = {1, 1, 1};
    for (int c = 0; c < nChunks; c++) { blockIdx.x = c; threadIdx.x = 0; k_ray_chunks(rmin.data(), rmax.data(), sortedMin.data(), nSorted, N, cmin.data(), cmax.data()); }
    blockIdx.x = 0;
    std::vector<float> from(3 * NR), to(3 * NR);
    for (int r = 0; r < NR; r++) {
        for (int c = 0; c < 3; c++) { from[3 * r + c] = uf(-2, 14); to[3 * r + c] = uf(-2, 14); }
        from[3 * r + 1] = uf(-3, 8); to[3 * r + 1] = uf(-3, 8);
        if (r % 5 == 0) { to[3 * r] = from[3 * r]; to[3 * r + 2] = from[3 * r + 2]; from[3 * r + 1] = 8; to[3 * r + 1] = -4; }   // straight down
    }
    std::vector<RayOut> out(NR); uint32_t overflow = 0;
    const int group = 1, mask = argc > 4 ? (-1 ^ 4) : -1;
    const uint32_t cbFilter = ((uint32_t)(uint16_t)group) | ((uint32_t)(uint16_t)mask << 16);
    k_ray_test(B, shapes.data(), hull.data(), meshes.data(), children.data(), sortedMin.data(), nSorted, cmin.data(), cmax.data(), N, rmin.data(), rmax.data(), from.data(), to.data(), NR, cbFilter, out.data(), &overflow);
    int hits = 0, hitMesh = 0, hitPlane = 0, hitComp = 0;
    for (int r = 0; r < NR; r++) {
        orc::RayHit h = W.rayTestClosest(orc::V3(from[3 * r], from[3 * r + 1], from[3 * r + 2]), orc::V3(to[3 * r], to[3 * r + 1], to[3 * r + 2]), group, mask);
        const RayOut& g = out[r];
        float of[7] = {h.fraction, h.normal.x, h.normal.y, h.normal.z, h.point.x, h.point.y, h.point.z};
        float gf[7] = {g.fraction, g.normal[0], g.normal[1], g.normal[2], g.point[0], g.point[1], g.point[2]};
        if (g.uid != h.uid || (h.uid && memcmp(of, gf, 28)) || (!h.uid && g.fraction != 1.f)) {
            printf("ray %d differs: uid %d/%d frac %.9g/%.9g n (%g %g %g)/(%g %g %g)\n", r, g.uid, h.uid, g.fraction, h.fraction, g.normal[0], g.normal[1], g.normal[2], h.normal.x, h.normal.y, h.normal.z);
            return 1;
        }
        if (h.uid) { hits++; int t = shapes[bodyShape[h.uid - 1]].type; hitMesh += t == SH_MESH; hitPlane += t == SH_PLANE; hitComp += t == SH_COMPOUND; }
    }
    // ---- convex sweeps (convexcast.cuh): sphere / box / hull casts with random bases; the callback mask leaves the static
    // plane out (group 2) except for every 50th sweep, which must report the reference's unsupported branch (uid -1)
    const int NS = NR / 4;
    int sHits = 0, sMesh = 0, sComp = 0, sUnsup = 0;
    const int castKinds[3] = {sS, bx, hl};
    for (int pass = 0; pass < 2; pass++) {
        const int smask = pass == 0 ? (-1 ^ 2) : -1;
        const int ns = pass == 0 ? NS : NS / 50 + 1;
        std::vector<int> castShape(ns); std::vector<float> basis(9 * ns), sf(3 * ns), st(3 * ns);
        for (int r = 0; r < ns; r++) {
            castShape[r] = castKinds[rng() % 3];
            randRot(&basis[9 * r]);
            if (r % 3 == 0) { float I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}; memcpy(&basis[9 * r], I, 36); }
            for (int c = 0; c < 3; c++) { sf[3 * r + c] = uf(-1, 13); st[3 * r + c] = sf[3 * r + c] + uf(-4, 4); }
            sf[3 * r + 1] = uf(0, 7); st[3 * r + 1] = sf[3 * r + 1] + uf(-6, 2);
        }
        std::vector<RayOut> so(ns); uint32_t sov = 0;
        const uint32_t sFilter = ((uint32_t)(uint16_t)group) | ((uint32_t)(uint16_t)smask << 16);
        k_convex_sweep(B, shapes.data(), hull.data(), meshes.data(), children.data(), sortedMin.data(), nSorted, cmin.data(), cmax.data(), N, rmin.data(), rmax.data(), castShape.data(), basis.data(), sf.data(), st.data(), ns, sFilter, 0.04f, so.data(), &sov, SweepNotMe{});
        if (pass == 0 && sov <= 64) { printf("no sweep had more candidates (%u) than fit beside one more chunk (a mid-sweep flush)\n", sov); return 1; }
        for (int r = 0; r < ns; r++) {
            orc::Xf f, t;
            for (int a = 0; a < 3; a++) for (int c = 0; c < 3; c++) f.basis.m[a][c] = t.basis.m[a][c] = basis[9 * r + 3 * a + c];
            f.origin.set(sf[3 * r], sf[3 * r + 1], sf[3 * r + 2]); t.origin.set(st[3 * r], st[3 * r + 1], st[3 * r + 2]);
            orc::ConvexSweepHit h = W.convexSweepClosest(castShape[r], f, t, group, smask, 0.04f);
            const RayOut& g = so[r];
            const int huid = h.unsupported ? -1 : h.uid;
            float of[7] = {h.fraction, h.normal.x, h.normal.y, h.normal.z, h.point.x, h.point.y, h.point.z};
            float gf[7] = {g.fraction, g.normal[0], g.normal[1], g.normal[2], g.point[0], g.point[1], g.point[2]};
            if (g.uid != huid || (huid > 0 && memcmp(of, gf, 28)) || (huid == 0 && g.fraction != 1.f)) {
                printf("sweep %d/%d differs: uid %d/%d frac %.9g/%.9g n (%g %g %g)/(%g %g %g) p (%g %g %g)/(%g %g %g)\n", pass, r, g.uid, huid, g.fraction, h.fraction, g.normal[0], g.normal[1], g.normal[2], h.normal.x, h.normal.y, h.normal.z, g.point[0], g.point[1], g.point[2], h.point.x, h.point.y, h.point.z);
                return 1;
            }
            if (huid > 0) { sHits++; int ty = shapes[bodyShape[huid - 1]].type; sMesh += ty == SH_MESH; sComp += ty == SH_COMPOUND; }
            sUnsup += huid < 0;
        }
        if (pass == 1 && sUnsup == 0) { printf("no sweep reached the static plane branch\n"); return 1; }
    }
    if (overflow <= 12) { printf("no ray met more boxes (%u) than one round holds: the tile path was not exercised\n", overflow); return 1; }
    // ---- CCD motion-clamping sweeps (ClosestNotMeConvexResultCallback): a sphere swept from every 3rd dynamic body's own
    // transform; the plane is left out by giving those bodies a mask without group 2 on both sides
    int cHits = 0;
    {
        std::vector<int> me; std::vector<float> rad, cto;
        for (int i = 2; i < N; i += 3) {
            me.push_back(i); rad.push_back(uf(0.1f, 0.35f));
            for (int c = 0; c < 3; c++) cto.push_back(xf4[3 * i + c].w + uf(-2.5f, 2.5f));
        }
        const int nc = (int)me.size();
        std::vector<uint32_t> filtSave = filt;
        for (int i : me) { filt[i] = (filt[i] & 0xffffu) | ((uint32_t)(uint16_t)(-1 ^ 2) << 16); W.bodies[i].mask = (short)(-1 ^ 2); }
        B.filt = filt.data();
        std::vector<RayOut> co(nc); uint32_t cov = 0;
        SweepNotMe nm{}; nm.me = me.data(); nm.radius = rad.data();
        k_convex_sweep(B, shapes.data(), hull.data(), meshes.data(), children.data(), sortedMin.data(), nSorted, cmin.data(), cmax.data(), N, rmin.data(), rmax.data(), nullptr, nullptr, nullptr, cto.data(), nc, 0u, 0.04f, co.data(), &cov, nm);
        for (int r = 0; r < nc; r++) {
            orc::ConvexSweepHit h = W.ccdSweepNotMe(me[r] + 1, rad[r], orc::V3(cto[3 * r], cto[3 * r + 1], cto[3 * r + 2]), 0.04f);
            const RayOut& g = co[r];
            const int huid = h.unsupported ? -1 : h.uid;
            float of[7] = {h.fraction, h.normal.x, h.normal.y, h.normal.z, h.point.x, h.point.y, h.point.z};
            float gf[7] = {g.fraction, g.normal[0], g.normal[1], g.normal[2], g.point[0], g.point[1], g.point[2]};
            if (g.uid != huid || (huid > 0 && memcmp(of, gf, 28)) || (huid == 0 && g.fraction != 1.f) || huid == me[r] + 1) {
                printf("ccd sweep %d (body %d) differs: uid %d/%d frac %.9g/%.9g\n", r, me[r] + 1, g.uid, huid, g.fraction, h.fraction);
                return 1;
            }
            cHits += huid > 0;
        }
        if (cHits < nc / 20) { printf("too few CCD sweeps hit anything (%d of %d)\n", cHits, nc); return 1; }
        printf("ccd sweeps %d hits %d\n", nc, cHits);
    }
    printf("ALL OK rays %d hits %d mesh %d plane %d compound %d overflow %u | sweeps %d hits %d mesh %d compound %d unsupported %d\n", NR, hits, hitMesh, hitPlane, hitComp, overflow, NS, sHits, sMesh, sComp, sUnsup);
    return 0;
}
