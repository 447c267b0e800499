// Host-mirror demo used by tests/test_host_mirror.py: the reference's call sequences (disp/CollisionWorld.java:123-151)
// through include/b2c_host.hpp, exactly as INTEGRATION.md §4 lists them:
//   part 1  the BasicDemo-style drop-in sequence on a 3-box stack (hand-derivable contacts);
//   part 2  a moving scene stepped twice over — once with the drop-in sequence (set transforms, updateAabbs,
//           calculateOverlappingPairs + pair list + deltas, dispatchAllCollisionPairs + contact stream, islands), once with the
//           fast path (prefetched uid-keyed packed stream + pair deltas around b2c_step_device) — and compared step by step.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include "b2c_host.hpp"

using namespace b2c_host;

static void planesOf(const std::vector<Transform>& xf, std::vector<float>& planes) {
    const size_t n = xf.size();
    planes.resize(12 * n);
    for (size_t i = 0; i < n; i++) {
        for (int k = 0; k < 9; k++) planes[k * n + i] = xf[i].basis[k];
        for (int k = 0; k < 3; k++) planes[(9 + k) * n + i] = xf[i].origin[k];
    }
}

static int sequences() {
    b2c_config cfg;
    b2c_default_config(&cfg);
    cfg.max_bodies = 256; cfg.max_pairs = 4096;
    GpuCollisionWorld a(cfg), b(cfg);       // a: drop-in sequence, b: fast path
    std::vector<Transform> xf;
    for (GpuCollisionWorld* w : {&a, &b}) {
        int32_t ground = w->BoxShape({20, 1, 20}), box = w->BoxShape({0.5f, 0.5f, 0.5f}), ball = w->SphereShape(0.5f);
        xf.clear();
        Transform t = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, -1, 0}};
        w->addCollisionObject(ground, t, 2, (int16_t)(-1 ^ 2), true);
        xf.push_back(t);
        for (int i = 0; i < 120; i++) {
            Transform o = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0.95f * (float)(i % 8) - 3.f, 0.45f + 0.9f * (float)(i / 64), 0.95f * (float)((i / 8) % 8) - 3.f}};
            w->addCollisionObject((i % 3) ? box : ball, o);
            xf.push_back(o);
        }
    }
    b.enableFastPath();
    const int32_t n = (int32_t)xf.size();
    std::set<std::pair<int, int>> mirror;   // the host's pair-cache mirror, maintained from the fast path's deltas only
    std::vector<float> planes;
    for (int step = 0; step < 5; step++) {
        for (int i = 1; i < n; i++) {       // a small drift: pairs come and go
            xf[(size_t)i].origin[0] += 0.03f * (float)((i * 7 + step) % 5 - 2);
            xf[(size_t)i].origin[2] += 0.02f * (float)((i * 3 + step) % 7 - 3);
        }
        planesOf(xf, planes);
        // --- drop-in: three reference calls -> three ABI calls + downloads
        a.setWorldTransformPlanes(n, planes.data());
        a.updateAabbs();
        a.getBroadphase()->calculateOverlappingPairs(a.getDispatcher());
        const std::vector<BroadphasePair> pairs = a.getPairCache()->getOverlappingPairArray();
        std::vector<BroadphasePair> add, rem;
        a.getPairCache()->getPairDeltas(add, rem);
        a.getDispatcher()->dispatchAllCollisionPairs(a.getPairCache(), nullptr, a.getDispatcher());
        std::vector<b2c_contact_header> ch;
        std::vector<b2c_manifold_point> cp;
        a.getDispatcher()->getContacts(ch, cp);
        std::vector<int32_t> tags;
        const int islands = a.computeIslands(tags);
        // --- fast path
        GpuCollisionWorld::FastStep fs;
        b.stepFast(n, planes.data(), 4096, fs);
        for (auto& p : fs.removed) mirror.erase({p.proxy0, p.proxy1});
        for (auto& p : fs.added) mirror.insert({p.proxy0, p.proxy1});
        // --- the two must tell the same story
        if (fs.numPairs != (int32_t)pairs.size() || mirror.size() != pairs.size()) { std::printf("FAIL pairs step %d: %d %zu %zu\n", step, fs.numPairs, pairs.size(), mirror.size()); return 1; }
        for (auto& p : pairs) if (!mirror.count({p.proxy0, p.proxy1})) { std::printf("FAIL mirror step %d\n", step); return 1; }
        if (add.size() != fs.added.size() || rem.size() != fs.removed.size()) { std::printf("FAIL deltas step %d\n", step); return 1; }
        if (ch.size() != fs.headers.size() || cp.size() != fs.points.size()) { std::printf("FAIL contact counts step %d\n", step); return 1; }
        std::map<std::pair<int, int>, const b2c_contact_header*> byPair;
        for (auto& h : ch) byPair[{h.pair_uid0, h.pair_uid1}] = &h;
        int32_t first = 0;                  // uid-keyed stream: a manifold's first point is the running sum of num_contacts
        for (auto& h : fs.headers) {
            const int nc = h.info & 0xff;
            auto it = byPair.find({h.pair_uid0, h.pair_uid1});
            if (it == byPair.end() || it->second->num_contacts != nc) { std::printf("FAIL header step %d\n", step); return 1; }
            for (int k = 0; k < nc; k++) {
                const b2c_manifold_point& q = cp[(size_t)(it->second->first_point + k)];
                const b2c_packed_point& r = fs.points[(size_t)(first + k)];
                if (std::memcmp(q.world_a, r.world_a, 12) || std::memcmp(q.world_b, r.world_b, 12) || std::memcmp(q.normal_on_b, r.normal_on_b, 12) ||
                    std::memcmp(&q.distance, &r.distance, 4) || (r.life_src >> 8) != q.life_time || (r.life_src & 0xff) - 1 != q.src_slot) {
                    std::printf("FAIL point step %d\n", step);
                    return 1;
                }
            }
            first += nc;
        }
        std::printf("step %d pairs %zu +%zu -%zu manifolds %zu points %zu islands %d\n", step, pairs.size(), add.size(), rem.size(), ch.size(), cp.size(), islands);
    }
    Vector3 mn, mx;
    a.getBroadphase()->getBroadphaseAabb(mn, mx);
    std::printf("broadphase aabb %.0e %.0e\n", (double)mn.x, (double)mx.x);
    return 0;
}

int main() {
    if (b2c_device_count() < 1) { std::printf("NO_DEVICE\n"); return 3; }
    b2c_config cfg;
    b2c_default_config(&cfg);
    cfg.max_bodies = 64; cfg.max_pairs = 1024;
    {
        GpuCollisionWorld world(cfg);
        int32_t ground = world.BoxShape({50, 50, 50});
        int32_t box = world.BoxShape({1, 1, 1});
        Transform t = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, -50, 0}};
        world.addCollisionObject(ground, t, 2, (int16_t)(-1 ^ 2), true);
        for (int i = 0; i < 3; i++) {
            Transform b = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {0, 1.0f + 2.0f * i, 0}};
            world.addCollisionObject(box, b);
        }
        world.performDiscreteCollisionDetection();
        auto& pairs = world.getPairCache()->getOverlappingPairArray();
        std::printf("pairs %d manifolds %d\n", (int)pairs.size(), world.getDispatcher()->getNumManifolds());
        for (auto& p : pairs) std::printf("pair %d %d\n", p.proxy0, p.proxy1);
        const b2c_manifold& m = world.getDispatcher()->getManifoldByIndexInternal(0);
        std::printf("m0 bodies %d %d contacts %d normal %.3f %.3f %.3f depth %.6f\n", m.body0, m.body1, m.num_contacts,
                    m.points[0].normal_on_b[0], m.points[0].normal_on_b[1], m.points[0].normal_on_b[2], m.points[0].distance);
        std::vector<BroadphasePair> added, removed;
        world.getPairCache()->getPairDeltas(added, removed);
        std::vector<int32_t> tags;
        int islands = world.computeIslands(tags);
        std::printf("deltas +%d -%d islands %d tags %d %d %d %d\n", (int)added.size(), (int)removed.size(), islands, tags[0], tags[1], tags[2],
                    tags[3]);
        // the queries a character controller issues: one ray and one convex sweep (a small sphere) straight down onto the stack
        int32_t probe = world.SphereShape(0.25f);
        const float from[3] = {0.f, 10.f, 0.f}, to[3] = {0.f, -10.f, 0.f}, basis[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        int32_t ruid = 0, suid = 0;
        float rfrac = 0.f, sfrac = 0.f, rn[3], rp[3], sn[3], sp[3];
        world.rayTestClosest(1, from, to, &ruid, &rfrac, rn, rp);
        world.convexSweepTestClosest(1, &probe, basis, from, to, &suid, &sfrac, sn, sp);
        std::printf("ray uid %d y %.2f | sweep uid %d fraction %.3f normal %.2f %.2f %.2f\n", ruid, (double)rp[1], suid, (double)sfrac,
                    (double)sn[0], (double)sn[1], (double)sn[2]);
    }
    return sequences();
}
