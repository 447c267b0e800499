// Host-mirror demo used by tests/test_host_mirror.py: the BasicDemo-style call sequence of the reference
// (disp/CollisionWorld.java:123-151) through include/b2c_host.hpp.  Prints pairs / manifolds / first contact.
#include <cstdio>
#include "b2c_host.hpp"

using namespace b2c_host;

int main() {
    if (b2c_device_count() < 1) { std::printf("NO_DEVICE\n"); return 3; }
    b2c_config cfg;
    b2c_default_config(&cfg);
    cfg.max_bodies = 64; cfg.max_pairs = 1024;
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
    return 0;
}
