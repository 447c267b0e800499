// compound_flatten.h — host side of the compound child table (common.cuh CompoundChildDev): a CompoundShape whose children
// may themselves be CompoundShapes (sh/CompoundShape.java:50-82 accepts any CollisionShape) becomes a run of FRAME entries
// (one per nested compound occurrence: its local transform, chained to the frame above it) followed by its LEAVES in the
// depth-first order in which disp/CompoundCollisionAlgorithm.java:57-75,100-125 creates and runs the child algorithms.
// Plain C++ (no CUDA): used by b2c_api.cu at registration and by tests/emu/ to build the same table.
#pragma once
#include <functional>
#include <vector>

namespace b2c {

struct CompoundDirectChild {   // one addChildShape(localTransform, shape) call
    int shape;
    float xf12[12];            // 9 row-major basis floats + origin
};

// directOf(shape) = the direct children of a registered compound shape, or nullptr for a leaf shape.
// Appends to `table`; firstLeaf / numLeaves describe the new compound.  Returns false when the nesting is deeper than
// COMPOUND_MAX_DEPTH frames.
inline bool flattenCompound(std::vector<CompoundChildDev>& table, const std::vector<CompoundDirectChild>& direct,
                            const std::function<const std::vector<CompoundDirectChild>*(int)>& directOf, int& firstLeaf, int& numLeaves) {
    struct Tmp { CompoundDirectChild c; int parentFrame; };   // parentFrame: index into `frames`, -1 = none
    std::vector<Tmp> frames, leaves;
    bool ok = true;
    std::function<void(const std::vector<CompoundDirectChild>&, int, int)> walk = [&](const std::vector<CompoundDirectChild>& kids, int parent,
                                                                                    int depth) {
        for (const CompoundDirectChild& k : kids) {
            const std::vector<CompoundDirectChild>* sub = directOf(k.shape);
            if (!sub) {
                leaves.push_back(Tmp{k, parent});
            } else {
                if (depth >= COMPOUND_MAX_DEPTH) { ok = false; return; }
                frames.push_back(Tmp{k, parent});
                walk(*sub, (int)frames.size() - 1, depth + 1);
                if (!ok) return;
            }
        }
    };
    walk(direct, -1, 0);
    if (!ok) return false;
    const int base = (int)table.size();
    auto put = [&](const Tmp& t, int shape) {
        CompoundChildDev ch{};
        for (int k = 0; k < 9; k++) ch.m[k] = t.c.xf12[k];
        for (int k = 0; k < 3; k++) ch.o[k] = t.c.xf12[9 + k];
        ch.shape = shape;
        ch.parent1 = t.parentFrame < 0 ? 0 : base + t.parentFrame + 1;
        table.push_back(ch);
    };
    for (const Tmp& f : frames) put(f, -1);
    firstLeaf = (int)table.size();
    for (const Tmp& l : leaves) put(l, l.c.shape);
    numLeaves = (int)leaves.size();
    return true;
}

}  // namespace b2c
