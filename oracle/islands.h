// ORACLE (test infrastructure only; never linked into the product) — "parity unpinned": the reference ships no tests.
//
// islands.h — CPU restatement of the island labelling that directly follows the collision path:
//   disp/UnionFind.java:77-140 (reset, unite without weighting, find with path halving)
//   disp/SimulationIslandManager.java:57-70 findUnions (every BROADPHASE pair whose two objects merge islands),
//   :72-90 updateActivationState (tag = object index), :92-110 storeIslandActivationState (tag = find(i), -1 for
//   static/kinematic objects).
// The root a component ends up with depends on the pair order (the reference's is insertion order, which is not
// reproducible — SURVEY Q2), so callers compare PARTITIONS, not tag values.
#pragma once
#include <utility>
#include <vector>

namespace orc {

struct UnionFindJ {
    std::vector<int> id, sz;
    void reset(int n) {  // disp/UnionFind.java:77-84
        id.resize(n);
        sz.assign(n, 1);
        for (int i = 0; i < n; i++) id[i] = i;
    }
    int find(int x) {  // :126-140
        while (x != id[x]) {
            id[x] = id[id[x]];
            x = id[x];
        }
        return x;
    }
    void unite(int p, int q) {  // :104-124 (USE_PATH_COMPRESSION branch)
        int i = find(p), j = find(q);
        if (i == j) return;
        id[i] = j;
        sz[j] += sz[i];
    }
};

// pairs: (uid0, uid1) with uid = object index + 1; isStatic[i] for object i.  tagsOut[i] = island tag or -1.
inline int islandTags(const std::vector<std::pair<int, int>>& pairs, const std::vector<char>& mergesIslands, int* tagsOut) {
    const int n = (int)mergesIslands.size();
    UnionFindJ uf;
    uf.reset(n);
    for (auto& pr : pairs) {
        int a = pr.first - 1, b = pr.second - 1;
        if (mergesIslands[a] && mergesIslands[b]) uf.unite(a, b);  // disp/SimulationIslandManager.java:64-67
    }
    int islands = 0;
    for (int i = 0; i < n; i++) {
        if (mergesIslands[i]) {
            tagsOut[i] = uf.find(i);
            if (tagsOut[i] == i) islands++;
        } else {
            tagsOut[i] = -1;
        }
    }
    return islands;
}

}  // namespace orc
