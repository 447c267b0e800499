// ORACLE — TEST INFRASTRUCTURE ONLY (see jmath.h header).  PARITY UNPINNED.
// dbvt_literal.h — a literal restatement of the reference's dynamic-AABB-tree broadphase, tree and all:
//   bp/Dbvt.java            insert/update/remove (:116-197), insertleaf (:603-646), removeleaf (:648-681),
//                           collideTT (:247-282), optimizeIncremental (:94-115), sort (:794-826)
//   bp/DbvtBroadphase.java  collide (:89-150), createProxy (:173-182), destroyProxy (:184-194), setAabb (:196-228)
//   bp/DbvtTreeCollider.java:41-56, bp/DbvtAabbMm.java (Merge, Proximity, Contain, Expand, SignedExpand, Intersect)
//   bp/HashedOverlappingPairCache.java (as a hash set; the reference's O(P) indexOf removal, SURVEY Q5, is not kept)
// Purpose: (1) differential pin of the claim the fast oracle and the CUDA path rely on — after collide() the
// pair cache equals {filtered pairs whose effective AABBs intersect}; (2) the
// "reference-algorithm" CPU timing of bench.py.  Java's identity hashCode ordering is replaced by
// node/proxy creation order; it only influences tree shape and insertion order, never the pair set.
#pragma once
#include <algorithm>
#include <cstdint>
#include <unordered_map>
#include <vector>
#include "jmath.h"

namespace orc {

struct LAabb {  // bp/DbvtAabbMm.java
    V3 mi, mx;
    void set(const LAabb& o) { mi = o.mi; mx = o.mx; }
    V3 Center() const { V3 o; o.set(mi).add(mx); o.scl(0.5f); return o; }
    void Expand(const V3& e) { mi.sub(e); mx.add(e); }
    void SignedExpand(const V3& e) {
        if (e.x > 0) mx.x += e.x; else mi.x += e.x;
        if (e.y > 0) mx.y += e.y; else mi.y += e.y;
        if (e.z > 0) mx.z += e.z; else mi.z += e.z;
    }
    bool Contain(const LAabb& a) const {
        return (mi.x <= a.mi.x) && (mi.y <= a.mi.y) && (mi.z <= a.mi.z) && (mx.x >= a.mx.x) && (mx.y >= a.mx.y) && (mx.z >= a.mx.z);
    }
    static bool Intersect(const LAabb& a, const LAabb& b) {
        return (a.mi.x <= b.mx.x) && (a.mx.x >= b.mi.x) && (a.mi.y <= b.mx.y) && (a.mx.y >= b.mi.y) && (a.mi.z <= b.mx.z) &&
               (a.mx.z >= b.mi.z);
    }
    static float Proximity(const LAabb& a, const LAabb& b) {
        V3 d, t;
        d.set(a.mi).add(a.mx);
        t.set(b.mi).add(b.mx);
        d.sub(t);
        return jabsf(d.x) + jabsf(d.y) + jabsf(d.z);
    }
    static void Merge(const LAabb& a, const LAabb& b, LAabb& r) {
        for (int i = 0; i < 3; i++) {
            r.mi.setc(i, a.mi.get(i) < b.mi.get(i) ? a.mi.get(i) : b.mi.get(i));
            r.mx.setc(i, a.mx.get(i) > b.mx.get(i) ? a.mx.get(i) : b.mx.get(i));
        }
    }
};

struct LNode {
    LAabb volume;
    LNode* parent = nullptr;
    LNode* childs[2] = {nullptr, nullptr};
    void* data = nullptr;
    uint64_t seq = 0;  // stands in for Object.hashCode()
    bool isleaf() const { return childs[1] == nullptr; }
    bool isinternal() const { return !isleaf(); }
};

struct LDbvt {
    LNode* root = nullptr;
    LNode* free_ = nullptr;
    int lkhd = -1;
    int leaves = 0;
    unsigned opath = 0;
    uint64_t seqCounter = 0;
    std::vector<LNode*> all;
    ~LDbvt() { for (LNode* n : all) delete n; }

    static int indexof(const LNode* n) { return (n->parent->childs[1] == n) ? 1 : 0; }
    void deletenode(LNode* n) { free_ = n; }
    LNode* createnode(LNode* parent, const LAabb& vol, void* data) {
        LNode* n;
        if (free_) { n = free_; free_ = nullptr; }
        else { n = new LNode(); n->seq = ++seqCounter; all.push_back(n); }
        n->parent = parent;
        n->volume.set(vol);
        n->data = data;
        n->childs[1] = nullptr;
        return n;
    }
    void insertleaf(LNode* rt, LNode* leaf) {
        if (!root) {
            root = leaf;
            leaf->parent = nullptr;
            return;
        }
        if (!rt->isleaf()) {
            do {
                if (LAabb::Proximity(rt->childs[0]->volume, leaf->volume) < LAabb::Proximity(rt->childs[1]->volume, leaf->volume))
                    rt = rt->childs[0];
                else
                    rt = rt->childs[1];
            } while (!rt->isleaf());
        }
        LNode* prev = rt->parent;
        LAabb m;
        LAabb::Merge(leaf->volume, rt->volume, m);
        LNode* node = createnode(prev, m, nullptr);
        if (prev) {
            prev->childs[indexof(rt)] = node;
            node->childs[0] = rt; rt->parent = node;
            node->childs[1] = leaf; leaf->parent = node;
            do {
                if (!prev->volume.Contain(node->volume)) LAabb::Merge(prev->childs[0]->volume, prev->childs[1]->volume, prev->volume);
                else break;
                node = prev;
            } while (nullptr != (prev = node->parent));
        } else {
            node->childs[0] = rt; rt->parent = node;
            node->childs[1] = leaf; leaf->parent = node;
            root = node;
        }
    }
    LNode* removeleaf(LNode* leaf) {
        if (leaf == root) { root = nullptr; return nullptr; }
        LNode* parent = leaf->parent;
        LNode* prev = parent->parent;
        LNode* sibling = parent->childs[1 - indexof(leaf)];
        if (prev) {
            prev->childs[indexof(parent)] = sibling;
            sibling->parent = prev;
            deletenode(parent);
            // bp/Dbvt.java:661-669: `pb` aliases prev.volume, so NotEqual is always false and the refit stops
            // after one level
            LAabb::Merge(prev->childs[0]->volume, prev->childs[1]->volume, prev->volume);
            return prev;
        }
        root = sibling;
        sibling->parent = nullptr;
        deletenode(parent);
        return root;
    }
    LNode* insert(const LAabb& box, void* data) {
        LNode* leaf = createnode(nullptr, box, data);
        insertleaf(root, leaf);
        leaves++;
        return leaf;
    }
    void update(LNode* leaf, int lookahead = -1) {  // :124-136
        LNode* rt = removeleaf(leaf);
        if (rt) {
            if (lookahead >= 0) { for (int i = 0; (i < lookahead) && rt->parent; i++) rt = rt->parent; }
            else rt = root;
        }
        insertleaf(rt, leaf);
    }
    void updateVolume(LNode* leaf, const LAabb& volume) {  // :138-151
        LNode* rt = removeleaf(leaf);
        if (rt) {
            if (lkhd >= 0) { for (int i = 0; (i < lkhd) && rt->parent; i++) rt = rt->parent; }
            else rt = root;
        }
        leaf->volume.set(volume);
        insertleaf(rt, leaf);
    }
    bool updateMoving(LNode* leaf, LAabb& volume, const V3& velocity, float margin) {  // :157-169 (volume IN PLACE)
        if (leaf->volume.Contain(volume)) return false;
        volume.Expand(V3(margin, margin, margin));
        volume.SignedExpand(velocity);
        updateVolume(leaf, volume);
        return true;
    }
    void remove(LNode* leaf) {
        removeleaf(leaf);
        deletenode(leaf);
        leaves--;
    }
    LNode* sort(LNode* n, LNode*& r) {  // :794-826
        LNode* p = n->parent;
        if (p && p->seq > n->seq) {
            int i = indexof(n), j = 1 - i;
            LNode* s = p->childs[j];
            LNode* q = p->parent;
            if (q) q->childs[indexof(p)] = n; else r = n;
            s->parent = n; p->parent = n; n->parent = q;
            p->childs[0] = n->childs[0]; p->childs[1] = n->childs[1];
            n->childs[0]->parent = p; n->childs[1]->parent = p;
            n->childs[i] = p; n->childs[j] = s;
            std::swap(p->volume, n->volume);
            return p;
        }
        return n;
    }
    void optimizeIncremental(int passes) {  // :94-115
        if (passes < 0) passes = leaves;
        if (root && passes > 0) {
            do {
                LNode* node = root;
                int bit = 0;
                while (node->isinternal()) {
                    node = sort(node, root)->childs[(opath >> bit) & 1];
                    bit = (bit + 1) & 31;
                }
                update(node);
                ++opath;
            } while (--passes);
        }
    }
    template <class F>
    static void collideTT(LNode* root0, LNode* root1, F process, std::vector<std::pair<LNode*, LNode*>>& stack) {  // :247-282
        if (!root0 || !root1) return;
        stack.clear();
        stack.push_back({root0, root1});
        do {
            auto p = stack.back();
            stack.pop_back();
            if (p.first == p.second) {
                if (p.first->isinternal()) {
                    stack.push_back({p.first->childs[0], p.first->childs[0]});
                    stack.push_back({p.first->childs[1], p.first->childs[1]});
                    stack.push_back({p.first->childs[0], p.first->childs[1]});
                }
            } else if (LAabb::Intersect(p.first->volume, p.second->volume)) {
                if (p.first->isinternal()) {
                    if (p.second->isinternal()) {
                        stack.push_back({p.first->childs[0], p.second->childs[0]});
                        stack.push_back({p.first->childs[1], p.second->childs[0]});
                        stack.push_back({p.first->childs[0], p.second->childs[1]});
                        stack.push_back({p.first->childs[1], p.second->childs[1]});
                    } else {
                        stack.push_back({p.first->childs[0], p.second});
                        stack.push_back({p.first->childs[1], p.second});
                    }
                } else {
                    if (p.second->isinternal()) {
                        stack.push_back({p.first, p.second->childs[0]});
                        stack.push_back({p.first, p.second->childs[1]});
                    } else {
                        process(p.first, p.second);
                    }
                }
            }
        } while (!stack.empty());
    }
};

struct LProxy {  // bp/DbvtProxy.java
    LAabb aabb;
    LNode* leaf = nullptr;
    LProxy* links[2] = {nullptr, nullptr};
    int stage = 0;
    int uid = 0;
    int16_t group = 1, mask = -1;
    int world = 0;
};

struct LDbvtBroadphase {
    static constexpr int STAGECOUNT = 2;
    LDbvt sets[2];
    LProxy* stageRoots[STAGECOUNT + 1] = {nullptr, nullptr, nullptr};
    float predictedframes = 2;
    float margin = 0.05f;
    int stageCurrent = 0, fupdates = 1, dupdates = 1, gid = 0;
    // pair cache
    std::vector<std::pair<LProxy*, LProxy*>> pairArray;
    std::unordered_map<uint64_t, int> pairIndex;
    std::vector<std::pair<LNode*, LNode*>> stk;
    std::vector<LProxy*> owned;
    ~LDbvtBroadphase() { for (LProxy* p : owned) delete p; }

    static uint64_t key(const LProxy* a, const LProxy* b) { return ((uint64_t)(uint32_t)a->uid << 32) | (uint32_t)b->uid; }
    static LProxy* listappend(LProxy* item, LProxy* list) {
        item->links[0] = nullptr;
        item->links[1] = list;
        if (list) list->links[0] = item;
        return item;
    }
    static LProxy* listremove(LProxy* item, LProxy* list) {
        if (item->links[0]) item->links[0]->links[1] = item->links[1];
        else list = item->links[1];
        if (item->links[1]) item->links[1]->links[0] = item->links[0];
        return list;
    }
    void addPair(LProxy* a, LProxy* b) {  // bp/HashedOverlappingPairCache.java:67-75,179-188,291-347
        if (a->world != b->world) return;
        bool collides = (a->group & b->mask) != 0;
        collides = collides && (b->group & a->mask) != 0;
        if (!collides) return;
        if (a->uid > b->uid) std::swap(a, b);
        uint64_t k = key(a, b);
        if (pairIndex.find(k) != pairIndex.end()) return;
        pairIndex[k] = (int)pairArray.size();
        pairArray.push_back({a, b});
    }
    void removePairAt(int i) {
        uint64_t k = key(pairArray[i].first, pairArray[i].second);
        pairIndex.erase(k);
        int last = (int)pairArray.size() - 1;
        if (i != last) {
            pairArray[i] = pairArray[last];
            pairIndex[key(pairArray[i].first, pairArray[i].second)] = i;
        }
        pairArray.pop_back();
    }
    void process(LNode* na, LNode* nb) {  // bp/DbvtTreeCollider.java:41-56
        LProxy* pa = (LProxy*)na->data;
        LProxy* pb = (LProxy*)nb->data;
        if (LAabb::Intersect(pa->aabb, pb->aabb)) addPair(pa, pb);
    }
    LProxy* createProxy(const V3& mn, const V3& mx, int group, int mask, int world) {  // :173-182
        LProxy* p = new LProxy();
        owned.push_back(p);
        p->group = (int16_t)group; p->mask = (int16_t)mask; p->world = world;
        p->aabb.mi = mn; p->aabb.mx = mx;
        p->leaf = sets[0].insert(p->aabb, p);
        p->stage = stageCurrent;
        p->uid = ++gid;
        stageRoots[stageCurrent] = listappend(p, stageRoots[stageCurrent]);
        return p;
    }
    void destroyProxy(LProxy* p) {  // :184-194
        if (p->stage == STAGECOUNT) sets[1].remove(p->leaf);
        else sets[0].remove(p->leaf);
        stageRoots[p->stage] = listremove(p, stageRoots[p->stage]);
        for (int i = 0; i < (int)pairArray.size();) {
            if (pairArray[i].first == p || pairArray[i].second == p) removePairAt(i);
            else i++;
        }
    }
    void setAabb(LProxy* proxy, const V3& aabbMin, const V3& aabbMax) {  // :196-228
        LAabb aabb;
        aabb.mi = aabbMin; aabb.mx = aabbMax;
        if (proxy->stage == STAGECOUNT) {
            sets[1].remove(proxy->leaf);
            proxy->leaf = sets[0].insert(aabb, proxy);
        } else {
            if (LAabb::Intersect(proxy->leaf->volume, aabb)) {
                V3 delta; delta.set(aabbMin).add(aabbMax);
                delta.scl(0.5f);
                delta.sub(proxy->aabb.Center());
                delta.scl(predictedframes);
                sets[0].updateMoving(proxy->leaf, aabb, delta, margin);
            } else {
                sets[0].updateVolume(proxy->leaf, aabb);
            }
        }
        stageRoots[proxy->stage] = listremove(proxy, stageRoots[proxy->stage]);
        proxy->aabb.set(aabb);
        proxy->stage = stageCurrent;
        stageRoots[stageCurrent] = listappend(proxy, stageRoots[stageCurrent]);
    }
    void collide() {  // :89-150
        sets[0].optimizeIncremental(1 + (sets[0].leaves * dupdates) / 100);
        sets[1].optimizeIncremental(1 + (sets[1].leaves * fupdates) / 100);
        stageCurrent = (stageCurrent + 1) % STAGECOUNT;
        LProxy* current = stageRoots[stageCurrent];
        auto proc = [this](LNode* a, LNode* b) { process(a, b); };
        if (current) {
            do {
                LProxy* next = current->links[1];
                stageRoots[current->stage] = listremove(current, stageRoots[current->stage]);
                stageRoots[STAGECOUNT] = listappend(current, stageRoots[STAGECOUNT]);
                LDbvt::collideTT(sets[1].root, current->leaf, proc, stk);
                sets[0].remove(current->leaf);
                current->leaf = sets[1].insert(current->aabb, current);
                current->stage = STAGECOUNT;
                current = next;
            } while (current);
        }
        LDbvt::collideTT(sets[0].root, sets[1].root, proc, stk);
        LDbvt::collideTT(sets[0].root, sets[0].root, proc, stk);
        for (int i = 0, ni = (int)pairArray.size(); i < ni; i++) {
            if (!LAabb::Intersect(pairArray[i].first->aabb, pairArray[i].second->aabb)) {
                removePairAt(i);
                ni--;
                i--;
            }
        }
    }
};

}  // namespace orc
