"""CPU: the oracle's island labelling and pair deltas (oracle/islands.h, orc_pair_deltas) on hand-derived cases."""
import numpy as np

import orc
import scenes


def canon_partition(tags):
    """Relabel every island by its smallest member index (statics stay -1): partitions compare equal iff these do."""
    t = np.asarray(tags).copy()
    dyn = t >= 0
    if not dyn.any():
        return t
    _, inv = np.unique(t[dyn], return_inverse=True)
    mins = np.full(inv.max() + 1, 1 << 30, dtype=np.int64)
    np.minimum.at(mins, inv, np.nonzero(dyn)[0])
    t[dyn] = mins[inv].astype(t.dtype)
    return t


def _chain_world():
    """ground (static) + 5 unit spheres on a line: 0-1 touching, 1-2 touching, 3-4 touching, 2|3 apart."""
    w = orc.OracleWorld(orc.TIGHT)
    g = w.box(50, 1, 50)
    s = w.sphere(0.5)
    w.body(g, orc.xf12(origin=(0, -1.0, 0)), group=2, mask=-1 ^ 2, static=True)
    for x in (0.0, 0.9, 1.8, 5.0, 5.9):
        w.body(s, orc.xf12(origin=(x, 0.5, 0)))
    return w


def test_union_find_partition_kat():
    w = _chain_world()
    w.step()
    tags, n = w.islands()
    # every sphere overlaps the static ground, but statics do not merge islands (disp/SimulationIslandManager.java:64-67)
    assert tags[0] == -1
    c = canon_partition(tags)
    assert list(c) == [-1, 1, 1, 1, 4, 4]
    assert n == 2


def test_unite_direction_matches_reference():
    """disp/UnionFind.java:104-124: unite(p, q) hangs find(p) under find(q) — the tag of {1,2} united as (1,2) is 2."""
    w = _chain_world()
    w.step()
    tags, _ = w.islands()
    # pairs in sorted order: (2,3) -> id[1]=2 ; (3,4) -> find(2)=2 under 3 -> root 3 ; objects are uid-1
    assert tags[1] == tags[2] == tags[3] == 3
    assert tags[4] == tags[5] == 5


def test_pair_deltas_kat():
    w = _chain_world()
    w.step()
    a, r = w.pair_deltas()
    assert len(r) == 0 and len(a) == 5 + 3          # 5 ground pairs + 3 touching sphere pairs, all new
    xf = np.stack([orc.xf12(origin=(0, -1.0, 0))] + [orc.xf12(origin=(x, 0.5, 0)) for x in (0.0, 0.9, 3.4, 4.2, 5.9)])
    w.step(xf)
    a, r = w.pair_deltas()
    assert r.tolist() == [[3, 4], [5, 6]]            # sphere 3 left 2, sphere 4 left 5
    assert a.tolist() == [[4, 5]]                    # ... and now touches sphere 4 (uids 4 and 5)


def test_deltas_compose_to_the_pair_set():
    sc = scenes.bin_scene(n=600, seed=9)
    sc.vel *= 4.0
    ow = scenes.build_oracle(sc, orc.DBVT)
    cur = set()
    for step in range(5):
        pairs = ow.step(sc.transforms(step))
        a, r = ow.pair_deltas()
        cur = (cur - set(map(tuple, r.tolist()))) | set(map(tuple, a.tolist()))
        assert cur == set(map(tuple, pairs.tolist()))


def test_island_partition_equals_scipy_connected_components():
    """Independent check: the island partition over the pairs of non-static bodies is the connected-component partition of
    that graph (scipy), on a scene with thousands of pairs."""
    import scenes
    from scipy.sparse import coo_matrix
    from scipy.sparse.csgraph import connected_components
    sc = scenes.bin_scene(n=1500, seed=33, spacing=1.05)
    ow = scenes.build_oracle(sc, orc.DBVT)
    pairs = ow.step(sc.transforms(0))
    tags, n_islands = ow.islands()
    static = np.asarray(sc.static, dtype=bool)
    dyn = ~static
    u, v = pairs[:, 0] - 1, pairs[:, 1] - 1
    keep = dyn[u] & dyn[v]
    g = coo_matrix((np.ones(keep.sum()), (u[keep], v[keep])), shape=(sc.n, sc.n))
    ncomp, lab = connected_components(g, directed=False)
    assert (tags[static] == -1).all() and (tags[dyn] >= 0).all()
    # same partition: the map tag -> scipy label is a bijection on the dynamic bodies
    pairs_tl = set(zip(tags[dyn].tolist(), lab[dyn].tolist()))
    assert len(pairs_tl) == len(set(tags[dyn].tolist())) == len(set(lab[dyn].tolist()))
    assert n_islands == len(set(tags[dyn].tolist()))
