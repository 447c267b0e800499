"""CPU checks of the oracle's CompoundShape / CompoundCollisionAlgorithm restatement (SURVEY §8f rank 3;
sh/CompoundShape.java:50-160, disp/CompoundCollisionAlgorithm.java:49-129).  No GPU."""
import numpy as np

import orc
import scenes


def _xf(pos, rot=None):
    return orc.xf12(rot, pos)


def test_compound_local_aabb_and_world_aabb_kat():
    """addChildShape extends the local box by each child's AABB under its local transform; getAabb is
    center/extents of that box under the world transform, margin 0 (sh/CompoundShape.java:50-82, 124-160)."""
    w = orc.OracleWorld(orc.TIGHT)
    b = w.box(0.5, 0.25, 0.25)      # AABB half extents = (0.5, 0.25, 0.25) exactly (core + margin)
    s = w.sphere(0.5)
    c = w.compound([b, s], np.stack([_xf((1.0, 0.0, 0.0)), _xf((-1.0, 0.5, 0.0))]))
    # local box: x in [-1.5, 1.5], y in [-0.25, 1.0], z in [-0.5, 0.5]
    a = w.shape_aabb(c, _xf((10.0, 20.0, 30.0)))
    # the box half extent is (0.5 - 0.04) + 0.04 in float32, so compare with that arithmetic
    hx = np.float32(np.float32(0.5) - np.float32(0.04)) + np.float32(0.04)
    hy = np.float32(np.float32(0.25) - np.float32(0.04)) + np.float32(0.04)
    lo = np.array([-1.5, min(-hy, 0.0), -0.5], np.float32)
    hi = np.array([1.0 + hx, 1.0, 0.5], np.float32)
    he = (hi - lo) * np.float32(0.5)
    ce = (hi + lo) * np.float32(0.5) + np.array([10.0, 20.0, 30.0], np.float32)
    assert np.array_equal(a[:3], ce - he) and np.array_equal(a[3:], ce + he)
    # rotating the body by 90 degrees about z swaps the x / y extents
    rot = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]
    r = w.shape_aabb(c, _xf((0.0, 0.0, 0.0), rot))
    assert np.allclose(r[3:] - r[:3], [2 * he[1], 2 * he[0], 2 * he[2]], atol=1e-6)


def test_child_manifolds_match_separate_bodies():
    """Each child algorithm sees (child shape, orgTrans * childTrans): with an identity body rotation and dyadic offsets
    the child's world transform is exact, so the detector output of every child must be BIT-identical to a plain body
    placed there (disp/CompoundCollisionAlgorithm.java:104-120)."""
    offs = [(-0.5, 0.0, 0.0), (0.5, 0.0, 0.0), (0.0, 0.5, 0.25)]
    org = (2.0, 1.0, -3.0)
    others = [("sphere", (2.25, 1.5, -3.0)), ("box", (2.0, 0.5, -2.75)), ("hull", (1.5, 1.25, -3.25))]
    rng = np.random.default_rng(5)
    hp = scenes.hull_points(rng, 0.4)
    for oname, opos in others:
        # world A: one compound body + the other body
        wa = orc.OracleWorld(orc.TIGHT)
        kids = [wa.sphere(0.3), wa.box(0.3, 0.2, 0.25), wa.hull(hp)]
        comp = wa.compound(kids, np.stack([_xf(o) for o in offs]))
        oth = {"sphere": lambda w: w.sphere(0.35), "box": lambda w: w.box(0.3, 0.3, 0.3), "hull": lambda w: w.hull(hp * 0.9)}[oname]
        oa = oth(wa)
        wa.body(comp, _xf(org))
        wa.body(oa, _xf(opos))
        wa.step()
        ri, rf = wa.raw()
        assert len(ri) == 3 and ri[:, 2].tolist() == [-2, -3, -4]
        hdr, pts, pint = wa.manifolds()
        assert len(hdr) == 3 and hdr[:, 5].tolist() == [0, 1, 2] and (hdr[:, 6] == -1).all()
        assert (hdr[:, 2] == 1).all() and (hdr[:, 3] == 2).all()
        # world B: the three children as plain bodies (never colliding with each other) + the other body
        for k in range(3):
            wb = orc.OracleWorld(orc.TIGHT)
            kb = [wb.sphere(0.3), wb.box(0.3, 0.2, 0.25), wb.hull(hp)]
            ob = oth(wb)
            pos = tuple(np.float32(org[i]) + np.float32(offs[k][i]) for i in range(3))
            wb.body(kb[k], _xf(pos))
            wb.body(ob, _xf(opos))
            # the child alone need not overlap the other body's AABB (the compound's box did): force the pair
            for uid in (1, 2):
                wb.set_aabb(uid, (-100.0, -100.0, -100.0), (100.0, 100.0, 100.0))
            wb.calculate_overlapping_pairs()
            wb.dispatch_all_pairs()
            qi, qf = wb.raw()
            assert len(qi) == 1
            assert qi[0, 3] == ri[k, 3] and qi[0, 4] == ri[k, 4] and qi[0, 5] == ri[k, 5], (oname, k)
            assert np.array_equal(qf[0].view(np.uint32), rf[k].view(np.uint32)), (oname, k)
            h2, p2, _ = wb.manifolds()
            assert h2[0, 4] == hdr[k, 4]
            if h2[0, 4]:
                # world points and distance agree bit for bit; local points differ (projected with the compound's transform)
                assert np.array_equal(p2[0, 0, 6:16].view(np.uint32), pts[k, 0, 6:16].view(np.uint32))
                # localPointA is relative to the compound body: child-local point + child offset
                assert np.allclose(pts[k, 0, 0:3], p2[0, 0, 0:3] + np.asarray(offs[k], np.float32), atol=1e-6)


def test_swapped_and_nested_order():
    """The compound may be the pair's second object (swappedCompoundCreateFunc) and both objects may be compounds: child
    algorithms then run i (first compound) outer, j (second) inner, with manifold bodies (second, first)
    (disp/CompoundCollisionAlgorithm.java:57-75; disp/DefaultCollisionConfiguration.java:198-204)."""
    w = orc.OracleWorld(orc.TIGHT)
    s = w.sphere(0.4)
    b = w.box(0.3, 0.3, 0.3)
    c2 = w.compound([s, b], np.stack([_xf((-0.4, 0, 0)), _xf((0.4, 0, 0))]))
    c3 = w.compound([b, s, s], np.stack([_xf((0, 0, 0)), _xf((0, 0.5, 0)), _xf((0, -0.5, 0))]))
    w.body(s, _xf((0.0, 0.0, 0.0)))          # uid 1: plain sphere
    w.body(c2, _xf((0.5, 0.0, 0.0)))         # uid 2: compound second in pair (1,2)
    w.body(c3, _xf((0.7, 0.2, 0.0)))         # uid 3: pair (2,3) compound x compound; pair (1,3)
    w.step()
    hdr, pts, pint = w.manifolds()
    rows = {}
    for h in hdr.tolist():
        rows.setdefault((h[0], h[1]), []).append(h[2:])
    assert [r[:2] + r[3:] for r in rows[(1, 2)]] == [[2, 1, 0, -1], [2, 1, 1, -1]]       # bodies (compound, other)
    assert [r[:2] + r[3:] for r in rows[(1, 3)]] == [[3, 1, 0, -1], [3, 1, 1, -1], [3, 1, 2, -1]]
    # (2,3): outer loop over body 2's children, inner over body 3's; manifold bodies (3, 2); child0 = child of body 3
    assert [r[:2] + r[3:] for r in rows[(2, 3)]] == [[3, 2, j, i] for i in range(2) for j in range(3)]
    ri, _ = w.raw()
    assert sorted(ri[(ri[:, 0] == 2) & (ri[:, 1] == 3), 2].tolist(), reverse=True) == [-2, -3, -4, -5, -6, -7]


def test_compound_manifolds_persist_and_die_with_the_pair():
    sc = scenes.compound_scene(n=60, seed=3)
    ow = scenes.build_oracle(sc, orc.DBVT)
    counts = []
    for step in range(4):
        ow.step(sc.transforms(step))
        hdr, pts, pint = ow.manifolds()
        counts.append(len(hdr))
        kid = hdr[:, 5] >= 0
        assert kid.any()
        if step:
            # lifetimes grow for points that persist in child manifolds
            live = np.arange(4)[None, :] < hdr[:, 4][:, None]
            assert (pint[kid][:, :, 0][live[kid]] >= 1).all()
    c = ow.counters()
    assert c["gjk_checks"] > 0 and c["added_contacts"] > 0
    assert counts[-1] > 0


def test_nested_compounds_in_the_oracle():
    """A child that is a CompoundShape: (1) wrapped at the identity it changes no bit (multiplying by the identity is exact);
    (2) with a real frame transform, every leaf sits where ((org * frame) * child) puts it — checked against a flat compound
    whose child transforms were composed in float64 and rounded, contact points within 1e-5."""
    import scenes
    flat = scenes.compound_scene(n=120, seed=11, plane_ground=True)
    wrap = scenes.compound_scene(n=120, seed=11, plane_ground=True, nested="identity")
    w0 = scenes.build_oracle(flat, orc.DBVT)
    w1 = scenes.build_oracle(wrap, orc.DBVT)
    for step in range(3):
        p0 = w0.step(flat.transforms(step))
        p1 = w1.step(wrap.transforms(step))
        assert np.array_equal(p0, p1)
        (h0, f0, i0), (h1, f1, i1) = w0.manifolds(), w1.manifolds()
        assert np.array_equal(h0, h1) and f0.tobytes() == f1.tobytes() and np.array_equal(i0, i1)
    assert len(h0) > 300
    # (2) one nested body above a ground box: dumbbell inside a rotated, shifted frame
    w = orc.OracleWorld(mode=orc.TIGHT)
    wf = orc.OracleWorld(mode=orc.TIGHT)
    c, s = np.cos(0.6), np.sin(0.6)
    R = np.asarray([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    frame_o = np.asarray([0.1, 0.3, -0.2])
    offs = [np.asarray([-0.5, 0, 0]), np.asarray([0.5, 0, 0])]
    for world, nested in ((w, True), (wf, False)):
        g = world.box(10, 0.5, 10)
        sp = world.sphere(0.3)
        world.body(g, orc.xf12(origin=(0, -0.5, 0)), 2, -1 ^ 2, True, 0)
        if nested:
            inner = world.compound([sp, sp], np.stack([_xf(tuple(o)) for o in offs]))
            outer = world.compound([inner], np.stack([orc.xf12(R, frame_o)]))
        else:
            outer = world.compound([sp, sp], np.stack([orc.xf12(R, frame_o + R @ o) for o in offs]))
        world.body(outer, orc.xf12(origin=(0, 0.25, 0)), 1, -1, False, 0)
    xf = np.stack([orc.xf12(origin=(0, -0.5, 0)), orc.xf12(origin=(0, 0.25, 0))])
    w.step(xf); wf.step(xf)
    (h, f, _), (hf, ff, _) = w.manifolds(), wf.manifolds()
    assert np.array_equal(h[:, :5], hf[:, :5]) and h[:, 4].sum() >= 1
    live = np.arange(4)[None, :] < h[:, 4][:, None]
    assert np.allclose(f[live], ff[live], atol=1e-5)
