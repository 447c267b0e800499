"""-m gpu: batched closest-hit convex sweeps (b2c_convex_sweep_closest, SURVEY §8f rank 4 — the CCD query) against the
oracle's sequential CollisionWorld.convexSweepTest + ClosestConvexResultCallback: hit body, fraction, normal and point must be
BIT-identical (both sides run the same IEEE binary32 sequences: -fmad=false on the device, -ffp-contract=off on the host)."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu
EYE = np.eye(3, dtype=np.float32)


def _compare(gw, ow, shapes_g, shapes_o, basis, f, t, group=1, mask=-1, allowed=0.04):
    gu, gf, gn, gp = gw.convexSweepTestClosest(shapes_g, basis, f, t, group, mask, allowed)
    ou, of, on, op = ow.convex_sweep_closest(shapes_o, basis, f, t, group, mask, allowed)
    assert np.array_equal(gu, ou), f"hit bodies differ for sweeps {np.nonzero(gu != ou)[0][:8]}: gpu {gu[gu != ou][:8]} oracle {ou[gu != ou][:8]}"
    ok = gu >= 0
    assert np.array_equal(gf[ok].view(np.uint32), of[ok].view(np.uint32)), "hit fractions are not bit-identical"
    hit = gu > 0
    assert np.array_equal(gn[hit].view(np.uint32), on[hit].view(np.uint32)), "hit normals are not bit-identical"
    assert np.array_equal(gp[hit].view(np.uint32), op[hit].view(np.uint32)), "hit points are not bit-identical"
    return gu


def _casts(sc):
    return [sc.add_shape("sphere", 0.25), sc.add_shape("box", (0.3, 0.2, 0.25)), sc.add_shape("hull", scenes.hull_points(np.random.default_rng(4), 0.3))]


def _sweeps(rng, n, lo, hi, reach):
    f = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    t = (f + rng.uniform(-reach, reach, size=(n, 3))).astype(np.float32)
    basis = scenes.random_rotations(rng, n).astype(np.float32)
    basis[::3] = EYE
    return basis, f, t


def test_sweeps_through_a_bin_of_mixed_shapes(gpu_pkg):
    sc = scenes.bin_scene(n=3000, seed=52)
    casts = _casts(sc)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    xf = sc.transforms(2)
    gw.setWorldTransforms(xf)
    ow.set_transforms(xf)
    rng = np.random.default_rng(9)
    ext = float(sc.extent)
    n = 500
    basis, f, t = _sweeps(rng, n, -0.1 * ext, 1.0 * ext, 0.4 * ext)
    f[:, 1] = np.abs(f[:, 1]) + 1.0
    t[:, 1] = f[:, 1] - rng.uniform(0.5, 8.0, size=n).astype(np.float32)
    which = rng.integers(3, size=n)
    sg = np.asarray([gw.scene_shape_ids[casts[k]] for k in which], np.int32)
    so = np.asarray([ow.scene_shape_ids[casts[k]] for k in which], np.int32)
    gu = _compare(gw, ow, sg, so, basis, f, t)
    assert (gu > 0).sum() > 150
    # only the dynamic bodies (group 1) answer a callback whose mask is 1; a larger allowed penetration rejects grazing hits
    gd = _compare(gw, ow, sg, so, basis, f, t, group=1, mask=1, allowed=0.0)
    assert 0 < (gd > 0).sum() <= (gu > 0).sum() + 5


def test_sweeps_after_a_step_and_with_removed_bodies(gpu_pkg):
    sc = scenes.stack_scene(n_side=4, extra=True, seed=6)
    casts = _casts(sc)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0)
    rng = np.random.default_rng(3)
    n = 300
    basis, f, t = _sweeps(rng, n, -8.0, 8.0, 6.0)
    f[:, 1] = np.abs(f[:, 1]) + 6.0
    t[:, 1] = f[:, 1] - rng.uniform(2.0, 16.0, size=n).astype(np.float32)
    sg = np.full(n, gw.scene_shape_ids[casts[1]], np.int32)
    so = np.full(n, ow.scene_shape_ids[casts[1]], np.int32)
    for step in range(2):
        xf = sc.transforms(step)
        gw.setWorldTransforms(xf); gw.step()
        ow.step(xf)
        _compare(gw, ow, sg, so, basis, f, t)
    for uid in (5, 20, 33):
        gw.removeCollisionObject(uid)
        ow.destroy_body(uid)
    gu = _compare(gw, ow, sg, so, basis, f, t)
    assert (gu > 0).sum() > 50 and not np.isin(gu, (5, 20, 33)).any()


def test_long_sweeps_with_more_candidates_than_one_round_holds(gpu_pkg):
    """The reference expands every body's box by the cast shape's box INCLUDING its whole motion, so a long sweep through a
    dense scene has thousands of candidates: k_convex_sweep takes them in rounds of SWEEP_MAX_CAND (4096)."""
    sc = scenes.bin_scene(n=9000, seed=53)
    cast = sc.add_shape("box", (0.4, 0.4, 0.4))
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    ext = float(sc.extent)
    rng = np.random.default_rng(12)
    n = 12
    f = rng.uniform(0.0, 0.15 * ext, size=(n, 3)).astype(np.float32)
    t = rng.uniform(0.85 * ext, ext, size=(n, 3)).astype(np.float32)
    f[:, 1] += 0.5
    t[:, 1] = rng.uniform(0.5, 0.9 * ext, size=n)
    basis = scenes.random_rotations(rng, n).astype(np.float32)
    sg = np.full(n, gw.scene_shape_ids[cast], np.int32)
    so = np.full(n, ow.scene_shape_ids[cast], np.int32)
    gu = _compare(gw, ow, sg, so, basis, f, t, group=1, mask=1)
    assert (gu > 0).sum() >= n - 2
    _compare(gw, ow, sg, so, basis, t, f, group=1, mask=1)          # and back


def test_sweep_kats(gpu_pkg):
    sc = scenes.stack_scene(n_side=1, extra=False, seed=1)      # ground box (top at y=0) + one unit box resting on it
    cast = sc.add_shape("sphere", 0.25)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=0)
    c = sc.base[1, 9:]
    sid = gw.scene_shape_ids[cast]
    uid, frac, nrm, pt = gw.convexSweepTestClosest(sid, EYE, [(c[0], 10.0, c[2]), (30.0, 10.0, 30.0), (c[0], 10.0, c[2])],
                                                   [(c[0], -10.0, c[2]), (30.0, -10.0, 30.0), (c[0], 9.0, c[2])])
    assert uid.tolist() == [2, 1, 0]                              # the box, the ground beside it, a sweep that stops short
    assert abs(frac[0] - (10.0 - (c[1] + 1.0) - 0.25) / 20.0) < 2e-3 and abs(frac[1] - (10.0 - 0.25) / 20.0) < 2e-3 and frac[2] == 1.0
    assert np.allclose(nrm[:2], [[0, 1, 0], [0, 1, 0]], atol=1e-2)
    assert abs(pt[1][1]) < 1e-2
    with pytest.raises(Exception):                                # the cast shape must be convex
        gw.convexSweepTestClosest(10 ** 6, EYE, [(0, 1, 0)], [(0, 2, 0)])


def test_sweeps_against_terrain_mesh_compounds_and_the_plane_branch(gpu_pkg):
    """objectQuerySingle's concave and compound branches (disp/CollisionWorld.java:429-456, 528-545): BVH box-cast walk +
    SubsimplexConvexCast per triangle for the mesh, every child of a compound; a static plane that passes the callback's
    filter is the reference's throwing branch (uid -1)."""
    sc = scenes.terrain_scene(cells=48, n=300, seed=14)
    rng = np.random.default_rng(22)
    pl = sc.add_shape("plane", (0.1, 1.0, -0.05), -6.0)
    sph = sc.add_shape("sphere", 0.35)
    bar = sc.add_shape("box", (0.5, 0.12, 0.12))
    eye = np.eye(3)
    dumb = sc.add_shape("compound", [sph, bar, sph], scenes.make_xf(np.stack([eye] * 3), np.asarray([(-0.5, 0, 0), (0, 0, 0), (0.5, 0, 0)])))
    casts = _casts(sc)
    extra_pos, extra_rot, extra_shape = [(0.0, 0.0, 0.0)], [eye], [pl]
    for _ in range(80):
        extra_pos.append(tuple(rng.uniform((1, 5, 1), (23, 8, 23))))
        extra_rot.append(scenes.random_rotations(rng, 1)[0])
        extra_shape.append(dumb)
    for k, sid in enumerate(extra_shape):
        sc.body_shape.append(sid); sc.static.append(k == 0); sc.group.append(2 if k == 0 else 1); sc.mask.append(-1 ^ 2 if k == 0 else -1); sc.world.append(0)
    # the terrain answers as group 8 so that a callback can leave the plane (group 2) out and keep the mesh
    sc.group[0] = 8
    sc.base = np.concatenate([sc.base, scenes.make_xf(np.asarray(extra_rot), np.asarray(extra_pos))], axis=0)
    sc.vel = None
    sc.spin = None
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    n = 900
    basis, f, t = _sweeps(rng, n, 0.0, 24.0, 5.0)
    f[:, 1] = rng.uniform(3.0, 11.0, size=n)
    t[:, 1] = f[:, 1] - rng.uniform(1.0, 12.0, size=n).astype(np.float32)
    which = rng.integers(3, size=n)
    sg = np.asarray([gw.scene_shape_ids[casts[k]] for k in which], np.int32)
    so = np.asarray([ow.scene_shape_ids[casts[k]] for k in which], np.int32)
    gu = _compare(gw, ow, sg, so, basis, f, t, group=1, mask=-1 ^ 2)
    kinds = np.asarray([sc.shapes[sc.body_shape[u - 1]][0] for u in gu[gu > 0]])
    assert (gu > 0).sum() > 400 and (kinds == "mesh").sum() > 100 and (kinds == "compound").sum() > 30, (len(kinds), (kinds == "mesh").sum(), (kinds == "compound").sum())
    # with the plane inside the filter every sweep whose box reaches it is the unsupported branch
    gu2 = _compare(gw, ow, sg[:100], so[:100], basis[:100], f[:100], t[:100])
    assert (gu2 == -1).all()
    # long sweeps from below: the mesh is hit from its back side
    _compare(gw, ow, sg[:300], so[:300], basis[:300], t[:300] - np.asarray([0, 6.0, 0], np.float32), f[:300], group=1, mask=-1 ^ 2)


def test_sweeps_against_a_multi_part_mesh(gpu_pkg):
    sc = scenes.terrain_scene(cells=32, n=50, seed=19)
    scenes.split_mesh_into_parts(sc, nparts=3, short_parts=(1,))
    casts = _casts(sc)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    rng = np.random.default_rng(6)
    n = 400
    basis, f, t = _sweeps(rng, n, 0.0, 16.0, 3.0)
    f[:, 1] = rng.uniform(3.0, 8.0, size=n)
    t[:, 1] = f[:, 1] - rng.uniform(2.0, 10.0, size=n).astype(np.float32)
    sg = np.full(n, gw.scene_shape_ids[casts[2]], np.int32)
    so = np.full(n, ow.scene_shape_ids[casts[2]], np.int32)
    gu = _compare(gw, ow, sg, so, basis, f, t)
    assert (gu == 1).sum() > 150


def test_ccd_motion_clamping_sweeps(gpu_pkg):
    """b2c_ccd_sweep_not_me == the oracle's DiscreteDynamicsWorld.integrateTransforms CCD query (ClosestNotMeConvexResult-
    Callback: not the body itself, not objects it already touches, not results whose normal follows the motion), after a real
    step so that the pair cache and its manifolds exist — bin of mixed shapes, and compounds / convex bodies on a terrain mesh."""
    rng = np.random.default_rng(41)
    for sc, mode in ((scenes.bin_scene(n=2500, seed=54), 1), (scenes.terrain_compound_scene(cells=32, n=150, seed=16), 1)):
        gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode)
        for step in range(2):
            xf = sc.transforms(step)
            gw.setWorldTransforms(xf); gw.step()
            ow.step(xf)
        dyn = np.asarray([k + 1 for k in range(sc.n) if not sc.static[k]], np.int32)
        me = rng.choice(dyn, size=min(400, len(dyn)), replace=False).astype(np.int32)
        radius = rng.uniform(0.1, 0.3, size=len(me)).astype(np.float32)
        to = (xf[me - 1, 9:] + rng.uniform(-2.0, 2.0, size=(len(me), 3))).astype(np.float32)
        to[::2, 1] = xf[me[::2] - 1, 10] - rng.uniform(0.5, 3.0, size=len(me[::2])).astype(np.float32)   # half of them downwards
        gu, gf, gn, gp = gw.ccdSweepNotMe(me, radius, to)
        ou, of, on, op = ow.ccd_sweep_not_me(me, radius, to)
        assert np.array_equal(gu, ou), f"hit bodies differ: {np.nonzero(gu != ou)[0][:8]} gpu {gu[gu != ou][:8]} oracle {ou[gu != ou][:8]}"
        assert np.array_equal(gf.view(np.uint32), of.view(np.uint32))
        hit = gu > 0
        assert np.array_equal(gn[hit].view(np.uint32), on[hit].view(np.uint32)) and np.array_equal(gp[hit].view(np.uint32), op[hit].view(np.uint32))
        assert hit.sum() > 40 and not (gu == me).any()
        # the exclusion of touching objects matters: without the pair cache (fresh worlds, no step) more sweeps report a hit
        g2, o2 = scenes.build_both(gpu_pkg, sc, mode=mode)
        g2.setWorldTransforms(xf); o2.set_transforms(xf)
        hu, hf, _, _ = g2.ccdSweepNotMe(me, radius, to)
        pu, pf, _, _ = o2.ccd_sweep_not_me(me, radius, to)
        assert np.array_equal(hu, pu) and np.array_equal(hf.view(np.uint32), pf.view(np.uint32))
        assert (hu > 0).sum() >= hit.sum()
    with pytest.raises(Exception):
        gw.ccdSweepNotMe([10 ** 6], 0.2, [(0, 0, 0)])
