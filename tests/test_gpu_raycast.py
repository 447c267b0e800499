"""-m gpu: batched closest-hit ray tests (b2c_ray_test_closest) against the oracle's sequential
CollisionWorld.rayTest + ClosestRayResultCallback on mixed box / sphere / hull scenes."""
import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


def _rays(rng, n, lo, hi):
    f = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    t = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    return f, t


def _compare(gw, ow, f, t, group=1, mask=-1):
    gu, gf, gn, gp = gw.rayTestClosest(f, t, group, mask)
    ou, of, on, op = ow.ray_test_closest(f, t, group, mask)
    assert np.array_equal(gu, ou), f"hit bodies differ for rays {np.nonzero(gu != ou)[0][:8]}: gpu {gu[gu != ou][:8]} oracle {ou[gu != ou][:8]}"
    assert np.array_equal(gf.view(np.uint32), of.view(np.uint32)), "hit fractions are not bit-identical"
    hit = gu > 0
    assert np.array_equal(gn[hit].view(np.uint32), on[hit].view(np.uint32)), "hit normals are not bit-identical"
    assert np.array_equal(gp[hit].view(np.uint32), op[hit].view(np.uint32)), "hit points are not bit-identical"
    return int(hit.sum())


def test_rays_through_a_bin_of_mixed_shapes(gpu_pkg):
    sc = scenes.bin_scene(n=3000, seed=51)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    xf = sc.transforms(2)
    gw.setWorldTransforms(xf)
    ow.set_transforms(xf)
    rng = np.random.default_rng(8)
    ext = float(sc.extent)
    f, t = _rays(rng, 600, -0.6 * ext, 0.6 * ext)
    f[:, 1] = np.abs(f[:, 1]) + 2.0           # most rays start above the pile
    hits = _compare(gw, ow, f, t)
    assert hits > 200
    # the callback's filter: only dynamic bodies (group 1) answer a ray whose mask is 1
    hits_dyn = _compare(gw, ow, f, t, group=1, mask=1)
    assert 0 < hits_dyn <= hits


def test_rays_after_a_step_and_with_removed_bodies(gpu_pkg):
    sc = scenes.stack_scene(n_side=4, extra=True, seed=6)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0)
    rng = np.random.default_rng(3)
    f, t = _rays(rng, 300, -10.0, 10.0)
    f[:, 1] += 12.0
    for step in range(2):
        xf = sc.transforms(step)
        gw.setWorldTransforms(xf); gw.step()
        ow.step(xf)
        _compare(gw, ow, f, t)
    for uid in (5, 20, 33):
        gw.removeCollisionObject(uid)
        ow.destroy_body(uid)
    hits = _compare(gw, ow, f, t)
    assert hits > 50


def test_ray_kats(gpu_pkg):
    sc = scenes.stack_scene(n_side=1, extra=False, seed=1)      # ground box (top at y=0) + one unit box resting on it
    gw = scenes.build_gpu(gpu_pkg, sc, mode=0)
    c = sc.base[1, 9:]
    uid, frac, nrm, pt = gw.rayTestClosest([(c[0], 10.0, c[2]), (30.0, 10.0, 30.0), (c[0], 10.0, c[2])],
                                           [(c[0], -10.0, c[2]), (30.0, -10.0, 30.0), (c[0], 9.0, c[2])])
    assert uid.tolist() == [2, 1, 0]                              # the box, the ground beside it, a ray that stops short
    assert abs(frac[0] - (10.0 - (c[1] + 1.0)) / 20.0) < 1e-3 and abs(frac[1] - 0.5) < 1e-3 and frac[2] == 1.0
    assert np.allclose(nrm[:2], [[0, 1, 0], [0, 1, 0]], atol=1e-2)
    assert abs(pt[1][1]) < 1e-2


def test_rays_against_terrain_mesh_plane_and_compounds(gpu_pkg):
    """rayTestSingle's concave and compound branches (disp/CollisionWorld.java:301-356): BVH ray walk + triangle test for the
    mesh, the two generated triangles for a static plane, every child of a compound — bit-identical to the oracle."""
    sc = scenes.terrain_scene(cells=48, n=300, seed=14)
    # add a static plane far below the terrain and a layer of compounds above it
    rng = np.random.default_rng(21)
    pl = sc.add_shape("plane", (0.1, 1.0, -0.05), -6.0)
    sph = sc.add_shape("sphere", 0.35)
    bar = sc.add_shape("box", (0.5, 0.12, 0.12))
    eye = np.eye(3)
    dumb = sc.add_shape("compound", [sph, bar, sph], scenes.make_xf(np.stack([eye] * 3), np.asarray([(-0.5, 0, 0), (0, 0, 0), (0.5, 0, 0)])))
    extra_pos, extra_rot, extra_shape = [(0.0, 0.0, 0.0)], [eye], [pl]
    for _ in range(80):
        extra_pos.append(tuple(rng.uniform((1, 5, 1), (23, 8, 23))))
        extra_rot.append(scenes.random_rotations(rng, 1)[0])
        extra_shape.append(dumb)
    for k, sid in enumerate(extra_shape):
        sc.body_shape.append(sid); sc.static.append(k == 0); sc.group.append(2 if k == 0 else 1); sc.mask.append(-1 ^ 2 if k == 0 else -1); sc.world.append(0)
    sc.base = np.concatenate([sc.base, scenes.make_xf(np.asarray(extra_rot), np.asarray(extra_pos))], axis=0)
    sc.vel = None
    sc.spin = None
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    f, t = _rays(rng, 1500, -2.0, 26.0)
    f[:, 1] = rng.uniform(4.0, 12.0, size=len(f))
    t[:, 1] = rng.uniform(-9.0, 2.0, size=len(t))
    f[::7] = t[::7] + np.asarray([0.0, 15.0, 0.0], np.float32)     # some rays straight down
    hits = _compare(gw, ow, f, t)
    gu = gw.rayTestClosest(f, t)[0]
    kinds = np.asarray([sc.shapes[sc.body_shape[u - 1]][0] for u in gu[gu > 0]])
    assert hits > 1000 and (kinds == "mesh").sum() > 300 and (kinds == "plane").sum() > 10 and (kinds == "compound").sum() > 50
    # rays from below the terrain: the mesh reports the flipped triangle normal
    hits_up = _compare(gw, ow, t, f)
    assert hits_up > 1000
    # only the dynamic bodies answer (callback mask 1): compounds and convex bodies, no mesh / plane
    _compare(gw, ow, f, t, group=1, mask=1)


def test_rays_against_a_compound_heavy_scene(gpu_pkg):
    sc = scenes.compound_scene(n=400, seed=17)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    rng = np.random.default_rng(5)
    ext = float(sc.extent)
    f, t = _rays(rng, 800, -1.0, ext)
    f[:, 1] = rng.uniform(2.0, 8.0, size=len(f))
    t[:, 1] = rng.uniform(-3.0, 1.0, size=len(t))
    assert _compare(gw, ow, f, t) > 600


def test_a_ray_through_more_boxes_than_one_round_holds(gpu_pkg):
    """1500 spheres crowded into one spot: a ray through the cluster meets more than the 1024 candidate boxes a block keeps
    and takes the index-ordered tile path of k_ray_test; same hits as the oracle's sequential loop, bit for bit."""
    rng = np.random.default_rng(31)
    sc = scenes.Scene()
    sph = sc.add_shape("sphere", 0.5)
    bx = sc.add_shape("box", (0.3, 0.3, 0.3))
    n = 1500
    pos = rng.uniform(-0.2, 0.2, size=(n, 3))
    pos[::10] += rng.uniform(-6.0, 6.0, size=(len(pos[::10]), 3))      # a few bodies away from the cluster
    for k in range(n):
        sc.body_shape.append(sph if k % 3 else bx); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    sc.base = scenes.make_xf(scenes.random_rotations(rng, n), pos)
    sc.vel = None
    sc.spin = None
    sc.extent = 12.0
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0)
    f = rng.uniform(-8.0, 8.0, size=(200, 3)).astype(np.float32)
    t = (-f + rng.uniform(-0.3, 0.3, size=(200, 3))).astype(np.float32)   # through the middle of the cluster
    f[100:], t[100:] = _rays(rng, 100, -8.0, 8.0)
    hits = _compare(gw, ow, f, t)
    assert hits >= 100
