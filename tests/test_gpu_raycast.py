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
