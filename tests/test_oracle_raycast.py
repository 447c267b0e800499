"""CPU: the oracle's ray test (oracle/raycast.h) against closed-form ray/box and ray/sphere intersections and hand-derived
ordering cases of the ClosestRayResultCallback loop (disp/CollisionWorld.java:553-590)."""
import numpy as np

import orc


def _world():
    w = orc.OracleWorld(orc.TIGHT)
    box = w.box(1, 1, 1)          # core 0.96 + margin 0.04: the cast sees the sharp unit box
    sph = w.sphere(0.5)
    w.body(box, orc.xf12(origin=(0, 0, 0)), group=2, mask=-1 ^ 2, static=True)   # uid 1
    w.body(sph, orc.xf12(origin=(3, 0, 0)))                                       # uid 2
    w.body(box, orc.xf12(origin=(0, -4, 0)))                                      # uid 3, hidden below uid 1
    return w


def test_ray_hits_box_top_face_and_sphere_pole():
    w = _world()
    uid, frac, nrm, pt = w.ray_test_closest([(0.2, 5, 0.1), (3, 5, 0)], [(0.2, -5, 0.1), (3, -5, 0)])
    assert uid.tolist() == [1, 2]
    assert abs(frac[0] - 0.4) < 1e-3 and abs(frac[1] - 0.45) < 1e-3
    assert np.allclose(nrm, [[0, 1, 0], [0, 1, 0]], atol=2e-2)
    assert np.allclose(pt[0], [0.2, 1.0, 0.1], atol=1e-2) and np.allclose(pt[1], [3, 0.5, 0], atol=1e-2)


def test_miss_filter_and_closest_of_two():
    w = _world()
    # 1: passes beside everything; 2: from below hits uid 3 first; 3: the callback's mask excludes group 2 (static) -> sees uid 3
    uid, frac, _, _ = w.ray_test_closest([(10, 5, 0), (0, -9, 0)], [(10, -5, 0), (0, 9, 0)])
    assert uid.tolist() == [0, 3] and frac[0] == 1.0 and abs(frac[1] - 4.0 / 18.0) < 1e-3
    uid, frac, _, _ = w.ray_test_closest([(0, 5, 0)], [(0, -9, 0)], group=1, mask=1)
    assert uid.tolist() == [3] and abs(frac[0] - 8.0 / 14.0) < 1e-3


def test_random_rays_against_analytic_spheres():
    rng = np.random.default_rng(4)
    w = orc.OracleWorld(orc.TIGHT)
    n = 40
    centers = rng.uniform(-6, 6, size=(n, 3))
    radii = rng.uniform(0.3, 0.9, size=n).astype(np.float32)
    for c, r in zip(centers, radii):
        w.body(w.sphere(float(r)), orc.xf12(origin=c))
    f = rng.uniform(-10, 10, size=(300, 3)).astype(np.float32)
    t = rng.uniform(-10, 10, size=(300, 3)).astype(np.float32)
    uid, frac, _, _ = w.ray_test_closest(f, t)
    d = (t - f).astype(np.float64)
    best = np.full(len(f), 1.0)
    who = np.zeros(len(f), dtype=int)
    for k in range(n):
        oc = f.astype(np.float64) - centers[k]
        a = (d * d).sum(1); b = 2 * (oc * d).sum(1); c = (oc * oc).sum(1) - float(radii[k]) ** 2
        disc = b * b - 4 * a * c
        s = np.where(disc >= 0, (-b - np.sqrt(np.maximum(disc, 0))) / (2 * a), np.inf)
        s = np.where((c <= 0), np.inf, s)          # starting inside: the cast reports no time of impact we compare against
        ok = (s >= 0) & (s < best)
        best = np.where(ok, s, best); who = np.where(ok, k + 1, who)
    inside = np.zeros(len(f), bool)
    for k in range(n):
        inside |= ((f.astype(np.float64) - centers[k]) ** 2).sum(1) <= float(radii[k]) ** 2 * 1.02
    chk = ~inside
    assert (uid[chk] == who[chk]).mean() > 0.99            # grazing rays may differ by the 1e-4 stopping criterion
    hit = chk & (uid == who) & (who > 0)
    assert hit.sum() > 30 and np.abs(frac[hit] - best[hit]).max() < 5e-3
