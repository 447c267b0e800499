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


# ---- rayTestSingle's concave and compound branches (disp/CollisionWorld.java:301-356) ----------------------------------
def test_ray_against_static_plane_kat():
    """StaticPlaneShape.processAllTriangles builds two triangles from the ray's local AABB (sh/StaticPlaneShape.java:60-122);
    the hit fraction is where the segment crosses the plane, the normal is the (unnormalised) triangle normal facing the ray."""
    w = orc.OracleWorld(orc.TIGHT)
    p = w.plane((0.0, 1.0, 0.0), 0.0)
    w.body(p, orc.xf12(origin=(0, 0, 0)), group=2, mask=-1 ^ 2, static=True)
    uid, frac, nrm, pt = w.ray_test_closest([(1.0, 5.0, 2.0), (1.0, -3.0, 2.0), (1.0, 5.0, 2.0)], [(1.0, -5.0, 2.0), (1.0, 1.0, 2.0), (4.0, 1.0, 2.0)])
    assert uid.tolist() == [1, 1, 0]
    assert abs(frac[0] - 0.5) < 1e-6 and abs(frac[1] - 0.75) < 1e-6 and frac[2] == 1.0
    n0 = nrm[0] / np.linalg.norm(nrm[0]); n1 = nrm[1] / np.linalg.norm(nrm[1])
    assert np.allclose(n0, [0, 1, 0], atol=1e-6) and np.allclose(n1, [0, -1, 0], atol=1e-6)   # from below: the flipped normal
    assert np.allclose(pt[0], [1, 0, 2], atol=1e-6)
    # a plane body moved and tilted: the test runs in the plane's local space, the normal comes back through the body's basis
    rot = [[0, -1, 0], [1, 0, 0], [0, 0, 1]]           # local +y -> world -x
    w2 = orc.OracleWorld(orc.TIGHT)
    w2.body(w2.plane((0.0, 1.0, 0.0), 0.0), orc.xf12(rot, (2.0, 0.0, 0.0)), static=True)
    uid, frac, nrm, pt = w2.ray_test_closest([(-4.0, 0.3, 0.1)], [(8.0, 0.3, 0.1)])
    assert uid.tolist() == [1] and abs(frac[0] - 0.5) < 1e-6
    assert np.allclose(nrm[0] / np.linalg.norm(nrm[0]), [-1, 0, 0], atol=1e-6) and np.allclose(pt[0], [2.0, 0.3, 0.1], atol=1e-5)


def test_ray_against_mesh_matches_brute_force_triangles():
    """performRaycast walks the quantised BVH with rayAabb per node (sh/OptimizedBvh.java:817-931): its answer must be the
    nearest triangle crossing found by testing every triangle (Moeller-Trumbore in float64)."""
    import scenes
    verts, tris, _ = scenes.heightfield(24, cell=0.5, amp=2.0, seed=3)
    w = orc.OracleWorld(orc.TIGHT)
    w.body(w.mesh(verts, tris), orc.xf12(origin=(0, 0, 0)), group=2, mask=-1 ^ 2, static=True)
    rng = np.random.default_rng(12)
    f = rng.uniform(-1, 13, size=(400, 3)).astype(np.float32); f[:, 1] = rng.uniform(2.5, 6, size=400)
    t = rng.uniform(-1, 13, size=(400, 3)).astype(np.float32); t[:, 1] = rng.uniform(-6, -2.5, size=400)
    uid, frac, nrm, pt = w.ray_test_closest(f, t)
    v = verts.astype(np.float64)
    a, b, c = v[tris[:, 0]], v[tris[:, 1]], v[tris[:, 2]]
    best = np.full(len(f), np.inf)
    for r in range(len(f)):
        o = f[r].astype(np.float64); d = t[r].astype(np.float64) - o
        e1, e2 = b - a, c - a
        pv = np.cross(d, e2); det = (e1 * pv).sum(1)
        ok = np.abs(det) > 1e-12
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        tv = o - a
        u = (tv * pv).sum(1) * inv
        qv = np.cross(tv, e1)
        vv = (qv * d).sum(1) * inv
        s = (e2 * qv).sum(1) * inv
        hit = ok & (u >= -1e-9) & (vv >= -1e-9) & (u + vv <= 1 + 1e-9) & (s > 0) & (s < 1)
        if hit.any():
            best[r] = s[hit].min()
    has = np.isfinite(best)
    assert has.sum() > 250
    assert np.array_equal(uid > 0, has)
    assert np.abs(frac[has] - best[has]).max() < 1e-5
    assert (nrm[has][:, 1] > 0).all()                         # rays come from above: the reported normal faces them


def test_ray_against_compound_equals_its_children_as_bodies():
    """The compound branch casts every child with colObjWorldTransform * childTrans (:333-352): with an identity body
    rotation and dyadic offsets that product is exact, so a hit on a child equals the hit on that child placed as a plain body."""
    offs = [(-0.5, 0.0, 0.0), (0.5, 0.0, 0.0), (0.0, 0.75, 0.25)]
    org = (2.0, 1.0, -3.0)
    wa = orc.OracleWorld(orc.TIGHT)
    kids = [wa.sphere(0.3), wa.box(0.3, 0.2, 0.25), wa.sphere(0.4)]
    wa.body(wa.compound(kids, np.stack([orc.xf12(origin=o) for o in offs])), orc.xf12(origin=org))
    wb = orc.OracleWorld(orc.TIGHT)
    kb = [wb.sphere(0.3), wb.box(0.3, 0.2, 0.25), wb.sphere(0.4)]
    for k in range(3):
        wb.body(kb[k], orc.xf12(origin=tuple(np.float32(org[i]) + np.float32(offs[k][i]) for i in range(3))))
    rng = np.random.default_rng(2)
    f = (np.asarray(org) + rng.uniform(-3, 3, size=(300, 3))).astype(np.float32)
    t = (np.asarray(org) + rng.uniform(-1, 1, size=(300, 3))).astype(np.float32)
    ua, fa, na, pa = wa.ray_test_closest(f, t)
    ub, fb, nb, pb = wb.ray_test_closest(f, t)
    assert (ub > 0).sum() > 50
    # the compound is tested behind ITS (larger) AABB, the plain bodies each behind their own: a cast the tight box would
    # have skipped can report a grazing time of impact (SubsimplexConvexCast stops within 1e-2 of the shape), so the
    # compound may see a few more hits — never fewer, and the common ones are bit-identical
    assert not ((ub > 0) & (ua == 0)).any()
    both = (ub > 0) & (fa.view(np.uint32) == fb.view(np.uint32))
    assert both.sum() >= 0.95 * (ub > 0).sum()
    assert np.array_equal(na[both].view(np.uint32), nb[both].view(np.uint32)) and np.array_equal(pa[both].view(np.uint32), pb[both].view(np.uint32))
