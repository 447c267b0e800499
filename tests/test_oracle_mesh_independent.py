"""CPU: independent pins of the oracle's triangle-mesh path (sh/OptimizedBvh.java box query, disp/ConvexTriangleCallback.java
per-triangle detector) against brute-force numpy computations that share no code with the restatement."""
import numpy as np

import orc


def test_bvh_aabb_query_is_a_tight_superset_of_the_brute_force_overlaps():
    """sh/OptimizedBvh.java:709-740, 940-997: the quantised walk must report every triangle whose box overlaps the query box
    (quantisation only ever grows boxes) and nothing farther away than one quantisation step."""
    import scenes
    verts, tris, _ = scenes.heightfield(32, cell=0.5, amp=2.0, seed=9)
    w = orc.OracleWorld(orc.TIGHT)
    mesh = w.mesh(verts, tris)
    _, q = w.mesh_nodes(mesh)
    step = 1.0 / q[6:9]                                 # one quantisation step per axis
    tv = verts[tris]                                    # (T, 3, 3)
    tmin, tmax = tv.min(axis=1), tv.max(axis=1)
    rng = np.random.default_rng(3)
    for _ in range(60):
        c = rng.uniform((0, -2, 0), (16, 2, 16))
        h = rng.uniform(0.05, 1.5, size=3)
        mn, mx = (c - h).astype(np.float32), (c + h).astype(np.float32)
        got = set(w.bvh_query(mesh, mn, mx).tolist())
        exact = set(np.nonzero(((tmin <= mx) & (tmax >= mn)).all(axis=1))[0].tolist())
        assert exact <= got
        slack = 2.0 * step + 0.002                      # rounding of both boxes + the 0.002 padding of flat triangle boxes
        loose = set(np.nonzero(((tmin - slack <= mx) & (tmax + slack >= mn)).all(axis=1))[0].tolist())
        assert got <= loose


def _point_triangle_distance(p, a, b, c):
    """Closest distance from point p to triangle abc (float64, Ericson's region walk) — independent of the oracle."""
    ab, ac, ap = b - a, c - a, p - a
    d1, d2 = ab @ ap, ac @ ap
    if d1 <= 0 and d2 <= 0:
        return np.linalg.norm(ap)
    bp = p - b
    d3, d4 = ab @ bp, ac @ bp
    if d3 >= 0 and d4 <= d3:
        return np.linalg.norm(bp)
    vc = d1 * d4 - d3 * d2
    if vc <= 0 and d1 >= 0 and d3 <= 0:
        return np.linalg.norm(ap - ab * (d1 / (d1 - d3)))
    cp = p - c
    d5, d6 = ab @ cp, ac @ cp
    if d6 >= 0 and d5 <= d6:
        return np.linalg.norm(cp)
    vb = d5 * d2 - d1 * d6
    if vb <= 0 and d2 >= 0 and d6 <= 0:
        return np.linalg.norm(ap - ac * (d2 / (d2 - d6)))
    va = d3 * d6 - d5 * d4
    if va <= 0 and (d4 - d3) >= 0 and (d5 - d6) >= 0:
        w = (d4 - d3) / ((d4 - d3) + (d5 - d6))
        return np.linalg.norm(p - (b + (c - b) * w))
    n = np.cross(ab, ac)
    return abs(ap @ n) / np.linalg.norm(n)


def test_sphere_vs_mesh_triangle_contacts_match_point_triangle_distance():
    """ConvexConcave + per-triangle GJK (disp/ConvexTriangleCallback.java:111-172): for a sphere (core = its centre, margin =
    radius, triangle margin 0) the detector's distance must be dist(centre, triangle) - radius, and a contact exists exactly
    when that is below the manifold's breaking threshold (up to the detector's slack)."""
    import scenes
    verts, tris, h = scenes.heightfield(16, cell=0.5, amp=1.0, seed=5)
    rng = np.random.default_rng(8)
    checked = 0
    for trial in range(25):
        w = orc.OracleWorld(orc.TIGHT)
        w.body(w.mesh(verts, tris), orc.xf12(origin=(0, 0, 0)), group=2, mask=-1 ^ 2, static=True)
        r = float(np.float32(rng.uniform(0.25, 0.5)))
        x, z = rng.uniform(1.0, 7.0, size=2)
        y = h[int(x / 0.5), int(z / 0.5)] + r + rng.uniform(-0.05, 0.05)
        c = np.asarray([x, y, z], dtype=np.float32)
        w.body(w.sphere(r), orc.xf12(origin=c))
        w.step()
        ri, rf = w.raw()
        assert len(ri) > 0
        for k in range(len(ri)):
            t = tris[ri[k, 2]]
            d = _point_triangle_distance(c.astype(np.float64), *[verts[i].astype(np.float64) for i in t]) - r
            if ri[k, 3]:
                assert abs(rf[k, 6] - d) < 2e-4, (trial, k, rf[k, 6], d)
                checked += 1
            else:
                assert d > 0.02 - 2e-3, (trial, k, d)
    assert checked > 20


def test_multi_part_mesh_is_the_same_surface():
    """Splitting one IndexedMesh into several parts changes the ids a contact reports (partId, index inside the part), not
    the geometry: same pairs, same contact points; and the BVH leaves carry partId << 21 | index (sh/OptimizedBvh.java:278)."""
    import scenes
    sc1 = scenes.terrain_scene(cells=24, n=120, seed=31)
    sc3 = scenes.terrain_scene(cells=24, n=120, seed=31)
    scenes.split_mesh_into_parts(sc3, nparts=3, short_parts=(0, 2))
    o1 = scenes.build_oracle(sc1, 1)
    o3 = scenes.build_oracle(sc3, 1)
    n1, q1 = o1.mesh_nodes(0)
    n3, q3 = o3.mesh_nodes(0)
    assert np.array_equal(q1, q3) and np.array_equal(n1[:, :3], n3[:, :3])       # same boxes, same tree shape
    leaf = n1[:, 3] >= 0
    cuts = np.linspace(0, 2 * 24 * 24, 4).astype(int)
    part = np.searchsorted(cuts, n1[leaf, 3], side="right") - 1
    assert np.array_equal(n3[leaf, 3], (part << 21) | (n1[leaf, 3] - cuts[part]))
    for step in range(3):
        p1 = o1.step(sc1.transforms(step))
        p3 = o3.step(sc3.transforms(step))
        assert np.array_equal(p1, p3)
        h1, pt1, i1 = o1.manifolds()
        h3, pt3, i3 = o3.manifolds()
        assert np.array_equal(h1, h3) and np.array_equal(pt1.view(np.uint32), pt3.view(np.uint32))
        live = np.arange(4)[None, :] < h1[:, 4][:, None]
        mesh = (i1[:, :, 2] == -1) & live                                          # partId0 == -1 marks a mesh contact
        g = i1[:, :, 5][mesh]
        pp = np.searchsorted(cuts, g, side="right") - 1
        assert np.array_equal(i3[:, :, 3][mesh], pp) and np.array_equal(i3[:, :, 5][mesh], g - cuts[pp])
    assert mesh.sum() > 20
