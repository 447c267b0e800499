"""-m gpu: CompoundShape pairs (compound.cuh) against the oracle's restatement of
disp/CompoundCollisionAlgorithm.java:49-129 and sh/CompoundShape.java:50-160: AABBs and pairs bit-exact, one raw detector
record and one manifold per child algorithm, in the reference's order."""
import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,plane", [(1, True), (0, False), (2, False)])
def test_compound_scene_parity(gpu_pkg, mode, plane):
    sc = scenes.compound_scene(n=300, seed=8, plane_ground=plane)
    kw = dict(world_aabb=((-50.0, -50.0, -50.0), (50.0, 50.0, 50.0))) if mode >= 2 else {}
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode, **kw)
    kids = touching = deep = 0
    for step in range(6):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        m = gw.manifolds()
        kid = m["child0"] >= 0
        kids += int(kid.sum())
        touching += int((m["num_contacts"][kid] > 0).sum())
        nested = (m["child0"] >= 0) & (m["child1"] >= 0)
        assert nested.any(), "the scene must contain compound x compound pairs"
        # getNumManifolds counts every child algorithm's manifold
        assert gw.getDispatcher().getNumManifolds() == len(m) == len(ow.manifolds()[0])
        st = gw.stats()
        deep += st["deep_penetration_checks"]
        assert st["epa_failed"] == 0
    assert kids > 1000 and touching > 50 and deep > 0


def test_compound_activation_and_removal(gpu_pkg):
    """Pairs that are not dispatched (both objects inactive) keep their child manifolds untouched; pairs that leave the cache
    lose them; removed bodies disappear."""
    sc = scenes.compound_scene(n=200, seed=9)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    rng = np.random.default_rng(11)
    for step in range(8):
        active = (rng.uniform(size=sc.n) > 0.4).astype(np.uint8)
        if step == 4:
            for uid in (7, 8, 60):
                gw.removeCollisionObject(uid)
                ow.destroy_body(uid)
        parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent, active=active)


def test_compound_contact_stream_and_fused_step(gpu_pkg):
    """b2c_step (graph replay) + the compact contact streams carry the child manifolds too."""
    sc = scenes.compound_scene(n=250, seed=10)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    for step in range(5):
        xf = sc.transforms(step)
        npairs, nman, nadded = gw.step(np.ascontiguousarray(xf.T))
        op = ow.step(xf)
        parity.compare_pairs(gw.pairs(), op)
        assert npairs == len(op)
        parity.compare_raw(gw.raw_contacts(), ow.raw(), sc.extent)
        parity.compare_manifolds(gw.manifolds(), ow.manifolds(), sc.extent)
        assert nman == len(ow.manifolds()[0])
    m = gw.manifolds(only_touching=True)
    for hdr, pts in (gw.contacts(), gw.solver_contacts()):
        assert len(hdr) == len(m) and hdr["num_contacts"].sum() == len(pts) == m["num_contacts"].sum()
        key = lambda a: sorted(zip(a["pair_uid0"].tolist(), a["pair_uid1"].tolist(), a["body0"].tolist(), a["num_contacts"].tolist()))
        assert key(hdr) == key(m)
        # child manifolds carry their child indices in a negative pair_index
        v = -1 - hdr["pair_index"]
        c0 = np.where(hdr["pair_index"] < 0, (v & 0x7fff) - 1, -1)
        c1 = np.where(hdr["pair_index"] < 0, (v >> 15) - 1, -1)
        full = lambda u0, u1, a, b: sorted(zip(u0.tolist(), u1.tolist(), a.tolist(), b.tolist()))
        assert full(hdr["pair_uid0"], hdr["pair_uid1"], c0, c1) == full(m["pair_uid0"], m["pair_uid1"], m["child0"], m["child1"])
        assert (hdr["pair_index"] < 0).sum() == (m["child0"] >= 0).sum() > 0
        # points of one child manifold: world_a / distance identical to the manifold record
        k = int(np.nonzero(m["child0"] >= 0)[0][0])
        cand = np.nonzero((hdr["pair_uid0"] == m["pair_uid0"][k]) & (hdr["pair_uid1"] == m["pair_uid1"][k]))[0]
        dists = {float(pts["distance"][hdr["first_point"][c]]) for c in cand}
        assert float(m["points"][k, 0]["distance"]) in dists


def test_compound_graph_equals_direct(gpu_pkg, monkeypatch):
    sc = scenes.compound_scene(n=200, seed=12)
    worlds = []
    for env in ({"B2C_GRAPH": "1", "B2C_OVERLAP": "1"}, {"B2C_GRAPH": "0", "B2C_OVERLAP": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        worlds.append(scenes.build_gpu(gpu_pkg, sc, mode=1))
    for step in range(6):
        xf = sc.transforms(step)
        outs = []
        for w in worlds:
            w.setWorldTransforms(xf)
            w.step()
            outs.append((w.pairs().tobytes(), w.manifolds().tobytes(), w.raw_contacts().tobytes()))
        assert outs[0] == outs[1], f"graph / direct results differ at step {step}"


def test_compound_registered_after_first_steps_and_capacity(gpu_pkg):
    """A compound may be registered after the world has stepped (new launch signature); a too small item capacity is
    reported as B2C_ERR_CAPACITY, not as corruption."""
    gw = gpu_pkg.GpuCollisionWorld(mode=0, max_bodies=64, max_pairs=4096, max_compound_items=8)
    s = gw.SphereShape(0.5)
    b = gw.BoxShape((0.4, 0.4, 0.4))
    eye = np.eye(3, dtype=np.float32).reshape(9)
    xf = lambda p: np.concatenate([eye, np.asarray(p, np.float32)])
    gw.addCollisionObject(s, xf((0, 0, 0)))
    gw.addCollisionObject(b, xf((0.6, 0, 0)))
    gw.performDiscreteCollisionDetection()
    assert len(gw.manifolds()) == 1
    c = gw.CompoundShape([s, b, s], np.stack([xf((-0.5, 0, 0)), xf((0, 0, 0)), xf((0.5, 0, 0))]))
    gw.addCollisionObject(c, xf((0.2, 0.7, 0)))
    gw.performDiscreteCollisionDetection()
    m = gw.manifolds()
    assert len(m) == 1 + 3 + 3 and (m["child0"] >= 0).sum() == 6
    gw.addCollisionObject(c, xf((0.3, 1.2, 0)))   # compound x compound: 9 more items -> over the capacity of 8
    with pytest.raises(gpu_pkg.B2CError) as e:
        gw.performDiscreteCollisionDetection()
    assert e.value.code == -3
    pl = gw.StaticPlaneShape((0.0, 1.0, 0.0), 0.0)
    with pytest.raises(gpu_pkg.B2CError):          # children must be convex shapes or compounds
        gw.CompoundShape([pl], np.stack([xf((0, 0, 0))]))


def test_compounds_on_a_triangle_mesh(gpu_pkg):
    """compound x BvhTriangleMeshShape: ConvexConcave per child, per-triangle raw records keyed -2 - (child << 21 | triangle),
    the child manifolds folded in BVH order with the triangle index in index1."""
    sc = scenes.terrain_compound_scene(cells=32, n=150, seed=15)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    touching = 0
    for step in range(5):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        m = gw.manifolds()
        kid_mesh = (m["child0"] >= 0) & (m["algorithm"] == 4)
        touching += int((m["num_contacts"][kid_mesh] > 0).sum())
        assert gw.getDispatcher().getNumManifolds() == len(m)
        assert gw.stats()["epa_failed"] == 0
    assert kid_mesh.sum() > 150 and touching > 100


def test_nested_compounds_parity(gpu_pkg):
    """Children that are CompoundShapes themselves: the leaves are visited depth first with transforms composed level by level
    ((org * frame) * child), as the reference's nested CompoundCollisionAlgorithms do — AABBs, pairs, one raw record and one
    manifold per leaf combination, bit for bit against the oracle; rays and sweeps through them too."""
    sc = scenes.compound_scene(n=260, seed=9, plane_ground=False, nested=True)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1, max_pairs=1 << 15)
    kids = 0
    for step in range(5):
        parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        m = gw.manifolds()
        kids += int((m["child0"] >= 0).sum())
        assert int(m["child0"].max()) >= 4, "the two-level compound has 5 leaves, more than any flat one of the scene"
        assert gw.getDispatcher().getNumManifolds() == len(m) == len(ow.manifolds()[0])
    assert kids > 3000
    rng = np.random.default_rng(2)
    ext = float(sc.extent)
    f = rng.uniform(-1.0, ext, size=(400, 3)).astype(np.float32); f[:, 1] = rng.uniform(2.0, 8.0, size=400)
    t = rng.uniform(-1.0, ext, size=(400, 3)).astype(np.float32); t[:, 1] = rng.uniform(-1.0, 1.0, size=400)
    gu, gf, gn, gp = gw.rayTestClosest(f, t)
    ou, of, on, op = ow.ray_test_closest(f, t)
    assert np.array_equal(gu, ou) and np.array_equal(gf.view(np.uint32), of.view(np.uint32)) and (gu > 0).sum() > 200
    cast = gw.SphereShape(0.2), ow.sphere(0.2)
    eye = np.eye(3, dtype=np.float32)
    su, sf, _, sp = gw.convexSweepTestClosest(cast[0], eye, f, t, 1, 1)
    qu, qf, _, qp = ow.convex_sweep_closest(cast[1], eye, f, t, 1, 1)
    assert np.array_equal(su, qu) and np.array_equal(sf.view(np.uint32), qf.view(np.uint32)) and np.array_equal(sp.view(np.uint32), qp.view(np.uint32))


def test_identity_wrapped_compounds_equal_flat_ones(gpu_pkg):
    """An outer compound with ONE child at the identity transform changes nothing: multiplying by the identity is exact in
    binary32, so pairs, AABBs and every contact bit must equal the flat scene's — a check that does not involve the oracle."""
    flat = scenes.compound_scene(n=200, seed=10, plane_ground=True)
    wrap = scenes.compound_scene(n=200, seed=10, plane_ground=True, nested="identity")
    assert np.array_equal(flat.base, wrap.base)
    g0 = scenes.build_gpu(gpu_pkg, flat, mode=1)
    g1 = scenes.build_gpu(gpu_pkg, wrap, mode=1)
    for step in range(4):
        for g, sc in ((g0, flat), (g1, wrap)):
            g.setWorldTransforms(sc.transforms(step)); g.step()
        assert np.array_equal(g0.pairs(), g1.pairs())
        m0, m1 = g0.manifolds(), g1.manifolds()
        assert m0.tobytes() == m1.tobytes()
    assert len(m0) > 500


def test_compound_nesting_depth_is_limited(gpu_pkg):
    gw = gpu_pkg.GpuCollisionWorld(mode=1, max_bodies=16, max_pairs=64)
    sid = gw.SphereShape(0.2)
    xf = scenes.make_xf(np.eye(3)[None], np.zeros((1, 3)))
    for level in range(5):                       # leaf -> 5 compounds around it = 4 frames above the leaf: still fine
        sid = gw.CompoundShape([sid], xf)
    with pytest.raises(Exception, match="nested deeper"):
        gw.CompoundShape([sid], xf)
