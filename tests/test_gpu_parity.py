"""-m gpu: parity of the CUDA path (through the C ABI) against the CPU oracle on seeded scenes."""
import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
def test_c1_stack_parity(gpu_pkg, mode):
    sc = scenes.stack_scene(n_side=5, extra=True, seed=1)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode)
    tot = 0
    for step in range(12):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        tot += r["contacts"]
    assert r["pairs"] > 300
    assert tot > 0


def test_c1_plane_ground(gpu_pkg):
    sc = scenes.stack_scene(n_side=3, extra=True, seed=2, plane_ground=True)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    for step in range(6):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["manifolds"] > 0


@pytest.mark.parametrize("mode", [0, 1])
def test_c2_bin_parity_small(gpu_pkg, mode):
    sc = scenes.bin_scene(n=3000, seed=3)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode)
    for step in range(5):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 3000 and r["contacts"] > 100


def test_activation_and_removal(gpu_pkg):
    sc = scenes.bin_scene(n=800, seed=4)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    rng = np.random.default_rng(5)
    for step in range(8):
        active = (rng.uniform(size=sc.n) > 0.3).astype(np.uint8)
        if step == 4:
            for uid in (10, 11, 300):
                gw.removeCollisionObject(uid)
                ow.destroy_body(uid)
        parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent, active=active)


def test_c3_terrain_mesh_parity(gpu_pkg):
    """convex-vs-BvhTriangleMeshShape: BVH built by the product's host builder must equal the oracle's node array
    bit for bit, then pairs / per-triangle contacts / folded manifolds."""
    sc = scenes.terrain_scene(cells=48, n=400, seed=4)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    gn, gq = gw.mesh_bvh(0)
    on, oq = ow.mesh_nodes(0)
    assert np.array_equal(gq.view(np.uint32), oq.view(np.uint32)), "quantisation parameters differ"
    assert np.array_equal(gn, on), "BVH node arrays differ"
    for step in range(5):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["records"] > 1000 and r["contacts"] > 50
    assert gw.stats()["mesh_items"] > 1000


def test_c4_batched_worlds_parity(gpu_pkg):
    sc = scenes.worlds_scene(num_worlds=48, seed=5)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    for step in range(4):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 48 * 300
    p = gw.pairs()
    w = np.asarray(sc.world)
    assert (w[p[:, 0] - 1] == w[p[:, 1] - 1]).all(), "a pair crosses worlds"


@pytest.mark.parametrize("mode", [0, 1])
def test_c5_spheres_parity(gpu_pkg, mode):
    sc = scenes.spheres_scene(n=30000, seed=6)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode)
    for step in range(3):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 30000


def test_host_supplied_aabbs_drop_in(gpu_pkg):
    """BroadphaseInterface.setAabb drop-in: AABBs computed by the host (here: by the oracle) instead of b2c_update_aabbs."""
    sc = scenes.bin_scene(n=1200, seed=8)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    import orc
    tight = scenes.build_oracle(sc, orc.TIGHT)  # a second oracle world only to produce the tight AABBs a Java host would
    for step in range(4):
        xf = sc.transforms(step)
        tight.set_transforms(xf); tight.update_aabbs()
        t = tight.aabbs()
        ow.set_transforms(xf); ow.update_aabbs()
        gw.setWorldTransforms(xf)
        gw.getBroadphase().setAabbs(None, t[:, :3], t[:, 3:])
        parity.compare_aabbs(gw.aabbs(), ow.aabbs())
        gw.getBroadphase().calculateOverlappingPairs()
        parity.compare_pairs(gw.pairs(), ow.calculate_overlapping_pairs())


def test_empty_and_tiny_worlds(gpu_pkg):
    gw = gpu_pkg.GpuCollisionWorld(max_bodies=16, max_pairs=64)
    assert gw.getBroadphase().calculateOverlappingPairs() == 0
    gw.getDispatcher().dispatchAllCollisionPairs()
    assert len(gw.pairs()) == 0 and len(gw.manifolds()) == 0
    s = gw.SphereShape(1.0)
    import orc
    gw.addCollisionObject(s, orc.xf12())
    gw.performDiscreteCollisionDetection()
    assert len(gw.pairs()) == 0
    gw.addCollisionObject(s, orc.xf12(origin=(1.5, 0, 0)))
    gw.performDiscreteCollisionDetection()
    assert gw.pairs().tolist() == [[1, 2]]
    m = gw.manifolds()
    assert m["num_contacts"][0] == 1 and abs(m["points"][0, 0]["distance"] + 0.5) < 1e-7
    assert np.allclose(m["points"][0, 0]["normal_on_b"], [-1, 0, 0])


def test_pair_capacity_overflow_is_reported(gpu_pkg):
    sc = scenes.bin_scene(n=600, seed=3)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1, max_pairs=512)
    gw.setWorldTransforms(sc.transforms(0))
    gw.updateAabbs()
    with pytest.raises(gpu_pkg.B2CError) as e:
        gw.getBroadphase().calculateOverlappingPairs()
    assert e.value.code == -3 and "capacity" in str(e.value)


def test_contact_stream_matches_manifolds(gpu_pkg):
    sc = scenes.bin_scene(n=2000, seed=12)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1)
    for step in range(3):
        gw.setWorldTransforms(sc.transforms(step))
        gw.performDiscreteCollisionDetection()
    m = gw.manifolds(only_touching=True)
    hdr, pts = gw.contacts()
    assert len(hdr) == len(m) and len(pts) == int(m["num_contacts"].sum())
    order = np.lexsort((hdr["pair_uid1"], hdr["pair_uid0"]))
    hdr = hdr[order]
    assert np.array_equal(hdr["pair_uid0"], m["pair_uid0"]) and np.array_equal(hdr["num_contacts"], m["num_contacts"])
    k = 17 % len(hdr)
    first = hdr["first_point"][k]
    assert np.array_equal(pts[first]["world_b"], m["points"][k, 0]["world_b"])


def _check_packed_stream(gw, stream=None):
    """b2c_get_packed_contacts carries exactly the touching manifolds: every field bit-identical to b2c_get_manifolds."""
    m = gw.manifolds(only_touching=True)
    pairs = gw.pairs()
    hdr, pts = gw.packed_contacts() if stream is None else stream
    assert len(hdr) == len(m) and len(pts) == int(m["num_contacts"].sum())
    nc = hdr["info"] & 0xff
    alg = (hdr["info"] >> 8) & 0xff
    swapped = (hdr["info"] >> 16) & 1
    c0 = (hdr["children"] & 0xffff).astype(np.int16).astype(np.int32)
    c1 = (hdr["children"] >> 16).astype(np.int32)
    uid0, uid1 = pairs[hdr["pair_index"], 0], pairs[hdr["pair_index"], 1]
    order = np.lexsort((c1, c0, uid1, uid0))
    morder = np.lexsort((m["child1"], m["child0"], m["pair_uid1"], m["pair_uid0"]))
    mm = m[morder]
    assert np.array_equal(uid0[order], mm["pair_uid0"]) and np.array_equal(uid1[order], mm["pair_uid1"])
    assert np.array_equal(c0[order], mm["child0"]) and np.array_equal(c1[order], mm["child1"])
    assert np.array_equal(nc[order], mm["num_contacts"]) and np.array_equal(alg[order], mm["algorithm"])
    assert np.array_equal(swapped[order] == 1, mm["body0"] != mm["pair_uid0"])
    for k in range(4):
        sel = nc[order] > k
        pp = pts[hdr["first_point"][order][sel] + k]
        ref = mm["points"][sel, k]
        for f in ("world_a", "world_b", "normal_on_b", "distance"):
            assert np.array_equal(pp[f].view(np.uint32), ref[f].view(np.uint32)), f
        assert np.array_equal(pp["life_src"] >> 8, ref["life_time"]) and np.array_equal((pp["life_src"] & 0xff) - 1, ref["src_slot"])
        assert np.array_equal(pp["index1"], ref["index1"])
    return len(hdr)


def test_packed_contact_stream(gpu_pkg):
    sc = scenes.bin_scene(n=2000, seed=12)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1)
    for step in range(4):
        gw.setWorldTransforms(sc.transforms(step))
        gw.performDiscreteCollisionDetection()
    assert _check_packed_stream(gw) > 500
    # mesh pairs (triangle index) and compound pairs (child indices, pair index of the child manifolds)
    sc = scenes.terrain_scene(cells=32, n=150, seed=6)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1)
    for step in range(3):
        gw.setWorldTransforms(sc.transforms(step))
        gw.performDiscreteCollisionDetection()
    assert _check_packed_stream(gw) > 50
    sc = scenes.compound_scene(n=250, seed=10)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1)
    for step in range(3):
        gw.setWorldTransforms(sc.transforms(step))
        gw.performDiscreteCollisionDetection()
    assert _check_packed_stream(gw) > 100
    # compaction behind the dispatch (b2c_set_contact_prefetch): the getter only copies; another format still works
    gw.set_contact_prefetch(2)
    for step in range(3, 7):
        gw.step(np.ascontiguousarray(sc.transforms(step).T))
        assert _check_packed_stream(gw) > 100
        h1, p1 = gw.solver_contacts()
        h2, p2 = gw.packed_contacts()
        assert len(h1) == len(h2) and len(p1) == len(p2)
    # the two-part download: the manifolds that are final before the penetration bin ends start travelling early
    import ctypes as C
    _lib = gpu_pkg._lib
    hbuf = np.zeros(1 << 16, dtype=_lib.PACKED_HEADER_DTYPE)
    pbuf = np.zeros(1 << 17, dtype=_lib.PACKED_POINT_DTYPE)
    for step in range(7, 10):
        gw.setWorldTransforms(sc.transforms(step))
        gw.step_device()
        gw._ck(gw.L.b2c_begin_contact_download(gw.h, hbuf.ctypes.data_as(C.c_void_p), len(hbuf), pbuf.ctypes.data_as(C.c_void_p), len(pbuf)))
        gw.sync_counts()
        nh, npt = C.c_int32(), C.c_int32()
        gw._ck(gw.L.b2c_get_packed_contacts(gw.h, hbuf.ctypes.data_as(C.c_void_p), len(hbuf), pbuf.ctypes.data_as(C.c_void_p), len(pbuf),
                                            C.byref(nh), C.byref(npt)))
        assert _check_packed_stream(gw, (hbuf[: nh.value].copy(), pbuf[: npt.value].copy())) > 100
    gw.set_contact_prefetch(-1)
    with pytest.raises(gpu_pkg.B2CError):
        gw.step_device()
        gw._ck(gw.L.b2c_begin_contact_download(gw.h, hbuf.ctypes.data_as(C.c_void_p), len(hbuf), pbuf.ctypes.data_as(C.c_void_p), len(pbuf)))
    gw.sync_counts()
    gw.step(np.ascontiguousarray(sc.transforms(10).T))
    assert _check_packed_stream(gw) > 100


def test_c2_full_size_properties(gpu_pkg):
    """BASELINE size (100k bodies): size-independent properties instead of a full oracle run."""
    import bench
    sc = bench.make_scene(100000, seed=100)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1, max_pairs=3 << 20)
    prev = None
    for step in range(3):
        gw.setWorldTransforms(sc.transforms(step))
        gw.performDiscreteCollisionDetection()
        p = gw.pairs()
        assert (p[:, 0] < p[:, 1]).all()
        key = p[:, 0].astype(np.int64) << 32 | p[:, 1]
        assert (np.diff(key) > 0).all(), "pairs not strictly sorted / not unique"
        # every reported pair overlaps on its effective AABBs and passes the filter; a sample of non-pairs does not
        a = gw.aabbs()
        i, j = p[:, 0] - 1, p[:, 1] - 1
        assert ((a[i, :3] <= a[j, 3:]).all(axis=1) & (a[j, :3] <= a[i, 3:]).all(axis=1)).all()
        rng = np.random.default_rng(step)
        ii = rng.integers(5, sc.n, size=200000); jj = np.minimum(ii + rng.integers(1, 60, size=200000), sc.n - 1)
        ok = ii != jj
        ii, jj = ii[ok], jj[ok]
        ov = (a[ii, :3] <= a[jj, 3:]).all(axis=1) & (a[jj, :3] <= a[ii, 3:]).all(axis=1)
        k2 = np.minimum(ii, jj).astype(np.int64) + 1 << 32 | (np.maximum(ii, jj) + 1)
        assert np.array_equal(np.isin(k2, key), ov), "sampled overlap predicate disagrees with the pair list"
        hdr, pts = gw.contacts()
        assert (pts["distance"] <= 0.02 + 1e-7).all()
        assert np.allclose(np.linalg.norm(pts["normal_on_b"], axis=1), 1.0, atol=1e-4)
        # after refreshContactPoints: distance1 = (worldA - worldB) . normal (np/PersistentManifold.java:332-335)
        proj = np.sum((pts["world_a"].astype(np.float64) - pts["world_b"]) * pts["normal_on_b"], axis=1)
        assert np.allclose(proj, pts["distance"], atol=2e-5)
        st = gw.stats()
        assert st["epa_failed"] == 0
    # idempotence: stepping again with unchanged transforms keeps the pair list
    p1 = gw.pairs()
    gw.performDiscreteCollisionDetection()
    gw.performDiscreteCollisionDetection()
    p2 = gw.pairs()
    gw.performDiscreteCollisionDetection()
    assert np.array_equal(p2, gw.pairs())
