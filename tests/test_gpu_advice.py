"""-m gpu: regressions for the round-1 advisor findings and the A1 overflow guard."""
import numpy as np
import pytest

import orc
import parity
import scenes

pytestmark = pytest.mark.gpu


def test_set_aabb_for_a_subset_and_update_for_the_rest(gpu_pkg):
    """BroadphaseInterface.setAabb on some proxies and CollisionWorld.updateAabbs on the others are independent
    (bp/SimpleBroadphase.java:112-116, disp/CollisionWorld.java:231-245): one step must honour both."""
    sc = scenes.bin_scene(n=900, seed=21)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0)
    rng = np.random.default_rng(3)
    for step in range(4):
        xf = sc.transforms(step)
        gw.setWorldTransforms(xf)
        ow.set_transforms(xf)
        ow.update_aabbs()
        base = ow.aabbs()
        sub = np.sort(rng.choice(np.arange(6, sc.n + 1), size=120, replace=False)).astype(np.int32)
        grow = rng.uniform(0.0, 0.3, size=(len(sub), 1)).astype(np.float32)
        mn = base[sub - 1, :3] - grow
        mx = base[sub - 1, 3:] + grow
        for k, uid in enumerate(sub):
            ow.set_aabb(int(uid), mn[k], mx[k])
        gw.getBroadphase().setAabbs(sub, mn, mx)
        gw.updateAabbs()
        parity.compare_aabbs(gw.aabbs(), ow.aabbs())
        gw.getBroadphase().calculateOverlappingPairs()
        parity.compare_pairs(gw.pairs(), ow.calculate_overlapping_pairs())
    assert len(gw.pairs()) > 900


def test_destroyed_slots_are_recycled_when_the_table_is_full(gpu_pkg):
    """max_bodies bounds the LIVE proxies: a full table hands out the slots of proxies destroyed before the last pair
    calculation, and the newcomer inherits nothing from the proxy that had the uid before."""
    gw = gpu_pkg.GpuCollisionWorld(mode=0, max_bodies=8, max_pairs=256)
    s = gw.SphereShape(0.5)
    for k in range(8):
        assert gw.addCollisionObject(s, orc.xf12(origin=(0.8 * k, 0, 0))) == k + 1
    gw.performDiscreteCollisionDetection()
    assert gw.pairs().tolist() == [[k, k + 1] for k in range(1, 8)]
    with pytest.raises(gpu_pkg.B2CError) as e:
        gw.addCollisionObject(s, orc.xf12(origin=(0, 5, 0)))
    assert e.value.code == -3
    gw.removeCollisionObject(3)
    gw.removeCollisionObject(6)
    with pytest.raises(gpu_pkg.B2CError):   # not before a pair calculation has dropped the dead proxies' pairs
        gw.addCollisionObject(s, orc.xf12(origin=(0, 5, 0)))
    for _ in range(3):                      # manifolds of the survivors age
        gw.performDiscreteCollisionDetection()
    assert gw.pairs().tolist() == [[1, 2], [4, 5], [7, 8]]
    # lowest free slot first; the new proxy sits between 1 and 2 and touches both
    assert gw.addCollisionObject(s, orc.xf12(origin=(0.4, 0.3, 0))) == 3
    assert gw.addCollisionObject(s, orc.xf12(origin=(100, 0, 0))) == 6
    gw.performDiscreteCollisionDetection()
    assert gw.pairs().tolist() == [[1, 2], [1, 3], [2, 3], [4, 5], [7, 8]]
    m = gw.manifolds()
    life = {(int(a), int(b)): int(p[0]["life_time"]) for a, b, p in zip(m["pair_uid0"], m["pair_uid1"], m["points"])}
    assert life[(1, 2)] > life[(1, 3)] and life[(1, 3)] == life[(2, 3)], life
    with pytest.raises(gpu_pkg.B2CError):
        gw.addCollisionObject(s, orc.xf12(origin=(0, 9, 0)))


def _column_scene(n=24):
    """A column of unit boxes resting face to face; `_column_tilt` rocks every other box onto each of its four bottom corners
    in turn, so every pair collects a full 4-point manifold."""
    sc = scenes.Scene()
    b = sc.add_shape("box", (1.0, 1.0, 1.0))
    for k in range(n):
        sc.body_shape.append(b); sc.static.append(k == 0); sc.group.append(2 if k == 0 else 1)
        sc.mask.append(-1 ^ 2 if k == 0 else -1); sc.world.append(0)
    sc.base = scenes.make_xf(np.tile(np.eye(3), (n, 1, 1)), np.asarray([(0.0, 1.0 + 2.0 * k, 0.0) for k in range(n)]))
    sc.extent = 2.0 * n
    return sc


def _column_tilt(sc, step, a=0.004):
    ax, az = [(a, a), (a, -a), (-a, -a), (-a, a)][step % 4]
    rx = np.array([[1, 0, 0], [0, np.cos(ax), -np.sin(ax)], [0, np.sin(ax), np.cos(ax)]])
    rz = np.array([[np.cos(az), -np.sin(az), 0], [np.sin(az), np.cos(az), 0], [0, 0, 1]])
    rot = np.tile(np.eye(3), (sc.n, 1, 1))
    rot[1::2] = rx @ rz
    return scenes.make_xf(rot, sc.base[:, 9:])


def test_contact_stream_holds_four_points_per_pair(gpu_pkg):
    """Every pair of a rocking column ends with a 4-point manifold: with max_pairs == the pair count the contact stream must
    still hold 4 points per pair (it was sized 2 x max_pairs)."""
    sc = _column_scene(24)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1, max_pairs=23)
    for step in range(6):
        r = parity.step_and_compare(gw, ow, _column_tilt(sc, step), sc.extent)
    assert r["pairs"] == 23 and r["points"] == 92
    for getter in (gw.contacts, gw.solver_contacts, gw.packed_contacts):
        hdr, pts = getter()
        assert len(hdr) == 23 and len(pts) == 92


def test_aabb_overflow_guard_disables_the_object(gpu_pkg):
    """disp/CollisionWorld.java:212-218: a non-static object whose AABB diagonal^2 reaches 1e12 is taken out of the
    simulation (DISABLE_SIMULATION): its broadphase AABB is not updated any more and it stops being dispatched."""
    sc = scenes.bin_scene(n=300, seed=22)
    huge = sc.add_shape("box", (6.0e5, 1.0, 1.0))
    sc.body_shape.append(huge); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    row = scenes.make_xf(np.eye(3)[None], np.asarray([(1.0, 2.0, 1.0)]))
    sc.base = np.concatenate([sc.base, row], axis=0)
    sc.vel = np.concatenate([sc.vel, np.zeros((1, 3))], axis=0)
    sc.spin = np.concatenate([sc.spin, np.eye(3)[None]], axis=0)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    for step in range(4):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    a = gw.aabbs()
    created = a[-1].copy()
    xf = sc.transforms(4)
    xf[-1, 9:] += 50.0          # the disabled object no longer follows its transform
    parity.step_and_compare(gw, ow, xf, sc.extent)
    assert np.array_equal(gw.aabbs()[-1], created)
    assert r["pairs"] > 300


def test_long_rows_switch_to_the_radix_passes(gpu_pkg):
    """A world strung out along the sweep axis puts thousands of proxies into one grid row: the row-grouped ordering is
    quadratic there, so the library falls back to the radix passes from the next step on — same pairs either way."""
    n = 6000
    sc = scenes.Scene()
    s = sc.add_shape("sphere", 0.5)
    rng = np.random.default_rng(5)
    pos = np.stack([np.arange(n) * 0.9, rng.uniform(-0.05, 0.05, n), rng.uniform(-0.05, 0.05, n)], axis=1)
    for _ in range(n):
        sc.body_shape.append(s); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    sc.base = scenes.make_xf(np.tile(np.eye(3), (n, 1, 1)), pos)
    sc.vel = rng.uniform(-0.01, 0.01, size=(n, 3))
    sc.extent = float(n)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    launches = []
    for step in range(4):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        launches.append(gw.stats()["kernel_launches"])
    assert r["pairs"] >= n - 1
    assert launches[-1] != launches[0], "the ordering pipeline did not change after the long row was seen"


def test_multi_part_mesh_with_16_bit_indices(gpu_pkg):
    """N12: a TriangleIndexVertexArray with several IndexedMesh parts, one of them with ScalarType.SHORT indices
    (sh/ByteBufferVertexData.java:75-84, sh/TriangleIndexVertexArray.java:72-100): the BVH leaves name (partId, index), the node
    array equals the oracle's bit for bit, and contact points report the part and the triangle inside it."""
    sc = scenes.terrain_scene(cells=40, n=350, seed=23)
    scenes.split_mesh_into_parts(sc, nparts=3, short_parts=(1,))
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    gn, gq = gw.mesh_bvh(0)
    on, oq = ow.mesh_nodes(0)
    assert np.array_equal(gq.view(np.uint32), oq.view(np.uint32)) and np.array_equal(gn, on), "BVH differs"
    leaves = gn[gn[:, 3] >= 0, 3]
    assert set(np.unique(leaves >> 21).tolist()) == {0, 1, 2}
    parts_seen = set()
    for step in range(4):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        m = gw.manifolds(only_touching=True)
        for k in range(4):
            sel = (m["algorithm"] == 4) & (m["num_contacts"] > k)
            parts_seen |= set(np.unique(m["points"][sel, k]["part_id1"]).tolist())
        hdr, pts = gw.packed_contacts()
        assert len(pts) == int(m["num_contacts"].sum())
    assert r["records"] > 800 and r["contacts"] > 40
    assert parts_seen == {0, 1, 2}, parts_seen
    # rays against the multi-part mesh decode the leaf word the same way
    rng = np.random.default_rng(2)
    size = 40 * 0.5
    frm = np.stack([rng.uniform(1, size - 1, 200), np.full(200, 30.0), rng.uniform(1, size - 1, 200)], axis=1).astype(np.float32)
    to = frm.copy(); to[:, 1] = -30.0
    gu, gf, gnrm, gpt = gw.rayTestClosest(frm, to)
    ou, of, onrm, opt = ow.ray_test_closest(frm, to)
    assert np.array_equal(gu, ou) and np.array_equal(gf.view(np.uint32), of.view(np.uint32))


def test_raw_records_are_an_opt_in_inspection_channel(gpu_pkg):
    """b2c_set_raw_records: off by default — pairs, manifolds and the contact stream do not depend on it, and
    b2c_get_raw_contacts says why it has nothing to return."""
    sc = scenes.bin_scene(n=1500, seed=61)
    on = scenes.build_gpu(gpu_pkg, sc, mode=1)                       # the test helpers switch the channel on
    off = scenes.build_gpu(gpu_pkg, sc, mode=1, raw_records=False)   # the library default
    for step in range(3):
        for w in (on, off):
            w.setWorldTransforms(sc.transforms(step)); w.step()
        assert np.array_equal(on.pairs(), off.pairs())
        assert on.manifolds().tobytes() == off.manifolds().tobytes()
        h0, p0 = on.contacts(); h1, p1 = off.contacts()
        assert len(h0) == len(h1) and len(p0) == len(p1)
    assert len(on.raw_contacts()) == len(on.pairs())
    with pytest.raises(gpu_pkg.B2CError, match="inspection channel"):
        off.raw_contacts()
