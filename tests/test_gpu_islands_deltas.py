"""-m gpu: the two consumers right behind the pair list against the CPU oracle:
pair add/remove deltas (ghost pair callback events) and simulation islands (union-find partition)."""
import numpy as np
import pytest

import parity
import scenes
from test_oracle_islands import canon_partition

pytestmark = pytest.mark.gpu


def _check(gw, ow, sc, steps, active_fn=None):
    seen_added = seen_removed = 0
    for step in range(steps):
        parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        ga, gr = gw.pair_deltas()
        oa, orm = ow.pair_deltas()
        assert np.array_equal(ga, oa), f"added pairs differ at step {step}: gpu {len(ga)} oracle {len(oa)}"
        assert np.array_equal(gr, orm), f"removed pairs differ at step {step}: gpu {len(gr)} oracle {len(orm)}"
        seen_added += len(ga)
        seen_removed += len(gr)
        gt, gn = gw.islands()
        ot, on = ow.islands()
        assert gn == on, f"island count differs at step {step}: {gn} vs {on}"
        assert np.array_equal(gt, canon_partition(ot)), f"island partition differs at step {step}"
        assert np.array_equal(gt, canon_partition(gt)), "device tags are not the smallest member index"
    return seen_added, seen_removed


def test_deltas_and_islands_bin(gpu_pkg):
    sc = scenes.bin_scene(n=3000, seed=21)
    sc.vel *= 3.0
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    a, r = _check(gw, ow, sc, 6)
    assert a > 3000 and r > 0   # step 0 adds everything; later steps must both add and remove


def test_deltas_and_islands_tight_mode_stack(gpu_pkg):
    sc = scenes.stack_scene(n_side=4, extra=True, seed=5)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0)
    _check(gw, ow, sc, 5)


def test_islands_batched_worlds(gpu_pkg):
    sc = scenes.worlds_scene(num_worlds=48, seed=3)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    _check(gw, ow, sc, 3)
    gt, gn = gw.islands()
    world = np.asarray(sc.world)
    dyn = gt >= 0
    # an island never spans two worlds
    assert np.array_equal(world[dyn], world[gt[dyn]])
    assert gn >= sc.num_worlds


def test_islands_sparse_spheres_many_islands(gpu_pkg):
    sc = scenes.spheres_scene(n=6000, seed=2, fill=0.12)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0)
    _check(gw, ow, sc, 2)
    _, gn = gw.islands()
    assert gn > 100


def test_solver_contact_stream_matches_full_stream(gpu_pkg):
    """b2c_get_solver_contacts is a 64-byte projection of b2c_get_contacts: same manifolds, same field bits."""
    sc = scenes.bin_scene(n=2500, seed=17)
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1)
    for step in range(3):
        gw.setWorldTransforms(sc.transforms(step))
        gw.step()
    h1, p1 = gw.contacts()
    h2, p2 = gw.solver_contacts()
    assert len(h1) == len(h2) > 100 and len(p1) == len(p2)
    o1 = np.lexsort((h1["pair_uid1"], h1["pair_uid0"]))
    o2 = np.lexsort((h2["pair_uid1"], h2["pair_uid0"]))
    for f in ("pair_uid0", "pair_uid1", "body0", "body1", "num_contacts", "algorithm", "pair_index"):
        assert np.array_equal(h1[f][o1], h2[f][o2]), f
    for a, b in zip(o1, o2):
        n = h1["num_contacts"][a]
        q1 = p1[h1["first_point"][a]:h1["first_point"][a] + n]
        q2 = p2[h2["first_point"][b]:h2["first_point"][b] + n]
        for f in ("world_a", "world_b", "normal_on_b", "distance", "combined_friction", "combined_restitution", "life_time",
                  "src_slot", "part_id1", "index1"):
            assert q1[f].tobytes() == q2[f].tobytes(), f


def test_constraint_linked_pairs_are_not_dispatched(gpu_pkg):
    """disp/CollisionDispatcher.java:216-218 + dynamics/RigidBody.java:624-639: linked bodies stay in the pair cache but
    get no algorithm; an existing manifold is left as it is (never refreshed)."""
    sc = scenes.bin_scene(n=1500, seed=23)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    r = parity.step_and_compare(gw, ow, sc.transforms(0), sc.extent)
    pairs = gw.pairs()
    hdr, _ = gw.contacts()
    touching = np.stack([hdr["pair_uid0"], hdr["pair_uid1"]], axis=1)
    # link 40 touching pairs (reversed order on purpose) and 40 arbitrary overlapping pairs
    rng = np.random.default_rng(1)
    link = np.concatenate([touching[rng.choice(len(touching), 40, replace=False)][:, ::-1],
                           pairs[rng.choice(len(pairs), 40, replace=False)]])
    gw.setNoCollidePairs(link)
    ow.set_no_collide_pairs(link)
    for step in range(1, 4):
        r2 = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r2["records"] < r["records"]          # fewer detector runs than with everything dispatched
    gw.setNoCollidePairs(np.zeros((0, 2), np.int32))
    ow.set_no_collide_pairs(np.zeros((0, 2), np.int32))
    r3 = parity.step_and_compare(gw, ow, sc.transforms(4), sc.extent)
    assert r3["records"] > r2["records"]


def test_graph_replay_equals_direct_launches(gpu_pkg, monkeypatch):
    """The step replayed as a CUDA graph (default) and issued kernel by kernel (B2C_GRAPH=0), with and without the side
    streams, must leave identical pair lists, manifolds and raw records."""
    sc = scenes.bin_scene(n=2000, seed=29)
    worlds = []
    for env in ({"B2C_GRAPH": "1", "B2C_OVERLAP": "1"}, {"B2C_GRAPH": "0", "B2C_OVERLAP": "1"}, {"B2C_GRAPH": "0", "B2C_OVERLAP": "0"}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)       # read by b2c_create
        worlds.append(scenes.build_gpu(gpu_pkg, sc, mode=1))
    for step in range(6):                  # enough steps for both ping-pong parities to be replayed
        xf = sc.transforms(step)
        outs = []
        for w in worlds:
            w.setWorldTransforms(xf)
            w.step()
            raw = w.raw_contacts()
            raw = raw[np.lexsort((raw["uid1"], raw["uid0"]))]
            outs.append((w.pairs().tobytes(), w.manifolds().tobytes(), raw.tobytes()))
        assert outs[0] == outs[1] == outs[2], f"graph / direct / single-stream results differ at step {step}"
