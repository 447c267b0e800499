"""CPU: differential pins of the oracle.

* sort-and-sweep pair finding == independent O(N^2) loop (both modes);
* the effective-AABB state machine (+ "all filtered overlaps of eff") == a literal restatement of the
  reference's Dbvt tree + DbvtBroadphase stage lists + pair cache, over multi-step traces with bodies going
  to sleep, waking, teleporting and being removed.
"""
import numpy as np
import pytest

import orc
import scenes


def run_trace(sc, mode, brute=False, steps=8, seed=0, sleepy=True, removals=()):
    ow = scenes.build_oracle(sc, mode, brute_force=brute)
    rng = np.random.default_rng(seed)
    out = []
    asleep = np.zeros(sc.n, dtype=bool)
    for step in range(steps):
        xf = sc.transforms(step)
        if step in (3, 6):  # teleport a few bodies far away and back
            idx = np.arange(7, min(sc.n, 60), 9)
            xf[idx, 9:] += np.float32(7.5 if step == 3 else 0.0)
        ow.set_transforms(xf)
        if sleepy:
            flip = rng.uniform(size=sc.n) < 0.25
            asleep ^= flip
            ow.set_active((~asleep).astype(np.uint8))
        for (s, uid) in removals:
            if s == step:
                ow.destroy_body(uid)
        ow.update_aabbs()
        aabbs = ow.aabbs().copy()
        pairs = ow.calculate_overlapping_pairs()
        out.append((aabbs, pairs))
    return out


@pytest.mark.parametrize("mode", [orc.TIGHT, orc.DBVT])
def test_sweep_equals_brute_force(mode):
    sc = scenes.bin_scene(n=500, seed=21)
    a = run_trace(sc, mode, brute=False, seed=1)
    b = run_trace(sc, mode, brute=True, seed=1)
    for (aa, pa), (ab, pb) in zip(a, b):
        assert np.array_equal(aa.view(np.uint32), ab.view(np.uint32))
        assert np.array_equal(pa, pb)
    assert len(a[-1][1]) > 500


@pytest.mark.parametrize("scene", ["bin", "stack", "worlds"])
def test_state_machine_equals_literal_dbvt(scene):
    if scene == "bin":
        sc = scenes.bin_scene(n=700, seed=22)
    elif scene == "stack":
        sc = scenes.stack_scene(n_side=4, seed=23)
    else:
        sc = scenes.worlds_scene(num_worlds=6, seed=24)
    removals = ((4, 12), (4, 13), (5, 40))
    a = run_trace(sc, orc.DBVT, steps=10, seed=2, removals=removals)
    b = run_trace(sc, orc.DBVT_LITERAL, steps=10, seed=2, removals=removals)
    for k, ((aa, pa), (ab, pb)) in enumerate(zip(a, b)):
        alive = np.ones(sc.n, dtype=bool)
        for (s, uid) in removals:
            if s <= k:
                alive[uid - 1] = False
        assert np.array_equal(aa[alive].view(np.uint32), ab[alive].view(np.uint32)), f"effective AABBs differ at step {k}"
        assert np.array_equal(pa, pb), f"pair sets differ at step {k}: {len(pa)} vs {len(pb)}"
    assert len(a[-1][1]) > 100


def test_literal_all_active_no_sleep():
    sc = scenes.bin_scene(n=400, seed=25)
    a = run_trace(sc, orc.DBVT, steps=6, sleepy=False)
    b = run_trace(sc, orc.DBVT_LITERAL, steps=6, sleepy=False)
    for (aa, pa), (ab, pb) in zip(a, b):
        assert np.array_equal(pa, pb)
