"""CPU: the AxisSweep3 restatements (oracle/sap_literal.h) — quantiser KATs from bp/AxisSweep3Internal.java:201-216 and the
differential that justifies the stateless definition: the literal incremental algorithm (edge arrays, insertion sorts,
index-order overlap tests) ends every step with exactly the pairs whose QUANTISED boxes overlap."""
import numpy as np
import pytest

import orc
import scenes

WORLD = ((-60.0, -20.0, -60.0), (60.0, 100.0, 60.0))


def test_quantize_kat_16bit():
    w = orc.OracleWorld(orc.SAP16, world_aabb=((-100, -100, -100), (100, 100, 100)))
    # quantize = 65535 / 200 ; min edges even, max edges odd; clamped to the world box
    assert w.sap_quantize((-100, -100, -100), 0).tolist() == [0, 0, 0]
    assert w.sap_quantize((-100, -100, -100), 1).tolist() == [1, 1, 1]
    assert w.sap_quantize((100, 100, 100), 1).tolist() == [65535, 65535, 65535]
    assert w.sap_quantize((100, 100, 100), 0).tolist() == [65534, 65534, 65534]
    assert w.sap_quantize((1e9, -1e9, 0), 0).tolist() == [65534, 0, int(np.float32(100) * (np.float32(65535) / np.float32(200))) & 0xfffe]
    q = int(np.float32(np.float32(12.5) - np.float32(-100)) * (np.float32(65535) / np.float32(200)))
    assert w.sap_quantize((12.5, 12.5, 12.5), 0).tolist() == [q & 0xfffe] * 3
    assert w.sap_quantize((12.5, 12.5, 12.5), 1).tolist() == [(q & 0xfffe) | 1] * 3


def test_quantize_kat_32bit_saturates_like_java():
    w = orc.OracleWorld(orc.SAP32, world_aabb=((0, 0, 0), (1, 1, 1)))
    # (float)0x7fffffff rounds to 2^31, so the top of the world box maps to 2^31 -> Java's (int) cast saturates
    assert w.sap_quantize((1, 1, 1), 1).tolist() == [0x7fffffff] * 3
    assert w.sap_quantize((1, 1, 1), 0).tolist() == [0x7ffffffe] * 3
    assert w.sap_quantize((0.5, 0.5, 0.5), 0).tolist() == [0x40000000] * 3


@pytest.mark.parametrize("modes", [(orc.SAP16, orc.SAP16_LITERAL), (orc.SAP32, orc.SAP32_LITERAL)])
@pytest.mark.parametrize("scene", ["bin", "stack", "cluster"])
def test_stateless_predicate_equals_literal_axis_sweep(modes, scene):
    if scene == "bin":
        sc = scenes.bin_scene(n=700, seed=31)
        sc.vel *= 6.0
    elif scene == "stack":
        sc = scenes.stack_scene(n_side=4, extra=True, seed=3)
        sc.vel *= 8.0
    else:
        rng = np.random.default_rng(5)
        sc = scenes.spheres_scene(n=400, seed=9, fill=0.5)
        sc.vel = rng.uniform(-0.3, 0.3, size=(400, 3))   # fast: many edge crossings on all axes per step
    a = scenes.build_oracle(sc, modes[0], world_aabb=WORLD)
    b = scenes.build_oracle(sc, modes[1], world_aabb=WORLD)
    rng = np.random.default_rng(7)
    removed = set()
    tot = 0
    for step in range(10):
        xf = sc.transforms(step)
        if step in (3, 6):
            for uid in rng.choice(np.arange(2, sc.n), size=5, replace=False):
                if int(uid) not in removed:
                    removed.add(int(uid))
                    a.destroy_body(int(uid)); b.destroy_body(int(uid))
        active = (rng.uniform(size=sc.n) > 0.2).astype(np.uint8)   # some bodies skip their setAabb
        for w in (a, b):
            w.set_transforms(xf)
            w.set_active(active)
            w.update_aabbs()
        pa = a.calculate_overlapping_pairs()
        pb = b.calculate_overlapping_pairs()
        assert np.array_equal(pa, pb), f"{scene} step {step}: stateless {len(pa)} literal {len(pb)}"
        tot += len(pa)
    assert tot > 0


def test_sap_pairs_differ_from_float_pairs_only_by_quantisation():
    """Sanity: quantised boxes are the float boxes widened to the grid, so SAP pairs ⊇ most tight pairs and the
    difference is small for a fine 16-bit grid over a 120-unit world."""
    sc = scenes.bin_scene(n=600, seed=2)
    s = scenes.build_oracle(sc, orc.SAP16, world_aabb=WORLD)
    t = scenes.build_oracle(sc, orc.TIGHT)
    ps = set(map(tuple, s.step(sc.transforms(0)).tolist()))
    pt = set(map(tuple, t.step(sc.transforms(0)).tolist()))
    assert len(pt - ps) <= 0.01 * len(pt) + 2 and len(ps - pt) <= 0.05 * len(pt) + 5
