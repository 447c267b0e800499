"""-m gpu: the AxisSweep3 pair semantics (broadphase modes SAP16 / SAP32) through the C ABI against the
oracle's stateless quantised predicate (itself checked against the literal algorithm in tests/test_oracle_sap.py)."""
import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu

WORLD = ((-60.0, -20.0, -60.0), (60.0, 100.0, 60.0))


@pytest.mark.parametrize("mode", [2, 3])
def test_sap_bin_scene_parity(gpu_pkg, mode):
    sc = scenes.bin_scene(n=2500, seed=41)
    sc.vel *= 3.0
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode, world_aabb=WORLD)
    for step in range(5):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 2500 and r["contacts"] > 100


def test_sap16_coarse_grid_changes_the_pair_set_and_still_matches(gpu_pkg):
    """A huge world box makes the 16-bit grid coarse (quantum ~0.3): many more pairs than the float test finds, and
    bodies outside the box are clamped onto its faces — both sides must agree on every one of them."""
    sc = scenes.bin_scene(n=1500, seed=43)
    big = ((-10000.0, -10000.0, -10000.0), (10000.0, 10000.0, 10000.0))
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=2, world_aabb=big, max_pairs=1 << 18)
    tight = scenes.build_oracle(sc, 0)
    for step in range(3):
        xf = sc.transforms(step)
        r = parity.step_and_compare(gw, ow, xf, sc.extent)
        nt = len(tight.step(xf))
    assert r["pairs"] > 1.2 * nt


def test_sap_clamped_outside_world_box(gpu_pkg):
    sc = scenes.spheres_scene(n=3000, seed=12, fill=0.3)
    small = ((2.0, 2.0, 2.0), (9.0, 9.0, 9.0))     # most of the cube lies outside: those boxes collapse onto the faces
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=3, world_aabb=small, max_pairs=1 << 22)
    for step in range(2):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 10000


def test_sap_default_world_box_and_batched_worlds(gpu_pkg):
    sc = scenes.worlds_scene(num_worlds=24, seed=8)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=2)     # default box +-1000 on both sides
    for step in range(3):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 24 * 60
    mn, mx = gw.getBroadphase().getBroadphaseAabb()
    assert mn.tolist() == [-1000.0] * 3 and mx.tolist() == [1000.0] * 3
