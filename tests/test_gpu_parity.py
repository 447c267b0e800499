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
