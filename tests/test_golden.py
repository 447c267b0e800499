"""Golden fixtures: CPU — the oracle still reproduces them bit for bit; GPU — the CUDA path matches them through
the C ABI without executing anything under oracle/."""
import os

import numpy as np
import pytest

import parity
import scenes

CASES, GOLD_DIR = parity.golden_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_golden(name):
    make, mode, steps = CASES[name]
    gold = np.load(os.path.join(GOLD_DIR, name + ".npz"))
    sc = make()
    ow = scenes.build_oracle(sc, mode)
    for step in range(steps):
        ow.set_transforms(sc.transforms(step))
        ow.update_aabbs()
        assert np.array_equal(ow.aabbs().view(np.uint32), gold[f"aabb{step}"].view(np.uint32))
        assert np.array_equal(ow.calculate_overlapping_pairs(), gold[f"pairs{step}"])
        ow.dispatch_all_pairs()
        ints, fl = ow.raw()
        order = np.lexsort((ints[:, 2], ints[:, 1], ints[:, 0]))
        assert np.array_equal(ints[order][:, :5], gold[f"raw_i{step}"])
        assert np.array_equal(fl[order].view(np.uint32), gold[f"raw_f{step}"].view(np.uint32))
        hdr, pts, pint = ow.manifolds()
        assert np.array_equal(hdr[:, :5], gold[f"mf_hdr{step}"][:, :5]) and np.array_equal(pint, gold[f"mf_int{step}"])
        assert np.array_equal(pts.view(np.uint32), gold[f"mf_pts{step}"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_matches_golden(gpu_pkg, name):
    make, mode, steps = CASES[name]
    gold = np.load(os.path.join(GOLD_DIR, name + ".npz"))
    sc = make()
    gw = scenes.build_gpu(gpu_pkg, sc, mode=mode)
    parity.compare_gpu_to_golden(gw, sc, gold, steps, sc.extent)
