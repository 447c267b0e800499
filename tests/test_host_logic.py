"""CPU: host-side logic — layout conversion, scene determinism, bench bookkeeping, BVH build parity between the
product's host builder (through the C ABI, no GPU needed for registration? -> needs a ctx, so compared on GPU) and
pure-python helpers."""
import json
import os
import subprocess
import sys

import numpy as np

import scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_transforms_to_planes_layout(pkg):
    xf = np.arange(36, dtype=np.float32).reshape(3, 12)
    pl = pkg.transforms_to_planes(xf)
    assert pl.shape == (12, 3) and pl.flags.c_contiguous
    assert np.array_equal(pl[9], xf[:, 9]) and np.array_equal(pl[4], xf[:, 4])


def test_scenes_are_deterministic():
    a = scenes.bin_scene(n=300, seed=5)
    b = scenes.bin_scene(n=300, seed=5)
    assert np.array_equal(a.base.view(np.uint32), b.base.view(np.uint32))
    assert np.array_equal(a.transforms(3).view(np.uint32), b.transforms(3).view(np.uint32))
    c = scenes.bin_scene(n=300, seed=6)
    assert not np.array_equal(a.base, c.base)
    assert a.n == 305 and sum(a.static) == 5


def test_trace_keeps_statics_fixed_and_rotations_orthonormal():
    sc = scenes.stack_scene(n_side=3, seed=2)
    xf = sc.transforms(5)
    stat = np.asarray(sc.static)
    assert np.array_equal(xf[stat], sc.base[stat])
    r = xf[:, :9].reshape(-1, 3, 3).astype(np.float64)
    assert np.allclose(np.einsum("nij,nkj->nik", r, r), np.eye(3), atol=1e-5)


def test_heightfield_counts():
    v, t, h = scenes.heightfield(16)
    assert v.shape == (17 * 17, 3) and t.shape == (2 * 16 * 16, 3)
    assert t.max() < len(v) and t.min() >= 0


def test_frame_index_ping_pong():
    import bench
    seq = [bench.frame_index(k) for k in range(30)]
    assert max(seq) == bench.FRAMES - 1 and min(seq) == 0
    assert all(abs(a - b) == 1 for a, b in zip(seq, seq[1:]))


def test_algorithmic_bytes_formula():
    import bench
    st = dict(gjk_checks=1000, deep_penetration_checks=10, large_proxies=5)
    assert bench.algorithmic_bytes("sweep", 100, 500, st, 7, 5, 0) == 9 * 100 * 4 + 100 * 32 + 8 * 500
    assert bench.algorithmic_bytes("sort_pairs", 100, 500, st, 7, 5, 0) == 8 * 100 + 500 * (8 + 4 + 4 + 16)


def test_reference_arm_runs_and_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bodies", "1500", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["metric"] == "collision_phase_world_steps_per_s_100k_bodies"
