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
    assert bench.algorithmic_bytes("sweep", 100, 500, st, 7, 0) == 9 * 100 * 4 + 100 * 32 + 8 * 500
    assert bench.algorithmic_bytes("sort_pairs", 100, 500, st, 7, 0) == 8 * 100 + 500 * (8 + 4 + 4 + 16)
    assert bench.algorithmic_bytes("large", 100, 500, st, 7, 0) == 0


def test_both_arms_name_the_workload_identically():
    import bench
    a = bench.workload_name("c2", 100000, 4096)
    assert "100000" in a and a == bench.workload_name("c2", 100000, 1)
    assert bench.max_pairs_for(type("A", (), dict(max_pairs=3 << 20, c5_bodies=1000000, bodies=100000, workload="c2")), "c5") >= 9000000


def test_multiset_hash_is_order_independent_and_sensitive(pkg):
    from libgdx_jbullet_b200.partitioned import MASK64, multiset_hash
    rng = np.random.default_rng(1)
    a = rng.integers(0, 1 << 40, size=(5000, 7), dtype=np.uint64)
    h = multiset_hash(a)
    assert h == multiset_hash(a[rng.permutation(len(a))])
    # a partition of the rows sums (mod 2^64) to the digest of the whole: what partition_check relies on
    assert (multiset_hash(a[:1234]) + multiset_hash(a[1234:])) & MASK64 == h
    b = a.copy(); b[17, 3] ^= np.uint64(1)
    assert multiset_hash(b) != h
    assert multiset_hash(np.concatenate([a, a[:1]])) != h     # a duplicated row changes it
    assert multiset_hash(np.zeros((0, 7), np.uint64)) == 0


def test_reference_arm_runs_and_prints_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--bodies", "1500", "--steps", "2",
                          "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["metric"] == "collision_phase_world_steps_per_s_100k_bodies"
    assert d["steps"] == 2 and d["warmup"] == 1, "the reference arm must honour --steps / --warmup"
    assert "java" in d["cpu_baseline"]["jvm"]          # probed at run time, not a constant
    import bench
    assert d["config"]["workload"] == bench.workload_name("c2", 1500, 4096)


def test_reference_arm_steps_one_world_per_requested_gpu():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--bodies", "1200",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["n_gpus"] == 2 and d["cpu_baseline"]["cores"] == 2
    assert abs(d["value"] - 2 * 1000.0 / d["ms_per_step"]) < 1e-6 * d["value"]   # whole-job rate of 2 worlds on 2 cores


def test_compound_child_table_flattening(tmp_path):
    """csrc/compound_flatten.h + compoundChildWorld (common.cuh) on the host: depth-first leaf order of nested CompoundShapes,
    one frame entry per nested occurrence, leaf world transforms equal to the level-by-level composition bit for bit, and the
    nesting limit (tests/hostcc/flatten_check.cpp; test infrastructure only)."""
    import subprocess
    here = os.path.dirname(os.path.abspath(__file__))
    exe = str(tmp_path / "flatten_check")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-DB2C_HOST_EMULATION", "-I", os.path.join(here, "emu"), "-w",
                           "-o", exe, os.path.join(here, "hostcc", "flatten_check.cpp")])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout + r.stderr
