"""-m gpu: parity of the CUDA path against the CPU oracle at the FULL size of every BASELINE.json config.

The small-scene tests (test_gpu_parity.py) cover the edge cases; these run the configs the numbers are quoted on:
  C1  125 boxes + ground, DbvtBroadphase, 300 steps
  C2  100 005 proxies (the settled snapshot bench.py measures), 3 steps
  C3  10 001 proxies on a 1 002 528-triangle BvhTriangleMeshShape, 2 steps (BVH node array compared bit for bit)
  C4  4096 worlds x 64 bodies = 262 144 proxies, 2 steps
  C5  1 000 000 spheres, one world, 2 steps
Every step compares what parity.step_and_compare compares everywhere else: AABB bits, the sorted (uid0, uid1) pair list
(bit-exact), the raw detector records and the persistent manifolds (north_star tolerances, lifetimes and slots exact).
The oracle (tests/orc.py -> oracle/) is the checker only; the product path is libb2c.so through the C ABI.
"""
import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu


def _bench():
    import bench
    return bench


def test_c1_full_300_steps(gpu_pkg):
    sc = scenes.stack_scene(n_side=5, extra=False, seed=1)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1)
    deep = contacts = 0
    # the trace is replayed incrementally (Scene.transforms(step) recomputes the spin product from scratch)
    r64 = sc.base[:, :9].astype(np.float64).reshape(-1, 3, 3)
    stat = np.asarray(sc.static, dtype=bool)
    for step in range(300):
        xf = sc.base.copy()
        if step:
            r64 = np.einsum("nij,njk->nik", sc.spin, r64)
            xf[:, :9] = r64.reshape(-1, 9).astype(np.float32)
            xf[:, 9:] = (sc.base[:, 9:].astype(np.float64) + sc.vel * step).astype(np.float32)
            xf[stat] = sc.base[stat]
        r = parity.step_and_compare(gw, ow, xf, sc.extent)
        contacts += r["contacts"]
        deep += gw.stats()["deep_penetration_checks"]
    assert r["pairs"] > 300 and contacts > 10000
    assert deep > 0, "300 steps of drift must reach the penetration solver"


def test_c2_full_100k(gpu_pkg):
    bench = _bench()
    sc = bench.make_scene(100000, seed=100)
    assert bench.load_settled(sc, 100000, 100, 60), "committed settled snapshot missing"
    sc.vel *= 0.25
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1, max_pairs=3 << 20)
    for step in range(3):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert sc.n == 100005
    assert r["pairs"] > 800000 and r["contacts"] > 200000 and r["manifolds"] == r["pairs"]
    assert r["bit_exact"] == 1.0, "contact depth bits differ from the oracle"


def test_c3_full_1m_triangles(gpu_pkg):
    bench = _bench()
    sc = bench.make_scene(10000, seed=100, workload="c3")
    assert len(sc.shapes[0][2]) == 1002528
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1, max_pairs=1 << 20, max_mesh_items=1 << 21)
    gn, gq = gw.mesh_bvh(0)
    on, oq = ow.mesh_nodes(0)
    assert np.array_equal(gq.view(np.uint32), oq.view(np.uint32)), "quantisation parameters differ"
    assert np.array_equal(gn, on), "BVH node arrays differ"
    retries = 0
    for step in range(2):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        retries += gw.stats()["epa_retries"]
    st = gw.stats()
    assert st["mesh_items"] > 80000 and r["records"] > 80000 and r["contacts"] > 5000
    assert st["deep_penetration_checks"] > 100, "C3 must exercise the penetration solver on (hull, triangle) items"
    assert retries > 0, "C3 must exercise the large-pool retry tier"


def test_c4_full_4096_worlds(gpu_pkg):
    bench = _bench()
    sc = bench.make_scene(0, seed=100, workload="c4", worlds=4096)
    assert sc.n == 262144 and sc.num_worlds == 4096
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1, max_pairs=3 << 20)
    for step in range(2):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 1500000
    p = gw.pairs()
    w = np.asarray(sc.world)
    assert (w[p[:, 0] - 1] == w[p[:, 1] - 1]).all(), "a pair crosses worlds"


def test_c5_full_1m_spheres(gpu_pkg):
    bench = _bench()
    sc = bench.make_scene(1000000, seed=100, workload="c5")
    import orc
    gw = scenes.build_gpu(gpu_pkg, sc, mode=1, max_pairs=10 << 20)  # step 0 fattens every proxy: ~6.5 M pairs
    # the oracle in its LITERAL DbvtBroadphase mode (the reference's own tree, O(N log N)); its stateless single-axis sweep
    # is quadratic in a 100^3 lattice
    ow = scenes.build_oracle(sc, orc.DBVT_LITERAL)
    first = None
    for step in range(2):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
        first = first or r["pairs"]
    assert sc.n == 1000000 and first > 6000000 and r["pairs"] > 1000000 and r["contacts"] > 100000
