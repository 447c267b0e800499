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


# ---- queries behind the pair list: rays, convex sweeps, CCD sweeps on a stepped world ---------------------------------
def _query_cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLD_DIR, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg


def _compare_queries(out, gold):
    for q in ("ray", "sweep", "ccd"):
        assert np.array_equal(out[q + "_uid"], gold[q + "_uid"]), f"{q}: hit bodies differ"
        assert np.array_equal(out[q + "_frac"].view(np.uint32), gold[q + "_frac"].view(np.uint32)), f"{q}: fractions differ"
        hit = gold[q + "_uid"] > 0
        assert np.array_equal(out[q + "_nrm"][hit].view(np.uint32), gold[q + "_nrm"][hit].view(np.uint32)), f"{q}: normals differ"
        assert np.array_equal(out[q + "_pt"][hit].view(np.uint32), gold[q + "_pt"][hit].view(np.uint32)), f"{q}: points differ"
        assert hit.sum() >= 5, f"{q}: the fixture should contain hits"


@pytest.mark.parametrize("k,name", list(enumerate(sorted(_query_cases().QUERY_CASES))))
def test_oracle_reproduces_query_golden(k, name):
    mg = _query_cases()
    make, mode, steps = mg.QUERY_CASES[name]
    sc = make()
    _compare_queries(mg.run_queries(scenes.build_oracle(sc, mode), sc, steps, k, True), np.load(os.path.join(GOLD_DIR, name + ".npz")))


@pytest.mark.gpu
@pytest.mark.parametrize("k,name", list(enumerate(sorted(_query_cases().QUERY_CASES))))
def test_cuda_matches_query_golden(gpu_pkg, k, name):
    """Rays, convex sweeps and CCD sweeps of the CUDA path against stored oracle outputs (nothing under oracle/ runs)."""
    mg = _query_cases()
    make, mode, steps = mg.QUERY_CASES[name]
    sc = make()
    gw = scenes.build_gpu(gpu_pkg, sc, mode=mode, max_pairs=1 << 15)
    _compare_queries(mg.run_queries(gw, sc, steps, k, False), np.load(os.path.join(GOLD_DIR, name + ".npz")))


# ---- vectors from the REAL reference, when somebody has produced them (tools/javaref/run.sh on a box with a JDK + gdx jar) ----
JAVA_CASES = sorted(n for n in CASES if os.path.exists(os.path.join(GOLD_DIR, "java_" + n + ".npz")))


def _compare_with_java(get_step, gold, steps, extent, exact):
    """get_step(k) -> (aabbs, pairs, (hdr, pts, pint)) of our side.  Pairs are compared exactly; AABBs bit for bit; manifolds
    by the north_star tolerances (bit for bit when `exact`: the oracle claims the reference's float sequences).  The Java dump
    carries no src_slot (not a reference field) and child indices -1, so those columns are skipped."""
    for k in range(steps):
        aabb, pairs, (hdr, pts, pint) = get_step(k)
        assert np.array_equal(aabb.view(np.uint32), gold[f"aabb{k}"].view(np.uint32)), f"AABBs differ from the Java reference at step {k}"
        parity.compare_pairs(pairs, gold[f"pairs{k}"])
        jh, jp, ji = gold[f"mf_hdr{k}"], gold[f"mf_pts{k}"], gold[f"mf_int{k}"]
        plain = hdr[:, 5] < 0 if hdr.shape[1] >= 7 else np.ones(len(hdr), bool)   # child manifolds of compounds: matched by pair only
        assert np.array_equal(hdr[plain][:, :5], jh[: len(hdr)][plain][:, :5]) if len(jh) == len(hdr) else False, f"manifold headers differ at step {k}"
        live = np.arange(4)[None, :] < hdr[:, 4][:, None]
        assert np.array_equal(pint[:, :, 0][live], ji[:, :, 0][live]), "lifeTime differs"
        assert np.array_equal(pint[:, :, 2:][live], ji[:, :, 2:][live]), "shape identifiers differ"
        if exact:
            assert np.array_equal(pts.view(np.uint32)[live], jp.view(np.uint32)[live]), f"contact point bits differ at step {k}"
        else:
            d = np.abs(pts[live].astype(np.float64) - jp[live].astype(np.float64))
            assert d[:, :12].max(initial=0) <= parity.POS_TOL_REL * extent and d[:, 15].max(initial=0) <= parity.DEPTH_TOL
            assert np.sum(pts[live][:, 12:15] * jp[live][:, 12:15], axis=1).min(initial=1) >= parity.NORMAL_DOT_MIN


@pytest.mark.skipif(not JAVA_CASES, reason="no tests/golden/java_*.npz: the reference has not been run (tools/javaref/run.sh needs a JDK + gdx jar)")
@pytest.mark.parametrize("name", JAVA_CASES or ["none"])
def test_oracle_matches_the_java_reference(name):
    """THE pin: the oracle against vectors dumped from the reference's own CollisionWorld."""
    make, mode, steps = CASES[name]
    gold = np.load(os.path.join(GOLD_DIR, "java_" + name + ".npz"))
    sc = make()
    ow = scenes.build_oracle(sc, mode)

    def step(k):
        ow.set_transforms(sc.transforms(k)); ow.update_aabbs()
        a = ow.aabbs().copy()
        p = ow.calculate_overlapping_pairs()
        ow.dispatch_all_pairs()
        return a, p, ow.manifolds()
    _compare_with_java(step, gold, steps, sc.extent, exact=True)


@pytest.mark.gpu
@pytest.mark.skipif(not JAVA_CASES, reason="no tests/golden/java_*.npz: the reference has not been run (tools/javaref/run.sh needs a JDK + gdx jar)")
@pytest.mark.parametrize("name", JAVA_CASES or ["none"])
def test_cuda_matches_the_java_reference(gpu_pkg, name):
    make, mode, steps = CASES[name]
    gold = np.load(os.path.join(GOLD_DIR, "java_" + name + ".npz"))
    sc = make()
    gw = scenes.build_gpu(gpu_pkg, sc, mode=mode)

    def step(k):
        gw.setWorldTransforms(sc.transforms(k)); gw.updateAabbs()
        a = gw.aabbs()
        gw.getBroadphase().calculateOverlappingPairs()
        p = gw.pairs()
        gw.getDispatcher().dispatchAllCollisionPairs()
        m = gw.manifolds()
        hdr = np.stack([m["pair_uid0"], m["pair_uid1"], m["body0"], m["body1"], m["num_contacts"], m["child0"], m["child1"]], axis=1)
        P = m["points"]
        pts = np.concatenate([P["local_a"], P["local_b"], P["world_a"], P["world_b"], P["normal_on_b"], P["distance"][..., None],
                              P["combined_friction"][..., None], P["combined_restitution"][..., None]], axis=-1).astype(np.float32)
        zero = np.zeros_like(P["life_time"])
        mesh = (m["algorithm"] == 4)[:, None] * np.ones_like(zero)
        pint = np.stack([P["life_time"], P["src_slot"], np.where(mesh, -1, 0), P["part_id1"], np.where(mesh, -1, 0), P["index1"]], axis=-1)
        return a, p, (hdr, pts, pint)
    _compare_with_java(step, gold, steps, sc.extent, exact=False)


def test_exported_scenes_match_the_golden_cases():
    """tools/javaref/export_scenes.py writes what the Java driver reads: one scene file per golden case, same bodies, same
    per-step transforms (so a java_<case>.npz is comparable with <case>.npz step by step)."""
    sdir = os.path.join(GOLD_DIR, "scenes")
    for name, (make, mode, steps) in CASES.items():
        z = np.load(os.path.join(sdir, name + ".npz"))
        sc = make()
        assert int(z["mode"][0]) == mode and int(z["steps"][0]) == steps and len(z["body_shape"]) == sc.n
        assert np.array_equal(z[f"xf{steps - 1}"].view(np.uint32), sc.transforms(steps - 1).view(np.uint32))
        assert len(z["shape_kind"]) == len(sc.shapes)
