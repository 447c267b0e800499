"""Parity checks between the CUDA path (through the C ABI) and the CPU oracle.

Tolerances are the ones BASELINE.json's north_star states:
  pairs      : bit-exact sorted (proxyA, proxyB) list
  contacts   : |dposition| <= 1e-4 * scene extent, normal . normal_ref >= 1 - 1e-5, |ddepth| <= 1e-4
"""
import numpy as np

POS_TOL_REL = 1e-4
NORMAL_DOT_MIN = 1.0 - 1e-5
DEPTH_TOL = 1e-4


def compare_pairs(gp, op):
    gp = np.asarray(gp).reshape(-1, 2)
    op = np.asarray(op).reshape(-1, 2)
    assert gp.shape == op.shape, f"pair count differs: gpu {len(gp)} vs oracle {len(op)}; " + _pair_diff(gp, op)
    assert np.array_equal(gp, op), "pair lists differ: " + _pair_diff(gp, op)


def _pair_diff(gp, op):
    a = set(map(tuple, gp.tolist()))
    b = set(map(tuple, op.tolist()))
    return f"only gpu {sorted(a - b)[:8]} only oracle {sorted(b - a)[:8]}"


def compare_aabbs(ga, oa):
    ga = np.asarray(ga, dtype=np.float32)
    oa = np.asarray(oa, dtype=np.float32)
    assert ga.shape == oa.shape
    same = ga.view(np.uint32) == oa.view(np.uint32)
    # -0.0 == +0.0 is acceptable nowhere on this path: require identical bits
    assert same.all(), f"AABB bits differ at rows {np.unique(np.nonzero(~same)[0])[:8]}"


def compare_raw(graw, oraw, extent):
    """graw: structured RAW_DTYPE; oraw: (ints (n,6), floats (n,7)) from the oracle."""
    oi, of = oraw
    gk = np.stack([graw["uid0"], graw["uid1"], graw["tri"]], axis=1) if len(graw) else np.zeros((0, 3), np.int32)
    ok = oi[:, :3] if len(oi) else np.zeros((0, 3), np.int32)
    go = np.lexsort((gk[:, 2], gk[:, 1], gk[:, 0]))
    oo = np.lexsort((ok[:, 2], ok[:, 1], ok[:, 0]))
    assert len(gk) == len(ok), f"raw contact record count differs: gpu {len(gk)} oracle {len(ok)}"
    assert np.array_equal(gk[go], ok[oo]), "raw contact keys differ"
    g = graw[go]
    oi = oi[oo]
    of = of[oo]
    has_g = g["has_contact"]
    has_o = oi[:, 3]
    bad = np.nonzero(has_g != has_o)[0]
    assert len(bad) == 0, f"has_contact differs for {[(tuple(gk[go][b]), int(has_g[b]), int(has_o[b])) for b in bad[:6]]}"
    m = has_o == 1
    if m.any():
        dn = np.sum(g["normal"][m] * of[m, 0:3], axis=1)
        dp = np.linalg.norm(g["point"][m].astype(np.float64) - of[m, 3:6].astype(np.float64), axis=1)
        dd = np.abs(g["depth"][m].astype(np.float64) - of[m, 6].astype(np.float64))
        assert dn.min() >= NORMAL_DOT_MIN, f"normal dot {dn.min()} at {gk[go][m][np.argmin(dn)]}"
        assert dp.max() <= POS_TOL_REL * extent, f"point delta {dp.max()} at {gk[go][m][np.argmax(dp)]}"
        assert dd.max() <= DEPTH_TOL, f"depth delta {dd.max()} at {gk[go][m][np.argmax(dd)]}"
    return dict(records=int(len(gk)), contacts=int(m.sum()),
                method_match=float(np.mean(g["method"] == oi[:, 4])) if len(gk) else 1.0,
                bit_exact=float(np.mean((g["depth"][m].view(np.uint32) == of[m, 6].view(np.uint32)))) if m.any() else 1.0)


def compare_manifolds(gm, om, extent):
    """gm: structured MANIFOLD_DTYPE (all manifolds, pair order); om: (hdr, pts, pint) from the oracle."""
    hdr, pts, pint = om
    assert len(gm) == len(hdr), f"manifold count differs: gpu {len(gm)} oracle {len(hdr)}"
    if len(hdr) == 0:
        return dict(manifolds=0, points=0)
    assert np.array_equal(gm["pair_uid0"], hdr[:, 0]) and np.array_equal(gm["pair_uid1"], hdr[:, 1]), "manifold pair keys differ"
    assert np.array_equal(gm["body0"], hdr[:, 2]) and np.array_equal(gm["body1"], hdr[:, 3]), "manifold body order differs"
    if hdr.shape[1] >= 7:  # child manifolds of compound pairs: which child of body0's / body1's CompoundShape
        assert np.array_equal(gm["child0"], hdr[:, 5]) and np.array_equal(gm["child1"], hdr[:, 6]), "compound child indices differ"
    bad = np.nonzero(gm["num_contacts"] != hdr[:, 4])[0]
    assert len(bad) == 0, f"num_contacts differs at {[(int(hdr[b,0]), int(hdr[b,1]), int(gm['num_contacts'][b]), int(hdr[b,4])) for b in bad[:6]]}"
    total = 0
    for k in range(4):
        m = hdr[:, 4] > k
        if not m.any():
            continue
        total += int(m.sum())
        gp = gm["points"][m, k]
        op = pts[m, k]
        oi = pint[m, k]
        tol = POS_TOL_REL * extent
        for name, sl in (("local_a", slice(0, 3)), ("local_b", slice(3, 6)), ("world_a", slice(6, 9)), ("world_b", slice(9, 12))):
            d = np.linalg.norm(gp[name].astype(np.float64) - op[:, sl].astype(np.float64), axis=1)
            assert d.max() <= tol, f"{name}[{k}] delta {d.max()}"
        dn = np.sum(gp["normal_on_b"] * op[:, 12:15], axis=1)
        assert dn.min() >= NORMAL_DOT_MIN, f"manifold normal dot {dn.min()}"
        dd = np.abs(gp["distance"].astype(np.float64) - op[:, 15])
        assert dd.max() <= DEPTH_TOL, f"manifold distance delta {dd.max()}"
        assert np.allclose(gp["combined_friction"], op[:, 16]) and np.allclose(gp["combined_restitution"], op[:, 17])
        assert np.array_equal(gp["life_time"], oi[:, 0]), f"lifeTime differs in slot {k}"
        assert np.array_equal(gp["src_slot"], oi[:, 1]), f"src_slot differs in slot {k}"
        assert np.array_equal(gp["index1"], oi[:, 5]), f"triangle index differs in slot {k}"
        assert np.array_equal(gp["part_id1"], oi[:, 3]), f"mesh part id differs in slot {k}"
    return dict(manifolds=int(len(hdr)), points=total)


def step_and_compare(gw, ow, xf, extent, active=None, check_aabbs=True):
    """One performDiscreteCollisionDetection on both sides + all comparisons."""
    gw.setWorldTransforms(xf)
    ow.set_transforms(xf)
    if active is not None:
        gw.setActivation(active)
        ow.set_active(active)
    gw.updateAabbs()
    ow.update_aabbs()
    if check_aabbs:
        compare_aabbs(gw.aabbs(), ow.aabbs())
    gw.getBroadphase().calculateOverlappingPairs()
    op = ow.calculate_overlapping_pairs()
    gp = gw.pairs()
    compare_pairs(gp, op)
    gw.getDispatcher().dispatchAllCollisionPairs()
    ow.dispatch_all_pairs()
    r = compare_raw(gw.raw_contacts(), ow.raw(), extent)
    m = compare_manifolds(gw.manifolds(), ow.manifolds(), extent)
    r.update(m)
    r["pairs"] = int(len(gp))
    return r


# ---- golden fixtures (tests/golden/*.npz, produced by tests/golden/make_golden.py from the oracle) -----------
def golden_cases():
    import importlib.util
    import os
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(here, "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return mg.CASES, here


def compare_gpu_to_golden(gw, sc, gold, steps, extent):
    """Drive the CUDA world through the trace and compare every step with the stored oracle outputs."""
    for step in range(steps):
        gw.setWorldTransforms(sc.transforms(step))
        gw.updateAabbs()
        compare_aabbs(gw.aabbs(), gold[f"aabb{step}"])
        gw.getBroadphase().calculateOverlappingPairs()
        compare_pairs(gw.pairs(), gold[f"pairs{step}"])
        gw.getDispatcher().dispatchAllCollisionPairs()
        oi = np.zeros((len(gold[f"raw_i{step}"]), 6), np.int32)
        oi[:, :5] = gold[f"raw_i{step}"]
        compare_raw(gw.raw_contacts(), (oi, gold[f"raw_f{step}"]), extent)
        compare_manifolds(gw.manifolds(), (gold[f"mf_hdr{step}"], gold[f"mf_pts{step}"], gold[f"mf_int{step}"]), extent)
