"""CPU: an independent pin of the oracle's GJK / EPA restatement (np/GjkPairDetector.java, np/GjkEpaSolver.java).

The reference ships no golden vectors, so the detector output is checked against a separately computed quantity: the signed
distance of the two margin-rounded convex shapes written with support functions only,
    separation(A, B) = max over unit d of  -(h_A(d) + h_B(-d)),        h_X(d) = max over x in X of d . x  (+ margin),
which is the separating distance when positive and minus the penetration depth when negative.  It is evaluated by dense
direction sampling plus a local refinement (scipy), never by GJK."""
import numpy as np
from scipy.optimize import minimize

import orc
import scenes


SHARP_BOX = False  # the penetration solver sees BoxShape.localGetSupportingVertex: the SHARP full-size box (sh/BoxShape.java:73-86)


def _support(kind, param, rot, pos, margin, d):
    """h(d) of the world-space shape: core support + margin (sphere: core is a point, margin = radius)."""
    dl = d @ rot                       # R^T d for every direction (rows)
    if kind == "box" and SHARP_BOX:
        return np.abs(dl) @ np.asarray(param, np.float64) + d @ pos
    if kind == "box":
        core = np.abs(dl) @ (np.asarray(param, np.float64) - margin)
    elif kind == "sphere":
        core = np.zeros(len(d))
    else:
        core = (dl @ np.asarray(param, np.float64).T).max(axis=1)
    return core + d @ pos + margin


def _separation(a, b, n_dirs=40000, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n_dirs, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    f = -(_support(*a, d) + _support(*b, -d))
    best = d[np.argsort(f)[-5:]]

    def neg(x):
        v = x / np.linalg.norm(x)
        return float(_support(*a, v[None])[0] + _support(*b, -v[None])[0])

    vals = [(-minimize(neg, x0, method="Nelder-Mead", options=dict(xatol=1e-9, fatol=1e-10, maxiter=4000)).fun) for x0 in best]
    return max(max(vals), float(f.max()))


def test_gjk_and_epa_depth_match_the_support_function_distance():
    rng = np.random.default_rng(42)
    hp = scenes.hull_points(rng, 0.5)
    checked_sep = checked_pen = quirks = 0
    for trial in range(60):
        w = orc.OracleWorld(orc.TIGHT)
        kinds = []
        for _ in range(2):
            k = rng.integers(3)
            if k == 0:
                he = rng.uniform(0.3, 0.6, size=3).astype(np.float32)
                kinds.append(("box", he, w.box(*[float(v) for v in he]), 0.04))
            elif k == 1:
                r = float(np.float32(rng.uniform(0.3, 0.6)))
                kinds.append(("sphere", r, w.sphere(r), r))
            else:
                pts = (hp * rng.uniform(0.7, 1.2)).astype(np.float32)
                kinds.append(("hull", pts, w.hull(pts), 0.04))
        if kinds[0][0] == "sphere" and kinds[1][0] == "sphere":
            continue                      # sphere-sphere never reaches the GJK detector
        rots = scenes.random_rotations(rng, 2)
        pa = rng.uniform(-0.2, 0.2, size=3)
        # from overlapping to just apart
        pb = pa + rng.normal(size=3) / 1.0 * rng.uniform(0.25, 1.15)
        xa = scenes.make_xf(rots[:1], pa[None])[0]
        xb = scenes.make_xf(rots[1:], pb[None])[0]
        r = w.gjk_pair(kinds[0][2], xa, kinds[1][2], xb)
        A = (kinds[0][0], kinds[0][1], xa[:9].reshape(3, 3).astype(np.float64), xa[9:].astype(np.float64), kinds[0][3])
        B = (kinds[1][0], kinds[1][1], xb[:9].reshape(3, 3).astype(np.float64), xb[9:].astype(np.float64), kinds[1][3])
        global SHARP_BOX
        SHARP_BOX = r["has"] and r["method"] == 3
        sep = _separation(A, B, seed=trial)
        if not r["has"]:
            # no contact reported: the rounded shapes are farther apart than the breaking threshold (0.02), up to the
            # detector's early-out slack.  One reference quirk is tolerated and counted: ConvexHullShape's support mapping
            # replaces a search direction shorter than 1e-2 by (1, 0, 0) (sh/ConvexHullShape.java:83-89), so for hulls whose
            # CORES overlap the shrinking GJK direction can produce a bogus support point and trip the separating-axis
            # early-out (np/GjkPairDetector.java:154) — the restatement has to reproduce that, not fix it.
            if sep <= 0.02 - 2e-3:
                assert "hull" in (kinds[0][0], kinds[1][0]) and sep < -0.05, (trial, sep, kinds[0][0], kinds[1][0])
                quirks += 1
            continue
        if r["method"] == 1:              # plain GJK answer: cores apart, distance of the rounded shapes
            assert abs(r["depth"] - sep) < 2e-4, (trial, r["depth"], sep)
            checked_sep += 1
        elif r["method"] == 3:            # penetration solver (shapes grown by EPA_ACCURACY = 1e-3, stops within 1e-3 of the hull)
            assert r["depth"] < 0 and abs(r["depth"] - sep) < 5e-3, (trial, r["depth"], sep)
            checked_pen += 1
        # the reported normal is a unit vector along which the separation is attained (within the same slack)
        n = r["normal"].astype(np.float64)
        assert abs(np.linalg.norm(n) - 1.0) < 1e-5
        along = -(_support(*A, -n[None])[0] + _support(*B, n[None])[0])
        assert along > sep - 5e-3, (trial, along, sep)
    assert checked_sep >= 5 and checked_pen >= 5 and quirks <= 6, (checked_sep, checked_pen, quirks)


def test_convex_plane_distance_matches_the_support_function():
    """disp/ConvexPlaneCollisionAlgorithm.java:75-136: the reported distance is the signed distance of the deepest point of the
    shape (support mapping WITH margin along -normal: sharp full-size box, rounded hull, sphere) to the plane."""
    global SHARP_BOX
    rng = np.random.default_rng(7)
    hp = scenes.hull_points(rng, 0.5)
    checked = 0
    for trial in range(40):
        w = orc.OracleWorld(orc.TIGHT)
        n = rng.normal(size=3)
        n /= np.linalg.norm(n)
        n32 = n.astype(np.float32)
        cst = float(np.float32(rng.uniform(-1, 1)))
        plane = w.plane([float(v) for v in n32], cst)
        k = trial % 3
        if k == 0:
            he = rng.uniform(0.3, 0.6, size=3).astype(np.float32)
            shp = ("box", he, w.box(*[float(v) for v in he]), 0.04)
        elif k == 1:
            r = float(np.float32(rng.uniform(0.3, 0.6)))
            shp = ("sphere", r, w.sphere(r), r)
        else:
            pts = (hp * rng.uniform(0.7, 1.2)).astype(np.float32)
            shp = ("hull", pts, w.hull(pts), 0.04)
        rot = scenes.random_rotations(rng, 1)
        nn = n32.astype(np.float64) / np.linalg.norm(n32.astype(np.float64))   # StaticPlaneShape normalises its normal
        pos = nn * (cst + rng.uniform(0.2, 0.7)) + rng.normal(size=3) * 0.01
        xf = scenes.make_xf(rot, pos[None])[0]
        w.body(plane, orc.xf12(origin=(0, 0, 0)), group=2, mask=-1 ^ 2, static=True)
        w.body(shp[2], xf)
        w.step()
        ri, rf = w.raw()
        assert len(ri) == 1 and ri[0, 4] == 11
        SHARP_BOX = True
        A = (shp[0], shp[1], xf[:9].reshape(3, 3).astype(np.float64), xf[9:].astype(np.float64), shp[3])
        expected = -_support(*A, -nn[None])[0] - cst
        assert abs(rf[0, 6] - expected) < 1e-5, (trial, rf[0, 6], expected)
        assert bool(ri[0, 3]) == (rf[0, 6] < 0.02)
        checked += 1
    SHARP_BOX = False
    assert checked == 40
