"""Generate the golden fixtures in this directory from the CPU oracle (run from the repo root):

    python tests/golden/make_golden.py

The reference itself (Java, no JVM in the image) cannot produce vectors, so these are ORACLE outputs: they pin
the oracle against regressions and let the -m gpu tests check the CUDA path without executing oracle/ at all.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import orc  # noqa: E402
import scenes  # noqa: E402

CASES = {
    "c1_stack_dbvt": (lambda: scenes.stack_scene(n_side=5, extra=True, seed=1), orc.DBVT, 6),
    "c1_plane_tight": (lambda: scenes.stack_scene(n_side=3, extra=True, seed=2, plane_ground=True), orc.TIGHT, 4),
    "c2_bin_dbvt": (lambda: scenes.bin_scene(n=400, seed=9), orc.DBVT, 4),
    "c3_terrain_dbvt": (lambda: scenes.terrain_scene(cells=24, n=60, seed=4), orc.DBVT, 3),
    "c4_worlds_dbvt": (lambda: scenes.worlds_scene(num_worlds=4, seed=5), orc.DBVT, 3),
    # CompoundShape pairs (child manifolds carry child indices in mf_hdr columns 5, 6)
    "c6_compound_dbvt": (lambda: scenes.compound_scene(n=120, seed=8), orc.DBVT, 4),
    # compounds on a triangle mesh: ConvexConcave per child, raw keys -2 - (child << 21 | triangle)
    "c7_terrain_compound_dbvt": (lambda: scenes.terrain_compound_scene(cells=24, n=60, seed=15), orc.DBVT, 3),
    # children that are CompoundShapes themselves (one and two levels, rotated frames): leaves in depth-first order
    "c8_nested_compound_dbvt": (lambda: scenes.compound_scene(n=90, seed=12, plane_ground=False, nested=True), orc.DBVT, 3),
}

# Queries behind the pair list on a stepped world: rays, convex sweeps, the integrator's CCD sweeps.
QUERY_CASES = {
    "q1_queries_nested_compounds": (lambda: scenes.compound_scene(n=150, seed=13, plane_ground=False, nested=True), orc.DBVT, 2),
    "q2_queries_terrain": (lambda: scenes.terrain_compound_scene(cells=24, n=80, seed=17), orc.DBVT, 2),
}


def query_inputs(sc, seed):
    """Deterministic query set for a scene: (ray from, ray to, sweep basis, cast kind per sweep, ccd bodies, radii, targets)."""
    rng = np.random.default_rng(1000 + seed)
    ext = float(sc.extent)
    n = 160
    f = rng.uniform(-1.0, ext, size=(n, 3)).astype(np.float32)
    f[:, 1] = rng.uniform(2.0, 8.0, size=n)
    t = rng.uniform(-1.0, ext, size=(n, 3)).astype(np.float32)
    t[:, 1] = rng.uniform(-2.0, 1.0, size=n)
    basis = scenes.random_rotations(rng, n).astype(np.float32)
    basis[::3] = np.eye(3, dtype=np.float32)
    kind = rng.integers(2, size=n).astype(np.int32)          # 0: sphere 0.2, 1: box (0.3, 0.2, 0.25)
    dyn = np.asarray([k + 1 for k in range(sc.n) if not sc.static[k]], np.int32)
    me = rng.choice(dyn, size=min(60, len(dyn)), replace=False).astype(np.int32)
    rad = rng.uniform(0.1, 0.3, size=len(me)).astype(np.float32)
    return f, t, basis, kind, me, rad, rng.uniform(-2.0, 2.0, size=(len(me), 3)).astype(np.float32)


def run_queries(world, sc, steps, seed, is_oracle):
    """Step `steps` times, then ask the queries; the same code drives the oracle (is_oracle) and the CUDA world."""
    for step in range(steps):
        xf = sc.transforms(step)
        if is_oracle:
            world.step(xf)
        else:
            world.setWorldTransforms(xf); world.step()
    f, t, basis, kind, me, rad, dto = query_inputs(sc, seed)
    to = (xf[me - 1, 9:] + dto).astype(np.float32)
    if is_oracle:
        casts = np.asarray([world.sphere(0.2), world.box(0.3, 0.2, 0.25)], np.int32)
        ray = world.ray_test_closest(f, t, 1, 1)
        sweep = world.convex_sweep_closest(casts[kind], basis, f, t, 1, 1)
        ccd = world.ccd_sweep_not_me(me, rad, to)
    else:
        casts = np.asarray([world.SphereShape(0.2), world.BoxShape((0.3, 0.2, 0.25))], np.int32)
        ray = world.rayTestClosest(f, t, 1, 1)
        sweep = world.convexSweepTestClosest(casts[kind], basis, f, t, 1, 1)
        ccd = world.ccdSweepNotMe(me, rad, to)
    out = {}
    for name, (uid, frac, nrm, pt) in (("ray", ray), ("sweep", sweep), ("ccd", ccd)):
        out[name + "_uid"], out[name + "_frac"], out[name + "_nrm"], out[name + "_pt"] = uid, frac, nrm, pt
    return out


def run_case(make, mode, steps):
    sc = make()
    ow = scenes.build_oracle(sc, mode)
    out = {}
    for step in range(steps):
        xf = sc.transforms(step)
        ow.set_transforms(xf)
        ow.update_aabbs()
        out[f"aabb{step}"] = ow.aabbs().copy()
        out[f"pairs{step}"] = ow.calculate_overlapping_pairs()
        ow.dispatch_all_pairs()
        ints, fl = ow.raw()
        order = np.lexsort((ints[:, 2], ints[:, 1], ints[:, 0])) if len(ints) else np.zeros(0, int)
        out[f"raw_i{step}"] = ints[order][:, :5]
        out[f"raw_f{step}"] = fl[order]
        hdr, pts, pint = ow.manifolds()
        out[f"mf_hdr{step}"] = hdr
        out[f"mf_pts{step}"] = pts
        out[f"mf_int{step}"] = pint
    return out


def main():
    only = set(sys.argv[1:])  # optional: regenerate just the named cases
    for name, (make, mode, steps) in CASES.items():
        if only and name not in only:
            continue
        out = run_case(make, mode, steps)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {k: v.shape for k, v in out.items() if k.startswith("pairs")})
    for k, (name, (make, mode, steps)) in enumerate(QUERY_CASES.items()):
        if only and name not in only:
            continue
        sc = make()
        out = run_queries(scenes.build_oracle(sc, mode), sc, steps, k, True)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print(name, {q: int((out[q + "_uid"] > 0).sum()) for q in ("ray", "sweep", "ccd")})


if __name__ == "__main__":
    main()
