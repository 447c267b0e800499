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
}


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


if __name__ == "__main__":
    main()
