"""Extract the hot-path constants of SURVEY Appendix A from the REFERENCE sources into tests/golden/reference_constants.json.

    python tools/extract_reference_constants.py            (needs /root/reference; the JSON is committed)

The reference cannot run here (no JVM), but its source can be READ: this pins the literals the oracle and the device code
must reproduce to the lines they come from.  tests/test_reference_constants.py compares the fixture with the constants in
oracle/*.h and libgdx-jbullet_b200/csrc/*.cuh, and — when /root/reference is present — re-extracts and checks the fixture
is current.  Each entry: name -> (file under src/com/bulletphysics, regex with one group)."""
import json
import os
import re
import sys

REF = "/root/reference/src/com/bulletphysics"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "reference_constants.json")
NUM = r"([-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)f?"

SPEC = {
    "CONVEX_DISTANCE_MARGIN": ("BulletGlobals.java", r"CONVEX_DISTANCE_MARGIN\s*=\s*" + NUM),
    "FLT_EPSILON": ("BulletGlobals.java", r"FLT_EPSILON\s*=\s*" + NUM),
    "contactBreakingThreshold": ("BulletGlobals.java", r"private float contactBreakingThreshold\s*=\s*" + NUM),
    "GJK_REL_ERROR2": ("collision/narrowphase/GjkPairDetector.java", r"REL_ERROR2\s*=\s*" + NUM),
    "EPA_GJK_maxiterations": ("collision/narrowphase/GjkEpaSolver.java", r"GJK_maxiterations\s*=\s*" + NUM),
    "EPA_GJK_hashsize_log2": ("collision/narrowphase/GjkEpaSolver.java", r"GJK_hashsize\s*=\s*1\s*<<\s*" + NUM),
    "EPA_GJK_insimplex_eps": ("collision/narrowphase/GjkEpaSolver.java", r"GJK_insimplex_eps\s*=\s*" + NUM),
    "EPA_maxiterations": ("collision/narrowphase/GjkEpaSolver.java", r"EPA_maxiterations\s*=\s*" + NUM),
    "EPA_inface_eps": ("collision/narrowphase/GjkEpaSolver.java", r"EPA_inface_eps\s*=\s*" + NUM),
    "EPA_accuracy": ("collision/narrowphase/GjkEpaSolver.java", r"EPA_accuracy\s*=\s*" + NUM),
    "DBVT_BP_MARGIN": ("collision/broadphase/DbvtBroadphase.java", r"DBVT_BP_MARGIN\s*=\s*" + NUM),
    "DBVT_predictedframes": ("collision/broadphase/DbvtBroadphase.java", r"predictedframes\s*=\s*" + NUM),
    "MANIFOLD_CACHE_SIZE": ("collision/narrowphase/PersistentManifold.java", r"MANIFOLD_CACHE_SIZE\s*=\s*" + NUM),
    "MAX_FRICTION": ("collision/dispatch/ManifoldResult.java", r"MAX_FRICTION\s*=\s*" + NUM),
    "GjkConvexCast_MAX_ITERATIONS": ("collision/narrowphase/GjkConvexCast.java", r"^\s*private static final int MAX_ITERATIONS\s*=\s*" + NUM),
    "GjkConvexCast_radius": ("collision/narrowphase/GjkConvexCast.java", r"float radius\s*=\s*" + NUM),
    "SubsimplexConvexCast_MAX_ITERATIONS": ("collision/narrowphase/SubsimplexConvexCast.java", r"MAX_ITERATIONS\s*=\s*" + NUM),
    "SubsimplexConvexCast_epsilon": ("collision/narrowphase/SubsimplexConvexCast.java", r"^\s*float epsilon\s*=\s*" + NUM),
    "allowedCcdPenetration": ("collision/broadphase/DispatcherInfo.java", r"allowedCcdPenetration\s*=\s*" + NUM),
    "aabb_overflow_guard_len2": ("collision/dispatch/CollisionWorld.java", r"len2\(\)\s*<\s*" + NUM),
    "ccd_min_hit_fraction": ("dynamics/DiscreteDynamicsWorld.java", r"closestHitFraction\s*>\s*" + NUM),
    "GJK_max_iterations": ("collision/narrowphase/GjkPairDetector.java", r"gGjkMaxIter\s*=\s*" + NUM),
    "GJK_degenerate5_lenSqr": ("collision/narrowphase/GjkPairDetector.java", r"lenSqr\s*<\s*" + NUM),
    "GJK_catch_degenerate_distance": ("collision/narrowphase/GjkPairDetector.java", r"\(distance \+ margin\)\s*<\s*" + NUM),
    "Voronoi_degenerate_signd": ("collision/narrowphase/VoronoiSimplexSolver.java", r"signd \* signd < \(\(" + NUM),
    "Hull_tiny_direction_lenSqr": ("collision/shapes/ConvexHullShape.java", r"lenSqr\s*<\s*" + NUM),
    "Hull_maxDot_init": ("collision/shapes/ConvexHullShape.java", r"maxDot\s*=\s*" + NUM),
    "BVH_MAX_NUM_PARTS_IN_BITS": ("collision/shapes/OptimizedBvh.java", r"MAX_NUM_PARTS_IN_BITS\s*=\s*" + NUM),
    "FILTER_DEFAULT": ("collision/broadphase/CollisionFilterGroups.java", r"DEFAULT_FILTER\s*=\s*" + NUM),
    "FILTER_STATIC": ("collision/broadphase/CollisionFilterGroups.java", r"STATIC_FILTER\s*=\s*" + NUM),
    "FILTER_ALL": ("collision/broadphase/CollisionFilterGroups.java", r"ALL_FILTER\s*=\s*" + NUM),
}


def extract(ref=REF):
    out = {}
    for name, (rel, rx) in SPEC.items():
        path = os.path.join(ref, rel)
        text = open(path, encoding="utf-8", errors="replace").read()
        hits = [(m.group(1), text.count("\n", 0, m.start()) + 1) for m in re.finditer(rx, text, flags=re.M)]
        if not hits:
            raise SystemExit(f"{name}: pattern not found in {rel}")
        val, line = hits[0]
        out[name] = {"value": float(val), "source": f"{rel}:{line}"}
    return out


if __name__ == "__main__":
    data = extract(sys.argv[1] if len(sys.argv) > 1 else REF)
    json.dump(data, open(OUT, "w"), indent=1, sort_keys=True)
    for k, v in sorted(data.items()):
        print(f"{k:36s} {v['value']:<14g} {v['source']}")
