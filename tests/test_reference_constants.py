"""CPU: the hot-path constants of SURVEY Appendix A, pinned to the REFERENCE'S OWN SOURCE LINES.

tests/golden/reference_constants.json was extracted from /root/reference by tools/extract_reference_constants.py (value +
file:line).  The reference cannot be run here, but its literals can be read: every constant below must appear with the same
value in the oracle (oracle/*.h) AND in the device code (libgdx-jbullet_b200/csrc), so a typo in either restatement — or a
later "tuning" of a threshold — fails a test instead of silently changing results.  When /root/reference is present the
fixture itself is re-extracted and compared."""
import json
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = json.load(open(os.path.join(ROOT, "tests", "golden", "reference_constants.json")))
NUM = r"([-+]?[0-9]*\.?[0-9]+(?:[eE][-+]?[0-9]+)?)f?"
CS = "libgdx-jbullet_b200/csrc/"

# constant -> [(file, regex with one numeric group), ...]: every listed place must carry the reference's value
WHERE = {
    "FLT_EPSILON": [(CS + "gjk.cuh", r"B2C_FLT_EPSILON\s*=\s*" + NUM), ("oracle/jmath.h", r"FLT_EPSILON_\s*=\s*" + NUM)],
    "contactBreakingThreshold": [(CS + "b2c_api.cu", r"cfg->contact_breaking_threshold\s*=\s*" + NUM),
                                 ("oracle/world.h", r"float breakingThreshold\s*=\s*" + NUM)],
    "GJK_REL_ERROR2": [(CS + "gjk.cuh", r"GJK_REL_ERROR2\s*=\s*" + NUM), ("oracle/gjk.h", r"REL_ERROR2\s*=\s*" + NUM)],
    "EPA_GJK_maxiterations": [(CS + "epa.cuh", r"EPA_GJK_MAXIT\s*=\s*" + NUM), ("oracle/gjk.h", r"GJK_maxiterations\s*=\s*" + NUM)],
    "EPA_GJK_insimplex_eps": [(CS + "epa.cuh", r"EPA_INSIMPLEX_EPS\s*=\s*" + NUM), ("oracle/gjk.h", r"GJK_insimplex_eps\s*=\s*" + NUM)],
    "EPA_maxiterations": [(CS + "epa.cuh", r"EPA_MAXIT\s*=\s*" + NUM), ("oracle/gjk.h", r"EPA_maxiterations\s*=\s*" + NUM)],
    "EPA_inface_eps": [(CS + "epa.cuh", r"EPA_INFACE_EPS\s*=\s*" + NUM), ("oracle/gjk.h", r"EPA_inface_eps\s*=\s*" + NUM)],
    "EPA_accuracy": [(CS + "epa.cuh", r"EPA_ACCURACY\s*=\s*" + NUM), ("oracle/gjk.h", r"EPA_accuracy\s*=\s*" + NUM)],
    "DBVT_BP_MARGIN": [(CS + "b2c_api.cu", r"cfg->dbvt_margin\s*=\s*" + NUM), ("oracle/world.h", r"float dbvtMargin\s*=\s*" + NUM)],
    "DBVT_predictedframes": [(CS + "b2c_api.cu", r"cfg->dbvt_predicted_frames\s*=\s*" + NUM),
                             ("oracle/world.h", r"float predictedFrames\s*=\s*" + NUM)],
    "MANIFOLD_CACHE_SIZE": [("oracle/manifold.h", r"MANIFOLD_CACHE_SIZE\s*=\s*" + NUM), ("include/b2c.h", r"b2c_manifold_point points\[" + NUM)],
    "MAX_FRICTION": [("oracle/manifold.h", r"MAX_FRICTION\s*=\s*" + NUM), (CS + "narrowphase.cuh", r"if \(f > ([0-9]+)\.?f?\) f = ")],
    "GjkConvexCast_MAX_ITERATIONS": [(CS + "convexcast.cuh", r"if \(numIter > " + NUM), ("oracle/convexcast.h", r"const int maxIter\s*=\s*" + NUM)],
    "GjkConvexCast_radius": [(CS + "convexcast.cuh", r"const float radius\s*=\s*" + NUM), ("oracle/convexcast.h", r"const float radius\s*=\s*" + NUM)],
    "SubsimplexConvexCast_MAX_ITERATIONS": [(CS + "convexcast.cuh", r"int maxIter\s*=\s*" + NUM), (CS + "raycast.cuh", r"int maxIter\s*=\s*" + NUM),
                                            ("oracle/convexcast.h", r"^\s*int maxIter\s*=\s*" + NUM), ("oracle/raycast.h", r"int maxIter\s*=\s*" + NUM)],
    "SubsimplexConvexCast_epsilon": [(CS + "convexcast.cuh", r"const float epsilon\s*=\s*" + NUM), (CS + "raycast.cuh", r"const float epsilon\s*=\s*" + NUM),
                                     ("oracle/convexcast.h", r"const float epsilon\s*=\s*" + NUM), ("oracle/raycast.h", r"const float epsilon\s*=\s*" + NUM)],
    "allowedCcdPenetration": [("include/b2c_host.hpp", r"allowedCcdPenetration\s*=\s*" + NUM), ("libgdx-jbullet_b200/world.py", r"allowed_ccd_penetration=" + NUM)],
    "aabb_overflow_guard_len2": [(CS + "broadphase.cuh", r"len2_3\(d\) < " + NUM), ("oracle/world.h", r"tmp\.len2\(\) < " + NUM)],
    "GJK_max_iterations": [(CS + "gjk.cuh", r"curIter\+\+ > " + NUM), ("oracle/gjk.h", r"gGjkMaxIter\s*=\s*" + NUM)],
    "GJK_degenerate5_lenSqr": [(CS + "gjk.cuh", r"if \(lenSqr < " + NUM + r"\) degenerate = 5"), ("oracle/gjk.h", r"if \(lenSqr < " + NUM + r"\) degenerateSimplex = 5")],
    "GJK_catch_degenerate_distance": [(CS + "gjk.cuh", r"\(distance \+ margin\) < " + NUM), ("oracle/gjk.h", r"\(distance \+ margin\) < " + NUM)],
    "Voronoi_degenerate_signd": [(CS + "gjk.cuh", r"signd \* signd < \(\(" + NUM), ("oracle/voronoi.h", r"signd \* signd < \(\(" + NUM)],
    "Hull_tiny_direction_lenSqr": [(CS + "gjk.cuh", r"if \(l2 < " + NUM + r"\) v = mk3\(1"), ("oracle/shapes.h", r"if \(lenSqr < " + NUM)],
    "Hull_maxDot_init": [(CS + "gjk.cuh", r"float maxDot\s*=\s*" + NUM), ("oracle/shapes.h", r"newDot, maxDot\s*=\s*" + NUM)],
    "BVH_MAX_NUM_PARTS_IN_BITS": [],   # checked below: leaf word = partId << (31 - bits)
    "FILTER_DEFAULT": [("libgdx-jbullet_b200/world.py", r"DEFAULT_FILTER, STATIC_FILTER, ALL_FILTER = " + NUM)],
    "FILTER_STATIC": [("libgdx-jbullet_b200/world.py", r"DEFAULT_FILTER, STATIC_FILTER, ALL_FILTER = [-0-9]+, " + NUM)],
    "FILTER_ALL": [("libgdx-jbullet_b200/world.py", r"DEFAULT_FILTER, STATIC_FILTER, ALL_FILTER = [-0-9]+, [-0-9]+, " + NUM)],
    "CONVEX_DISTANCE_MARGIN": [(CS + "b2c_api.cu", r"s\.margin = margin >= 0\.f \? margin : " + NUM),
                               ("oracle/jmath.h", r"CONVEX_DISTANCE_MARGIN\s*=\s*" + NUM)],
}


@pytest.mark.parametrize("name", sorted(WHERE))
def test_constant_matches_the_reference_source(name):
    want = GOLD[name]["value"]
    for rel, rx in WHERE[name]:
        text = open(os.path.join(ROOT, rel), encoding="utf-8").read()
        hits = re.findall(rx, text, flags=re.M)
        assert hits, f"{name}: no literal found in {rel} (pattern {rx!r})"
        for h in hits:
            assert float(h) == want, f"{name}: {rel} has {h}, the reference has {want} at {GOLD[name]['source']}"


def test_mesh_leaf_word_uses_the_reference_part_bits():
    """sh/OptimizedBvh.java:65 MAX_NUM_PARTS_IN_BITS = 10: a leaf stores partId << (31 - 10) | triangleIndex."""
    shift = 31 - int(GOLD["BVH_MAX_NUM_PARTS_IN_BITS"]["value"])
    assert shift == 21
    assert f"<< {shift}) | (uint32_t)t" in open(os.path.join(ROOT, CS + "b2c_api.cu")).read()
    assert f"<< {shift}" in open(os.path.join(ROOT, "oracle", "world.h")).read()


def test_every_extracted_constant_is_checked_somewhere_or_documented():
    # the 64-entry ray hash of GjkEpaSolver is replaced by exact list membership on both sides (same answers: the hash only
    # speeds up "has this ray been seen"); the CCD clamp threshold is the host's (documented in b2c.h)
    only_documented = {"EPA_GJK_hashsize_log2", "ccd_min_hit_fraction"}
    assert set(GOLD) == set(WHERE) | only_documented
    assert "list membership" in open(os.path.join(ROOT, CS + "epa.cuh")).read() and "0.0001" in open(os.path.join(ROOT, "include", "b2c.h")).read()


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/com/bulletphysics"), reason="the reference tree is not on this box")
def test_fixture_is_current_with_the_reference_tree():
    import importlib.util
    spec = importlib.util.spec_from_file_location("extract", os.path.join(ROOT, "tools", "extract_reference_constants.py"))
    ex = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ex)
    assert ex.extract() == GOLD
