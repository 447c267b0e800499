"""-m gpu: the row-based pair ordering (csrc/pair_rows.cuh) on the shapes of input that stress it: very long rows
(every body overlaps every other), a long row whose uid range needs more than one bitmap chunk, rows right at the
short/long threshold, and bodies removed in the middle of the uid range."""
import numpy as np
import pytest

import parity
import scenes

pytestmark = pytest.mark.gpu


def _cluster_scene(n, spread, seed, radius=0.5):
    """n spheres scattered inside a cube of side `spread` (small spread -> all pairs overlap)."""
    rng = np.random.default_rng(scenes.SEED + seed)
    sc = scenes.Scene()
    s = sc.add_shape("sphere", radius)
    for _ in range(n):
        sc.body_shape.append(s); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    sc.base = scenes.make_xf(np.tile(np.eye(3), (n, 1, 1)), rng.uniform(0, spread, size=(n, 3)))
    sc.vel = rng.uniform(-0.02, 0.02, size=(n, 3))
    sc.spin = None
    sc.extent = float(spread + 1)
    return sc


@pytest.mark.parametrize("mode", [0, 1])
def test_all_pairs_overlap_every_row_is_long(gpu_pkg, mode):
    sc = _cluster_scene(300, 0.6, seed=1)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=mode, max_pairs=1 << 17)
    for step in range(3):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] == 300 * 299 // 2      # rows of 299, 298, ... 1: long and short rows, every length once


def test_rows_around_the_short_long_threshold(gpu_pkg):
    # 90 bodies in a tight cluster: row lengths 89 ... 1 cross the threshold (48) in both directions
    sc = _cluster_scene(90, 0.5, seed=2)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0, max_pairs=1 << 13)
    for step in range(2):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] == 90 * 89 // 2


def test_long_row_spanning_several_bitmap_chunks(gpu_pkg):
    """One static slab under 300 000 spheres: the slab's row holds ~300 k uids spread over a range wider than one
    shared-memory bitmap (262 144 uids), so the long-row kernel has to walk two chunks in order."""
    n = 300000
    rng = np.random.default_rng(scenes.SEED + 3)
    sc = scenes.Scene()
    slab = sc.add_shape("box", (400.0, 1.0, 400.0))
    sph = sc.add_shape("sphere", 0.3)
    sc.body_shape.append(slab); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(0)
    side = int(np.ceil(np.sqrt(n)))
    idx = np.arange(n)
    pos = np.stack([(idx % side) * 1.2 - 0.6 * side, np.full(n, 1.25), (idx // side) * 1.2 - 0.6 * side], axis=1)
    pos += rng.uniform(-0.25, 0.25, size=(n, 3)) * np.array([1.0, 0.1, 1.0])
    for _ in range(n):
        sc.body_shape.append(sph); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    sc.base = scenes.make_xf(np.tile(np.eye(3), (n + 1, 1, 1)), np.concatenate([np.zeros((1, 3)), pos]))
    sc.vel = None
    sc.spin = None
    sc.extent = 800.0
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=0, max_pairs=1 << 20)
    xf = sc.transforms(0)
    gw.setWorldTransforms(xf); ow.set_transforms(xf)
    gw.updateAabbs(); ow.update_aabbs()
    gw.getBroadphase().calculateOverlappingPairs()
    gp = gw.pairs()
    op = ow.calculate_overlapping_pairs()
    parity.compare_pairs(gp, op)
    slab_row = gp[gp[:, 0] == 1]
    assert len(slab_row) == n and slab_row[-1, 1] - slab_row[0, 1] > 262144
    assert np.all(np.diff(gp[:, 0].astype(np.int64) * (1 << 21) + gp[:, 1]) > 0), "pair list is not strictly sorted"


def test_rows_with_removed_bodies(gpu_pkg):
    sc = _cluster_scene(120, 1.5, seed=4)
    gw, ow = scenes.build_both(gpu_pkg, sc, mode=1, max_pairs=1 << 14)
    parity.step_and_compare(gw, ow, sc.transforms(0), sc.extent)
    for uid in (1, 2, 60, 61, 119, 120):
        gw.removeCollisionObject(uid)
        ow.destroy_body(uid)
    for step in range(1, 4):
        r = parity.step_and_compare(gw, ow, sc.transforms(step), sc.extent)
    assert r["pairs"] > 100
