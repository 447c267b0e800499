"""Ad-hoc staged debug run on the GPU box (not a test)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as ge
import scenes, parity
pkg = ge.load_package()
def log(*a):
    print(*a, flush=True)
which = sys.argv[1] if len(sys.argv) > 1 else "stack"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if which == "stack":
    sc = scenes.stack_scene(n_side=5, extra=True, seed=1)
elif which == "bin":
    sc = scenes.bin_scene(n=int(sys.argv[3]) if len(sys.argv) > 3 else 3000, seed=3)
log("bodies", sc.n)
gw, ow = scenes.build_both(pkg, sc, mode=mode)
log("built")
for step in range(3):
    xf = sc.transforms(step)
    gw.setWorldTransforms(xf); ow.set_transforms(xf)
    gw.updateAabbs(); ow.update_aabbs()
    ga = gw.aabbs(); log("aabbs got")
    parity.compare_aabbs(ga, ow.aabbs()); log("aabb ok")
    n = gw.getBroadphase().calculateOverlappingPairs(); log("pairs", n, gw.stats())
    op = ow.calculate_overlapping_pairs()
    parity.compare_pairs(gw.pairs(), op); log("pairs ok", len(op))
    gw.getDispatcher().dispatchAllCollisionPairs(); log("dispatched", gw.stats())
    ow.dispatch_all_pairs()
    r = parity.compare_raw(gw.raw_contacts(), ow.raw(), sc.extent); log("raw ok", r)
    m = parity.compare_manifolds(gw.manifolds(), ow.manifolds(), sc.extent); log("manifolds ok", m)
log("ALL OK")
