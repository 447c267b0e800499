"""CPU: known answers for the oracle's CollisionWorld.convexSweepTest restatement (oracle/convexcast.h) — the oracle is test
infrastructure; these pin it against closed-form times of impact before the -m gpu tests compare the device path with it."""
import numpy as np

import orc
import scenes

EYE = np.eye(3, dtype=np.float32)


def _world():
    w = orc.OracleWorld(mode=orc.TIGHT)
    return w


def test_sphere_swept_onto_a_sphere_and_a_box():
    w = _world()
    big = w.sphere(1.0)
    small = w.sphere(0.25)
    bx = w.box(1.0, 0.5, 1.0)
    w.body(big, orc.xf12(origin=(0, 0, 0)), 1, -1, False, 0)
    w.body(bx, orc.xf12(origin=(10, 0, 0)), 1, -1, False, 0)
    # straight down onto the sphere's pole: touching when the centres are 1.25 apart -> travelled 3.75 of 10
    uid, frac, nrm, pt = w.convex_sweep_closest(small, EYE, [(0, 5, 0), (10, 5, 0), (5, 5, 0)], [(0, -5, 0), (10, -5, 0), (5, -5, 0)])
    assert uid.tolist() == [1, 2, 0]
    assert abs(frac[0] - 0.375) < 2e-3 and abs(frac[1] - (5 - 0.75) / 10) < 2e-3 and frac[2] == 1.0
    assert np.allclose(nrm[:2], [[0, 1, 0], [0, 1, 0]], atol=1e-3)
    assert np.allclose(pt[0], [0, 1, 0], atol=5e-3) and abs(pt[1][1] - 0.5) < 5e-3
    # conservative advancement stops within its 0.001 radius BEFORE contact: never past the true time of impact
    assert frac[0] <= 0.375 + 1e-6 and frac[1] <= 0.425 + 1e-6


def test_closest_of_several_targets_filter_and_overlapping_start():
    w = _world()
    s = w.sphere(0.5)
    cast = w.box(0.2, 0.2, 0.2)
    for k, x in enumerate((6.0, 3.0, 9.0)):
        w.body(s, orc.xf12(origin=(x, 0, 0)), 1 if k != 1 else 4, -1, False, 0)
    f, t = [(0, 0, 0)], [(12, 0, 0)]
    uid, frac, nrm, _ = w.convex_sweep_closest(cast, EYE, f, t)
    assert uid[0] == 2 and abs(frac[0] - (3 - 0.5 - 0.2) / 12) < 2e-3 and nrm[0][0] < -0.99
    uid, frac, _, _ = w.convex_sweep_closest(cast, EYE, f, t, group=1, mask=-1 ^ 4)    # the callback ignores group 4
    assert uid[0] == 1 and abs(frac[0] - (6 - 0.5 - 0.2) / 12) < 2e-3
    # a sweep that starts in deep overlap: GjkPairDetector without a penetration solver has no valid result -> no hit from it
    uid, frac, _, _ = w.convex_sweep_closest(cast, EYE, [(3, 0, 0)], [(3, 5, 0)], group=1, mask=4)
    assert uid[0] == 0 and frac[0] == 1.0
    # moving away from a body it almost touches: n.r >= -allowedPenetration rejects
    uid, _, _, _ = w.convex_sweep_closest(cast, EYE, [(3, 0.7004, 0)], [(3, 5, 0)], group=1, mask=4)
    assert uid[0] == 0


def test_sweep_onto_a_mesh_a_compound_and_the_plane_branch():
    sc = scenes.terrain_scene(cells=16, n=0, seed=3)
    ow = scenes.build_oracle(sc, orc.TIGHT)
    sph = ow.sphere(0.3)
    bar = ow.box(0.5, 0.1, 0.1)
    comp = ow.compound([sph, bar], scenes.make_xf(np.stack([EYE, EYE]), np.asarray([(0, 0.5, 0), (0, 0, 0)])))
    ow.body(comp, orc.xf12(origin=(4, 6, 4)), 1, -1, False, 0)
    cast = ow.sphere(0.2)
    uid, frac, nrm, pt = ow.convex_sweep_closest(cast, EYE, [(4, 9, 4), (2, 9, 2)], [(4, -3, 4), (2, -3, 2)])
    # first sweep lands on the compound's upper sphere (top at y = 6.8), second on the terrain
    assert uid.tolist() == [2, 1]
    assert abs(frac[0] - (9 - 6.8 - 0.2) / 12) < 2e-3 and nrm[0][1] > 0.99
    assert 0.3 < frac[1] < 1.0 and nrm[1][1] > 0.5
    assert abs(pt[0][1] - 6.8) < 5e-3
    # the mesh reports a point ON the triangle side (hitB of the simplex), within the mesh margin of the surface
    pl = ow.plane([0.0, 1.0, 0.0], -5.0)
    ow.body(pl, orc.xf12(), 2, -1 ^ 2, True, 0)
    uid, _, _, _ = ow.convex_sweep_closest(cast, EYE, [(4, 9, 4)], [(4, -3, 4)])
    assert uid[0] == -1                       # the reference throws in its static-plane branch
    uid, _, _, _ = ow.convex_sweep_closest(cast, EYE, [(4, 9, 4)], [(4, -3, 4)], group=1, mask=-1 ^ 2)
    assert uid[0] == 2


def test_rotated_box_cast_uses_its_basis():
    w = _world()
    ground = w.box(5, 0.5, 5)
    w.body(ground, orc.xf12(origin=(0, -0.5, 0)), 1, -1, False, 0)
    cast = w.box(0.5, 0.5, 0.5)
    c, s = np.cos(np.pi / 4), np.sin(np.pi / 4)
    rot = np.asarray([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)    # 45 degrees about z: the lowest corner is sqrt(.5) down
    uid, frac, _, _ = w.convex_sweep_closest(cast, [EYE, rot], [(0, 5, 0), (1, 5, 0)], [(0, -5, 0), (1, -5, 0)])
    assert uid.tolist() == [1, 1]
    assert abs(frac[0] - 0.45) < 2e-3 and abs(frac[1] - (5 - np.sqrt(0.5)) / 10) < 3e-3


def test_ccd_motion_clamping_sweep_skips_me_and_touching_objects():
    """DiscreteDynamicsWorld.integrateTransforms' CCD query (dyn/DiscreteDynamicsWorld.java:700-729, 1129-1199)."""
    w = orc.OracleWorld(mode=orc.DBVT)
    ground = w.box(20, 0.5, 20)
    ball = w.sphere(0.5)
    wall = w.box(0.25, 3, 3)
    w.body(ground, orc.xf12(origin=(0, -0.5, 0)), 2, -1 ^ 2, True, 0)          # uid 1, top at y = 0
    w.body(ball, orc.xf12(origin=(0, 0.49, 0)), 1, -1, False, 0)               # uid 2, resting on the ground (touching)
    w.body(wall, orc.xf12(origin=(5, 3, 0)), 2, -1 ^ 2, True, 0)               # uid 3, face at x = 4.75
    w.body(ball, orc.xf12(origin=(0, 6, 0)), 1, -1, False, 0)                  # uid 4, free in the air
    xf = np.stack([orc.xf12(origin=(0, -0.5, 0)), orc.xf12(origin=(0, 0.49, 0)), orc.xf12(origin=(5, 3, 0)), orc.xf12(origin=(0, 6, 0))])
    w.step(xf)
    # body 2 flies at the wall: swept sphere of radius 0.2 from x = 0 to x = 10 meets the face at x = 4.75 - 0.2;
    # the ground it rests on (contact points already) is not reported although the sweep grazes it
    uid, frac, nrm, _ = w.ccd_sweep_not_me([2], 0.2, [(10, 0.49, 0)])
    assert uid[0] == 3 and abs(frac[0] - 4.55 / 10) < 2e-3 and nrm[0][0] < -0.99
    # straight down: the only thing below is the ground, which is excluded -> no hit, full motion allowed
    uid, frac, _, _ = w.ccd_sweep_not_me([2], 0.2, [(0, -3, 0)])
    assert uid[0] == 0 and frac[0] == 1.0
    # body 4 falls onto body 2 (not touching anything yet): top of the ball at y = 0.99
    uid, frac, nrm, _ = w.ccd_sweep_not_me([4], 0.2, [(0, -4, 0)])
    assert uid[0] == 2 and abs(frac[0] - (6 - 0.99 - 0.2) / 10) < 2e-3 and nrm[0][1] > 0.99
    # moving away from everything: nothing
    uid, _, _, _ = w.ccd_sweep_not_me([4], 0.2, [(0, 16, 0)])
    assert uid[0] == 0
    # the plain sweep of the same sphere from body 2's place does see the ground it rests on... unless it starts in contact
    s02 = w.sphere(0.2)
    uid, _, _, _ = w.convex_sweep_closest(s02, EYE, [(0, 2.0, 0)], [(0, -3, 0)])
    assert uid[0] == 2 or uid[0] == 1


def test_sphere_sweeps_onto_a_box_face_match_the_closed_form_time_of_impact():
    """Randomised: a sphere of radius r swept onto the top face of a large box from random heights and directions touches when
    its centre is r above the face: fraction = (y0 - r) / (y0 - y1).  Against a flat face the first conservative-advancement
    step of GjkConvexCast lands on the contact itself, so the reported fraction equals the closed form to float rounding
    (positions up to 20 units: ~1e-5); in general it may stop up to its 0.001 radius early."""
    rng = np.random.default_rng(77)
    w = orc.OracleWorld(mode=orc.TIGHT)
    slab = w.box(50.0, 1.0, 50.0)                       # top face at y = 0 (centre at y = -1)
    w.body(slab, orc.xf12(origin=(0, -1.0, 0)), 1, -1, True, 0)
    n = 200
    r = 0.3
    cast = w.sphere(r)
    f = np.stack([rng.uniform(-20, 20, n), rng.uniform(1.0, 6.0, n), rng.uniform(-20, 20, n)], axis=1).astype(np.float32)
    t = np.stack([f[:, 0] + rng.uniform(-5, 5, n), rng.uniform(-4.0, -0.5, n), f[:, 2] + rng.uniform(-5, 5, n)], axis=1).astype(np.float32)
    uid, frac, nrm, pt = w.convex_sweep_closest(cast, EYE, f, t)
    assert (uid == 1).all()
    exact = (f[:, 1].astype(np.float64) - r) / (f[:, 1].astype(np.float64) - t[:, 1].astype(np.float64))
    slack = 0.0011 / (f[:, 1] - t[:, 1]).astype(np.float64) + 1e-6
    assert (frac <= exact + 2e-5).all() and (frac >= exact - slack).all(), (np.abs(frac - exact).max(), slack.max())
    assert np.abs(frac - exact).max() < 2e-5
    assert np.allclose(nrm, [0, 1, 0], atol=1e-4)
    # the hit point lies on the face (y = 0) below the sphere centre at the time of impact
    centre = f + (t - f) * frac[:, None]
    assert np.abs(pt[:, 1]).max() < 2e-3 and np.abs(pt[:, [0, 2]] - centre[:, [0, 2]]).max() < 2e-3
    # a rotated box cast (45 degrees about z, lowest edge sqrt(2) * h below the centre) against the same face
    c, s = np.cos(np.pi / 4), np.sin(np.pi / 4)
    rot = np.asarray([[c, -s, 0], [s, c, 0], [0, 0, 1]], np.float32)
    h = 0.25
    bx = w.box(h, h, h)
    uid, frac, _, _ = w.convex_sweep_closest(bx, rot, f, t)
    # GJK works on the margin-less core (half extent h - 0.04) plus a 0.04 margin sphere: the lowest EDGE of the rotated box is rounded
    low = np.sqrt(2.0) * (h - 0.04) + 0.04
    exact = (f[:, 1].astype(np.float64) - low) / (f[:, 1].astype(np.float64) - t[:, 1].astype(np.float64))
    assert (uid == 1).all() and (frac <= exact + 1e-4).all() and (frac >= exact - slack - 1e-4).all()


def test_sphere_sweeps_onto_a_flat_triangle_mesh_match_the_closed_form():
    """BVH box-cast walk + SubsimplexConvexCast per triangle on a flat grid mesh at y = 0.5: a sphere of radius r touches the
    triangles' margin shell when its centre is r + mesh margin above the plane.  SubsimplexConvexCast stops once the squared
    distance falls under 1e-4, i.e. up to 0.01 early."""
    rng = np.random.default_rng(5)
    w = orc.OracleWorld(mode=orc.TIGHT)
    c = 12
    xs = np.arange(c + 1, dtype=np.float32)
    verts = np.asarray([(x, 0.5, z) for x in xs for z in xs], np.float32)
    tris = []
    for i in range(c):
        for j in range(c):
            v00, v10, v01, v11 = i * (c + 1) + j, (i + 1) * (c + 1) + j, i * (c + 1) + j + 1, (i + 1) * (c + 1) + j + 1
            tris += [(v00, v01, v10), (v10, v01, v11)]
    m = w.mesh(verts, np.asarray(tris, np.int32))
    w.body(m, orc.xf12(), 2, -1 ^ 2, True, 0)
    r = 0.25
    cast = w.sphere(r)
    n = 120
    f = np.stack([rng.uniform(1, 11, n), rng.uniform(2.0, 5.0, n), rng.uniform(1, 11, n)], axis=1).astype(np.float32)
    t = np.stack([np.clip(f[:, 0] + rng.uniform(-1.5, 1.5, n), 0.5, 11.5), rng.uniform(-2.0, 0.0, n),
                  np.clip(f[:, 2] + rng.uniform(-1.5, 1.5, n), 0.5, 11.5)], axis=1).astype(np.float32)
    uid, frac, nrm, pt = w.convex_sweep_closest(cast, EYE, f, t)
    assert (uid == 1).all()
    margin = 0.0   # BvhTriangleMeshShape: collisionMargin 0 (sh/TriangleMeshShape.java), the triangles get the mesh's margin
    dy = (f[:, 1] - t[:, 1]).astype(np.float64)
    exact = (f[:, 1].astype(np.float64) - (0.5 + r + margin)) / dy
    assert (frac <= exact + 1e-5).all() and (frac >= exact - 0.0101 / dy - 1e-5).all(), np.abs(frac - exact).max()
    # the reported normal is the simplex's separation vector at the LAST advancement (np/SubsimplexConvexCast.java:150), which
    # for an oblique approach is still a little off the face normal when the cast stops early
    assert (nrm[:, 1] > 0.95).all() and np.abs(np.linalg.norm(nrm, axis=1) - 1.0).max() < 1e-5
    assert np.abs(pt[:, 1] - 0.5).max() < 1e-3        # hit point = the simplex's point on the triangle
