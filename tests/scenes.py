"""Seeded synthetic scenes for the collision path and helpers that build the same scene in
the CUDA world and in the CPU oracle.  Both sides receive the same float32 arrays, so they see identical
bits.  Transforms are (n,12): 9 row-major basis floats + 3 origin floats.
"""
import numpy as np

SEED = 0x6A62756C6C6574  # "jbullet"


def quat_to_mat(q):
    """(n,4) xyzw unit quaternions -> (n,3,3) float64 rotation matrices."""
    x, y, z, w = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    m = np.empty((len(q), 3, 3))
    m[:, 0, 0] = 1 - 2 * (y * y + z * z); m[:, 0, 1] = 2 * (x * y - w * z); m[:, 0, 2] = 2 * (x * z + w * y)
    m[:, 1, 0] = 2 * (x * y + w * z); m[:, 1, 1] = 1 - 2 * (x * x + z * z); m[:, 1, 2] = 2 * (y * z - w * x)
    m[:, 2, 0] = 2 * (x * z - w * y); m[:, 2, 1] = 2 * (y * z + w * x); m[:, 2, 2] = 1 - 2 * (x * x + y * y)
    return m


def random_rotations(rng, n):
    q = rng.normal(size=(n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return quat_to_mat(q)


def make_xf(rot, pos):
    n = len(pos)
    xf = np.zeros((n, 12), dtype=np.float32)
    xf[:, :9] = np.asarray(rot, dtype=np.float64).reshape(n, 9).astype(np.float32)
    xf[:, 9:] = np.asarray(pos, dtype=np.float32)
    return xf


def small_rotation(rng, n, angle):
    """Random axis, fixed small angle -> (n,3,3)."""
    ax = rng.normal(size=(n, 3))
    ax /= np.linalg.norm(ax, axis=1, keepdims=True)
    h = 0.5 * angle
    q = np.concatenate([ax * np.sin(h), np.full((n, 1), np.cos(h))], axis=1)
    return quat_to_mat(q)


class Scene:
    def __init__(self):
        self.shapes = []      # tuples: ("box", he3) ("sphere", r) ("hull", pts) ("plane", n3, c) ("mesh", verts, idx)
                              #         ("compound", [indices of child shapes in this list], child transforms (k,12))
        self.body_shape = []  # index into shapes
        self.static = []
        self.group = []
        self.mask = []
        self.world = []
        self.base = None      # (n,12) float32
        self.vel = None       # (n,3) per-step drift of the origin
        self.spin = None      # (n,3,3) per-step rotation increment or None
        self.extent = 1.0
        self.num_worlds = 1

    @property
    def n(self):
        return len(self.body_shape)

    def add_shape(self, *s):
        self.shapes.append(s)
        return len(self.shapes) - 1

    def transforms(self, step):
        """Transform trace: a deterministic drift (the solver/integrator are not ours — SURVEY §8d)."""
        xf = self.base.copy()
        if step == 0:
            return xf
        if self.vel is not None:
            xf[:, 9:] = (self.base[:, 9:].astype(np.float64) + self.vel * step).astype(np.float32)
        if self.spin is not None:
            r = self.base[:, :9].astype(np.float64).reshape(-1, 3, 3)
            for _ in range(step):
                r = np.einsum("nij,njk->nik", self.spin, r)
            xf[:, :9] = r.reshape(-1, 9).astype(np.float32)
        stat = np.asarray(self.static, dtype=bool)
        xf[stat] = self.base[stat]
        return xf


def hull_points(rng, radius, npts=16):
    """npts points on a jittered sphere of the given radius."""
    p = rng.normal(size=(npts, 3))
    p /= np.linalg.norm(p, axis=1, keepdims=True)
    p *= radius * rng.uniform(0.8, 1.0, size=(npts, 1))
    return p.astype(np.float32)


def stack_scene(n_side=5, extra=True, seed=1, plane_ground=False):
    """C1: n^3 unit boxes (half extent 1) on a lattice of spacing 2.0 over a static ground."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    if plane_ground:
        g = sc.add_shape("plane", (0.0, 1.0, 0.0), 0.0)
        gpos = (0.0, 0.0, 0.0)
    else:
        g = sc.add_shape("box", (50.0, 50.0, 50.0))
        gpos = (0.0, -50.0, 0.0)
    b = sc.add_shape("box", (1.0, 1.0, 1.0))
    pos = [gpos]
    sc.body_shape.append(g); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(0)
    for i in range(n_side):
        for j in range(n_side):
            for k in range(n_side):
                pos.append((2.0 * i - n_side + 1.0, 1.0 + 2.0 * j, 2.0 * k - n_side + 1.0))
                sc.body_shape.append(b); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    if extra:
        s = sc.add_shape("sphere", 0.5)
        h = sc.add_shape("hull", hull_points(rng, 0.6))
        top = 2.0 * n_side
        for (sh, p) in [(s, (0.0, top + 0.49, 0.0)), (s, (0.9, top + 0.5, 0.1)), (h, (-2.0, top + 0.55, 0.0))]:
            pos.append(p)
            sc.body_shape.append(sh); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    n = len(pos)
    rot = np.tile(np.eye(3), (n, 1, 1))
    sc.base = make_xf(rot, np.asarray(pos))
    jitter = rng.uniform(-0.004, 0.004, size=(n, 3))
    sc.base[1:, 9:] += jitter[1:].astype(np.float32)
    sc.vel = rng.uniform(-0.003, 0.003, size=(n, 3))
    sc.spin = small_rotation(rng, n, 0.002)
    sc.extent = 2.0 * n_side + 4.0
    return sc


def bin_scene(n=2000, seed=2, spacing=0.82, footprint=None, rotate=True, mix=(0.4, 0.4, 0.2), shape_variants=64):
    """C2: mixed boxes / spheres / 16-point hulls on a jittered lattice inside a closed bin of 5 static boxes."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    if footprint is None:
        footprint = max(4, int(round((n ** (1.0 / 3.0)) * 0.9)))
    side = footprint * spacing
    layers = (n + footprint * footprint - 1) // (footprint * footprint)
    height = layers * spacing + 2.0
    t = 1.0  # wall half thickness
    walls = [
        ((side / 2 + 2 * t, t, side / 2 + 2 * t), (side / 2, -t, side / 2)),                       # floor
        ((t, height / 2, side / 2 + 2 * t), (-t, height / 2, side / 2)),                           # -x
        ((t, height / 2, side / 2 + 2 * t), (side + t, height / 2, side / 2)),                     # +x
        ((side / 2 + 2 * t, height / 2, t), (side / 2, height / 2, -t)),                           # -z
        ((side / 2 + 2 * t, height / 2, t), (side / 2, height / 2, side + t)),                     # +z
    ]
    pos, rots = [], []
    for he, p in walls:
        sid = sc.add_shape("box", he)
        sc.body_shape.append(sid); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(0)
        pos.append(p); rots.append(np.eye(3))
    # a palette of shapes (the reference shares shape objects between bodies too)
    boxes = [sc.add_shape("box", tuple(rng.uniform(0.25, 0.5, size=3))) for _ in range(shape_variants)]
    spheres = [sc.add_shape("sphere", float(rng.uniform(0.25, 0.5))) for _ in range(shape_variants)]
    hulls = [sc.add_shape("hull", hull_points(rng, float(rng.uniform(0.3, 0.5)))) for _ in range(shape_variants)]
    kind = rng.choice(3, size=n, p=list(mix))
    which = rng.integers(0, shape_variants, size=n)
    idx = np.arange(n)
    ix = idx % footprint
    iz = (idx // footprint) % footprint
    iy = idx // (footprint * footprint)
    p = np.stack([(ix + 0.5) * spacing, (iy + 0.5) * spacing + 0.02, (iz + 0.5) * spacing], axis=1)
    p += rng.uniform(-0.06, 0.06, size=(n, 3))
    r = random_rotations(rng, n) if rotate else np.tile(np.eye(3), (n, 1, 1))
    for k in range(n):
        sid = (boxes, spheres, hulls)[kind[k]][which[k]]
        sc.body_shape.append(sid); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    allpos = np.concatenate([np.asarray(pos), p], axis=0)
    allrot = np.concatenate([np.asarray(rots), r], axis=0)
    sc.base = make_xf(allrot, allpos)
    nb = len(allpos)
    sc.vel = rng.uniform(-0.004, 0.004, size=(nb, 3))
    sc.spin = small_rotation(rng, nb, 0.003)
    sc.extent = float(max(side, height))
    return sc


def heightfield(cells, cell=0.5, amp=3.0, seed=3):
    """(cells+1)^2 vertices, 2*cells^2 triangles, 2-octave value noise."""
    rng = np.random.default_rng(SEED + seed)
    nv = cells + 1

    def octave(freq):
        g = rng.uniform(-1, 1, size=(freq + 2, freq + 2))
        u = np.linspace(0, freq, nv)
        i = np.minimum(u.astype(int), freq)
        f = u - i
        f = f * f * (3 - 2 * f)
        a = g[np.ix_(i, i)]; b = g[np.ix_(i + 1, i)]; c = g[np.ix_(i, i + 1)]; d = g[np.ix_(i + 1, i + 1)]
        fx = f[:, None]; fz = f[None, :]
        return (a * (1 - fx) + b * fx) * (1 - fz) + (c * (1 - fx) + d * fx) * fz

    h = amp * (0.7 * octave(max(2, cells // 32)) + 0.3 * octave(max(4, cells // 8)))
    xs = np.arange(nv) * cell
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    verts = np.stack([X, h, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(cells), np.arange(cells), indexing="ij")
    v00 = (i * nv + j).ravel(); v10 = ((i + 1) * nv + j).ravel(); v01 = (i * nv + j + 1).ravel(); v11 = ((i + 1) * nv + j + 1).ravel()
    tris = np.concatenate([np.stack([v00, v01, v10], axis=1), np.stack([v10, v01, v11], axis=1)], axis=1).reshape(-1, 3)
    return verts, tris.astype(np.int32), h.astype(np.float32)


def terrain_scene(cells=64, n=200, seed=4, cell=0.5, mix=(0.45, 0.45, 0.10)):
    """C3: hulls and spheres resting on a BVH triangle-mesh heightfield."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    verts, tris, h = heightfield(cells, cell=cell, seed=seed)
    m = sc.add_shape("mesh", verts, tris)
    sc.body_shape.append(m); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(0)
    variants = 16
    hulls = [sc.add_shape("hull", hull_points(rng, float(rng.uniform(0.3, 0.5)))) for _ in range(variants)]
    sph = sc.add_shape("sphere", 0.4)
    boxs = [sc.add_shape("box", tuple(rng.uniform(0.25, 0.45, size=3))) for _ in range(variants)]
    size = cells * cell
    xz = rng.uniform(1.0, size - 1.0, size=(n, 2))
    gi = np.clip((xz / cell).astype(int), 0, cells - 1)
    ground = h[gi[:, 0], gi[:, 1]]
    kind = rng.choice(3, size=n, p=list(mix))   # hull / sphere / box shares
    y = ground + np.where(kind == 1, 0.4, 0.45) + rng.uniform(-0.05, 0.25, size=n)
    pos = np.stack([xz[:, 0], y, xz[:, 1]], axis=1)
    for k in range(n):
        sid = hulls[rng.integers(variants)] if kind[k] == 0 else (sph if kind[k] == 1 else boxs[rng.integers(variants)])
        sc.body_shape.append(sid); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    allpos = np.concatenate([np.zeros((1, 3)), pos], axis=0)
    allrot = np.concatenate([np.eye(3)[None], random_rotations(rng, n)], axis=0)
    sc.base = make_xf(allrot, allpos)
    sc.vel = rng.uniform(-0.004, 0.004, size=(n + 1, 3))
    sc.spin = small_rotation(rng, n + 1, 0.003)
    sc.extent = float(size)
    return sc


def split_mesh_into_parts(sc, nparts=3, short_parts=(1,)):
    """Replace the scene's one-part mesh (shape 0) by the same triangles split into `nparts` IndexedMesh parts with their own
    (compacted) vertex arrays; the parts listed in short_parts use 16-bit indices (ScalarType.SHORT)."""
    kind, verts, tris = sc.shapes[0]
    assert kind == "mesh"
    cuts = np.linspace(0, len(tris), nparts + 1).astype(int)
    parts = []
    for p in range(nparts):
        t = tris[cuts[p]:cuts[p + 1]]
        used, inv = np.unique(t.reshape(-1), return_inverse=True)
        idx = inv.reshape(-1, 3)
        if p in short_parts:
            assert len(used) < 65536
            idx = idx.astype(np.uint16)
        else:
            idx = idx.astype(np.int32)
        parts.append((np.ascontiguousarray(verts[used]), np.ascontiguousarray(idx)))
    sc.shapes[0] = ("meshparts", parts)
    return sc


def worlds_scene(num_worlds=32, seed=5):
    """C4: independent worlds of 64 bodies: one static floor box + 63 dice (half extent 0.5) in a jittered 4x4x4 lattice."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    sc.num_worlds = num_worlds
    floor = sc.add_shape("box", (4.0, 0.5, 4.0))
    die = sc.add_shape("box", (0.5, 0.5, 0.5))
    pos, rots = [], []
    for w in range(num_worlds):
        pos.append((2.0, -0.5, 2.0)); rots.append(np.eye(3))
        sc.body_shape.append(floor); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(w)
        cells = [(i, j, k) for j in range(4) for i in range(4) for k in range(4)][:63]
        p = np.asarray(cells, dtype=np.float64) * 1.02 + 0.5 + rng.uniform(-0.02, 0.02, size=(63, 3))
        p[:, 1] += 0.01
        r = small_rotation(rng, 63, 0.05)
        for q in range(63):
            pos.append(tuple(p[q])); rots.append(r[q])
            sc.body_shape.append(die); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(w)
    sc.base = make_xf(np.asarray(rots), np.asarray(pos))
    n = len(pos)
    sc.vel = rng.uniform(-0.003, 0.003, size=(n, 3))
    sc.spin = small_rotation(rng, n, 0.002)
    sc.extent = 8.0
    return sc


def spheres_scene(n=20000, seed=6, radius=0.5, fill=0.40):
    """C5: n spheres of one radius in a cube sized for the given packing fraction (jittered lattice)."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    s = sc.add_shape("sphere", radius)
    vol = n * (4.0 / 3.0) * np.pi * radius ** 3 / fill
    side = vol ** (1.0 / 3.0)
    m = int(np.ceil(n ** (1.0 / 3.0)))
    sp = side / m
    idx = np.arange(n)
    p = np.stack([idx % m, (idx // m) % m, idx // (m * m)], axis=1).astype(np.float64) * sp + 0.5 * sp
    p += rng.uniform(-0.5, 0.5, size=(n, 3)) * max(0.0, sp - 2 * radius * 0.9)
    for _ in range(n):
        sc.body_shape.append(s); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    sc.base = make_xf(np.tile(np.eye(3), (n, 1, 1)), p)
    sc.vel = rng.uniform(-0.01, 0.01, size=(n, 3))
    sc.spin = None
    sc.extent = float(side)
    return sc


def compound_scene(n=300, seed=8, plane_ground=True, spacing=0.9, compound_share=0.5, nested=False):
    """compounds (dumbbells, L-brackets, hull+sphere clusters, three-box crosses) mixed with plain boxes,
    spheres and hulls on a jittered lattice over a static plane (or box) floor, close enough that compound x {sphere, box,
    hull, plane / static box, compound} pairs all occur, some of them penetrating."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    if plane_ground:
        g = sc.add_shape("plane", (0.0, 1.0, 0.0), 0.0)
        gpos = (0.0, 0.0, 0.0)
    else:
        g = sc.add_shape("box", (40.0, 1.0, 40.0))
        gpos = (0.0, -1.0, 0.0)
    sc.body_shape.append(g); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(0)
    eye = np.eye(3)

    def cxf(rot, pos):
        return make_xf(np.asarray(rot).reshape(-1, 3, 3), np.asarray(pos, dtype=np.float64).reshape(-1, 3))

    s_small = sc.add_shape("sphere", 0.3)
    s_big = sc.add_shape("sphere", 0.4)
    bar = sc.add_shape("box", (0.45, 0.12, 0.12))
    slab = sc.add_shape("box", (0.4, 0.15, 0.3))
    post = sc.add_shape("box", (0.15, 0.4, 0.15))
    hl = sc.add_shape("hull", hull_points(rng, 0.35))
    rz = small_rotation(rng, 1, 0.7)[0]
    compounds = [
        sc.add_shape("compound", [s_small, bar, s_big], cxf([eye, eye, eye], [(-0.5, 0, 0), (0, 0, 0), (0.5, 0, 0)])),   # dumbbell
        sc.add_shape("compound", [slab, post], cxf([eye, eye], [(0, -0.25, 0), (0.25, 0.3, 0)])),                         # L-bracket
        sc.add_shape("compound", [hl, s_small], cxf([rz, eye], [(-0.2, 0, 0.1), (0.3, 0.1, -0.1)])),                      # hull + sphere
        sc.add_shape("compound", [bar, bar, bar], cxf([eye, [[0, -1, 0], [1, 0, 0], [0, 0, 1]], [[0, 0, 1], [0, 1, 0], [-1, 0, 0]]],
                                                      [(0, 0, 0), (0, 0, 0), (0, 0, 0)])),                                  # 3-axis cross
        sc.add_shape("compound", [s_big], cxf([eye], [(0.0, 0.2, 0.0)])),                                                 # single offset child
    ]
    if nested == "identity":
        # every compound wrapped in an outer compound with an identity child transform: the same leaves, the same floats
        compounds = [sc.add_shape("compound", [c], cxf([eye], [(0.0, 0.0, 0.0)])) for c in compounds]
    elif nested:
        # children that are CompoundShapes themselves (sh/CompoundShape.java accepts any CollisionShape), one and two levels
        r1, r2, r3 = small_rotation(rng, 3, 0.9)
        n1 = sc.add_shape("compound", [compounds[0], post], cxf([r1, eye], [(0.1, 0.35, 0.0), (-0.1, -0.3, 0.05)]))
        n2 = sc.add_shape("compound", [n1, s_small], cxf([r2, eye], [(0.0, 0.2, 0.1), (0.2, -0.4, 0.0)]))
        n3 = sc.add_shape("compound", [compounds[2], compounds[1]], cxf([r3, r1], [(-0.2, 0.0, 0.0), (0.25, 0.1, 0.0)]))
        compounds = compounds + [n1, n2, n3, n2, n3]
    plain = [sc.add_shape("box", tuple(rng.uniform(0.25, 0.45, size=3))) for _ in range(4)]
    plain += [s_small, s_big, hl, sc.add_shape("hull", hull_points(rng, 0.4))]
    m = max(2, int(np.ceil(n ** (1.0 / 3.0))))
    idx = np.arange(n)
    p = np.stack([(idx % m) * spacing, (idx // (m * m)) * spacing * 0.8 + 0.55, ((idx // m) % m) * spacing], axis=1).astype(np.float64)
    p += rng.uniform(-0.12, 0.12, size=(n, 3))
    is_c = rng.uniform(size=n) < compound_share
    for k in range(n):
        sid = compounds[rng.integers(len(compounds))] if is_c[k] else plain[rng.integers(len(plain))]
        sc.body_shape.append(sid); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    allpos = np.concatenate([np.asarray([gpos]), p], axis=0)
    allrot = np.concatenate([eye[None], random_rotations(rng, n)], axis=0)
    sc.base = make_xf(allrot, allpos)
    sc.vel = rng.uniform(-0.004, 0.004, size=(n + 1, 3))
    sc.spin = small_rotation(rng, n + 1, 0.004)
    sc.extent = float(m * spacing + 2.0)
    return sc


def terrain_compound_scene(cells=32, n=150, seed=15, cell=0.5):
    """Compounds (and a few plain convex bodies) resting on a BVH triangle-mesh heightfield: compound x mesh runs
    ConvexConcave per child (SURVEY §8f rank 3 with §8a N10)."""
    rng = np.random.default_rng(SEED + seed)
    sc = Scene()
    verts, tris, h = heightfield(cells, cell=cell, amp=1.5, seed=seed)
    m = sc.add_shape("mesh", verts, tris)
    sc.body_shape.append(m); sc.static.append(True); sc.group.append(2); sc.mask.append(-1 ^ 2); sc.world.append(0)
    eye = np.eye(3)
    sph = sc.add_shape("sphere", 0.3)
    bar = sc.add_shape("box", (0.45, 0.12, 0.12))
    slab = sc.add_shape("box", (0.35, 0.12, 0.3))
    hl = sc.add_shape("hull", hull_points(rng, 0.3))
    compounds = [
        sc.add_shape("compound", [sph, bar, sph], make_xf(np.stack([eye] * 3), np.asarray([(-0.5, 0, 0), (0, 0, 0), (0.5, 0, 0)]))),
        sc.add_shape("compound", [slab, hl], make_xf(np.stack([eye, small_rotation(rng, 1, 0.6)[0]]), np.asarray([(0, -0.1, 0), (0.2, 0.25, 0.1)]))),
    ]
    plain = [sph, slab, hl]
    size = cells * cell
    xz = rng.uniform(1.0, size - 1.0, size=(n, 2))
    gi = np.clip((xz / cell).astype(int), 0, cells - 1)
    y = h[gi[:, 0], gi[:, 1]] + 0.3 + rng.uniform(-0.1, 0.25, size=n)
    pos = np.stack([xz[:, 0], y, xz[:, 1]], axis=1)
    for k in range(n):
        sid = compounds[rng.integers(2)] if rng.uniform() < 0.7 else plain[rng.integers(3)]
        sc.body_shape.append(sid); sc.static.append(False); sc.group.append(1); sc.mask.append(-1); sc.world.append(0)
    sc.base = make_xf(np.concatenate([eye[None], random_rotations(rng, n)], axis=0), np.concatenate([np.zeros((1, 3)), pos], axis=0))
    sc.vel = rng.uniform(-0.004, 0.004, size=(n + 1, 3))
    sc.spin = small_rotation(rng, n + 1, 0.003)
    sc.extent = float(size)
    return sc


# ---- build the same scene on both sides ---------------------------------------------------------------
def build_gpu(pkg, sc, mode, max_pairs=None, **kw):
    n = sc.n
    kw.setdefault("raw_records", True)   # the parity helpers compare the raw detector records too
    gw = pkg.GpuCollisionWorld(mode=mode, max_bodies=max(n, 16), max_pairs=max_pairs or max(16 * n, 4096),
                               num_worlds=sc.num_worlds, **kw)
    ids = []
    for s in sc.shapes:
        if s[0] == "box":
            ids.append(gw.BoxShape(s[1]))
        elif s[0] == "sphere":
            ids.append(gw.SphereShape(s[1]))
        elif s[0] == "hull":
            ids.append(gw.ConvexHullShape(s[1]))
        elif s[0] == "plane":
            ids.append(gw.StaticPlaneShape(s[1], s[2]))
        elif s[0] == "mesh":
            ids.append(gw.BvhTriangleMeshShape(s[1], s[2]))
        elif s[0] == "meshparts":
            ids.append(gw.BvhTriangleMeshShapeParts(s[1]))
        elif s[0] == "compound":
            ids.append(gw.CompoundShape([ids[c] for c in s[1]], s[2]))
    shapes = np.asarray([ids[k] for k in sc.body_shape], dtype=np.int32)
    gw.addCollisionObjects(shapes, sc.base, sc.group, sc.mask, np.asarray(sc.static, dtype=np.int32), sc.world)
    gw.scene_shape_ids = ids          # scene shape index -> registered shape id
    return gw


def build_oracle(sc, mode, brute_force=False, world_aabb=None):
    import orc
    ow = orc.OracleWorld(mode=mode, brute_force=brute_force, world_aabb=world_aabb)
    ids = []
    for s in sc.shapes:
        if s[0] == "box":
            ids.append(ow.box(*[float(v) for v in np.asarray(s[1], dtype=np.float32)]))
        elif s[0] == "sphere":
            ids.append(ow.sphere(float(np.float32(s[1]))))
        elif s[0] == "hull":
            ids.append(ow.hull(s[1]))
        elif s[0] == "plane":
            ids.append(ow.plane([float(v) for v in np.asarray(s[1], dtype=np.float32)], float(s[2])))
        elif s[0] == "mesh":
            ids.append(ow.mesh(s[1], s[2]))
        elif s[0] == "meshparts":
            ids.append(ow.mesh_parts(s[1]))
        elif s[0] == "compound":
            ids.append(ow.compound([ids[c] for c in s[1]], s[2]))
    for k in range(sc.n):
        ow.body(ids[sc.body_shape[k]], sc.base[k], sc.group[k], sc.mask[k], sc.static[k], sc.world[k])
    ow.scene_shape_ids = ids
    return ow


def build_both(pkg, sc, mode, world_aabb=None, **kw):
    if world_aabb is not None:
        kw["world_aabb"] = world_aabb
    # device modes 2/3 are AxisSweep3 / AxisSweep3_32 = oracle modes 3/4
    return build_gpu(pkg, sc, mode, **kw), build_oracle(sc, {2: 3, 3: 4}.get(mode, mode), world_aabb=world_aabb)
