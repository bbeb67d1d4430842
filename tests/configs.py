"""BASELINE.json configs[0], [2], [3], [4] as seeded generators (configs[1], the 4M settled bed, is bench.py's).

Every generator takes a `scale` so that the same geometry exists at the size BASELINE.json names (parity against digests
of the UNMODIFIED reference, tests/golden/big_*.npz, and throughput in tools/config_runs.py) and at a size the CPU oracle
steps in seconds.  The beds are DENSE from step 0: particles sit on a simple-cubic lattice whose pitch is a hair below
two radii, clipped by the analytic surface the triangle mesh approximates, so pair contacts and particle-triangle
contacts exist at the first force evaluation -- a lattice dropped from above would spend its first 10^4 steps in free fall
(round 1's stand-ins measured exactly that).

  C1  10,648 monodisperse spheres settling in a box of primitive planes           (configs[0])
  C3  ~1.0 M monodisperse spheres in a conical hopper STL with an open outlet       (configs[2])
  C4  ~0.5 M bonded spheres (INL bond/nonlinear) under a moving stress plate        (configs[3])
  C5  rotating drum (fix move/mesh rotate) filled to its axis; `bricks` = the share of one of 8 GPUs by default (configs[4])
"""
import numpy as np
import cases

HERTZ_CDT = "model hertz tangential history rolling_friction cdt"


def _cubic(lo, hi, pitch):
    ax = [np.arange(lo[d] + 0.5 * pitch, hi[d], pitch) for d in range(3)]
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    return g


def _finish(c, x, rad, rho=2500.0, v=None, jitter=0.0, seed=None):
    n = len(x)
    if jitter:
        # a perfect lattice puts every interior sphere in exact force balance: its net force would be rounding noise of
        # contact forces ~50 weights and no relative tolerance could hold.  A seeded jitter of a thousandth of a radius
        # leaves net forces that are small against the contact forces but far above their rounding noise.
        x = x + np.random.default_rng(cases.SEED if seed is None else seed).uniform(-jitter, jitter, x.shape)
    c.update(tag=np.arange(1, n + 1, dtype=np.int32), type=np.ones(n, np.int32), mask=np.ones(n, np.int32), x=np.ascontiguousarray(x),
             v=np.zeros((n, 3)) if v is None else v, omega=np.zeros((n, 3)), radius=np.full(n, rad), density=np.full(n, rho))
    return c


def _base(name, lo, hi, skin=0.001, dt=1e-5, model=HERTZ_CDT):
    T = 1
    props = [("youngsModulus", "peratomtype", [5e6]), ("poissonsRatio", "peratomtype", [0.45]),
             ("coefficientRestitution", "peratomtypepair", [0.3]), ("coefficientFriction", "peratomtypepair", [0.5]),
             ("coefficientRollingFriction", "peratomtypepair", [0.1])]
    return dict(name=name, lo=list(lo), hi=list(hi), periodic=[0, 0, 0], ntypes=T, skin=skin, dt=dt, props=props, pair=model,
                walls=[], gravity=(9.81, [0.0, 0.0, -1.0]), freeze=0)


def C1(scale=1.0):
    """configs[0]: 22^3 = 10,648 spheres, r = 2.5 mm, falling into a box of five primitive planes"""
    n = max(4, int(round(22 * scale ** (1.0 / 3.0))))
    return cases.case_box(n3=(n, n, n), name="C1")


def mesh_hopper(cx, cy, Rc, Ro, zc, ztop, nseg, nz_cone, nz_cyl):
    """cylinder (radius Rc, zc..ztop) on a cone narrowing to an OPEN outlet of radius Ro at z = 0; nseg x (nz_cone + nz_cyl)
    quads = twice as many triangles"""
    t = []
    p = lambda r, a, z: [cx + r * np.cos(a), cy + r * np.sin(a), z]
    rings = [(Ro + (Rc - Ro) * k / nz_cone, zc * k / nz_cone) for k in range(nz_cone + 1)] + [(Rc, zc + (ztop - zc) * k / nz_cyl) for k in range(1, nz_cyl + 1)]
    for (r0, z0), (r1, z1) in zip(rings[:-1], rings[1:]):
        for k in range(nseg):
            a0, a1 = 2 * np.pi * k / nseg, 2 * np.pi * (k + 1) / nseg
            t += cases._quad(p(r0, a0, z0), p(r0, a1, z0), p(r1, a1, z1), p(r1, a0, z1))
    return np.asarray(t, np.float64)


def C3(scale=1.0, nseg=None, seed=None):
    """configs[2]: hopper discharge.  scale 1 -> ~1.0 M spheres of r = 2 mm filling a hopper of radius 0.2 m (cone half
    angle 30 deg, outlet radius 40 mm, ~16 k triangles); the outlet is open, the column above it starts to fall at once into
    a catch box of primitive planes below the hopper (a sphere that left a non-periodic box would be lost)"""
    rad = 0.002
    pitch = 1.995 * rad
    s = scale ** (1.0 / 3.0)
    Rc, Ro = 0.2 * s, max(0.04 * s, 6 * rad)
    zc = (Rc - Ro) / np.tan(np.radians(30.0))
    H = 0.41 * s          # fill height above the cone
    ztop = zc + H + 0.02
    nseg = nseg or max(16, int(256 * s))
    cx = cy = Rc + 0.01
    c = _base("C3", [0.0, 0.0, -0.25 * s - 0.02], [2 * cx, 2 * cx, ztop + 0.01])
    mesh = mesh_hopper(cx, cy, Rc, Ro, zc, ztop, nseg, max(4, int(24 * s)), max(2, int(8 * s)))
    g = _cubic([cx - Rc, cy - Rc, 0.0], [cx + Rc, cy + Rc, zc + H], pitch)
    rho = np.hypot(g[:, 0] - cx, g[:, 1] - cy)
    Rz = np.where(g[:, 2] < zc, Ro + (Rc - Ro) * g[:, 2] / zc, Rc)
    cosa = np.where(g[:, 2] < zc, np.cos(np.radians(30.0)), 1.0)
    # interior: lattice sites at least a pitch clear of the shell below; shell: one layer of spheres ON the wall (rings of
    # pitch-spaced spheres pressed 1 % of a radius into the cone and the cylinder), so that ~5 % of the bed touches triangles
    inside = (Rz - rho) * cosa >= 0.98 * rad + pitch
    inside &= g[:, 2] >= 0.6 * rad
    shell = []
    sl = np.hypot(Rc - Ro, zc)                                   # slant length of the cone
    for k in range(1, int(sl / pitch)):
        t = (k + 0.5) * pitch / sl
        zr = t * zc; rr = Ro + (Rc - Ro) * t
        # sphere centre: 0.99 rad off the cone surface along its inward normal (cos a, -sin a in the (r, z) plane)
        rr_c = rr - 0.99 * rad * np.cos(np.radians(30.0)); zr_c = zr + 0.99 * rad * np.sin(np.radians(30.0))
        m = max(3, int(2 * np.pi * rr_c / pitch))
        a = 2 * np.pi * (np.arange(m) + 0.5 * (k % 2)) / m
        shell.append(np.stack([cx + rr_c * np.cos(a), cy + rr_c * np.sin(a), np.full(m, zr_c)], 1))
    for k in range(int(H / pitch)):
        zr = zc + (k + 1.0) * pitch
        rr_c = Rc - 0.99 * rad
        m = int(2 * np.pi * rr_c / pitch)
        a = 2 * np.pi * (np.arange(m) + 0.5 * (k % 2)) / m
        shell.append(np.stack([cx + rr_c * np.cos(a), cy + rr_c * np.sin(a), np.full(m, zr)], 1))
    shell = np.concatenate(shell)
    shell = shell[shell[:, 2] >= 0.6 * rad]
    _finish(c, np.concatenate([g[inside], shell]), rad, jitter=0.001 * rad, seed=seed)
    zf = c["lo"][2] + 0.005
    c["walls"] = [("catch", HERTZ_CDT + " primitive type 1 zplane %.17g" % zf), ("cx0", HERTZ_CDT + " primitive type 1 xplane 0.001"),
                  ("cx1", HERTZ_CDT + " primitive type 1 xplane %.17g" % (2 * cx - 0.001)), ("cy0", HERTZ_CDT + " primitive type 1 yplane 0.001"),
                  ("cy1", HERTZ_CDT + " primitive type 1 yplane %.17g" % (2 * cx - 0.001))]
    c["meshes"] = [("hopper", 1, mesh)]
    c["mesh_walls"] = [("mw", HERTZ_CDT + " mesh n_meshes 1 meshes hopper")]
    return c


def C4(scale=1.0, plate_speed=-0.05):
    """configs[3]: bonded-sphere block under uniaxial compression.  scale 1 -> 80 x 80 x 78 = 499,200 spheres (r = 3 mm) on a
    lattice, INL `cohesion bond/nonlinear` bonds created at step 2 between face neighbours and in-plane diagonal neighbours
    (5 per sphere, ~2.4 M), frozen bottom layer on a floor plane, a `mesh/surface/stress` plate moving down onto the top"""
    s = scale ** (1.0 / 3.0)
    nx = max(4, int(round(80 * s))); nz = max(4, int(round(78 * s)))
    rad = 0.003
    p = 2.02 * rad; pz = 1.15 * p    # in-plane diagonals (1.414 p) bond, out-of-plane ones (1.52 p) do not
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(nx), np.arange(nz), indexing="ij"), -1).reshape(-1, 3).astype(np.float64)
    order = np.lexsort((g[:, 0], g[:, 1], g[:, 2]))   # bottom layer first: the frozen group is the first nx*nx tags
    g = g[order]
    x = np.stack([(g[:, 0] + 0.5) * p + 0.01, (g[:, 1] + 0.5) * p + 0.01, (g[:, 2]) * pz + 1.0005 * rad], 1)
    L = nx * p + 0.02
    top = x[:, 2].max() + 1.0002 * rad   # the plate touches the top layer after two steps
    c = _base("C4", [0.0, 0.0, 0.0], [L, L, top + 0.05], model="model hertz tangential history cohesion bond/nonlinear", skin=0.001, dt=1e-5)
    TT = lambda v: [float(v)]
    sfx = "nonlinear"
    c["props"] += [("radiusMultiplierBond" + sfx, "peratomtypepair", TT(0.8)),
                   ("dampingNormalForceBond" + sfx, "peratomtypepair", TT(0.1)), ("dampingTangentialForceBond" + sfx, "peratomtypepair", TT(0.1)),
                   ("dampingNormalTorqueBond" + sfx, "peratomtypepair", TT(0.1)), ("dampingTangentialTorqueBond" + sfx, "peratomtypepair", TT(0.1)),
                   ("tsCreateBond" + sfx, "scalar", [2]), ("createDistanceBond" + sfx, "peratomtypepair", TT(1.45 * p)),
                   ("maxDistanceBond" + sfx, "peratomtypepair", TT(1.6 * p))]
    k = 1e9
    # (the compression branch of the nonlinear law is k*sqrt(displacement) WITHOUT the area factor,
    # cohesion_model_bond_nonlinear.h:603-612: its constants are of order 10, the others of order 1e9)
    for nm, val in (("K_fn1", 20.0), ("Ku_fn1", 80.0), ("Kc_fn1", 10.0), ("K_fn2", k), ("Ku_fn2", 4 * k), ("Kc_fn2", 0.5 * k), ("K_ft", 0.5 * k),
                    ("K_tn", 0.4 * k), ("Ku_tn", 1.6 * k), ("Kc_tn", 0.2 * k), ("K_tt", 0.6 * k), ("Ku_tt", 2.4 * k), ("Kc_tt", 0.3 * k)):
        c["props"].append(("stiffnessPerUnitArea" + nm, "peratomtypepair", TT(val)))
    _finish(c, x, rad)
    c["mask"][: nx * nx] |= 2; c["freeze"] = 2
    c["walls"] = [("floor", "model hertz tangential history primitive type 1 zplane 0.0")]
    plate = np.asarray(cases._quad([0.005, 0.005, top], [L - 0.005, 0.005, top], [L - 0.005, L - 0.005, top], [0.005, L - 0.005, top]))
    c["meshes"] = [("plate", 1, plate)]
    c["mesh_stress"] = ["plate"]
    c["mesh_moves"] = [("plate", "linear 0. 0. %.17g" % plate_speed)]
    c["mesh_walls"] = [("mw", "model hertz tangential history mesh n_meshes 1 meshes plate")]
    return c


def C5(scale=1.0, bricks=8, period=2.0, nseg=None, part=None, yrange=None):
    """configs[4]: rotating drum (axis y), filled to its axis with r = 2.5 mm spheres on a dense lattice clipped by the
    mantle.  scale 1, bricks 8 -> one GPU's share of the 16.8 M spheres: a drum of radius 0.6 m and 1/8 of its 3.7 m.
    bricks 1, part (k, n) -> the FULL drum's geometry with only the spheres of the k-th of n slabs along the axis (what rank k
    of an n-GPU run holds; tags start at 1 in every part: the caller offsets them); yrange (lo, hi) -> the same for an explicit
    interval of the axis (a rank's sub-box of the engine's own decomposition)"""
    rad = 0.0025
    pitch = 1.995 * rad
    s = scale ** (1.0 / 3.0)
    R = 0.6 * s
    Ly = 3.7 * s / bricks   # the full drum is 3.7 m long (16.8 M spheres in its lower half); a brick holds 1/bricks of it
    nseg = nseg or max(16, int(256 * s))
    cx, cz = R + 0.02, R + 0.02
    y0, y1 = 0.0, Ly
    c = _base("C5", [0.0, -0.02, 0.0], [2 * cx, Ly + 0.02, 2 * cz])
    ya, yb = y0 + 0.2 * rad, y1 - 0.2 * rad
    ys_all = np.arange(ya + 0.5 * pitch, yb, pitch)
    ysh_all = np.arange(y0 + 1.0 * rad, y1 - 0.97 * rad, pitch)
    if part is not None:
        k, nparts = part
        lo_p, hi_p = y0 + (y1 - y0) * k / nparts, y0 + (y1 - y0) * (k + 1) / nparts
        yrange = (lo_p, hi_p)
    if yrange is not None:
        ys_all = ys_all[(ys_all >= yrange[0]) & (ys_all < yrange[1])]
        ysh_all = ysh_all[(ysh_all >= yrange[0]) & (ysh_all < yrange[1])]
    ax = [np.arange(cx - R + 0.5 * pitch, cx + R, pitch), ys_all, np.arange(cz - R + 0.5 * pitch, cz, pitch)]   # lower half
    g = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    rho = np.hypot(g[:, 0] - cx, g[:, 2] - cz)
    inside = (R - rho >= 0.98 * rad + pitch) & (g[:, 1] - y0 >= 0.97 * rad) & (y1 - g[:, 1] >= 0.97 * rad)
    # one layer of spheres ON the mantle (lower half), pressed 1 % of a radius into it
    rr_c = R - 0.99 * rad
    m = int(np.pi * rr_c / pitch)
    ys = ysh_all
    a = np.pi + np.pi * (np.arange(m) + 0.5) / m        # angles pi..2pi: below the axis
    A, Y = np.meshgrid(a, ys, indexing="ij")
    shell = np.stack([cx + rr_c * np.cos(A.ravel()), Y.ravel(), cz + rr_c * np.sin(A.ravel())], 1)
    _finish(c, np.concatenate([g[inside], shell]), rad, jitter=0.001 * rad, seed=None if yrange is None else cases.SEED + 1 + int(1000 * yrange[0]))
    c["meshes"] = [("drum", 1, cases.mesh_drum(cx, cz, R, y0, y1, nseg=nseg))]
    c["mesh_moves"] = [("drum", "rotate origin %.17g 0. %.17g axis 0. 1. 0. period %s" % (cx, cz, period))]
    c["mesh_walls"] = [("mw", HERTZ_CDT + " mesh n_meshes 1 meshes drum")]
    return c


CONFIGS = {"C1": C1, "C3": C3, "C4": C4, "C5": C5}
# sizes the CPU oracle steps in seconds (tests/test_gpu_configs.py, against the oracle) ...
MINI = {"C1": 0.1, "C3": 0.02, "C4": 0.02, "C5": 0.02}
# ... and the checkpoints of the full-size digests from the unmodified reference (tests/golden/make_golden_big.py)
BIG_CHECKPOINTS = {"C1": [0, 1, 2, 10, 200], "C3": [0, 1, 10], "C4": [0, 1, 2, 3, 10], "C5": [0, 1, 10]}
