"""Seeded test cases shared by the golden generator (reference), the oracle and the CUDA
engine.  A case is a plain dict; `to_deck` renders it as a LIGGGHTS input script + data
file for the unmodified reference, `apply` replays it on an Engine-like object (the CUDA
engine or, in tests only, the CPU oracle) through the C ABI."""
import os
import re
import numpy as np

SEED = 20261017


def lattice(n3, pitch, origin, jitter, rng):
    nx, ny, nz = n3
    g = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 3)
    return np.asarray(origin) + g * pitch + rng.uniform(-jitter, jitter, (len(g), 3))


def case_box(n3=(4, 4, 4), model="model hertz tangential history rolling_friction cdt", seed=SEED,
             poly=False, periodic=(0, 0, 0), ntypes=1, hooke=False, frozen=0, cyl=False, shear=False, name="box",
             settings="", bond=None):
    """particles on a jittered lattice falling under gravity onto a floor inside side walls"""
    rng = np.random.default_rng(seed)
    rad = 0.0025
    pitch = 2.05 * rad if not poly else 2.05 * 0.003
    L = max(n3[0], n3[1]) * pitch
    lo = [0.0, 0.0, 0.0]; hi = [L, L, 4 * n3[2] * pitch + 0.02]
    if cyl:  # the cylinder (radius 0.64 L about the box centre) reaches outside the lattice footprint
        lo = [-0.2 * L, -0.2 * L, 0.0]; hi = [1.2 * L, 1.2 * L, hi[2]]
    x = lattice(n3, pitch, [0.5 * pitch, 0.5 * pitch, 0.6 * pitch], 0.04 * rad, rng)
    n = len(x)
    radius = rng.uniform(0.0015, 0.003, n) if poly else np.full(n, rad)
    typ = (np.arange(n) % ntypes + 1).astype(np.int32)
    v = np.tile([0.0, 0.0, -0.5], (n, 1)) + rng.uniform(-0.3, 0.3, (n, 3))
    mask = np.ones(n, np.int32)
    if frozen:
        mask[:frozen] |= 2
        v[:frozen] = 0.0
    T = ntypes
    props = [("youngsModulus", "peratomtype", [5e6, 7e6, 9e6][:T]),
             ("poissonsRatio", "peratomtype", [0.45, 0.3, 0.25][:T]),
             ("coefficientRestitution", "peratomtypepair", (0.3 + 0.1 * np.add.outer(np.arange(T), np.arange(T))).ravel()),
             ("coefficientFriction", "peratomtypepair", (0.5 - 0.05 * np.add.outer(np.arange(T), np.arange(T))).ravel()),
             ("coefficientRollingFriction", "peratomtypepair", (0.1 + 0.02 * np.add.outer(np.arange(T), np.arange(T))).ravel())]
    if "epsd " in model + " ":
        props.append(("coefficientRollingViscousDamping", "peratomtypepair", np.full(T * T, 0.3)))
    if "hooke" in model:
        props.append(("characteristicVelocity", "scalar", [2.0]))
    if "epsd2" in model:  # registered by the reference's epsd2 model although unused
        pass
    wmodel = model
    if "hysteretic/nonlinear" in model:  # INL normal laws on the pair style, hertz on the walls (as in the reference's example decks)
        wmodel = model.replace("model hysteretic/nonlinear1", "model hertz").replace("model hysteretic/nonlinear2", "model hertz")
        wmodel = wmodel.replace("tangential hysteretic/nonlinear", "tangential history").replace("cdtnonlinear2", "cdt")
        full = lambda v: np.full(T * T, float(v))
        props += [("LoadingStiffness", "peratomtypepair", full(5e4)), ("UnloadingStiffness", "peratomtypepair", full(2.5)),
                  ("coefficientAdhesionStiffness", "peratomtypepair", full(0.0)), ("coefficientPlasticityDepth", "peratomtypepair", full(0.1)),
                  ("pullOffForce", "peratomtypepair", full(0.0)), ("alphaCustom", "peratomtypepair", full(20.0)), ("cinCustom", "peratomtypepair", full(1e-7)),
                  ("aoneCustom", "peratomtypepair", full(10e8)), ("atwoCustom", "peratomtypepair", full(4e4)), ("athreeCustom", "peratomtypepair", full(8.0)),
                  ("kcinCustom", "peratomtypepair", full(1e-4))]
    if bond:  # bonded-sphere decks: `cohesion bond|bond/nonlinear` on the pair style, plain contact model on the walls
        kind = bond.get("kind", "bond")
        wmodel = model
        model = model.replace("tangential history", "tangential history cohesion " + kind)
        sfx = "" if kind == "bond" else "nonlinear"
        TT = T * T
        full = lambda v: np.full(TT, float(v))
        props += [("radiusMultiplierBond" + sfx, "peratomtypepair", full(bond.get("lam", 0.8))),
                  ("dampingNormalForceBond" + sfx, "peratomtypepair", full(bond.get("damp", 0.1))),
                  ("dampingTangentialForceBond" + sfx, "peratomtypepair", full(bond.get("damp", 0.1))),
                  ("dampingNormalTorqueBond" + sfx, "peratomtypepair", full(bond.get("damp", 0.1))),
                  ("dampingTangentialTorqueBond" + sfx, "peratomtypepair", full(bond.get("damp", 0.1))),
                  ("tsCreateBond" + sfx, "scalar", [bond.get("ts", 2)]),
                  ("createDistanceBond" + sfx, "peratomtypepair", full(bond.get("create", 2.2 * 0.003 if poly else 2.2 * rad)))]
        if "dissipationBond on" in settings:  # time scales of the bond's relaxation (cohesion_model_bond.h:352-362)
            props += [("dissipationNormalForceBond", "peratomtypepair", full(2e-3)), ("dissipationTangentialForceBond", "peratomtypepair", full(1e-3)),
                      ("dissipationNormalTorqueBond", "peratomtypepair", full(4e-3)), ("dissipationTangentialTorqueBond", "peratomtypepair", full(3e-3))]
        if "stressBreak on" in settings:
            props += [("maxSigmaBond" + sfx, "peratomtypepair", full(bond.get("sigma", 2e5))), ("maxTauBond" + sfx, "peratomtypepair", full(bond.get("tau", 1e5)))]
        else:
            props += [("maxDistanceBond" + sfx, "peratomtypepair", full(bond.get("maxdist", 2.5 * 0.003 if poly else 2.5 * rad)))]
        if kind == "bond":
            props += [("normalBondStiffnessPerUnitArea", "peratomtypepair", full(bond.get("kn", 2e9))),
                      ("tangentialBondStiffnessPerUnitArea", "peratomtypepair", full(bond.get("kt", 1e9)))]
        else:
            k = bond.get("k", 1e9)
            # the compression branch is k*sqrt(displacement) without the area factor (cohesion_model_bond_nonlinear.h:603-612)
            for nm, val in (("K_fn1", 20.0), ("Ku_fn1", 80.0), ("Kc_fn1", 10.0)):
                props.append(("stiffnessPerUnitArea" + nm, "peratomtypepair", full(val)))
            for nm, f in (("K_fn2", 1.0), ("Ku_fn2", 4.0), ("Kc_fn2", 0.5), ("K_ft", 0.5),
                          ("K_tn", 0.4), ("Ku_tn", 1.6), ("Kc_tn", 0.2), ("K_tt", 0.6), ("Ku_tt", 2.4), ("Kc_tt", 0.3)):
                props.append(("stiffnessPerUnitArea" + nm, "peratomtypepair", full(k * f)))
    # wall/gran keyword order: model selection, wall keywords, then on/off settings (fix_wall_gran.cpp:150-342)
    st = (" " + settings) if settings else ""
    wst = (" " + " ".join(w for w in [settings] if bond is None)) if (settings and bond is None) else ""
    hyst = "hysteretic/nonlinear" in model
    pair_model = model
    model = wmodel
    st_pair = st
    st = wst if bond else st
    if hyst: st = ""  # (the pair style's settings are the INL model's own)
    walls = [("zw", model + " primitive type %d zplane 0.0" % T + (" shear x 0.2" if shear else "") + st)]
    if cyl:
        walls.append(("cw", model + " primitive type 1 zcylinder %.17g %.17g %.17g" % (0.64 * L, 0.5 * L, 0.5 * L)
                      + (" shear z 0.3" if shear else "") + st))
    else:
        if not periodic[0]:
            walls += [("x0", model + " primitive type 1 xplane 0.0" + st), ("x1", model + " primitive type 1 xplane %.17g" % L
                       + (" shear y 0.2" if shear else "") + st)]
        if not periodic[1]:
            walls += [("y0", model + " primitive type 1 yplane 0.0" + st), ("y1", model + " primitive type 1 yplane %.17g" % L + st)]
    return dict(name=name, lo=lo, hi=hi, periodic=list(periodic), ntypes=T, skin=0.001, dt=1e-5, props=props,
                pair=pair_model + st_pair, walls=walls, gravity=(9.81, [0.0, 0.0, -1.0]), freeze=2 if frozen else 0,
                tag=np.arange(1, n + 1, dtype=np.int32), type=typ, mask=mask, x=x, v=v,
                omega=rng.uniform(-5, 5, (n, 3)) * (0 if frozen else 1) + 0.0, radius=radius, density=np.full(n, 2500.0))


def _quad(p0, p1, p2, p3):
    """two triangles of the quad p0-p1-p2-p3 (shared diagonal p0-p2)"""
    return [[p0, p1, p2], [p0, p2, p3]]


def mesh_box(L, H, nf=3, z0=0.0):
    """open box: floor of nf x nf quads (coplanar neighbours), four side walls of one quad each"""
    t = []
    xs = np.linspace(0.0, L, nf + 1)
    for a in range(nf):
        for b in range(nf):
            t += _quad([xs[a], xs[b], z0], [xs[a + 1], xs[b], z0], [xs[a + 1], xs[b + 1], z0], [xs[a], xs[b + 1], z0])
    t += _quad([0, 0, z0], [L, 0, z0], [L, 0, H], [0, 0, H]) + _quad([0, L, z0], [0, L, H], [L, L, H], [L, L, z0])
    t += _quad([0, 0, z0], [0, 0, H], [0, L, H], [0, L, z0]) + _quad([L, 0, z0], [L, L, z0], [L, L, H], [L, 0, H])
    return np.asarray(t, np.float64)


def mesh_roof(L, zr, h):
    """ridge obstacle (two inclined quads meeting at a convex edge) + its two gable triangles: edge and corner contacts"""
    a, b, m = 0.2 * L, 0.8 * L, 0.5 * L
    t = _quad([a, a, zr], [a, b, zr], [m, b, zr + h], [m, a, zr + h]) + _quad([b, a, zr], [m, a, zr + h], [m, b, zr + h], [b, b, zr])
    t += [[[a, a, zr], [m, a, zr + h], [b, a, zr]], [[a, b, zr], [b, b, zr], [m, b, zr + h]]]
    return np.asarray(t, np.float64)


def mesh_funnel(L, z_top, z_bot, r_top, r_bot, nseg=12):
    """conical funnel (non-coplanar neighbours all round)"""
    c = 0.5 * L; t = []
    for k in range(nseg):
        a0, a1 = 2 * np.pi * k / nseg, 2 * np.pi * (k + 1) / nseg
        p = lambda r, a, z: [c + r * np.cos(a), c + r * np.sin(a), z]
        t += _quad(p(r_top, a0, z_top), p(r_top, a1, z_top), p(r_bot, a1, z_bot), p(r_bot, a0, z_bot))
    return np.asarray(t, np.float64)


def mesh_drum(cx, cz, R, y0, y1, nseg=12):
    """closed polygonal drum about the y axis through (cx, *, cz): mantle quads + two end-cap fans"""
    t = []
    p = lambda a, y: [cx + R * np.cos(a), y, cz + R * np.sin(a)]
    for k in range(nseg):
        a0, a1 = 2 * np.pi * k / nseg, 2 * np.pi * (k + 1) / nseg
        t += _quad(p(a0, y0), p(a1, y0), p(a1, y1), p(a0, y1))
        t += [[[cx, y0, cz], p(a0, y0), p(a1, y0)], [[cx, y1, cz], p(a1, y1), p(a0, y1)]]
    return np.asarray(t, np.float64)


def case_mesh(kind="box", n3=(4, 4, 4), model="model hertz tangential history rolling_friction cdt", seed=SEED, poly=True,
              name="mesh", move=None, settings="", nseg=12):
    """particles falling into triangle-mesh geometry (fix mesh/surface + fix wall/gran ... mesh)"""
    c = case_box(n3=n3, model=model, seed=seed, poly=poly, name=name, settings=settings)
    L = c["hi"][0]; H = c["hi"][2]
    st = (" " + settings) if settings else ""
    c["lo"] = [-0.25 * L, -0.25 * L, -0.01]; c["hi"] = [1.25 * L, 1.25 * L, H]
    c["walls"] = []
    if kind == "box":
        c["meshes"] = [("cad", 1, mesh_box(L, 0.9 * H))]
    elif kind == "roof":
        c["meshes"] = [("cad", 1, mesh_box(L, 0.9 * H, nf=2)), ("roof", 1, mesh_roof(L, 0.002, 0.4 * L))]
        c["x"][:, 2] += 0.45 * L
    elif kind == "funnel":
        c["meshes"] = [("cad", 1, mesh_box(L, 0.9 * H, nf=2)), ("fun", 1, mesh_funnel(L, 0.55 * L, 0.2 * L, 0.62 * L, 0.12 * L))]
        c["x"][:, 2] += 0.6 * L
        c["lo"][2] = -0.01
    elif kind == "plate":   # floor mesh + a plate moving down onto the particles (fix move/mesh linear)
        top = c["x"][:, 2].max() + 0.004
        plate = np.asarray(_quad([0.02 * L, 0.02 * L, top], [0.98 * L, 0.02 * L, top], [0.98 * L, 0.98 * L, top], [0.02 * L, 0.98 * L, top]))
        c["meshes"] = [("cad", 1, mesh_box(L, 0.9 * H, nf=2)), ("plate", 1, plate)]
        c["mesh_moves"] = [("plate", "linear 0. 0. %s" % (move if move is not None else -0.4))]
    elif kind == "drum":    # particles inside a closed drum rotating about its (y) axis (fix move/mesh rotate)
        cx, cz, R = 0.5 * L, 0.62 * L, 0.62 * L
        c["x"][:, 2] += 0.22 * L
        c["meshes"] = [("drum", 1, mesh_drum(cx, cz, R, -0.12 * L, 1.12 * L, nseg=nseg))]
        c["mesh_moves"] = [("drum", "rotate origin %.17g 0. %.17g axis 0. 1. 0. period %s" % (cx, cz, move if move is not None else 0.25))]
        c["lo"] = [cx - 1.1 * R, -0.2 * L, cz - 1.1 * R]; c["hi"] = [cx + 1.1 * R, 1.2 * L, cz + 1.1 * R]
    c["mesh_walls"] = [("mw", model + " mesh n_meshes %d meshes %s" % (len(c["meshes"]), " ".join(m[0] for m in c["meshes"])) + st)]
    c["hi"][2] = max(c["hi"][2], float(max(m[2][..., 2].max() for m in c["meshes"])) + 0.01, float(c["x"][:, 2].max()) + 0.01)
    return c


def write_stl(path, nodes, name="mesh"):
    """ASCII STL with round-trip (%.17g) vertices: the reference parses them with atof (input_mesh_tri.cpp:509-511)"""
    with open(path, "w") as f:
        f.write("solid %s\n" % name)
        for t in np.asarray(nodes, np.float64).reshape(-1, 3, 3):
            f.write(" facet normal 0 0 0\n  outer loop\n")
            for v in t:
                f.write("   vertex %.17g %.17g %.17g\n" % tuple(v))
            f.write("  endloop\n endfacet\n")
        f.write("endsolid %s\n" % name)


def to_deck(c, datafile):
    """LIGGGHTS deck + data file text for the reference (grammar: SURVEY.md 8b)"""
    n = len(c["tag"])
    data = ["synthetic case %s" % c["name"], "", "%d atoms" % n, "%d atom types" % c["ntypes"], "",
            "%.17g %.17g xlo xhi" % (c["lo"][0], c["hi"][0]), "%.17g %.17g ylo yhi" % (c["lo"][1], c["hi"][1]),
            "%.17g %.17g zlo zhi" % (c["lo"][2], c["hi"][2]), "", "Atoms", ""]
    for i in range(n):
        data.append("%d %d %.17g %.17g %.17g %.17g %.17g" % (c["tag"][i], c["type"][i], 2 * c["radius"][i],
                                                           c["density"][i], *c["x"][i]))
    data += ["", "Velocities", ""]
    for i in range(n):
        data.append("%d %.17g %.17g %.17g %.17g %.17g %.17g" % (c["tag"][i], *c["v"][i], *c["omega"][i]))
    b = " ".join("p" if p else "f" for p in c["periodic"])
    deck = ["units si", "atom_style sphere", "atom_modify map array sort 0 0", "boundary " + b, "newton off",
            "communicate single vel yes", "read_data " + datafile, "neighbor %.17g bin" % c["skin"],
            "neigh_modify delay %d every %d check %s" % (c.get("neigh", (1, 0, True))[1], c.get("neigh", (1, 0, True))[0],
                                                         "yes" if c.get("neigh", (1, 0, True))[2] else "no")
            + (" contact_distance_factor %.17g" % c["cdf"] if c.get("cdf") else "")]
    for k, (name, kind, vals) in enumerate(c["props"]):
        extra = " %d" % c["ntypes"] if kind == "peratomtypepair" else ""
        deck.append("fix m%d all property/global %s %s%s %s" % (k, name, kind, extra, " ".join("%.17g" % v for v in vals)))
    deck += ["pair_style gran " + c["pair"], "pair_coeff * *"]
    if "cohesion bond " in c["pair"] + " ":  # linear bond model: the reference's bond counter (compute_bond_counter.cpp)
        deck.append("compute bc all bond/counter")
    if c["gravity"]:
        deck.append("fix grav all gravity %.17g vector %g %g %g" % (c["gravity"][0], *c["gravity"][1]))
    for wid, text in c["walls"]:
        deck.append("fix %s all wall/gran %s" % (wid, text))
    for mid, mtype, nodes in c.get("meshes", []):
        stl = os.path.join(os.path.dirname(datafile), mid + ".stl")
        write_stl(stl, nodes, mid)
        style = "mesh/surface/stress" if mid in c.get("mesh_stress", []) else "mesh/surface"
        deck.append("fix %s all %s file %s type %d" % (mid, style, stl, mtype))
    for mid, text in c.get("mesh_moves", []):
        deck.append("fix mv_%s all move/mesh mesh %s %s" % (mid, mid, text))
    for wid, text in c.get("mesh_walls", []):
        deck.append("fix %s all wall/gran %s" % (wid, text))
    if c["freeze"]:
        ids = " ".join(str(t) for t in c["tag"][(c["mask"] & c["freeze"]) != 0])
        deck += ["group frozen id " + ids, "fix frz frozen freeze"]
    deck += ["fix integr all nve/sphere", "timestep %.17g" % c["dt"]]
    return "\n".join(deck), "\n".join(data) + "\n"


def apply(c, eng):
    """replay the case on an Engine (C ABI call sequence == deck order)"""
    eng.units("si")
    eng.box(c["lo"], c["hi"], c["periodic"])
    eng.ntypes(c["ntypes"])
    every, delay, check = c.get("neigh", (1, 0, True))  # neigh_modify every / delay / check
    eng.neighbor(c["skin"], every=every, delay=delay, check=check)
    if c.get("cdf"):
        eng.contact_distance_factor(c["cdf"])
    for name, kind, vals in c["props"]:
        eng.property_global(name, kind, vals)
    eng.pair_style(c["pair"])
    if c["gravity"]:
        eng.gravity(*c["gravity"])
    for wid, text in c["walls"]:
        eng.wall_primitive(wid, text)
    for mid, mtype, nodes in c.get("meshes", []):
        eng.mesh(mid, mtype, nodes, options="stress on" if mid in c.get("mesh_stress", []) else "")
    for mid, text in c.get("mesh_moves", []):
        eng.move_mesh(mid, text)
    for wid, text in c.get("mesh_walls", []):
        eng.wall_mesh(wid, text)
    if c["freeze"]:
        eng.freeze(c["freeze"])
    eng.integrate(1)
    eng.timestep(c["dt"])
    eng.upload(c["tag"], c["type"], c["x"], c["radius"], c["density"], v=c["v"], omega=c["omega"], mask=c["mask"])
    return eng


GOLDEN_CASES = {
    "box_hertz_cdt": dict(kw=dict(n3=(4, 4, 4)), checkpoints=[0, 1, 2, 10, 400, 2500]),
    "poly_hooke_epsd_cyl": dict(kw=dict(n3=(4, 4, 5), model="model hooke tangential history rolling_friction epsd",
                                        poly=True, ntypes=2, cyl=True, shear=True, frozen=6),
                                checkpoints=[0, 1, 2, 10, 400, 2500]),
    "periodic_epsd2": dict(kw=dict(n3=(5, 5, 4), model="model hertz tangential history rolling_friction epsd2", settings="limitForce on",
                                   poly=True, periodic=(1, 1, 0), ntypes=2, shear=True), checkpoints=[0, 1, 2, 10, 400, 2500]),
    # rolling_friction cdtnonlinear2 (rolling_model_cdtnonlinear2.h): CDT with the full normal force, walls included
    "box_hertz_cdtnl2": dict(kw=dict(n3=(4, 4, 4), model="model hertz tangential history rolling_friction cdtnonlinear2", poly=True,
                                     cyl=True, shear=True), checkpoints=[0, 1, 2, 10, 400, 2500]),
    # INL normal laws hysteretic/nonlinear1 and 2 (12 history values per pair), hertz walls
    "box_hyst1": dict(kw=dict(n3=(4, 4, 4), model="model hysteretic/nonlinear1 tangential history", poly=True), checkpoints=[0, 1, 2, 10, 400, 1000]),
    "box_hyst2_cdt": dict(kw=dict(n3=(4, 4, 3), model="model hysteretic/nonlinear2 tangential history rolling_friction cdt", ntypes=2),
                          checkpoints=[0, 1, 2, 10, 400, 1000]),
    # all INL laws together: hysteretic normal + hysteretic tangential (7 history values, plastic-range restart) + cdtnonlinear2
    "box_hyst1_thyst": dict(kw=dict(n3=(4, 4, 3), model="model hysteretic/nonlinear1 tangential hysteretic/nonlinear rolling_friction cdtnonlinear2", poly=True),
                            checkpoints=[0, 1, 2, 10, 400, 1000]),
    # particles added between two runs (create_atoms / fix insert/*): six spheres appear above the settling bed
    # before step 301; the history of the existing contacts must survive, the newcomers fall in and touch
    "box_insert": dict(kw=dict(n3=(4, 4, 3), poly=True), insert=dict(at=301, n=6), checkpoints=[0, 1, 10, 300, 301, 310, 900, 2500]),
    # neigh_modify contact_distance_factor on a plain contact model (neighbor.cpp:1922-1925): pairs inside the band that do not
    # touch run surfacesClose -- tangential / rolling history zeroed (the flag stays: the normal model keeps its bit)
    "box_cdf": dict(kw=dict(n3=(4, 4, 4), model="model hertz tangential history rolling_friction epsd", poly=True), cdf=1.15,
                    checkpoints=[0, 1, 10, 400, 1500]),
    "hertz_nodamp_notroll": dict(kw=dict(n3=(3, 3, 3), model="model hertz tangential history", settings="tangential_damping off",
                                         poly=True), checkpoints=[0, 1, 300, 1500]),
    # rebuild cadence other than `delay 0 every 1 check yes` (Neighbor::decide, neighbor.cpp:1362-1376)
    "box_neigh_every2_delay4": dict(kw=dict(n3=(4, 4, 4), poly=True), neigh=(2, 4, True), checkpoints=[0, 1, 10, 400, 1500]),
    "box_neigh_nocheck": dict(kw=dict(n3=(4, 4, 4)), neigh=(25, 0, False), checkpoints=[0, 1, 10, 400, 1500]),
    # triangle-mesh walls (fix mesh/surface + fix wall/gran mesh): coplanar floor grid, convex ridge, cone, moving plate
    "mesh_box": dict(mesh="box", kw=dict(n3=(4, 4, 4)), checkpoints=[0, 1, 10, 400, 2500]),
    "mesh_roof_epsd2": dict(mesh="roof", kw=dict(n3=(4, 4, 4), model="model hertz tangential history rolling_friction epsd2"),
                            checkpoints=[0, 1, 10, 1500, 3000]),
    "mesh_funnel_hooke": dict(mesh="funnel", kw=dict(n3=(4, 4, 3), model="model hooke tangential history rolling_friction cdt"),
                              checkpoints=[0, 1, 10, 1500, 3000]),
    "mesh_plate_moving": dict(mesh="plate", kw=dict(n3=(4, 4, 3)), checkpoints=[0, 1, 10, 1500, 3000]),
    "mesh_drum_rotating": dict(mesh="drum", kw=dict(n3=(4, 4, 3)), checkpoints=[0, 1, 10, 1500, 3000]),
    # fix mesh/surface/stress: total force / torque on a static floor, a moving plate and a rotating drum (reference point travels)
    "mesh_plate_stress": dict(mesh="plate", kw=dict(n3=(4, 4, 3)), stress=["cad", "plate"], checkpoints=[0, 1, 10, 1500, 3000]),
    "mesh_drum_stress": dict(mesh="drum", kw=dict(n3=(4, 4, 3)), stress=["drum"], checkpoints=[0, 1, 600, 1500, 3000]),
    # `fix move/mesh` issued between two runs (the t01a tutorial deck starts its mesh after the settling run)
    "mesh_plate_late_move": dict(mesh="plate", kw=dict(n3=(4, 4, 3)), checkpoints=[0, 1, 200, 201, 210, 800, 2000], late_move_at=201),
    # bonded spheres (INL bond models): bonds form at step 2, stretch, some break
    # (sgn()-type bond damping makes these trajectories diverge from rounding noise within a few hundred steps: short horizons)
    "bond_linear": dict(kw=dict(n3=(4, 4, 4), model="model hertz tangential history", poly=True, bond=dict(kind="bond", maxdist=2.1 * 0.003)),
                        checkpoints=[0, 1, 2, 3, 10, 100]),
    "bond_linear_stress": dict(kw=dict(n3=(4, 4, 3), model="model hertz tangential history rolling_friction cdt", settings="stressBreak on",
                                       bond=dict(kind="bond", sigma=4e4, tau=2e4)), checkpoints=[0, 1, 2, 3, 10, 100]),
    "bond_linear_dissipation": dict(kw=dict(n3=(4, 4, 3), model="model hertz tangential history", settings="dissipationBond on", poly=True,
                                            bond=dict(kind="bond", maxdist=2.1 * 0.003)), checkpoints=[0, 1, 2, 3, 10, 100]),
    "bond_nonlinear": dict(kw=dict(n3=(4, 4, 4), model="model hertz tangential history", poly=True, bond=dict(kind="bond/nonlinear")),
                           checkpoints=[0, 1, 2, 3, 10, 100, 400]),
}


def make_case(name):
    g = GOLDEN_CASES[name]
    if "mesh" in g:
        c = case_mesh(kind=g["mesh"], name=name, **g["kw"])
        if "stress" in g:
            c["mesh_stress"] = list(g["stress"])
        if "late_move_at" in g:  # the movers are not part of the initial deck: late_commands() issues them before that checkpoint
            c["late_moves"] = (g["late_move_at"], c.pop("mesh_moves"))
        return c
    c = case_box(name=name, **g["kw"])
    if "neigh" in g:
        c["neigh"] = g["neigh"]
    if "cdf" in g:
        c["cdf"] = g["cdf"]
    if "insert" in g:  # newcomers on a small lattice above the initial bed, numbered behind the existing tags
        k = g["insert"]["n"]
        rng = np.random.default_rng(777)
        L = c["hi"][0] - c["lo"][0]
        xs = np.array([[c["lo"][0] + L * (0.25 + 0.25 * (q % 3)), c["lo"][1] + L * (0.3 + 0.4 * ((q // 3) % 2)), 0.6 * c["hi"][2] + 0.008 * (q // 6)] for q in range(k)])
        xs += rng.uniform(-2e-4, 2e-4, xs.shape)
        n0 = len(c["tag"])
        c["inserts"] = {g["insert"]["at"]: dict(tag=np.arange(n0 + 1, n0 + k + 1, dtype=np.int32), type=np.ones(k, np.int32), x=xs,
                                                radius=rng.uniform(0.0015, 0.003, k), density=np.full(k, 2500.0))}
    return c


def late_commands(c, cp):
    """deck lines to issue before running on to checkpoint cp (empty for most cases)"""
    at, moves = c.get("late_moves", (None, []))
    lines = ["fix mv_%s all move/mesh mesh %s %s" % (mid, mid, text) for mid, text in moves] if cp == at else []
    ins = c.get("inserts", {}).get(cp)
    if ins:  # particles added between two runs: create_atoms single + set (the reference numbers them max tag + 1, ...)
        for k in range(len(ins["tag"])):
            lines.append("create_atoms %d single %.17g %.17g %.17g units box" % (ins["type"][k], *ins["x"][k]))
            lines.append("set atom %d diameter %.17g density %.17g" % (ins["tag"][k], 2.0 * ins["radius"][k], ins["density"][k]))
    return lines


def apply_late(c, eng, cp):
    """the same through the Engine API"""
    at, moves = c.get("late_moves", (None, []))
    if cp == at:
        for mid, text in moves:
            eng.move_mesh(mid, text)
    ins = c.get("inserts", {}).get(cp)
    if ins:
        eng.insert(ins["tag"], ins["type"], ins["x"], ins["radius"], ins["density"])


def snapshot(eng, c):
    """state + bookkeeping of an Engine-like object as a flat dict of arrays"""
    out = dict(eng.atoms(("x", "v", "f", "omega", "torque")))
    p = eng.pairs()
    out.update(pair_lo=p["lo"], pair_hi=p["hi"], pair_flag=(p["flag"] != 0).astype(np.int32), pair_hist=p["hist"])
    if "cohesion bond " in c["pair"] + " " and hasattr(eng, "bond_counter"):  # compute bond/counter as the reference returns it between runs
        out["bondcounter"] = eng.bond_counter()
    if "cohesion" in c["pair"]:  # contactPos (history 2..4) is written at bond creation and read for wall bonds only: not compared
        out["pair_hist"] = out["pair_hist"].copy(); out["pair_hist"][:, 2:5] = 0.0
    for wid, text in c["walls"]:
        dn = 3 + (3 if ("epsd" in text) else 0)
        out["wall_" + wid] = eng.wall_history(wid, dn)
    for mid, mtype, nodes in c.get("meshes", []):
        m = eng.mesh_contacts(mid)
        out["mesh_%s_tag" % mid] = m["tag"]; out["mesh_%s_tri" % mid] = m["tri"]; out["mesh_%s_hist" % mid] = m["hist"]
        if mid in c.get("mesh_stress", []):
            out["meshforce_%s" % mid] = eng.mesh_force(mid)
    return out


# fix insert/pack decks (tests/golden/in.<name>, goldens by tests/golden/make_golden_insert.py): name -> checkpoints
#  a: cylinder region, one template, volumefraction_region, insert_every once (the shape of tutorial t01a's insertion)
#  b: block region, two templates (mass based), all_in yes, particles_in_region topped up every 400 steps against the spheres
#     already there, vel uniform, omega constant
#  c: cylinder along x, number based, all_in no, mass_in_region every 300 steps, vel gaussian, maxattempt
#  lattice_a / lattice_b: lattice + create_atoms box|region, region INF/EDGE, group region|union|subtract, velocity set (no insertion fix)
INSERT_DECKS = {"insert_pack_a": [1, 2, 200, 1000], "insert_pack_b": [1, 400, 401, 801, 2500], "insert_pack_c": [1, 301, 601, 2000],
                "lattice_a": [0, 1, 300, 1500], "lattice_b": [0, 1, 300, 1500],
                "insert_pack_d": [1, 251, 501, 1200]}  # d: three templates, random_distribute uncorrelated, overlapcheck no (CPU only)
INSERT_DECKS_GPU = ["insert_pack_a", "insert_pack_b", "insert_pack_c", "lattice_a", "lattice_b"]  # (the set that was run on the B200)

# the reference's own INL example decks that the deck front end runs unchanged (read from the reference tree where they lie,
# build container only; only the length of their `run` is cut): path under examples/LIGGGHTS -> steps.  Goldens:
# tests/golden/inl_examples.npz (make_golden_insert.py examples)
INL_EXAMPLES = "/root/reference/examples/LIGGGHTS"
INL_EXAMPLE_DECKS = {
    "Tutorials_public/contactModels/in.newModels": 1500,  # (1,800 spheres from fix insert/pack into a cylinder, hertz/history, plane + cylinder walls; 2 runs)
    "Tutorials_public/cohesion/in.noCohesion": 3000,  # (fix insert/pack every 3000 steps: 3 x 250 spheres, group region, unfix of the insertion; 3 runs)
    "Tutorials_premium/mesh_force_eval/in.testmeshforce": 1500,  # (500 inserted spheres on an STL plate with mesh/surface/stress; 3 runs)
    "INL/cohesive_bond/chain_bending_test/in.chain_bending.lmp": 100000,  # (its full length: linear bond, fix addforce, fix viscous, fix freeze)
    "INL/cohesive_bond_nonlinear_compression/chain_bending_mm_1/in.chain_bending.lmp": 200000,
    "INL/cohesive_bond_nonlinear_compression/chain_bending_mm_2/in.chain_bending.lmp": 200000,
    "INL/cohesive_bond_nonlinear_compression/chain_bending_um_1/in.chain_bending.lmp": 200000,
    "INL/cohesive_bond_nonlinear_compression/chain_bending_um_2/in.chain_bending.lmp": 200000,
}


# decks that only run under the `-suffix b200` shim (they use script features the deck front end does not have: `print` of
# compute-based variables, `compute displace/atom` ...): 5,850 spheres from fix insert/pack (mass_in_region) into a cylinder
SHIM_ONLY_DECKS = {"Tutorials_premium/dump_custom_vtk/in.dump_custom_vtk": 500}
ALL_EXAMPLE_DECKS = dict(INL_EXAMPLE_DECKS, **SHIM_ONLY_DECKS)


def example_deck_text(rel, nsteps):
    """an example deck of the reference with its run length replaced and its dump lines dropped (no files into the read-only tree)"""
    import re
    text = re.sub(r"&[ \t]*\n", " ", open(os.path.join(INL_EXAMPLES, rel)).read())
    total = [0]

    def shorten(m):  # every `run N` becomes `run nsteps`, every `run N upto` ends nsteps later than the run before it
        total[0] += nsteps
        return m.group(1) + (str(total[0]) + m.group(3) if m.group(3) else str(nsteps))
    text = re.sub(r"^(run\s+)(\d+)(\s+upto)?", shorten, text, flags=re.M)
    text = re.sub(r"^(run\s+)\$\{\w+\}", lambda m: m.group(1) + str(nsteps), text, flags=re.M)  # (`run ${nDump}`)
    return "\n".join(l for l in text.splitlines() if not l.strip().startswith(("dump", "fix\t\tprint", "fix print")))
