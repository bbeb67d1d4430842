#!/usr/bin/env python
"""Long-run observables (BASELINE.json north_star, third correctness leg; SURVEY.md 8d iii): DEM trajectories are chaotic,
so beyond the first steps the engine is compared with the reference through what a user of either code would measure --

  packing   packing fraction of a bed of 2,000 polydisperse spheres settled in a box of primitive planes (C1 / C2 physics)
  discharge discharge rate [particles/s] of the conical hopper of configs.C3 at reduced size (~19 k spheres), outlet open
  repose    static angle of repose [deg] of the heap left when a tube (STL mesh, `fix move/mesh linear` arriving between two
            runs) is lifted off a column of 4,000 spheres -- the flow of the reference's tutorial deck
            examples/LIGGGHTS/INL_tutorials/t01a_static_angle_of_repose_monosphere/in.staticAOR_MonoSphere

each over the seeds 1, 2, 3 (jitter / initial velocities of the generators).

  python tools/observables.py --reference            (build container: the UNMODIFIED reference, oracle/_ref; ~10 minutes)
        -> tests/golden/observables_ref.json (committed fixture)
  python tools/observables.py --engine [--json out]  (GPU box: the CUDA engine through the C ABI; compares with the fixture)
The GPU test tests/test_gpu_observables.py runs the engine leg and asserts the tolerances written below."""
import json
import os
import sys
import tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "liggghts-inl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import cases  # noqa: E402
import configs  # noqa: E402

FIXTURE = os.path.join(ROOT, "tests", "golden", "observables_ref.json")
SEEDS = (1, 2, 3)
# statistical tolerances on the mean over the seeds (engine vs reference).  The spread between seeds (the fixture's `std`)
# is the natural scale: packing ~2e-3, discharge ~2 %, repose ~1 deg.
TOL = {"packing": 0.01, "discharge": 0.08, "repose": 2.5}   # absolute packing fraction, relative rate, degrees


# ------------------------------------------------------------------------------------------------ cases
def case_packing(seed):
    c = cases.case_box(n3=(10, 10, 20), poly=True, seed=100 + seed, name="obs_packing")
    return c, [("run", 40000)]


def case_discharge(seed):
    c = configs.C3(configs.MINI["C3"], seed=200 + seed)
    return c, [("run", 15000), ("mark", "t0"), ("run", 45000), ("mark", "t1")]   # rate over 0.45 s (~700 spheres: counting noise ~4 % per seed)


def mesh_tube(cx, cy, R, z0, z1, nseg=24):
    t = []
    p = lambda a, z: [cx + R * np.cos(a), cy + R * np.sin(a), z]
    for k in range(nseg):
        a0, a1 = 2 * np.pi * k / nseg, 2 * np.pi * (k + 1) / nseg
        t += cases._quad(p(a0, z0), p(a1, z0), p(a1, z1), p(a0, z1))
    return np.asarray(t, np.float64)


def case_repose(seed):
    """4,000 spheres (r = 2.5 mm) dropped into a tube of radius 30 mm standing on a floor plane; after settling the tube is
    lifted at 0.25 m/s and the column collapses into a heap"""
    rng = np.random.default_rng(300 + seed)
    rad, Rt = 0.0025, 0.030
    L = 0.40
    cx = cy = 0.5 * L
    pitch = 2.1 * rad
    g = configs._cubic([cx - Rt, cy - Rt, 0.0], [cx + Rt, cy + Rt, 1.2], pitch)
    g = g[np.hypot(g[:, 0] - cx, g[:, 1] - cy) <= Rt - 1.15 * rad]
    g = g[np.argsort(g[:, 2], kind="stable")][:4000]   # the lowest 4,000 sites: a column of ~48 layers
    g = g + rng.uniform(-0.04 * rad, 0.04 * rad, g.shape)
    c = configs._base("obs_repose", [0.0, 0.0, 0.0], [L, L, g[:, 2].max() + 0.45])   # (tall enough for the lifted tube: a mesh must stay inside the box)
    configs._finish(c, g, rad, v=np.tile([0.0, 0.0, -0.3], (len(g), 1)) + rng.uniform(-0.05, 0.05, (len(g), 3)))
    c["props"][3] = ("coefficientFriction", "peratomtypepair", [0.6])
    c["props"][4] = ("coefficientRollingFriction", "peratomtypepair", [0.3])
    c["walls"] = [("floor", configs.HERTZ_CDT + " primitive type 1 zplane 0.0")]
    c["meshes"] = [("tube", 1, mesh_tube(cx, cy, Rt, 0.0005, g[:, 2].max() + 0.02))]
    c["mesh_walls"] = [("mw", configs.HERTZ_CDT + " mesh n_meshes 1 meshes tube")]
    return c, [("run", 45000), ("move", ("tube", "linear 0. 0. 0.25")), ("run", 150000)]


CASES = {"packing": case_packing, "discharge": case_discharge, "repose": case_repose}


# ------------------------------------------------------------------------------------------------ observables
def packing_fraction(c, x):
    """solid fraction of the lower 60 % of the bed, one diameter clear of the side walls and the floor"""
    r = c["radius"]; L = c["hi"][0]
    ztop = np.percentile(x[:, 2], 98)
    m = 3 * r.max()
    z0, z1 = m, m + 0.6 * (ztop - m)
    inside = (x[:, 0] > m) & (x[:, 0] < L - m) & (x[:, 1] > m) & (x[:, 1] < L - m) & (x[:, 2] > z0) & (x[:, 2] < z1)
    return float((4.0 / 3.0 * np.pi * r[inside] ** 3).sum() / ((L - 2 * m) ** 2 * (z1 - z0)))


def discharged(c, x):
    return int((x[:, 2] < -2 * c["radius"][0]).sum())


def repose_angle(c, x):
    """slope of the heap's surface: highest sphere top per annulus around the heap's axis, least-squares line over the flank"""
    r = c["radius"][0]
    cx, cy = np.median(x[:, 0]), np.median(x[:, 1])
    rho = np.hypot(x[:, 0] - cx, x[:, 1] - cy)
    edges = np.arange(0.0, np.percentile(rho, 99), 2 * r)
    rr, hh = [], []
    for a, b in zip(edges[:-1], edges[1:]):
        m = (rho >= a) & (rho < b)
        if m.sum() >= 3:
            rr.append(0.5 * (a + b)); hh.append(np.percentile(x[m, 2], 95) + r)
    rr, hh = np.asarray(rr), np.asarray(hh)
    k = (rr > 0.2 * rr.max()) & (rr < 0.85 * rr.max())
    slope = np.polyfit(rr[k], hh[k], 1)[0]
    return float(np.degrees(np.arctan(-slope)))


def evaluate(name, c, marks, x_final, dt):
    if name == "packing":
        return packing_fraction(c, x_final)
    if name == "discharge":
        (s0, x0), (s1, x1) = marks["t0"], marks["t1"]
        return (discharged(c, x1) - discharged(c, x0)) / ((s1 - s0) * dt)
    return repose_angle(c, x_final)


# ------------------------------------------------------------------------------------------------ runners
def run_engine(c, prog, engine):
    eng = cases.apply(c, engine)
    marks, step = {}, 0
    eng.setup()
    for op, arg in prog:
        if op == "run":
            eng.run(arg); step += arg
        elif op == "move":
            eng.move_mesh(*arg); eng.setup()
        elif op == "mark":
            marks[arg] = (step, eng.download("x"))
    x = eng.download("x")
    eng.close()
    return marks, x


def run_reference(c, prog):
    import ref_driver
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "case.data"))
    open(os.path.join(tmp, "case.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    marks, step = {}, 0
    for op, arg in prog:
        if op == "run":
            r.cmd("run %d" % arg); step += arg
        elif op == "move":
            r.cmd("fix mv_%s all move/mesh mesh %s %s" % (arg[0], arg[0], arg[1]))
        elif op == "mark":
            marks[arg] = (step, r.atoms()["x"])
    x = r.atoms()["x"]
    r.close()
    return marks, x


def one(name, seed, impl):
    c, prog = CASES[name](seed)
    if impl == "reference":
        marks, x = run_reference(c, prog)
    else:
        import dem_b200
        marks, x = run_engine(c, prog, dem_b200.Engine(device=0))
    assert np.isfinite(x).all() and len(x) == len(c["tag"])
    return evaluate(name, c, marks, x, c["dt"])


def summarize(vals):
    return {k: {"values": v, "mean": float(np.mean(v)), "std": float(np.std(v))} for k, v in vals.items()}


def compare(got, ref):
    """per observable: |mean difference| in the unit of TOL, and whether it is inside the tolerance"""
    out = {}
    for k in ref:
        d = got[k]["mean"] - ref[k]["mean"]
        err = abs(d) / abs(ref[k]["mean"]) if k == "discharge" else abs(d)
        out[k] = {"engine": got[k]["mean"], "reference": ref[k]["mean"], "ref_std": ref[k]["std"], "err": err, "tol": TOL[k], "ok": bool(err <= TOL[k])}
    return out


def main():
    import subprocess
    args = sys.argv[1:]
    if args and args[0] == "--one":   # one reference run per process: the reference keeps global registries
        print(json.dumps({"value": one(args[1], int(args[2]), args[3])})); return
    impl = "reference" if "--reference" in args else "engine"
    names = [a for a in args if a in CASES] or list(CASES)
    vals = {}
    for name in names:
        vals[name] = []
        for seed in SEEDS:
            if impl == "reference":
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name, str(seed), impl], capture_output=True, text=True, check=True).stdout
                v = json.loads(out.strip().splitlines()[-1])["value"]
            else:
                v = one(name, seed, impl)
            vals[name].append(v)
            print(impl, name, "seed", seed, "->", v, flush=True)
        if impl == "reference":   # the fixture grows observable by observable (a reference run takes minutes)
            old = json.load(open(FIXTURE)) if os.path.exists(FIXTURE) else {}
            old.update(summarize({name: vals[name]}))
            json.dump(old, open(FIXTURE, "w"), indent=1)
            print("wrote", name, "to", FIXTURE, flush=True)
    res = summarize(vals)
    if impl == "reference":
        pass
    else:
        ref = json.load(open(FIXTURE))
        cmp_ = compare(res, {k: ref[k] for k in res})
        print(json.dumps(cmp_, indent=1))
        if "--json" in args:
            json.dump({"engine": res, "comparison": cmp_}, open(args[args.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
