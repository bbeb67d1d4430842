"""GPU: the CUDA engine through the C ABI against (a) golden vectors from the unmodified
reference and (b) the CPU oracle on the same seeded inputs."""
import os
import numpy as np
import pytest
import cases
import parity

pytestmark = pytest.mark.gpu


def gpu_engine(owner_list=0, **kw):
    """owner_list: the measured alternative of dem_pairs.cuh (every pair evaluated once by its owner, k_pairs + k_finish)
    instead of the default full list (k_step)"""
    import dem_b200
    e = dem_b200.Engine(device=0, **kw)
    if owner_list:
        e.option("owner_list", 1)
    return e


def tol_at(cp):
    # trajectories are chaotic: rounding-level differences (FMA contraction, summation order)
    # grow with the step count; the first steps carry the 1e-10 bar of BASELINE.json
    return 1e-10 if cp <= 10 else (1e-6 if cp <= 400 else 1e-4)


@pytest.mark.parametrize("owner_list", [0, 1])
@pytest.mark.parametrize("name", list(cases.GOLDEN_CASES))
def test_engine_matches_reference_golden(name, owner_list):
    c = cases.make_case(name)
    g = parity.golden(name)
    e = cases.apply(c, gpu_engine(owner_list=owner_list))
    done = 0
    for cp in cases.GOLDEN_CASES[name]["checkpoints"]:
        cases.apply_late(c, e, cp)
        e.setup()
        if done == 0:
            parity.compare_topology(e, c, g)
        e.run(cp - done); done = cp
        ref = parity.golden_at(g, cp)
        parity.compare_snapshot(cases.snapshot(e, c), ref, g["rmass"], tol=parity.tol_for(c, cp, gpu=True), label="%s@%d" % (name, cp))
        assert e.stats().nbuilds == int(ref["nbuilds"]), "rebuild cadence differs at %d" % cp
    assert e.stats().kernel_launches > 0
    e.close()


@pytest.mark.parametrize("kw", [
    dict(n3=(12, 12, 14), poly=True, ntypes=2, model="model hertz tangential history rolling_friction cdt"),
    dict(n3=(10, 10, 10), poly=True, periodic=(1, 1, 0), model="model hertz tangential history rolling_friction epsd2"),
    dict(n3=(8, 8, 8), model="model hooke tangential history rolling_friction epsd", frozen=20, cyl=True),
])
@pytest.mark.parametrize("owner_list", [0, 1])
def test_engine_matches_oracle_bed(kw, owner_list):
    """~2k particle beds, 600 steps incl. rebuilds: pair set / flags bit-exact, forces to tolerance"""
    c = cases.case_box(name="bed", seed=7, **kw)
    rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    got = cases.apply(c, gpu_engine(owner_list=owner_list))
    ref = cases.apply(c, parity.oracle_engine())
    done = 0
    for cp in (0, 1, 2, 10, 200, 600):
        for eng in (got, ref):
            eng.setup(); eng.run(cp - done)
        done = cp
        parity.compare_snapshot(cases.snapshot(got, c), cases.snapshot(ref, c), rmass, tol=tol_at(cp), label="bed@%d" % cp)
        assert got.stats().nbuilds == ref.stats().nbuilds
    got.close(); ref.close()


@pytest.mark.parametrize("kind,kw", [
    ("box", dict(n3=(10, 10, 8))),
    ("roof", dict(n3=(9, 9, 8), model="model hertz tangential history rolling_friction epsd")),
    ("funnel", dict(n3=(9, 9, 7))),
    ("plate", dict(n3=(8, 8, 6), model="model hooke tangential history rolling_friction epsd2")),
    ("drum", dict(n3=(8, 8, 6), model="model hertz tangential history rolling_friction epsd", move=0.2)),
    ("drum", dict(n3=(8, 8, 6), move=0.3, nseg=40)),  # end caps = flat fans of 40 triangles: 39 coplanar node-neighbours per triangle
])
def test_mesh_walls_match_oracle(kind, kw):
    """~700 particles in triangle-mesh geometry, 3000 steps incl. rebuilds (a moving plate, a rotating drum): mesh contact
    rows (particle, triangle) bit-exact, topology flags identical, moved mesh geometry bit-exact, forces to tolerance"""
    c = cases.case_mesh(kind=kind, name="mesh_" + kind, seed=11, **kw)
    rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    got = cases.apply(c, gpu_engine())
    ref = cases.apply(c, parity.oracle_engine())
    done = 0
    for cp in (0, 1, 10, 500, 1500, 3000):
        for eng in (got, ref):
            eng.setup(); eng.run(cp - done)
        if done == 0:
            for mid, mt, nodes in c["meshes"]:
                for f in ("edge_active", "corner_active", "obtuse", "nneighs"):
                    assert np.array_equal(got.mesh_field(mid, f, len(nodes)), ref.mesh_field(mid, f, len(nodes))), (mid, f)
        done = cp
        for mid, text in c.get("mesh_moves", []):  # moved / rotated geometry: same arithmetic as the reference -> bit-exact
            nt = len([m for m in c["meshes"] if m[0] == mid][0][2])
            for f in ("nodes", "center", "edge_vec", "edge_norm", "surf_norm"):
                assert np.array_equal(got.mesh_field(mid, f, nt), ref.mesh_field(mid, f, nt)), (mid, f, cp)
        parity.compare_snapshot(cases.snapshot(got, c), cases.snapshot(ref, c), rmass, tol=tol_at(cp) if cp <= 500 else 1e-3, label="%s@%d" % (kind, cp))
        assert got.stats().nbuilds == ref.stats().nbuilds
        if cp == 3000:
            sg = cases.snapshot(got, c)
            assert sum(len(sg["mesh_%s_tag" % m[0]]) for m in c["meshes"]) > 0, "no mesh contact was exercised"
    got.close(); ref.close()


@pytest.mark.parametrize("kw", [
    dict(n3=(9, 9, 8), poly=True, model="model hertz tangential history", bond=dict(kind="bond", maxdist=2.1 * 0.003)),
    dict(n3=(8, 8, 8), model="model hertz tangential history rolling_friction epsd2", settings="stressBreak on", periodic=(1, 1, 0),
         bond=dict(kind="bond", sigma=4e4, tau=2e4)),
    dict(n3=(8, 8, 7), poly=True, model="model hooke tangential history rolling_friction cdt", bond=dict(kind="bond/nonlinear")),
])
def test_bonded_spheres_match_oracle(kw):
    """~600 bonded spheres: bonds form at step 2, some break; pair set, flags, bond bookkeeping bit-exact, forces to tolerance"""
    c = cases.case_box(name="bonded", seed=5, **kw)
    rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    got = cases.apply(c, gpu_engine())
    ref = cases.apply(c, parity.oracle_engine())
    done = 0
    for cp in (0, 1, 2, 3, 10, 30, 60):
        for eng in (got, ref):
            eng.setup(); eng.run(cp - done)
        done = cp
        sg, sr = cases.snapshot(got, c), cases.snapshot(ref, c)
        parity.compare_snapshot(sg, sr, rmass, tol=parity.tol_for(c, cp, gpu=True), label="bonded@%d" % cp)
        assert np.array_equal(sg["pair_hist"][:, 0] > 0, sr["pair_hist"][:, 0] > 0), "bond flags differ at %d" % cp
        assert got.stats().nbuilds == ref.stats().nbuilds
    assert (sg["pair_hist"][:, 0] > 0).sum() > 50, "no bonds were exercised"
    got.close(); ref.close()


def test_insertion_that_outgrows_the_lists_matches_oracle():
    """dem_insert_particles with more newcomers than the particle arrays and the neighbour rows have room for (48 -> 1,548
    particles): capacity growth and the re-striding of the old list must keep the history of the contacts that already exist"""
    c = cases.case_box(n3=(4, 4, 3), name="grow", seed=21, poly=True)
    c["hi"][2] = 0.75  # room for the column of newcomers
    got = cases.apply(c, gpu_engine())
    ref = cases.apply(c, parity.oracle_engine())
    for eng in (got, ref):
        eng.setup(); eng.run(400)
    before = cases.snapshot(got, c)
    assert int(before["pair_flag"].sum()) > 0, "no contact to preserve"
    k = 1500
    L = c["hi"][0] - c["lo"][0]
    nx = 4
    q = np.arange(k)
    xs = np.stack([c["lo"][0] + L * (0.14 + 0.24 * (q % nx)), c["lo"][1] + L * (0.14 + 0.24 * ((q // nx) % nx)), 0.05 + 0.0065 * (q // (nx * nx))], 1)
    xs += np.random.default_rng(5).uniform(-1e-4, 1e-4, xs.shape)
    assert xs[:, 2].max() + 0.003 < c["hi"][2]
    n0 = len(c["tag"])
    new = dict(tag=np.arange(n0 + 1, n0 + k + 1, dtype=np.int32), type=np.ones(k, np.int32), x=xs, radius=np.full(k, 0.0025), density=np.full(k, 2500.0))
    rmass = np.concatenate([4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"], 4.0 * np.pi / 3.0 * new["radius"] ** 3 * new["density"]])
    done = 0
    for eng in (got, ref):
        eng.insert(new["tag"], new["type"], new["x"], new["radius"], new["density"])
    for cp in (0, 1, 10, 300):
        for eng in (got, ref):
            eng.setup(); eng.run(cp - done)
        done = cp
        sg = cases.snapshot(got, c)
        assert len(sg["x"]) == n0 + k
        parity.compare_snapshot(sg, cases.snapshot(ref, c), rmass, tol=tol_at(cp), label="grow@%d" % cp)
        assert got.stats().nbuilds == ref.stats().nbuilds
    got.close(); ref.close()


def test_many_contacts_on_one_particle_match_oracle():
    """a sphere three times larger than the 40 small ones sitting on its surface: all 40 contacts form in the first step -- far
    more than the 12 contacts the step kernel stages in shared memory and than the default 16 history slots.  Exercises the
    on-the-spot evaluation of surplus contacts and the sizing of the history rows from the entries inside the contact band."""
    c = cases.case_box(n3=(4, 4, 3), name="shell", seed=13)
    rs, R, nshell = 0.0025, 0.0075, 40
    L = c["hi"][0]
    ctr = np.array([0.5 * L, 0.5 * L, 0.03])
    k = np.arange(nshell) + 0.5
    phi = np.arccos(1.0 - 2.0 * k / nshell); th = np.pi * (1.0 + 5.0 ** 0.5) * k   # Fibonacci sphere
    shell = ctr + (R + rs) * 0.999 * np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
    n = nshell + 1
    c.update(tag=np.arange(1, n + 1, dtype=np.int32), type=np.ones(n, np.int32), mask=np.ones(n, np.int32),
             x=np.vstack([ctr, shell]), v=np.zeros((n, 3)), omega=np.zeros((n, 3)),
             radius=np.concatenate([[R], np.full(nshell, rs)]), density=np.full(n, c["density"][0]))
    c["hi"][2] = max(c["hi"][2], 0.06)
    rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    got = cases.apply(c, gpu_engine())
    ref = cases.apply(c, parity.oracle_engine())
    done = 0
    for cp in (0, 1, 2, 10, 300, 1500):
        for eng in (got, ref):
            eng.setup(); eng.run(cp - done)
        done = cp
        sg = cases.snapshot(got, c)
        parity.compare_snapshot(sg, cases.snapshot(ref, c), rmass, tol=tol_at(cp), label="shell@%d" % cp)
        assert got.stats().nbuilds == ref.stats().nbuilds
        if cp == 1:
            on_big = int(((sg["pair_lo"] == 1) & (sg["pair_flag"] != 0)).sum())
            assert on_big == nshell, "the big sphere holds %d contacts" % on_big
    got.close(); ref.close()


@pytest.mark.parametrize("kw", [
    dict(n3=(10, 10, 10), poly=True, model="model hertz tangential history rolling_friction cdt"),
    dict(n3=(8, 8, 8), poly=True, periodic=(1, 1, 0), ntypes=2, model="model hertz tangential history rolling_friction epsd2"),
    dict(n3=(8, 8, 8), model="model hooke tangential history rolling_friction epsd", frozen=20),
])
def test_fp32_mode_within_1e_5_of_the_fp64_oracle(kw):
    """option fp32 (BASELINE.json north_star: "1e-5 in fp32 mode"): the contact law in single precision, state / geometry /
    sums in fp64.  Pair sets and contact flags stay bit-exact (the predicates are fp64), per-particle forces and torques of
    the first steps agree with the fp64 oracle to 1e-5 (relative, floor 1e-7 of a weight x radius for torques)"""
    c = cases.case_box(name="fp32", seed=9, **kw)
    rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    got = cases.apply(c, gpu_engine())
    got.option("fp32", 1)
    ref = cases.apply(c, parity.oracle_engine())
    done = 0
    worst = 0.0
    for cp in (0, 1, 2, 10, 60):
        for eng in (got, ref):
            eng.setup(); eng.run(cp - done)
        done = cp
        sg, sr = cases.snapshot(got, c), cases.snapshot(ref, c)
        parity.compare_bookkeeping(sg, sr, label="fp32@%d" % cp)
        mg = rmass * 9.81
        ef = parity.rel_err(sg["f"], sr["f"], 1e-7 * mg)
        et = parity.rel_err(sg["torque"], sr["torque"], np.maximum(1e-7 * mg * c["radius"], 1e-5 * np.linalg.norm(sr["f"], axis=1) * c["radius"]))
        worst = max(worst, ef, et)
        tol = 1e-5 if cp <= 10 else 1e-3   # (after 60 steps the single-precision history has drifted: chaotic growth)
        assert ef <= tol and et <= tol, "fp32@%d: force %.2e torque %.2e" % (cp, ef, et)
    print("fp32 mode: worst relative force/torque error over the first 60 steps %.2e" % worst)
    got.close(); ref.close()


def test_settings_changed_between_runs_match_oracle():
    """`neighbor` and `fix property/global` between two runs (the deck front end forwards them and calls setup again): the
    material tables, the neighbour cutoff and the cell grid are derived again -- a doubled skin must not lose pairs beyond
    the old cells, a changed friction coefficient must reach the contact law"""
    c = cases.case_box(n3=(9, 9, 9), poly=True, name="resetup", seed=21)
    rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
    got = cases.apply(c, gpu_engine())
    ref = cases.apply(c, parity.oracle_engine())
    for eng in (got, ref):
        eng.setup(); eng.run(100)
    parity.compare_snapshot(cases.snapshot(got, c), cases.snapshot(ref, c), rmass, tol=1e-6, label="resetup@100")
    for eng in (got, ref):
        eng.neighbor(3.0 * c["skin"], every=1, delay=0, check=True)
        eng.property_global("coefficientFriction", "peratomtypepair", [0.15])
        eng.setup()
    s0g, s0r = cases.snapshot(got, c), cases.snapshot(ref, c)
    assert len(s0g["pair_lo"]) == len(s0r["pair_lo"]) and np.array_equal(s0g["pair_lo"], s0r["pair_lo"]) and np.array_equal(s0g["pair_hi"], s0r["pair_hi"])
    for eng in (got, ref):
        eng.run(300)
    parity.compare_snapshot(cases.snapshot(got, c), cases.snapshot(ref, c), rmass, tol=1e-6, label="resetup@400")
    assert got.stats().nbuilds == ref.stats().nbuilds
    got.close(); ref.close()


def test_history_overflow_is_reported_under_check_no():
    """forced rebuilds (`neigh_modify every 2 check no`) must not swallow the history-slot overflow of the step before the
    rebuild: 40 small spheres just off the surface of a big one fly inwards and all touch it in step 3 -- more than the 16
    history rows sized at the (contact-free) build before step 2 -- and dem_run has to say so although step 4 starts with a
    forced rebuild that clears the step flags"""
    import dem_b200
    c = cases.case_box(n3=(4, 4, 3), name="ovf", seed=13)
    rs, R, nshell = 0.001, 0.004, 40   # (shell radius 5 mm: clear of the side walls, all 40 gaps close in the same step)
    L = c["hi"][0]
    ctr = np.array([0.5 * L, 0.5 * L, 0.03])
    k = np.arange(nshell) + 0.5
    phi = np.arccos(1.0 - 2.0 * k / nshell); th = np.pi * (1.0 + 5.0 ** 0.5) * k
    u = np.stack([np.cos(th) * np.sin(phi), np.sin(th) * np.sin(phi), np.cos(phi)], 1)
    n = nshell + 1
    c.update(tag=np.arange(1, n + 1, dtype=np.int32), type=np.ones(n, np.int32), mask=np.ones(n, np.int32),
             x=np.vstack([ctr, ctr + (R + rs) * 1.00225 * u]), v=np.vstack([np.zeros(3), -0.5 * u]), omega=np.zeros((n, 3)),
             radius=np.concatenate([[R], np.full(nshell, rs)]), density=np.full(n, c["density"][0]))
    c["hi"][2] = max(c["hi"][2], 0.06)
    c["neigh"] = (2, 0, False)
    e = cases.apply(c, gpu_engine())
    e.option("histslots", 16)   # (the default -- one history row per list entry -- cannot overflow)
    e.setup()
    with pytest.raises(dem_b200.DemError, match="history"):
        e.run(20)
    e.close()


def test_engine_is_deterministic():
    c = cases.case_box(n3=(8, 8, 8), poly=True, name="det", seed=3)
    snaps = []
    for rep in range(2):
        e = cases.apply(c, gpu_engine())
        e.setup(); e.run(500)
        snaps.append(cases.snapshot(e, c)); e.close()
    for k in snaps[0]:
        assert np.array_equal(snaps[0][k], snaps[1][k]), "run-to-run difference in " + k


def test_error_paths():
    import dem_b200
    e = gpu_engine()
    with pytest.raises(dem_b200.DemError):
        e.run(1)                      # run before setup
    with pytest.raises(dem_b200.DemError):
        e.pair_style("model luding tangential history")   # outside the hot-path scope
    with pytest.raises(dem_b200.DemError):
        e.property_global("youngsModulus", "peratomtype", [1.0, 2.0, 3.0])  # wrong count
    e.close()


def test_multi_gpu_bricks_match_oracle():
    """2 GPUs: brick decomposition + NCCL halo + migration against the single-process oracle"""
    import os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    here = os.path.dirname(os.path.abspath(__file__))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(here, "run_multi.py")], capture_output=True, text=True, timeout=600)
    assert "MULTI-GPU PARITY OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.parametrize("name", ["box_hertz_cdt", "periodic_epsd2", "poly_hooke_epsd_cyl"])
def test_contact_output_matches_reference_compute_pair_gran_local(name):
    """dem_download_contacts against `compute pair/gran/local id force torque` of the unmodified reference
    (tests/golden/contacts_*.npz, the values of the last step of a run): the same touching pairs, force and torque on the
    reference's first particle to 1e-10 -- and the rows of a particle add up to its pair force"""
    g = np.load(os.path.join(parity.ROOT, "tests", "golden", "contacts_%s.npz" % name))
    c = cases.make_case(name)
    e = cases.apply(c, gpu_engine())
    e.option("contact_output", 1)
    e.setup(); e.run(int(g["steps"]))
    ct = e.contacts()
    key_g = ct["tag"].astype(np.int64) * (1 << 32) + ct["partner"]
    pos = {k: r for r, k in enumerate(key_g)}
    key_r = g["id1"].astype(np.int64) * (1 << 32) + g["id2"]
    # the reference lists a pair once (twice across a periodic face); the engine lists both particles' views
    und_g = set(zip(np.minimum(ct["tag"], ct["partner"]).tolist(), np.maximum(ct["tag"], ct["partner"]).tolist()))
    und_r = set(zip(np.minimum(g["id1"], g["id2"]).tolist(), np.maximum(g["id1"], g["id2"]).tolist()))
    assert und_g == und_r, "touching pair sets differ"
    worst = 0.0
    for r in range(len(key_r)):
        q = pos[key_r[r]]
        sf = max(np.linalg.norm(g["force"][r]), 1e-300)
        worst = max(worst, np.linalg.norm(ct["force"][q] - g["force"][r]) / sf,
                    np.linalg.norm(ct["torque"][q] - g["torque"][r]) / max(np.linalg.norm(g["torque"][r]), 1e-3 * sf * c["radius"].min()))
    # (the rows follow a run of hundreds of steps: the bound is the one of the per-particle forces at that horizon)
    tol = parity.tol_for(c, int(g["steps"]), gpu=True)
    assert worst <= tol, "per-contact force / torque differ from the reference by %.2e (tol %.0e)" % (worst, tol)
    # Newton's third law between the two views of a pair of owned particles
    for (a, b) in list(und_g)[:200]:
        if (a * (1 << 32) + b) in pos and (b * (1 << 32) + a) in pos:
            fa, fb = ct["force"][pos[a * (1 << 32) + b]], ct["force"][pos[b * (1 << 32) + a]]
            if any(c["periodic"]):  # a pair across a periodic face: each side sees the other's shifted image, rounded once more
                assert np.allclose(fa, -fb, rtol=1e-9, atol=0.0)
            else:
                assert np.array_equal(fa, -fb)
    e.close()
