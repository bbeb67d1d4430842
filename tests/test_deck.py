"""The input-script front end (csrc/dem_deck.cpp): the SAME deck text that tests/golden/make_golden.py feeds to the
unmodified reference is parsed by the product's front end.  CPU: the parser source re-targeted at the oracle
(oracle/libdeck_oracle.so, tests only) against the reference's golden vectors; GPU: the shipped library."""
import ctypes
import os
import subprocess
import numpy as np
import pytest
import cases
import parity


def oracle_deck():
    import dem_b200
    so = os.path.join(parity.ROOT, "oracle", "libdeck_oracle.so")
    subprocess.run(["make", "-C", os.path.join(parity.ROOT, "oracle"), "libdeck_oracle.so"], check=True, capture_output=True)
    eng = parity.oracle_engine()
    return eng, dem_b200.Deck(eng, lib=ctypes.CDLL(so), prefix="orc_deck_")


def write_deck(c, tmp_path, lines_extra=()):
    deck, data = cases.to_deck(c, str(tmp_path / "case.data"))
    (tmp_path / "case.data").write_text(data)
    # the material lines live in a second script (`include`, Input::include)
    mat = [l for l in deck.splitlines() if " property/global " in l]
    if mat:
        (tmp_path / "in.materials").write_text("\n".join(mat) + "\n")
        first = deck.index(mat[0])
        deck = deck[:first] + "include in.materials\n" + "\n".join(l for l in deck[first:].splitlines() if l not in mat)
    # exercise the file reader: a comment, a continuation line and a variable
    deck = deck.replace("timestep ", "variable dt equal ").replace("fix integr all nve/sphere", "fix integr all &\n   nve/sphere   # integrator")
    deck += "\ntimestep ${dt}\nthermo 1000\ncompute 1 all erotate\n" + "\n".join(lines_extra) + "\n"
    (tmp_path / "in.case").write_text(deck)
    return str(tmp_path / "in.case")


def follow_golden(name, eng, deck, path, gpu):
    c = cases.make_case(name)
    g = parity.golden(name)
    deck.file(path)
    done = 0
    for cp in cases.GOLDEN_CASES[name]["checkpoints"]:
        for line in cases.late_commands(c, cp):
            deck.command(line)
        deck.command("run %d upto" % cp if cp else "run 0")
        if done == 0:
            parity.compare_topology(eng, c, g)
        done = cp
        parity.compare_snapshot(cases.snapshot(eng, c), parity.golden_at(g, cp), g["rmass"], tol=parity.tol_for(c, cp, gpu=gpu), label="deck %s@%d" % (name, cp))
        assert eng.stats().nbuilds == int(parity.golden_at(g, cp)["nbuilds"])
    assert deck.ntimestep == done
    assert "compute ignored" in deck.warnings and "Step" in deck.output
    deck.close(); eng.close()


DECK_CASES = ["box_hertz_cdt", "poly_hooke_epsd_cyl", "periodic_epsd2", "mesh_funnel_hooke", "mesh_plate_moving", "mesh_drum_rotating",
              "mesh_plate_late_move", "bond_nonlinear", "box_neigh_every2_delay4", "box_neigh_nocheck", "mesh_plate_stress", "mesh_drum_stress",
              "box_insert", "box_hyst1_thyst", "bond_linear"]  # (create_atoms single + set atom between two runs; all INL laws; bond counter)


@pytest.mark.parametrize("name", DECK_CASES)
def test_deck_on_oracle_matches_reference_golden(name, tmp_path):
    path = write_deck(cases.make_case(name), tmp_path)
    eng, deck = oracle_deck()
    follow_golden(name, eng, deck, path, gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", DECK_CASES)
def test_deck_on_engine_matches_reference_golden(name, tmp_path):
    import dem_b200
    path = write_deck(cases.make_case(name), tmp_path)
    eng = dem_b200.Engine(device=0)
    follow_golden(name, eng, dem_b200.Deck(eng), path, gpu=True)


def follow_insert_golden(name, eng, deck, gpu):
    """(also the lattice decks: lattice + create_atoms box|region, region INF / EDGE, group region|union|subtract, velocity set)
    fix insert/pack decks: the spheres the deck front end draws (Park-Miller streams, Monte-Carlo region volume, overlap search)
    must be the reference's spheres -- ids, types, radii and masses bit-exact, positions / velocities at the insertion step to
    1e-10 -- and stay on the reference's trajectory afterwards (tolerances of parity.tol_for)"""
    g = parity.golden(name)
    deck.file(os.path.join(parity.ROOT, "tests", "golden", "in." + name))
    for cp in cases.INSERT_DECKS[name]:
        deck.command("run %d upto" % cp)
        ref = parity.golden_at(g, cp)
        n = len(ref["tag"])
        assert eng.nlocal == n, "%s@%d: %d spheres, reference %d" % (name, cp, eng.nlocal, n)
        assert np.array_equal(eng.download("tag"), ref["tag"]) and np.array_equal(eng.download("type"), ref["type"])
        assert np.array_equal(eng.download("radius"), ref["radius"]) and np.array_equal(eng.download("rmass"), ref["rmass"]), "%s@%d: radius / mass" % (name, cp)
        tol = (1e-12 if not gpu else 1e-10) if cp <= 10 else parity.tol_for({"pair": "hertz"}, cp, gpu=gpu)
        mg = ref["rmass"] * 9.81
        floors = {"x": 1e-3, "v": 1e-3, "omega": 1e-2, "f": 1e-12 * mg, "torque": np.maximum(1e-15 * mg, 1e-9 * np.linalg.norm(ref["f"], axis=1))}
        for k in ("x", "v", "omega", "f", "torque"):
            err = parity.rel_err(eng.download(k).reshape(ref[k].shape), ref[k], floors[k])
            assert err <= tol, "%s@%d: %s rel err %.3e" % (name, cp, k, err)
        assert eng.stats().nbuilds == int(ref["nbuilds"])
    assert "Particle insertion ins: inserted" in deck.output or not name.startswith("insert_pack")
    deck.close(); eng.close()


@pytest.mark.parametrize("name", sorted(cases.INSERT_DECKS))
def test_insert_pack_deck_on_oracle_matches_reference(name):
    eng, deck = oracle_deck()
    follow_insert_golden(name, eng, deck, gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", cases.INSERT_DECKS_GPU)
def test_insert_pack_deck_on_engine_matches_reference(name):
    import dem_b200
    eng = dem_b200.Engine(device=0)
    follow_insert_golden(name, eng, dem_b200.Deck(eng), gpu=True)


TUTORIAL = "/root/reference/examples/LIGGGHTS/INL_tutorials/t01a_static_angle_of_repose_monosphere"


@pytest.mark.skipif(not os.path.isdir(TUTORIAL), reason="the reference's tutorial deck and STL file exist in the build container only")
def test_reference_tutorial_t01a_deck_runs_unchanged_on_oracle(tmp_path):
    """the reference's own tutorial deck t01a (static angle of repose: boundary m m m, cylinder region, fix insert/pack with
    volumefraction_region 0.6 -> 10,802 spheres, STL tube, primitive floor, thermo / dump / compute lines, the tube lifted by a
    late fix move/mesh) through the deck front end, only its two long runs shortened; compared with what the unmodified
    reference makes of the same text (tests/golden/tutorial_t01a.npz: every 16th sphere and the sums over all)"""
    sys_path = os.path.join(parity.ROOT, "tests", "golden")
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden_insert", os.path.join(sys_path, "make_golden_insert.py"))
    gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
    text = gen.tutorial_text()
    os.symlink(os.path.join(TUTORIAL, "Cylinder_35cm_7cm.stl"), tmp_path / "Cylinder_35cm_7cm.stl")
    head, tail = text.split("unfix", 1)
    (tmp_path / "in.head").write_text(head); (tmp_path / "in.tail").write_text("unfix" + tail)
    g = parity.golden("tutorial_t01a")
    eng, dk = oracle_deck()
    for cp, part, tol in ((1, "in.head", 1e-12), (500, "in.tail", 1e-6)):
        dk.file(str(tmp_path / part))
        assert dk.ntimestep == cp and eng.nlocal == int(g["s%d_n" % cp])
        for k in ("radius", "rmass"):
            assert np.array_equal(eng.download(k)[::16], g["s%d_%s" % (cp, k)])
        for k, floor in (("x", 1e-3), ("v", 1e-3), ("f", 1e-12 * 9.81 * g["s%d_rmass" % cp])):
            got, ref = eng.download(k), g["s%d_%s" % (cp, k)]
            assert parity.rel_err(got[::16], ref, floor) <= tol, "%s@%d" % (k, cp)
            assert np.allclose(got.sum(axis=0), g["s%d_sum_%s" % (cp, k)], rtol=1e-6, atol=1e-9 * len(got)), "sum %s@%d" % (k, cp)
    assert "inserted 10802 particle templates" in dk.output
    assert list(tmp_path.glob("out.*.dump")), "dump custom with dump_modify first yes"
    dk.close(); eng.close()


@pytest.mark.skipif(not os.path.isdir(cases.INL_EXAMPLES), reason="the reference's example decks exist in the build container only")
@pytest.mark.parametrize("rel", sorted(cases.INL_EXAMPLE_DECKS))
def test_reference_inl_example_deck_runs_unchanged_on_oracle(rel, tmp_path):
    """the reference's own INL example decks (bond/nonlinear chain bending in SI and micro units: create_atoms single, group id,
    set group, fix freeze, velocity set, 30 property/global lines through ${variables}; linear-bond chain bending: fix addforce,
    fix viscous, its full 100,000 steps) through the deck front end, only the run length cut; result bit-identical to the
    unmodified reference's (tests/golden/inl_examples.npz)"""
    (tmp_path / "in.deck").write_text(cases.example_deck_text(rel, cases.INL_EXAMPLE_DECKS[rel]))
    src = os.path.dirname(os.path.join(cases.INL_EXAMPLES, rel))
    for f in os.listdir(src):  # (mesh files are named relative to the deck)
        if os.path.isdir(os.path.join(src, f)):
            os.symlink(os.path.join(src, f), tmp_path / f)
    g = parity.golden("inl_examples")
    eng, dk = oracle_deck()
    dk.file(str(tmp_path / "in.deck"))
    assert dk.ntimestep > 0 and dk.ntimestep % cases.INL_EXAMPLE_DECKS[rel] == 0  # (one slice per `run` of the deck)
    key = rel.replace("/", "|")
    assert np.array_equal(eng.download("tag"), g[key + ":tag"])
    for k in ("radius", "rmass", "x", "v", "omega", "f", "torque"):
        ref, got = g[key + ":" + k], eng.download(k)
        if len(ref) <= 16 or k in ("radius", "rmass"):  # the small decks are bit-identical
            assert np.array_equal(got.reshape(ref.shape), ref), "%s: %s" % (rel, k)
        else:  # (1,800 spheres: the order of a particle's force sum differs in the last bit now and then)
            assert parity.rel_err(got.reshape(ref.shape), ref, 1e-9 * max(np.abs(ref).max(), 1e-300)) <= 1e-9, "%s: %s" % (rel, k)
    dk.close(); eng.close()


CHAIN_DECK = """
units si
atom_style granular
boundary f f f
newton off
communicate single vel yes
region dom block -0.005 0.005 -0.005 0.005 0.0 0.008 units box
create_box 1 dom
neighbor 5e-4 bin
neigh_modify delay 0
fix m1 all property/global youngsModulus peratomtype 1e7
fix m2 all property/global poissonsRatio peratomtype 0.3
fix m3 all property/global coefficientRestitution peratomtypepair 1 0.5
fix m4 all property/global coefficientFriction peratomtypepair 1 0.5
fix m5 all property/global radiusMultiplierBond peratomtypepair 1 0.9
fix m6 all property/global normalBondStiffnessPerUnitArea peratomtypepair 1 2e10
fix m7 all property/global tangentialBondStiffnessPerUnitArea peratomtypepair 1 8e9
fix m8 all property/global maxDistanceBond peratomtypepair 1 0.002
fix m9 all property/global dampingNormalForceBond peratomtypepair 1 0.
fix m10 all property/global dampingTangentialForceBond peratomtypepair 1 0.
fix m11 all property/global dampingNormalTorqueBond peratomtypepair 1 0.
fix m12 all property/global dampingTangentialTorqueBond peratomtypepair 1 0.
fix m13 all property/global tsCreateBond scalar 1
fix m14 all property/global createDistanceBond peratomtypepair 1 0.0011
pair_style gran model hertz tangential history cohesion bond
pair_coeff * *
create_atoms 1 single 0 0 0.001
create_atoms 1 single 0 0 0.002
create_atoms 1 single 0 0 0.003
create_atoms 1 single 0 0 0.004
create_atoms 1 single 0 0 0.005
set group all density 2500 diameter 0.001
group head id 1
group end id 5
fix bound_head head freeze
fix pull end addforce 2e-4 0.0 -1e-4
fix drag all viscous 2e-5
fix integr all nve/sphere
velocity end set 0.0 0.01 NULL
timestep 2e-8
"""


def run_chain(eng, dk, tmp_path):
    (tmp_path / "in.chain").write_text(CHAIN_DECK)
    dk.file(str(tmp_path / "in.chain"))
    out = {}
    dk.command("run 1000")
    out.update({k + "@1000": eng.download(k) for k in ("x", "v", "omega", "f", "torque")})
    dk.command("run 20000 upto"); dk.command("unfix pull"); dk.command("run 5000")
    out.update({k: eng.download(k) for k in ("x", "v", "omega", "f", "torque")})
    assert dk.ntimestep == 25000
    dk.close(); eng.close()
    return out


def test_chain_deck_addforce_viscous_velocity_on_oracle(tmp_path):
    """fix addforce / fix viscous / velocity set / set group / unfix of an addforce: the bent chain keeps its bonds, the pulled
    end moves the way the force points (goldens of these commands against the reference: test_reference_inl_example_deck...)"""
    eng, dk = oracle_deck()
    out = run_chain(eng, dk, tmp_path)
    assert out["x"][4, 0] > 1e-7 and abs(out["x"][0]).max() <= 0.001 and np.all(out["v"][0] == 0.0)


@pytest.mark.gpu
def test_chain_deck_addforce_viscous_velocity_on_engine_matches_oracle(tmp_path):
    import dem_b200
    eng, dk = oracle_deck()
    ref = run_chain(eng, dk, tmp_path)
    eng = dem_b200.Engine(device=0)
    got = run_chain(eng, dem_b200.Deck(eng), tmp_path)
    errs = {k: float(np.abs(got[k] - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-300)) for k in ref}
    print("chain deck, engine vs oracle (fraction of each field's scale):", errs)
    # the undamped stiff chain amplifies rounding noise: on the oracle alone a 1e-15 change of the initial velocity is O(1) in the
    # forces by step 10,000.  Tight at step 1000, positions only at the end.
    for k in ref:
        if k.endswith("@1000"):
            assert errs[k] <= 1e-9, "%s differs by %.2e of its scale (all: %s)" % (k, errs[k], errs)
    assert errs["x"] <= 1e-3, errs


def test_insertion_commands_error_classes():
    """argument errors carry the reference's message text (fix_insert.cpp / fix_insert_pack.cpp / fix_template_sphere.cpp /
    fix_particledistribution_discrete.cpp `error->fix_error`), requests outside the path are DEM_ERR_UNSUPPORTED"""
    import dem_b200
    eng, deck = oracle_deck()
    for line in ("units si", "boundary f f f", "region dom block 0 1 0 1 0 1 units box", "create_box 1 dom",
                 "region cyl cylinder z 0.5 0.5 0.2 0.1 0.9 units box", "region ball sphere 0.5 0.5 0.5 0.2 units box"):
        deck.command(line)
    bad = [
        (r"\(-2\).*prime numbers > 10000", "fix t all particletemplate/sphere 1 atom_type 1 density constant 2500 radius constant 0.01"),
        (r"\(-1\).*have to define 'density'", "fix t all particletemplate/sphere 15485863 atom_type 1 radius constant 0.01"),
        (r"\(-1\).*invalid radius random style", "fix t all particletemplate/sphere 15485863 atom_type 1 density constant 2500 radius uniform 0.01 0.02"),
        (r"\(-2\).*outside the hot-path scope", "fix t all particletemplate/multisphere 15485863 atom_type 1 density constant 2500 nspheres 2"),
        (r"\(-1\).*invalid ID for fix particletemplate", "fix pdd all particledistribution/discrete 15485867 1 nope 1.0"),
    ]
    for pat, line in bad:
        with pytest.raises(dem_b200.DemError, match=pat):
            deck.command(line)
    deck.command("fix t all particletemplate/sphere 15485863 atom_type 1 density constant 2500 radius constant 0.01")
    deck.command("fix pdd all particledistribution/discrete 15485867 1 t 1.0")
    bad = [
        (r"\(-1\).*# of templates does not match", "fix p2 all particledistribution/discrete 32452843 2 t 1.0"),
        (r"\(-1\).*expecting keyword 'seed'", "fix ins all insert/pack distributiontemplate pdd region cyl insert_every once particles_in_region 5"),
        (r"\(-1\).*must define an insertion region", "fix ins all insert/pack seed 32452843 distributiontemplate pdd insert_every once particles_in_region 5"),
        (r"\(-1\).*must define 'insert_every'", "fix ins all insert/pack seed 32452843 distributiontemplate pdd region cyl particles_in_region 5"),
        (r"\(-1\).*must define exactly one keyword", "fix ins all insert/pack seed 32452843 distributiontemplate pdd region cyl insert_every once particles_in_region 5 volumefraction_region 0.1"),
        (r"\(-1\).*ntry_mc must be > 1000", "fix ins all insert/pack seed 32452843 distributiontemplate pdd region cyl insert_every once particles_in_region 5 ntry_mc 10"),
        (r"\(-1\).*region ID does not exist", "fix ins all insert/pack seed 32452843 distributiontemplate pdd region nowhere insert_every once particles_in_region 5"),
        (r"\(-2\).*neither a block nor a cylinder", "fix ins all insert/pack seed 32452843 distributiontemplate pdd region ball insert_every once particles_in_region 5"),
        (r"\(-2\).*outside the hot-path scope", "fix ins all insert/stream seed 32452843 distributiontemplate pdd nparticles 10"),
        (r"\(-2\).*outside the hot-path scope", "fix f all addforce 1.0 0.0 v_fz"),
        (r"\(-1\).*Illegal fix viscous command", "fix v all viscous"),
    ]
    for pat, line in bad:
        with pytest.raises(dem_b200.DemError, match=pat):
            deck.command(line)
    deck.close(); eng.close()


def test_deck_mesh_load_transforms_and_errors(tmp_path):
    """`fix mesh/surface ... move/rotate/scale` act on the nodes like FixMesh::moveMesh/rotateMesh/scaleMesh; error classes"""
    import dem_b200
    c = cases.make_case("mesh_funnel_hooke")
    path = write_deck(c, tmp_path)
    text = open(path).read()
    stl = str(tmp_path / "fun.stl")
    text = text.replace("file %s type 1" % stl, "file %s type 1 move 0.001 0. 0. rotate axis 0. 0. 1. angle 90. scale 0.5" % stl)
    open(path, "w").write(text)
    eng, deck = oracle_deck()
    deck.file(path)
    nodes = np.asarray([m for m in c["meshes"] if m[0] == "fun"][0][2]).reshape(-1, 3)
    moved = nodes + [0.001, 0., 0.]
    a = 90. * 3.14159265 / 180.
    want = 0.5 * np.stack([np.cos(a) * moved[:, 0] - np.sin(a) * moved[:, 1], np.sin(a) * moved[:, 0] + np.cos(a) * moved[:, 1], moved[:, 2]], 1)
    got = eng.mesh_field("fun", "nodes", len(nodes) // 3).reshape(-1, 3)
    assert np.abs(got - want).max() < 1e-15
    # error classes: unsupported (-2) vs bad argument (-1), with the reference's message text where one exists
    with pytest.raises(dem_b200.DemError, match=r"\(-2\).*outside the hot-path scope"):
        deck.command("fix ins all insert/stream seed 1")
    with pytest.raises(dem_b200.DemError, match=r"\(-1\).*Expected floating point parameter"):
        deck.command("timestep abc")
    with pytest.raises(dem_b200.DemError, match=r"\(-1\).*Substitution for illegal variable"):
        deck.command("timestep ${nope}")
    with pytest.raises(dem_b200.DemError, match=r"\(-2\)"):
        deck.command("pair_style lj/cut 2.5")
    with pytest.raises(dem_b200.DemError, match=r"Could not find fix group ID"):
        deck.command("fix g nobody gravity 9.81 vector 0 0 -1")
    deck.close(); eng.close()


@pytest.mark.gpu
def test_lmp_b200_cli_runs_a_deck(tmp_path):
    """`lmp_b200 -in deck`: the reference's `lmp_<machine> -in deck` for the hot-path commands"""
    path = write_deck(cases.make_case("box_hertz_cdt"), tmp_path, lines_extra=["run 300"])
    exe = os.path.join(parity.ROOT, "liggghts-inl_b200", "lmp_b200")
    r = subprocess.run([exe, "-in", path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "for 300 steps with 64 atoms" in r.stdout, r.stdout
    bad = tmp_path / "in.bad"; bad.write_text("units si\nfix ins all insert/stream seed 1\n")
    r = subprocess.run([exe, "-in", str(bad)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 1 and "outside the hot-path scope" in r.stderr and "line 2" in r.stderr


def test_tutorial_style_deck_on_oracle(tmp_path):
    """a deck written the way the INL tutorial decks are (tabs, trailing comments, continuation lines, create_box from a region,
    two atom types with the wall as the last one, a diagnostic fix that is later unfixed, `run N upto`, a mesh that starts to
    move between two runs); particles from read_data here, fix insert/pack has its own tests above"""
    import dem_b200
    c = cases.case_mesh(kind="plate", n3=(4, 4, 3), name="tut")
    _, data = cases.to_deck(c, str(tmp_path / "case.data"))
    (tmp_path / "case.data").write_text(data)
    for mid, mtype, nodes in c["meshes"]:
        cases.write_stl(str(tmp_path / (mid + ".stl")), nodes, mid)
    lo, hi = c["lo"], c["hi"]
    deck = """
units\t\tsi
atom_style\tsphere
atom_modify\tmap array
boundary\tm m m # f f f
newton\t\toff
communicate\tsingle vel yes
region\t\tdomain block %.17g %.17g %.17g %.17g %.17g %.17g units box # the simulation box
create_box\t2 domain # the last type is the wall
read_data\tcase.data
neighbor\t0.001 bin
neigh_modify\tdelay 0
variable\tyoung equal 5e6
fix\t\tm1 all property/global youngsModulus peratomtype ${young}  1e9
fix\t\tm2 all property/global poissonsRatio peratomtype 0.45 0.3
fix\t\tm3 all property/global coefficientRestitution peratomtypepair 2 &
\t\t0.3 0.3 &
\t\t0.3 0.3
fix\t\tm4 all property/global coefficientFriction peratomtypepair 2 &
\t\t0.5 0.3 &
\t\t0.3 0.3
fix\t\tm11 all property/global coefficientRollingFriction peratomtypepair 2 0.1 0.1 0.1 0.1
fix\t\tcad all mesh/surface file cad.stl    type 2
fix\t\tplate all mesh/surface file plate.stl type 2
fix\t\tgeo all wall/gran model hertz tangential history rolling_friction cdt mesh n_meshes 2 meshes cad plate
pair_style\tgran model hertz tangential history rolling_friction cdt
pair_coeff\t* *
fix\t\tintegr1 all nve/sphere
fix\t\tgrav all gravity 9.81 vector 0.0 0.0 -1.0
timestep\t0.00001
compute\t\t1 all erotate
thermo_style\tcustom step atoms c_1 cpu
thermo\t\t50000
fix\t\tctg all check/timestep/gran 1 0.01 0.01
region\t\tfactory cylinder z 0 0 0.033 0.002 1.0 units box
fix\t\tbal all balance 100 xyz 20 1.2
run\t\t1
unfix\t\tctg
dump\t\tdmp all custom 50000 out.*.dump id type x y z radius
run\t\t300 upto
fix\t\tmove all move/mesh mesh plate linear 0. 0. -0.4
run\t\t200
""" % (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])
    (tmp_path / "in.tut").write_text(deck)
    eng, dk = oracle_deck()
    dk.file(str(tmp_path / "in.tut"))
    assert dk.ntimestep == 500
    assert "check/timestep/gran ignored" in dk.warnings and "fix balance ignored" in dk.warnings
    assert not list(tmp_path.glob("out.*.dump")), "the dump is defined at step 1: no multiple of 50000 is reached in 500 steps"
    # the same through the API calls, step for step
    ref = parity.oracle_engine()
    ref.units("si"); ref.box(lo, hi, [0, 0, 0]); ref.ntypes(2); ref.neighbor(0.001, every=1, delay=0, check=True)
    ref.property_global("youngsModulus", "peratomtype", [5e6, 1e9]); ref.property_global("poissonsRatio", "peratomtype", [0.45, 0.3])
    ref.property_global("coefficientRestitution", "peratomtypepair", [0.3] * 4); ref.property_global("coefficientFriction", "peratomtypepair", [0.5, 0.3, 0.3, 0.3])
    ref.property_global("coefficientRollingFriction", "peratomtypepair", [0.1] * 4)
    for mid, mtype, nodes in c["meshes"]:
        ref.mesh(mid, 2, nodes)
    ref.wall_mesh("geo", "model hertz tangential history rolling_friction cdt mesh n_meshes 2 meshes cad plate")
    ref.pair_style("model hertz tangential history rolling_friction cdt")
    ref.integrate(1); ref.gravity(9.81, [0.0, 0.0, -1.0]); ref.timestep(0.00001)
    ref.upload(c["tag"], c["type"], c["x"], c["radius"], c["density"], v=c["v"], omega=c["omega"], mask=c["mask"])
    ref.setup(); ref.run(1); ref.setup(); ref.run(299)
    ref.move_mesh("plate", "linear 0. 0. -0.4")
    ref.setup(); ref.run(200)
    for k in ("x", "v", "f", "omega", "torque"):
        assert np.array_equal(eng.download(k), ref.download(k)), k
    dk.close(); eng.close(); ref.close()


def test_deck_variable_formulas(tmp_path):
    """equal-style formulas as the INL example decks use them (`variable max_dist equal 1.5*${d}`), evaluated at substitution
    and printed %.15g like Variable::retrieve; output-only variables (vx[1], time, f_mesh[3]) are accepted and only fail when a
    hot-path command asks for their value"""
    import dem_b200
    c = cases.make_case("box_hertz_cdt")
    path = write_deck(c, tmp_path)
    text = open(path).read()
    text = text.replace("variable dt equal", "variable dt_unused equal").replace("timestep ${dt}", "")
    text += "\n".join(["variable d equal 2", "variable half equal 0.5*${d}", "variable vz1 equal vz[1]", "variable top_Fz equal f_mesh_top[3]",
                       "variable tc equal time", "variable dt equal ${half}*1e-5*(v_d^2-3)+sqrt(16)*0-abs(-0)+(2^3^2-64)+(-2^2-4)",  # ^ is left-associative and binds less than unary minus (variable.cpp:133-141)
                       "timestep ${dt}", "run 10"]) + "\n"
    open(path, "w").write(text)
    eng, deck = oracle_deck()
    deck.file(path)
    ref = cases.apply(c, parity.oracle_engine())   # 0.5*2 * 1e-5 * (4-3) == the case's 1e-5
    ref.setup(); ref.run(10)
    for k in ("x", "v", "f"):
        assert np.array_equal(eng.download(k), ref.download(k)), k
    with pytest.raises(dem_b200.DemError, match=r"\(-2\).*vz.*outside the hot-path scope"):
        deck.command("timestep ${vz1}")
    with pytest.raises(dem_b200.DemError, match=r"\(-1\).*Divide by 0"):
        deck.command("variable z equal 1/0"); deck.command("timestep ${z}")
    with pytest.raises(dem_b200.DemError, match=r"\(-1\).*Invalid syntax"):
        deck.command("variable y equal 2*(3"); deck.command("timestep ${y}")
    deck.close(); eng.close(); ref.close()


# ---- dump custom: the text snapshots of the deck front end against the reference's own dump file -------------------------------
DUMP_FIELDS = "id type x y z vx vy vz fx fy fz omegax omegay omegaz radius"


def _parse_dump(text):
    snaps, lines, k = [], text.splitlines(), 0
    while k < len(lines):
        assert lines[k] == "ITEM: TIMESTEP"
        step = int(lines[k + 1]); n = int(lines[k + 3])
        head = lines[k + 2:k + 9]
        cols = lines[k + 8].split()[2:]
        rows = np.array([[float(v) for v in ln.split()] for ln in lines[k + 9:k + 9 + n]])
        snaps.append((step, head, cols, rows))
        k += 9 + n
    return snaps


def _follow_dump(eng, deck, tmp_path, tol):
    c = cases.make_case("box_hertz_cdt")
    path = write_deck(c, tmp_path)
    out = os.path.join(str(tmp_path), "dump.txt")
    deck.file(path)
    deck.command("dump d1 all custom 100 %s %s" % (out, DUMP_FIELDS))
    deck.command("dump_modify d1 sort id")
    deck.command("thermo_style custom step atoms ke erotate")
    deck.command("thermo 100")
    mark = len(deck.output)
    deck.command("run 250")
    # thermo lines of the run against the reference's log (header text identical, values printed %14.8g)
    tg = deck.output[mark:].splitlines()
    tr = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thermo_box_hertz_cdt.txt")).read().splitlines()
    assert tg[0] == tr[0], "thermo header %r != %r" % (tg[0], tr[0])
    assert len(tg) == len(tr)
    for a, b in zip(tg[1:], tr[1:]):
        fa, fb = [float(v) for v in a.split()], [float(v) for v in b.split()]
        assert fa[:2] == fb[:2] and np.allclose(fa[2:], fb[2:], rtol=max(tol, 1e-6), atol=0.0), "thermo line %r != %r" % (a, b)
    got = _parse_dump(open(out).read())
    ref = _parse_dump(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dump_box_hertz_cdt.txt")).read())
    assert [s[0] for s in got] == [s[0] for s in ref] == [0, 100, 200]
    for (gs, gh, gc, gr), (rs, rh, rc, rr) in zip(got, ref):
        assert gh == rh, "snapshot header differs at step %d: %s != %s" % (gs, gh, rh)
        assert gc == rc and gr.shape == rr.shape
        assert np.array_equal(gr[:, :2], rr[:, :2])
        scale = np.maximum(np.abs(rr[:, 2:]).max(axis=0), 1e-300)
        assert (np.abs(gr[:, 2:] - rr[:, 2:]) / scale).max() <= tol, "dump values differ at step %d" % gs   # (%g prints six digits)
    assert deck.ntimestep == 250
    deck.close(); eng.close()


def test_dump_custom_on_oracle_matches_reference_file(tmp_path):
    eng, deck = oracle_deck()
    _follow_dump(eng, deck, tmp_path, 2e-6)


@pytest.mark.gpu
def test_dump_custom_on_engine_matches_reference_file(tmp_path):
    import dem_b200
    eng = dem_b200.Engine(device=0)
    _follow_dump(eng, dem_b200.Deck(eng), tmp_path, 2e-6)
