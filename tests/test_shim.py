"""The `-suffix b200` binding inside the reference tree (integration/b200_shim.{h,cpp}, built by integration/Makefile into
oracle/_ref/): an UNMODIFIED input deck, run by the reference binary with `-suffix b200`, hands its timestep loop to the
engine through the C ABI's input-script front end and must reproduce the plain reference run.
  CPU (not gpu): the shim bound to the CPU oracle's orc_* entry points (libliggghts_ref_orc.so) -- covers the binding itself;
  GPU: the shim bound to libdem_b200.so (libliggghts_ref_b200.so)."""
import os
import subprocess
import sys
import numpy as np
import pytest
import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFDIR = os.path.join(ROOT, "oracle", "_ref")
WORKER = r'''
import sys, os, tempfile
sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/oracle"); sys.path.insert(0, sys.argv[1] + "/liggghts-inl_b200")
import numpy as np, cases, ref_driver
name, lib, out = sys.argv[2], sys.argv[3], sys.argv[4]
c = cases.make_case(name)
tmp = tempfile.mkdtemp(); os.chdir(tmp)
deck, data = cases.to_deck(c, os.path.join(tmp, "case.data")); open(os.path.join(tmp, "case.data"), "w").write(data)
r = ref_driver.Ref(lib=lib, extra_args=["-suffix", "b200"]) if lib != "plain" else ref_driver.Ref()
r.cmd(deck)
done = 0
for cp in (1, 200, 500):   # three `run` commands: the engine lives across them
    for line in cases.late_commands(c, cp): r.cmd(line)
    r.cmd("run %d" % (cp - done)); done = cp
a = r.atoms()
np.savez(out, **{k: a[k] for k in ("x", "v", "f", "omega", "torque")})
'''


def run_ref(name, lib, tmp_path):
    """one reference process per run: the reference keeps global registries (and exit()s on errors)"""
    out = str(tmp_path / ("%s_%s.npz" % (name, os.path.basename(lib))))
    r = subprocess.run([sys.executable, "-c", WORKER, ROOT, name, lib, out], capture_output=True, text=True, timeout=600)
    assert os.path.exists(out), "reference run failed: " + r.stdout[-1500:] + r.stderr[-1500:]
    return np.load(out)


def check(name, lib, tmp_path):
    a, b = run_ref(name, "plain", tmp_path), run_ref(name, lib, tmp_path)
    for k in a.files:
        scale = max(np.abs(a[k]).max(), 1e-300)
        assert np.abs(a[k] - b[k]).max() <= 1e-6 * scale, "%s: %s differs by %.2e of its scale" % (name, k, np.abs(a[k] - b[k]).max() / scale)


@pytest.mark.parametrize("name", ["box_hertz_cdt", "mesh_plate_late_move", "poly_hooke_epsd_cyl"])
def test_suffix_b200_deck_on_oracle_binding(name, tmp_path):
    lib = os.path.join(REFDIR, "libliggghts_ref_orc.so")
    if not (os.path.exists(lib) and os.path.exists(os.path.join(REFDIR, "libliggghts_ref.so"))):
        pytest.skip("the shim build of the reference tree is not here (make -C integration orc)")
    check(name, lib, tmp_path)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["box_hertz_cdt", "mesh_plate_late_move", "mesh_drum_rotating", "periodic_epsd2"])
def test_suffix_b200_deck_on_gpu_engine(name, tmp_path):
    lib = os.path.join(REFDIR, "libliggghts_ref_b200.so")
    if not (os.path.exists(lib) and os.path.exists(os.path.join(REFDIR, "libliggghts_ref.so"))):
        pytest.skip("the shim build of the reference tree is not here (make -C integration)")
    check(name, lib, tmp_path)


# ---- insertion under `-suffix b200`: the reference's own fix insert/pack keeps drawing the spheres (its streams, regions and
# templates), VerletB200::insertion_step hands them to the engine inside the timestep (dem_insert_step_begin / _end)
INS_WORKER = r'''
import sys, os
sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/oracle")
import numpy as np, cases, ref_driver
name, lib, out = sys.argv[2], sys.argv[3], sys.argv[4]
r = ref_driver.Ref(lib=lib, extra_args=["-suffix", "b200"])
r.cmd(open(os.path.join(sys.argv[1], "tests", "golden", "in." + name)).read())
res = {}
for cp in cases.INSERT_DECKS[name]:
    r.cmd("run %d upto" % cp)
    a = r.atoms()
    for k in ("tag", "radius", "rmass", "x", "v", "omega", "f", "torque"):
        res["s%d_%s" % (cp, k)] = a[k]
np.savez(out, **res)
'''


def check_insertion(name, lib, tmp_path, gpu):
    import parity
    out = str(tmp_path / (name + ".npz"))
    r = subprocess.run([sys.executable, "-c", INS_WORKER, ROOT, name, lib, out], capture_output=True, text=True, timeout=600)
    assert os.path.exists(out), "reference run failed: " + r.stdout[-1500:] + r.stderr[-1500:]
    a, g = np.load(out), parity.golden(name)
    for cp in cases.INSERT_DECKS[name]:
        for k in ("tag", "radius", "rmass"):
            assert np.array_equal(a["s%d_%s" % (cp, k)], g["s%d_%s" % (cp, k)]), "%s@%d: %s" % (name, cp, k)
        tol = (1e-10 if gpu else 1e-12) if cp <= 10 else parity.tol_for({"pair": "hertz"}, cp, gpu=gpu)
        mg = g["s%d_rmass" % cp] * 9.81
        floors = {"x": 1e-3, "v": 1e-3, "omega": 1e-2, "f": 1e-12 * mg, "torque": np.maximum(1e-15 * mg, 1e-9 * np.linalg.norm(g["s%d_f" % cp], axis=1))}
        for k in ("x", "v", "omega", "f", "torque"):
            err = parity.rel_err(a["s%d_%s" % (cp, k)], g["s%d_%s" % (cp, k)], floors[k])
            assert err <= tol, "%s@%d: %s rel err %.3e" % (name, cp, k, err)


@pytest.mark.parametrize("name", sorted(cases.INSERT_DECKS))
def test_suffix_b200_insertion_deck_on_oracle_binding(name, tmp_path):
    lib = os.path.join(REFDIR, "libliggghts_ref_orc.so")
    if not os.path.exists(lib):
        pytest.skip("the shim build of the reference tree is not here (make -C integration orc)")
    check_insertion(name, lib, tmp_path, gpu=False)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["insert_pack_a", "insert_pack_b"])
def test_suffix_b200_insertion_deck_on_gpu_engine(name, tmp_path):
    lib = os.path.join(REFDIR, "libliggghts_ref_b200.so")
    if not os.path.exists(lib):
        pytest.skip("the shim build of the reference tree is not here (make -C integration)")
    check_insertion(name, lib, tmp_path, gpu=True)


EX_WORKER = r'''
import sys, os
sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/oracle")
import numpy as np, cases, ref_driver
rel, lib, out = sys.argv[2], sys.argv[3], sys.argv[4]
os.chdir(os.path.dirname(os.path.join(cases.INL_EXAMPLES, rel)))  # (mesh files are named relative to the deck; dump lines are dropped)
r = ref_driver.Ref(lib=lib, extra_args=["-suffix", "b200"])
r.cmd(cases.example_deck_text(rel, cases.ALL_EXAMPLE_DECKS[rel]))
a = r.atoms()
np.savez(out, **{k: a[k] for k in ("tag", "x", "v", "omega", "f", "torque")})
'''


@pytest.mark.skipif(not os.path.isdir(cases.INL_EXAMPLES), reason="the reference's example decks exist in the build container only")
@pytest.mark.parametrize("rel", sorted(cases.ALL_EXAMPLE_DECKS))
def test_suffix_b200_inl_example_deck_on_oracle_binding(rel, tmp_path):
    """the reference's INL example decks (bonded chains: fix addforce / viscous / freeze, velocity set, set group) run by the
    reference binary with `-suffix b200` on the oracle binding: bit-identical to the plain reference (inl_examples.npz)"""
    import parity
    lib = os.path.join(REFDIR, "libliggghts_ref_orc.so")
    if not os.path.exists(lib):
        pytest.skip("the shim build of the reference tree is not here (make -C integration orc)")
    out = str(tmp_path / "ex.npz")
    r = subprocess.run([sys.executable, "-c", EX_WORKER, ROOT, rel, lib, out], capture_output=True, text=True, timeout=600)
    assert os.path.exists(out), "reference run failed: " + r.stdout[-1500:] + r.stderr[-1500:]
    a, g = np.load(out), parity.golden("inl_examples")
    key = rel.replace("/", "|")
    for k in a.files:
        ref = g[key + ":" + k]
        if len(ref) <= 16 or k == "tag":
            assert np.array_equal(a[k], ref), "%s: %s" % (rel, k)
        else:
            assert parity.rel_err(a[k], ref, 1e-9 * max(np.abs(ref).max(), 1e-300)) <= 1e-9, "%s: %s" % (rel, k)


TUT_WORKER = r'''
import sys, os, importlib.util
sys.path.insert(0, sys.argv[1] + "/tests"); sys.path.insert(0, sys.argv[1] + "/oracle")
import numpy as np, ref_driver
spec = importlib.util.spec_from_file_location("mgi", sys.argv[1] + "/tests/golden/make_golden_insert.py")
gen = importlib.util.module_from_spec(spec); spec.loader.exec_module(gen)
os.chdir(gen.TUTORIAL)  # (the deck names its STL file relative to itself; dump lines dropped: nothing is written there)
text = "\n".join(l for l in gen.tutorial_text().replace("&\n", " ").splitlines() if not l.strip().startswith("dump"))
r = ref_driver.Ref(lib=sys.argv[2], extra_args=["-suffix", "b200"])
r.cmd(text)
a = r.atoms()
np.savez(sys.argv[3], n=len(a["tag"]), **{k: a[k][::16] for k in ("x", "v", "f", "radius", "rmass")})
'''


@pytest.mark.skipif(not os.path.isdir(cases.INL_EXAMPLES), reason="the reference's tutorial deck exists in the build container only")
def test_suffix_b200_tutorial_t01a_on_oracle_binding(tmp_path):
    """the reference binary with `-suffix b200` on its own tutorial deck t01a (runs cut to 500 steps): 10,802 spheres from the
    reference's fix insert/pack enter the engine inside timestep 1, STL tube, late fix move/mesh, thermo with c_pe -- against the
    plain reference's result (tutorial_t01a.npz)"""
    import parity
    lib = os.path.join(REFDIR, "libliggghts_ref_orc.so")
    if not os.path.exists(lib):
        pytest.skip("the shim build of the reference tree is not here (make -C integration orc)")
    out = str(tmp_path / "tut.npz")
    r = subprocess.run([sys.executable, "-c", TUT_WORKER, ROOT, lib, out], capture_output=True, text=True, timeout=600)
    assert os.path.exists(out), "reference run failed: " + r.stdout[-1500:] + r.stderr[-1500:]
    a, g = np.load(out), parity.golden("tutorial_t01a")
    assert int(a["n"]) == int(g["s500_n"]) == 10802
    assert np.array_equal(a["radius"], g["s500_radius"]) and np.array_equal(a["rmass"], g["s500_rmass"])
    for k, floor in (("x", 1e-3), ("v", 1e-3), ("f", 1e-12 * 9.81 * g["s500_rmass"])):
        assert parity.rel_err(a[k], g["s500_" + k], floor) <= 1e-9, k
