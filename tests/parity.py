"""Parity helpers (tests only).  Tolerances: SURVEY.md 8d / BASELINE.json north_star --
pair sets and history bookkeeping bit-exact; forces/torques rel. err <= 1e-10 (fp64)."""
import ctypes
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "liboracle.so")
FTOL = 1e-10


def oracle_lib():
    """build (if needed) and load the CPU restatement -- the checker, never the product"""
    src = os.path.join(ROOT, "oracle", "dem_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle.so"], check=True, capture_output=True)
    return ctypes.CDLL(ORACLE_SO)


def oracle_engine():
    import dem_b200
    return dem_b200.Engine(lib=oracle_lib(), prefix="orc_")


def unique_pairs(lo, hi, flag, hist):
    """the reference lists an owned/periodic-ghost pair on both sides: keep one row per pair"""
    key = lo.astype(np.int64) * (1 << 32) + hi.astype(np.int64)
    _, idx = np.unique(key, return_index=True)
    return lo[idx], hi[idx], flag[idx], hist[idx]


def rel_err(a, b, floor):
    """max |a-b| / max(|b|_row, floor): per-particle vector error with an absolute floor"""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    den = np.maximum(np.linalg.norm(b.reshape(len(b), -1), axis=1), floor)
    return float(np.max(np.linalg.norm((a - b).reshape(len(b), -1), axis=1) / den)) if len(b) else 0.0


def compare_snapshot(got, ref, rmass, tol=FTOL, tol_state=None, hist_tol=None, label=""):
    """got/ref: dicts with x,v,f,omega,torque,pair_lo,pair_hi,pair_flag,pair_hist,(wall_*)"""
    tol_state = tol if tol_state is None else tol_state
    hist_tol = tol if hist_tol is None else hist_tol
    mg = rmass * 9.81
    glo, ghi, gfl, gh = unique_pairs(got["pair_lo"], got["pair_hi"], got["pair_flag"], got["pair_hist"])
    rlo, rhi, rfl, rh = unique_pairs(ref["pair_lo"], ref["pair_hi"], ref["pair_flag"], ref["pair_hist"])
    assert len(glo) == len(rlo) and np.array_equal(glo, rlo) and np.array_equal(ghi, rhi), label + ": pair set differs"
    assert np.array_equal(gfl != 0, rfl != 0), label + ": contact flags differ"
    errs = {}
    if rh.size:
        scale = max(np.abs(rh).max(), 1e-300)
        errs["hist"] = float(np.abs(gh - rh).max() / scale)
        assert errs["hist"] <= hist_tol, "%s: history rel err %.3e" % (label, errs["hist"])
    errs["f"] = rel_err(got["f"], ref["f"], 1e-12 * mg)
    rad_t = 1e-12 * mg * 1e-3
    errs["torque"] = rel_err(got["torque"], ref["torque"], np.maximum(rad_t, 1e-6 * np.linalg.norm(ref["f"], axis=1) * 1e-3))
    assert errs["f"] <= tol, "%s: force rel err %.3e" % (label, errs["f"])
    assert errs["torque"] <= tol, "%s: torque rel err %.3e" % (label, errs["torque"])
    for k, fl in (("x", 1e-3), ("v", 1e-3), ("omega", 1e-2)):
        errs[k] = rel_err(got[k], ref[k], fl)
        assert errs[k] <= tol_state, "%s: %s rel err %.3e" % (label, k, errs[k])
    for k in ref:
        if k.startswith("wall_") and k in got:
            scale = max(np.abs(ref[k]).max(), 1e-300)
            errs[k] = float(np.abs(got[k] - ref[k]).max() / scale)
            tiny = 1e-12 * scale  # projections leave 1e-40-size residues: "zeroed" means far below scale
            assert np.array_equal(np.abs(got[k]) > tiny, np.abs(ref[k]) > tiny), "%s: %s bookkeeping differs" % (label, k)
            assert errs[k] <= hist_tol, "%s: %s rel err %.3e" % (label, k, errs[k])
    for k in ref:  # triangle-mesh contact rows: (tag, triangle) sets bit-exact, history to tolerance
        if k.startswith("mesh_") and k.endswith("_tag"):
            mid = k[5:-4]
            gt, gi = got["mesh_%s_tag" % mid], got["mesh_%s_tri" % mid]
            assert len(gt) == len(ref[k]) and np.array_equal(gt, ref[k]) and np.array_equal(gi, ref["mesh_%s_tri" % mid]), \
                "%s: mesh %s contact rows differ" % (label, mid)
            rh, gh = ref["mesh_%s_hist" % mid], got["mesh_%s_hist" % mid]
            if rh.size:
                errs["mesh_" + mid] = float(np.abs(gh - rh).max() / max(np.abs(rh).max(), 1e-300))
                assert errs["mesh_" + mid] <= hist_tol, "%s: mesh %s history rel err %.3e" % (label, mid, errs["mesh_" + mid])
    for k in ref:  # fix mesh/surface/stress: total force, total torque, reference point of a mesh (f_<id>[1..9])
        if k.startswith("meshforce_"):
            assert k in got, label + ": " + k + " missing"
            gv, rv = np.asarray(got[k]), np.asarray(ref[k])
            fs = max(np.abs(rv[:3]).max(), 1e-12 * float(mg.max()))
            errs[k] = float(np.abs(gv[:3] - rv[:3]).max() / fs)
            assert errs[k] <= max(tol, 1e-12), "%s: %s force rel err %.3e" % (label, k, errs[k])
            ts = max(np.abs(rv[3:6]).max(), fs * 1e-3)
            assert np.abs(gv[3:6] - rv[3:6]).max() / ts <= max(tol, 1e-12), "%s: %s torque differs" % (label, k)
            assert np.abs(gv[6:9] - rv[6:9]).max() <= 1e-14 * max(1.0, np.abs(rv[6:9]).max()), "%s: %s reference point differs" % (label, k)
    return errs


def compare_topology(eng, c, g):
    """active edge / corner flags and neighbour counts of every mesh against the reference's (golden `topo_*`)"""
    for mid, mtype, nodes in c.get("meshes", []):
        for k in ("edge_active", "corner_active", "nneighs"):
            got = eng.mesh_field(mid, k, len(nodes))
            assert np.array_equal(got, g["topo_%s_%s" % (mid, k)]), "mesh %s: %s differs from the reference" % (mid, k)


def tol_for(c, cp, gpu=False):
    """force/state tolerance at checkpoint `cp`.  The first steps carry the 1e-10 bar of BASELINE.json; DEM trajectories are
    chaotic, so rounding-level differences (summation order, FMA contraction on the GPU) grow with the step count.  The bond
    models damp with sgn(v)*|F| (cohesion_model_bond.h:700-712): a sign flip of a ~1e-16 velocity component changes a force by
    2*damping*|F|, so bonded decks lose digits much faster and are only compared over short horizons."""
    if cp <= 10:
        return 1e-10
    if "cohesion" in c["pair"]:
        return 1e-5 if cp <= 100 else 1e-3
    if gpu:
        return 1e-6 if cp <= 400 else 1e-4
    return 1e-7 if cp <= 400 else 1e-5


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def golden_at(g, cp):
    pre = "s%d_" % cp
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}
