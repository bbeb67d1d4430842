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


def compare_snapshot(got, ref, rmass, tol=FTOL, tol_state=None, hist_tol=None, label="", hist_floor=1e-300):
    """got/ref: dicts with x,v,f,omega,torque,pair_lo,pair_hi,pair_flag,pair_hist,(wall_*); hist_floor: absolute scale below
    which history values are rounding noise (e.g. the tangential spring of a plate that only moves along its normal)"""
    tol_state = tol if tol_state is None else tol_state
    hist_tol = tol if hist_tol is None else hist_tol
    rmass = np.asarray(rmass)[:len(ref["x"])]  # (cases that insert particles later: the fixture's masses are those of the final set, tag order)
    if "bondcounter" in got and "bondcounter" in ref:  # bonds created / broken since the previous checkpoint: exact
        assert np.array_equal(np.asarray(got["bondcounter"]), np.asarray(ref["bondcounter"])), "%s: bond counter %s != %s" % (label, got["bondcounter"], ref["bondcounter"])
    mg = rmass * 9.81
    glo, ghi, gfl, gh = unique_pairs(got["pair_lo"], got["pair_hi"], got["pair_flag"], got["pair_hist"])
    rlo, rhi, rfl, rh = unique_pairs(ref["pair_lo"], ref["pair_hi"], ref["pair_flag"], ref["pair_hist"])
    assert len(glo) == len(rlo) and np.array_equal(glo, rlo) and np.array_equal(ghi, rhi), label + ": pair set differs"
    assert np.array_equal(gfl != 0, rfl != 0), label + ": contact flags differ"
    errs = {}
    if rh.size:
        scale = max(np.abs(rh).max(), hist_floor)
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
                errs["mesh_" + mid] = float(np.abs(gh - rh).max() / max(np.abs(rh).max(), hist_floor))
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


def compare_bookkeeping(got, ref, label=""):
    """pair set, contact flags and mesh contact rows only (bit-exact) -- for horizons at which per-particle values of a
    chaotic bed have lost their meaning"""
    glo, ghi, gfl, _ = unique_pairs(got["pair_lo"], got["pair_hi"], got["pair_flag"], got["pair_hist"])
    rlo, rhi, rfl, _ = unique_pairs(ref["pair_lo"], ref["pair_hi"], ref["pair_flag"], ref["pair_hist"])
    assert len(glo) == len(rlo) and np.array_equal(glo, rlo) and np.array_equal(ghi, rhi), label + ": pair set differs"
    assert np.array_equal(gfl != 0, rfl != 0), label + ": contact flags differ"
    for k in ref:
        if k.startswith("mesh_") and k.endswith("_tag"):
            mid = k[5:-4]
            assert np.array_equal(got[k], ref[k]) and np.array_equal(got["mesh_%s_tri" % mid], ref["mesh_%s_tri" % mid]), "%s: mesh %s contact rows differ" % (label, mid)


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
    if "hysteretic" in c["pair"]:  # loading / unloading branches switch on `deltan >= delta_old` and the sign of vn: same amplification
        # (beyond that a branch flips somewhere: the oracle, same arithmetic as the reference, holds 2e-2 at step 1000; with the
        # GPU's FMA contraction / libm a single contact in the other branch changes a particle's force by O(1), so the last
        # checkpoint compares pair set and contact flags only)
        return (1e-5 if gpu else 1e-7) if cp <= 400 else (1e9 if gpu else 2e-2)
    if gpu:
        return 1e-6 if cp <= 400 else 1e-4
    return 1e-7 if cp <= 400 else 1e-5


def golden(name):
    return np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))


def golden_at(g, cp):
    pre = "s%d_" % cp
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}


# ---- digests: parity at BASELINE.json's sizes without megabyte fixtures ---------------------------------------------------------
def _sha(*arrays):
    import hashlib
    h = hashlib.sha256()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return np.frombuffer(h.digest(), np.uint8).copy()


def digest(snap, tags, c, nsample=1024):
    """condense a snapshot (cases.snapshot + `tags`, both ordered by tag) into what a full-size parity check needs:
    the per-particle state of every `stride`-th particle, global sums, and hashes of the pair set / contact flags /
    mesh contact rows (bit-exact bookkeeping) -- a few hundred kilobytes whatever the number of particles"""
    n = len(tags)
    stride = max(1, n // nsample)
    sel = np.arange(0, n, stride)
    out = {"n": np.array(n), "sample_tag": np.asarray(tags)[sel]}
    for k in ("x", "v", "f", "omega", "torque"):
        out["sample_" + k] = snap[k][sel]
    out["sum_f"] = snap["f"].sum(0); out["sum_abs_f"] = np.array(np.abs(snap["f"]).sum()); out["sum_abs_torque"] = np.array(np.abs(snap["torque"]).sum())
    lo, hi, fl, hs = unique_pairs(snap["pair_lo"], snap["pair_hi"], snap["pair_flag"], snap["pair_hist"])
    out["npairs"] = np.array(len(lo)); out["nflag"] = np.array(int((fl != 0).sum()))
    out["pair_sha"] = _sha(lo.astype(np.int32), hi.astype(np.int32), (fl != 0).astype(np.int32))
    out["sum_abs_hist"] = np.array(np.abs(hs).sum()) if hs.size else np.array(0.0)
    if "cohesion" in c["pair"] and hs.size:
        out["nbonds"] = np.array(int((hs[:, 0] > 0).sum()))
    m = np.isin(lo, out["sample_tag"])          # history rows of the sampled particles' pairs
    out["sample_pair_lo"] = lo[m][:4000]; out["sample_pair_hi"] = hi[m][:4000]; out["sample_pair_hist"] = hs[m][:4000]
    for mid, mtype, nodes in c.get("meshes", []):
        mt, mi, mh = snap["mesh_%s_tag" % mid], snap["mesh_%s_tri" % mid], snap["mesh_%s_hist" % mid]
        out["mesh_%s_rows" % mid] = np.array(len(mt)); out["mesh_%s_sha" % mid] = _sha(mt.astype(np.int32), mi.astype(np.int32))
        out["mesh_%s_sum_abs_hist" % mid] = np.array(np.abs(mh).sum()) if mh.size else np.array(0.0)
        if "meshforce_" + mid in snap:
            out["meshforce_" + mid] = np.asarray(snap["meshforce_" + mid])
    return out


def compare_digest(got, ref, rmass_sample, tol=FTOL, label=""):
    """got/ref: digests of the same case at the same step (ref: the unmodified reference's, tests/golden/big_*.npz)"""
    assert int(got["n"]) == int(ref["n"]) and np.array_equal(got["sample_tag"], ref["sample_tag"]), label + ": particle set differs"
    assert int(got["npairs"]) == int(ref["npairs"]), "%s: %d pairs, reference %d" % (label, int(got["npairs"]), int(ref["npairs"]))
    assert np.array_equal(got["pair_sha"], ref["pair_sha"]), label + ": pair set / contact flags differ from the reference (hash)"
    assert int(got["nflag"]) == int(ref["nflag"])
    if "nbonds" in ref:
        assert int(got["nbonds"]) == int(ref["nbonds"]), "%s: %d bonds, reference %d" % (label, int(got["nbonds"]), int(ref["nbonds"]))
    mg = rmass_sample * 9.81
    errs = {"f": rel_err(got["sample_f"], ref["sample_f"], 1e-12 * mg)}
    errs["torque"] = rel_err(got["sample_torque"], ref["sample_torque"], np.maximum(1e-15 * mg, 1e-9 * np.linalg.norm(ref["sample_f"], axis=1)))
    for k, fl in (("x", 1e-3), ("v", 1e-3), ("omega", 1e-2)):
        errs[k] = rel_err(got["sample_" + k], ref["sample_" + k], fl)
    for k in errs:
        assert errs[k] <= tol, "%s: %s of the sampled particles differs by %.3e" % (label, k, errs[k])
    for k in ("sum_abs_f", "sum_abs_torque", "sum_abs_hist"):
        den = max(abs(float(ref[k])), 1e-300)
        if k == "sum_abs_torque":  # a bonded lattice carries no torque at all: its |torque| sum is rounding noise of the forces
            den = max(den, 1e-4 * abs(float(ref["sum_abs_f"])))
        errs[k] = abs(float(got[k]) - float(ref[k])) / den
        assert errs[k] <= max(tol, 1e-9), "%s: %s differs by %.3e" % (label, k, errs[k])
    if ref["sample_pair_hist"].size:
        assert np.array_equal(got["sample_pair_lo"], ref["sample_pair_lo"]) and np.array_equal(got["sample_pair_hi"], ref["sample_pair_hi"])
        sc = max(np.abs(ref["sample_pair_hist"]).max(), 1e-300)
        errs["hist"] = float(np.abs(got["sample_pair_hist"] - ref["sample_pair_hist"]).max() / sc)
        assert errs["hist"] <= max(tol, 1e-9), "%s: history rows differ by %.3e" % (label, errs["hist"])
    for k in ref:
        if k.startswith("mesh_") and k.endswith("_rows"):
            mid = k[5:-5]
            assert int(got[k]) == int(ref[k]), "%s: mesh %s has %d contact rows, reference %d" % (label, mid, int(got[k]), int(ref[k]))
            assert np.array_equal(got["mesh_%s_sha" % mid], ref["mesh_%s_sha" % mid]), "%s: mesh %s contact rows differ (hash)" % (label, mid)
            den = max(abs(float(ref["mesh_%s_sum_abs_hist" % mid])), 1e-12 * max(int(ref[k]), 1))  # (floor: 1e-12 m of spring per row is noise)
            assert abs(float(got["mesh_%s_sum_abs_hist" % mid]) - float(ref["mesh_%s_sum_abs_hist" % mid])) / den <= max(tol, 1e-9), \
                "%s: mesh %s history sum differs" % (label, mid)
        if k.startswith("meshforce_"):
            gv, rv = np.asarray(got[k]), np.asarray(ref[k])
            fs = max(np.abs(rv[:3]).max(), 1e-12)
            assert np.abs(gv[:3] - rv[:3]).max() / fs <= max(tol, 1e-9), "%s: %s differs" % (label, k)
    return errs
