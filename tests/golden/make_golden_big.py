"""tests/golden/make_golden_big.py -- digests of the UNMODIFIED reference (oracle/_ref) on BASELINE.json's configs at their
named sizes (tests/configs.py): tests/golden/big_<config>.npz.  A digest (tests/parity.py: digest) holds the state of
every ~250th particle, global sums and SHA-256 hashes of the pair set, the contact flags and the mesh contact rows, so a
million-particle parity check needs a few hundred kilobytes of fixture.  Build container only (minutes per config):
    make -C oracle ref && python tests/golden/make_golden_big.py [C1 C3 C4 C5]"""
import os
import sys
import tempfile
import time
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import cases  # noqa: E402
import configs  # noqa: E402
import parity  # noqa: E402
import ref_driver  # noqa: E402


def ref_snapshot(r, c):
    a = r.atoms()
    out = {k: a[k] for k in ("x", "v", "f", "omega", "torque")}
    p = r.pairs()
    if "cohesion" in c["pair"]:
        p["hist"][:, 2:5] = 0.0
    out.update(pair_lo=p["lo"], pair_hi=p["hi"], pair_flag=(p["flag"] != 0).astype(np.int32), pair_hist=p["hist"])
    for mid, mtype, nodes in c.get("meshes", []):
        m = r.mesh_contacts(mid)
        out["mesh_%s_tag" % mid] = m["tag"]; out["mesh_%s_tri" % mid] = m["tri"]; out["mesh_%s_hist" % mid] = m["hist"]
    for mid in c.get("mesh_stress", []):
        out["meshforce_%s" % mid] = r.fix_vector(mid, 9)
    return out, a["tag"], a["rmass"]


def run(name, scale=1.0, out_name=None):
    t0 = time.time()
    c = configs.CONFIGS[name](scale)
    cps = configs.BIG_CHECKPOINTS[name]
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "case.data"))
    open(os.path.join(tmp, "case.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    out = {"scale": np.array(scale)}
    done = 0
    for cp in cps:
        r.cmd("run %d" % (cp - done) if cp else "run 0")
        done = cp
        snap, tags, rmass = ref_snapshot(r, c)
        d = parity.digest(snap, tags, c)
        for k, v in d.items():
            out["s%d_%s" % (cp, k)] = v
        out["s%d_nbuilds" % cp] = np.array(r.neigh_builds)
        stride = max(1, len(tags) // 1024)
        out["sample_rmass"] = rmass[::stride]
        print(name, "step", cp, "pairs", int(d["npairs"]), "flagged", int(d["nflag"]), "bonds", int(d.get("nbonds", 0)),
              "mesh rows", {k: int(v) for k, v in d.items() if k.endswith("_rows")}, "builds", r.neigh_builds, "%.0f s" % (time.time() - t0), flush=True)
    r.close()
    np.savez_compressed(os.path.join(HERE, (out_name or "big_" + name) + ".npz"), **out)


if __name__ == "__main__":
    args = sys.argv[1:]
    if len(args) >= 1 and args[0] == "--one":
        run(args[1], float(args[2]) if len(args) > 2 else 1.0, args[3] if len(args) > 3 else None)
    else:
        import subprocess
        for name in (args or list(configs.CONFIGS)):   # one process per case: the reference keeps global registries
            subprocess.run([sys.executable, os.path.abspath(__file__), "--one", name], check=True)
