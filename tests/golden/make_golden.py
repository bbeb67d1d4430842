"""tests/golden/make_golden.py -- regenerates tests/golden/*.npz by running the UNMODIFIED
reference (oracle/_ref/libliggghts_ref.so, built from /root/reference by oracle/Makefile.ref)
on the seeded cases of tests/cases.py.  Run in the build container only:
    make -C oracle ref && python tests/golden/make_golden.py
The .npz files are committed; the GPU box never needs /root/reference."""
import os
import sys
import tempfile
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle"))
import cases  # noqa: E402
import ref_driver  # noqa: E402


def run_case(name):
    c = cases.make_case(name)
    cps = cases.GOLDEN_CASES[name]["checkpoints"]
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "case.data"))
    open(os.path.join(tmp, "case.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    out = {}
    done = 0
    for cp in cps:
        for line in cases.late_commands(c, cp):
            r.cmd(line)
        r.cmd("run %d" % (cp - done) if cp else "run 0")
        if cp and done == 0 and cps[0] != 0:
            pass
        done = cp
        a = r.atoms()
        for k in ("x", "v", "f", "omega", "torque"):
            out["s%d_%s" % (cp, k)] = a[k]
        p = r.pairs()
        out["s%d_pair_lo" % cp] = p["lo"]; out["s%d_pair_hi" % cp] = p["hi"]
        if "cohesion" in c["pair"]:
            p["hist"][:, 2:5] = 0.0  # contactPos: written at bond creation, read for wall bonds only -> not part of the parity contract
        out["s%d_pair_flag" % cp] = (p["flag"] != 0).astype(np.int32); out["s%d_pair_hist" % cp] = p["hist"]
        for wid, text in c["walls"]:
            dn = 3 + (3 if "epsd" in text else 0)
            out["s%d_wall_%s" % (cp, wid)] = r.fix_peratom_array("history_" + wid, dn)
        for mid, mtype, nodes in c.get("meshes", []):
            m = r.mesh_contacts(mid)
            out["s%d_mesh_%s_tag" % (cp, mid)] = m["tag"]; out["s%d_mesh_%s_tri" % (cp, mid)] = m["tri"]; out["s%d_mesh_%s_hist" % (cp, mid)] = m["hist"]
        for mid in c.get("mesh_stress", []):
            out["s%d_meshforce_%s" % (cp, mid)] = r.fix_vector(mid, 9)
        if "cohesion bond " in c["pair"] + " " and cp > 0:  # (at `run 0` the compute has not been invoked: its vector is uninitialised)
            out["s%d_bondcounter" % cp] = r.compute_vector("bc", 6)
        out["s%d_nbuilds" % cp] = np.array(r.neigh_builds)
    for mid, mtype, nodes in c.get("meshes", []):
        t = r.mesh_topology(mid)
        assert len(t["nodes"]) == len(nodes), "the reference dropped triangles of mesh " + mid
        for k in ("edge_active", "corner_active", "nneighs"):
            out["topo_%s_%s" % (mid, k)] = t[k]
    out["rmass"] = r.atoms()["rmass"]
    r.close()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "npairs", len(out["s%d_pair_lo" % cps[-1]]), "contacts", int(out["s%d_pair_flag" % cps[-1]].sum()),
          "builds", int(out["s%d_nbuilds" % cps[-1]]))


if __name__ == "__main__":
    # one process per case: the reference keeps global registries, a second LAMMPS
    # instance in the same process crashes (SURVEY.md 8b "Threading")
    if len(sys.argv) == 2:
        run_case(sys.argv[1])
    else:
        import subprocess
        for name in cases.GOLDEN_CASES:
            subprocess.run([sys.executable, os.path.abspath(__file__), name], check=True)
