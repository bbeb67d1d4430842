#!/usr/bin/env python
"""dev tool (build container only): run one seeded case on the UNMODIFIED reference and on the C oracle and
compare state, pair bookkeeping, mesh topology and mesh contact rows at a few checkpoints.
usage: tools/pin_case.py <kind> [steps ...]   kind in box|roof|funnel|plate (mesh cases)"""
import os, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for d in ("liggghts-inl_b200", "tests", "oracle"):
    sys.path.insert(0, os.path.join(ROOT, d))
import cases, parity, ref_driver

kind = sys.argv[1]
cps = [int(a) for a in sys.argv[2:]] or [0, 1, 2, 10, 400, 2500]
c = cases.make_case(kind) if kind in cases.GOLDEN_CASES else cases.case_mesh(kind=kind, name=kind)
tmp = tempfile.mkdtemp()
deck, data = cases.to_deck(c, os.path.join(tmp, "case.data"))
open(os.path.join(tmp, "case.data"), "w").write(data)
r = ref_driver.Ref(log=os.path.join(tmp, "log.liggghts"))
r.cmd(deck)
o = cases.apply(c, parity.oracle_engine())
rmass = 4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"]
done = 0
for cp in cps:
    r.cmd("run %d" % (cp - done))
    o.setup(); o.run(cp - done); done = cp
    if cp == cps[0]:
        for mid, mt, nodes in c.get("meshes", []):
            tr = r.mesh_topology(mid); n = len(nodes)
            assert len(tr["nodes"]) == n, "reference kept %d of %d triangles" % (len(tr["nodes"]), n)
            for f in ("edge_active", "corner_active", "nneighs"):
                got = o.mesh_field(mid, f, n)
                bad = np.argwhere(got != tr[f])
                print("  topology %s/%s: %s" % (mid, f, "OK" if not len(bad) else "MISMATCH at %s (oracle %s ref %s)" % (bad[:6].tolist(), got[tuple(bad[0])], tr[f][tuple(bad[0])])))
    a = r.atoms(); p = r.pairs()
    ref = dict(x=a["x"], v=a["v"], f=a["f"], omega=a["omega"], torque=a["torque"], pair_lo=p["lo"], pair_hi=p["hi"],
               pair_flag=(p["flag"] != 0).astype(np.int32), pair_hist=p["hist"])
    if "cohesion" in c["pair"]:
        ref["pair_hist"][:, 2:5] = 0.0
        msg_b = "bonds ref %d orc %d" % (int((ref["pair_hist"][:, 0] > 0).sum()), int((cases.snapshot(o, c)["pair_hist"][:, 0] > 0).sum()))
    else:
        msg_b = ""
    got = cases.snapshot(o, c)
    msg = []
    try:
        errs = parity.compare_snapshot(got, ref, rmass, tol=1e-9 if cp <= 10 else 1e-5, label="%s@%d" % (kind, cp))
        msg.append("state OK f=%.1e x=%.1e" % (errs["f"], errs["x"]))
    except AssertionError as ex:
        msg.append("STATE FAIL " + str(ex))
    for mid, mt, nodes in c.get("meshes", []):
        m = r.mesh_contacts(mid)
        gt, gi, gh = got["mesh_%s_tag" % mid], got["mesh_%s_tri" % mid], got["mesh_%s_hist" % mid]
        same = len(gt) == len(m["tag"]) and np.array_equal(gt, m["tag"]) and np.array_equal(gi, m["tri"])
        herr = float(np.abs(gh - m["hist"]).max() / max(np.abs(m["hist"]).max(), 1e-300)) if same and gh.size else 0.0
        msg.append("mesh %s rows %d/%d %s hist %.1e" % (mid, len(gt), len(m["tag"]), "OK" if same else "ROWS DIFFER", herr))
    print("step %5d builds ref %d orc %d | %s %s" % (cp, r.neigh_builds, o.stats().nbuilds, " | ".join(msg), msg_b))
print("log:", os.path.join(tmp, "log.liggghts"))
