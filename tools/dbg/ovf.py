import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, "liggghts-inl_b200")
import numpy as np, cases, dem_b200
c = cases.case_box(n3=(4, 4, 3), name="ovf", seed=13)
rs, R, nshell = 0.001, 0.004, 40
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
e = cases.apply(c, dem_b200.Engine(device=0))
e.setup()
for s in range(12):
    try:
        e.run(1)
    except dem_b200.DemError as ex:
        print("step", s + 1, "ERR", ex)
    p = e.pairs(); st = e.stats()
    x = e.download("x"); d = np.linalg.norm(x[1:] - x[0], axis=1) - (R + rs)
    print("step", s + 1, "pairs", len(p["lo"]), "flagged", int((p["flag"] != 0).sum()), "builds", st.nbuilds, "min gap %.3e max gap %.3e" % (d.min(), d.max()))
