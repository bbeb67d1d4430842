import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, "liggghts-inl_b200")
import numpy as np, cases, dem_b200
name = "box_hertz_cdt"
g = np.load("tests/golden/contacts_%s.npz" % name)
c = cases.make_case(name)
e = cases.apply(c, dem_b200.Engine(device=0))
e.setup(); e.run(int(g["steps"])); e.setup()
ct = e.contacts()
np.set_printoptions(precision=6, linewidth=200)
for r in range(4):
    i1, i2 = g["id1"][r], g["id2"][r]
    q = np.where((ct["lo"] == min(i1, i2)) & (ct["hi"] == max(i1, i2)))[0][0]
    print("ref", i1, i2, g["force"][r], g["torque"][r])
    print("got", ct["lo"][q], ct["hi"][q], ct["force_lo"][q], ct["torque_lo"][q], ct["torque_hi"][q])
