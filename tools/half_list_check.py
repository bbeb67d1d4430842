#!/usr/bin/env python
"""The north_star's alternative, measured: half list with fp64 reductions (option "half_list": every pair evaluated once,
the partner's share sent with RED.E.ADD.F64, separate integration kernel) against the default full list without atomics,
on the bench bed.  Prints agreement after 10 steps and the step-kernel times.  usage (GPU box): python tools/half_list_check.py [tiles]"""
import json
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "liggghts-inl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import bench, cases, dem_b200  # noqa: E402

tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 16
c = bench.bed_case(tiles, tiles)
n = len(c["tag"])
out = {"particles": n}
state = {}
for name, half in (("full_list_no_atomics", 0), ("half_list_fp64_red", 1)):
    e = cases.apply(c, dem_b200.Engine(device=0))
    e.option("half_list", half); e.option("time_kernels", 1)
    e.setup(); e.run(10)
    state[name] = {k: e.download(k) for k in ("x", "v", "omega", "f", "torque")}
    e.run(200)
    st = e.stats()
    out[name + "_step_kernels_ms"] = st.step_kernel_ms / max(st.step_kernel_calls, 1)
    e.close()
a, b = state["full_list_no_atomics"], state["half_list_fp64_red"]
w = (4.0 * np.pi / 3.0 * c["radius"] ** 3 * c["density"] * 9.81)[:, None]
out["max_force_difference_over_weight"] = float((np.abs(a["f"] - b["f"]) / w).max())
out["max_position_difference_m"] = float(np.abs(a["x"] - b["x"]).max())
out["max_velocity_difference"] = float(np.abs(a["v"] - b["v"]).max())
print(json.dumps(out))
