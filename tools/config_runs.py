#!/usr/bin/env python
"""Scale runs of the other BASELINE.json configs through the C ABI (supplementary to bench.py, which measures configs[1]):
  C1  10,648 monodisperse spheres settling in a box of primitive planes (hertz/history/cdt)
  C3  1,000,000 spheres falling into triangle-mesh geometry (box of 2 x 40 x 40 floor triangles + walls, 128-segment
      funnel = 256 triangles), hertz/history/cdt on `fix wall/gran ... mesh`
  C4  499,200 bonded spheres (INL bond/nonlinear, bonds created at step 2), hertz/history
  C5brick  2,252,800 spheres in a rotating 256-segment drum (`fix move/mesh rotate`): one GPU's share of configs[4]
Each prints one JSON line: particle-steps/s over the timed window, list/contact statistics, rebuilds, and sanity checks
(no particle lost, finite state).  usage (GPU box): python tools/config_runs.py [C1 C3 C4]"""
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "liggghts-inl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import dem_b200  # noqa: E402


def make(name):
    if name == "C1":
        return cases.case_box(n3=(22, 22, 22), name="C1"), 20, 2000
    if name == "C3":
        c = cases.case_mesh(kind="funnel", n3=(100, 100, 100), name="C3", poly=True)
        L = c["hi"][0] / 1.25
        H = c["hi"][2]
        c["meshes"] = [("cad", 1, cases.mesh_box(L, 0.9 * H, nf=40)), ("fun", 1, cases.mesh_funnel(L, 0.55 * L, 0.2 * L, 0.62 * L, 0.12 * L, nseg=128))]
        return c, 20, 300
    if name == "C4":
        return cases.case_box(n3=(80, 80, 78), model="model hertz tangential history", poly=True, name="C4", bond=dict(kind="bond/nonlinear")), 10, 200
    if name == "C5brick":  # one GPU's share of the 16M-sphere drum of configs[4]: 2.25M spheres in a rotating 256-segment drum
        c = cases.case_mesh(kind="drum", n3=(160, 160, 88), name="C5brick", poly=True, move=2.0)
        L = 160 * 2.05 * 0.003
        c["x"][:, 2] += 0.62 * L - 0.5 * (c["x"][:, 2].min() + c["x"][:, 2].max())  # bed centred on the drum axis
        c["meshes"] = [("drum", 1, cases.mesh_drum(0.5 * L, 0.62 * L, 0.62 * L, -0.12 * L, 1.12 * L, nseg=256))]
        return c, 20, 300
    raise SystemExit("unknown config " + name)


def main():
    import torch
    for name in (sys.argv[1:] or ["C1", "C3", "C4"]):
        c, warm, steps = make(name)
        n = len(c["tag"])
        eng = cases.apply(c, dem_b200.Engine(device=0))
        if name == "C4":
            eng.option("maxneigh", 40)
        eng.option("time_kernels", 1)
        eng.setup(); eng.run(warm)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.run(steps)
        st = eng.stats()  # synchronises
        dt = time.perf_counter() - t0
        x = eng.download("x"); v = eng.download("v"); tag = eng.download("tag")
        ok = bool(np.isfinite(x).all() and np.isfinite(v).all() and np.array_equal(tag, np.sort(c["tag"])))
        out = {"config": name, "particles": n, "steps": steps, "particle_steps_per_s": n * steps / dt, "ms_per_step": 1e3 * dt / steps,
               "step_kernel_ms": st.step_kernel_ms / max(st.step_kernel_calls, 1), "rebuilds": int(st.nbuilds),
               "halflist_per_particle": st.npairs_full / 2.0 / n, "contacts_per_particle": st.ncontacts_full / 2.0 / n,
               "pair_style": c["pair"], "triangles": int(sum(len(m[2]) for m in c.get("meshes", []))), "state_ok": ok}
        if c.get("meshes"):
            out["mesh_contact_rows"] = int(sum(len(eng.mesh_contacts(m[0])["tag"]) for m in c["meshes"]))
        if "cohesion" in c["pair"]:
            p = eng.pairs()
            out["bonds"] = int((p["hist"][:, 0] > 0).sum())
        print(json.dumps(out), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
