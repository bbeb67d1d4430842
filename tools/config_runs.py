#!/usr/bin/env python
"""Throughput of the other BASELINE.json configs at their named sizes through the C ABI (supplementary to bench.py, which
measures configs[1]); the beds are those of tests/configs.py, i.e. the ones whose first steps are checked against digests
of the unmodified reference (tests/test_gpu_configs.py):
  C1  10,648 monodisperse spheres settling in a box of primitive planes (hertz/history/cdt)
  C3  1,013,189 spheres in a conical hopper STL (16,384 triangles), outlet open: discharge (~50 k particle-triangle rows)
  C4  499,200 bonded spheres (INL bond/nonlinear, 2.45 M bonds created at step 2) compressed by a moving stress plate
  C5  2,067,792 spheres in a rotating drum (1,024 triangles): one GPU's share of the 16.8 M-sphere drum of configs[4]
Each prints one JSON line: particle-steps/s over the timed window (rebuilds included), list/contact statistics, rebuilds,
mesh contact rows and sanity checks (no particle lost, finite state).  usage (GPU box): python tools/config_runs.py [C1 C3 C4 C5]"""
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
    import configs
    c = configs.CONFIGS[name]()
    warm, steps = {"C1": (200, 2000), "C3": (200, 1000), "C4": (20, 300), "C5": (200, 1000)}[name]
    return c, warm, steps


def multi_c5():
    """configs[4] at its full size on all ranks of a torchrun launch: 16.65 M spheres in the rotating drum, slabs along the drum
    axis (one per GPU); every rank generates and uploads the spheres of its own sub-box"""
    import torch
    import torch.distributed as dist
    import configs
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    box = configs.C5(bricks=1, yrange=(0.0, 0.0))
    lay = dem_b200.brick_layout(world, rank, box["lo"], box["hi"], box["periodic"], procgrid=[1, world, 1])
    c = configs.C5(bricks=1, yrange=(lay["sublo"][1], lay["subhi"][1]))
    counts = [None] * world
    dist.all_gather_object(counts, len(c["tag"]))
    c["tag"] = (c["tag"].astype(np.int64) + sum(counts[:rank])).astype(np.int32)
    n = sum(counts)
    buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf.copy_(torch.frombuffer(bytearray(dem_b200.Engine.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(buf, 0)
    eng = dem_b200.Engine(device=local, rank=rank, nranks=world, nccl_id=buf.cpu().numpy().tobytes())
    eng.box(c["lo"], c["hi"], c["periodic"]); eng.processors(1, world, 1)
    eng = cases.apply(c, eng)
    eng.option("time_kernels", 1)
    warm, steps = 100, 1000
    eng.setup(); eng.run(warm)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(steps); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    st = eng.stats()
    x = eng.download("x")
    agg = torch.tensor([float(st.nlocal), float(len(eng.mesh_contacts("drum")["tag"])), float(np.isfinite(x).all()), float(st.npairs_full), float(st.ncontacts_full)], dtype=torch.float64, device="cuda")
    dist.all_reduce(agg, op=dist.ReduceOp.SUM)
    if rank == 0:
        print(json.dumps({"config": "C5 (full, %d GPUs)" % world, "particles": n, "steps": steps, "particle_steps_per_s": n * steps / (ms * 1e-3), "ms_per_step": ms / steps,
                          "step_kernel_ms_rank0": st.step_kernel_ms / max(st.step_kernel_calls, 1), "rebuilds": int(st.nbuilds), "particles_after": int(agg[0].item()),
                          "mesh_contact_rows": int(agg[1].item()), "state_ok": bool(agg[2].item() == world and int(agg[0].item()) == n),
                          "halflist_per_particle": agg[3].item() / 2.0 / n, "contacts_per_particle": agg[4].item() / 2.0 / n, "triangles": len(c["meshes"][0][2])}), flush=True)
    eng.close()
    dist.barrier(); dist.destroy_process_group()


def main():
    import torch
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return multi_c5()
    for name in (sys.argv[1:] or ["C1", "C3", "C4", "C5"]):
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
        for mid in c.get("mesh_stress", []):
            out["force_on_" + mid] = [float(v) for v in eng.mesh_force(mid)[:3]]
        print(json.dumps(out), flush=True)
        eng.close()


if __name__ == "__main__":
    main()
