#!/usr/bin/env python
"""bench.py -- particle-steps/s of the DEM hot path on BASELINE.json configs[1]:
4 194 304 polydisperse spheres, Hertz/history + rolling friction (cdt), settled bed on a floor,
periodic in x,y (a 16 384-sphere tile settled by the reference, bench_data/, replicated 16x16).

  python bench.py --gpus 1 --steps K --warmup W            -> own arm (CUDA engine through the C ABI)
  python bench.py --impl reference --gpus 1 --steps K ...  -> reference arm: the UNMODIFIED reference
         (oracle/_ref) on all host cores, one independent replica of the tile per core (the image
         has no MPI; SURVEY.md 8d "P-replica proxy")
One JSON line on stdout (rank 0)."""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "liggghts-inl_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

try:  # the metric name is BASELINE.json's, verbatim
    METRIC = json.load(open(os.path.join(ROOT, "BASELINE.json")))["metric"]
except Exception:
    METRIC = "particle-steps/s (Hertz/history, 4M spheres) at 1/2/4/8 B200; % HBM roofline"
MODEL = "model hertz tangential history rolling_friction cdt"
PROPS = [("youngsModulus", "peratomtype", [5e6]), ("poissonsRatio", "peratomtype", [0.45]),
         ("coefficientRestitution", "peratomtypepair", [0.3]), ("coefficientFriction", "peratomtypepair", [0.5]),
         ("coefficientRollingFriction", "peratomtypepair", [0.1])]


def load_tile():
    t = np.load(os.path.join(ROOT, "bench_data", "tile16k.npz"))
    return {k: t[k] for k in t.files}


def bed_case(tiles_x, tiles_y, name="bed"):
    """the settled tile replicated tiles_x x tiles_y (the reference's `replicate` idea)"""
    t = load_tile()
    n0 = len(t["radius"])
    Lx, Ly = t["hi"][0] - t["lo"][0], t["hi"][1] - t["lo"][1]
    nt = tiles_x * tiles_y
    x = np.empty((nt * n0, 3)); k = 0
    for ix in range(tiles_x):
        for iy in range(tiles_y):
            x[k * n0:(k + 1) * n0] = t["x"] + np.array([ix * Lx, iy * Ly, 0.0]); k += 1
    rep = lambda a: np.tile(a, (nt,) + (1,) * (a.ndim - 1))
    n = nt * n0
    return dict(name=name, lo=[t["lo"][0], t["lo"][1], t["lo"][2]], hi=[t["lo"][0] + tiles_x * Lx, t["lo"][1] + tiles_y * Ly, t["hi"][2]],
                periodic=[1, 1, 0], ntypes=1, skin=0.001, dt=1e-5, props=PROPS, pair=MODEL,
                walls=[("floor", MODEL + " primitive type 1 zplane 0.0")], gravity=(9.81, [0.0, 0.0, -1.0]), freeze=0,
                tag=np.arange(1, n + 1, dtype=np.int32), type=np.ones(n, np.int32), mask=np.ones(n, np.int32), x=x,
                v=rep(t["v"]), omega=rep(t["omega"]), radius=rep(t["radius"]), density=rep(t["density"]))


# --------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate(); self.p.wait()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [s.strip() for s in line.split(",")]
            if len(c) < 7:
                continue
            try:
                sm.append(float(c[0])); mx.append(float(c[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------
def ref_worker(tiles, steps, out):
    """one reference process: `tiles x tiles` tile bed, `run steps`; prints loop seconds"""
    import cases
    import ref_driver
    c = bed_case(tiles, tiles, name="cpu")
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "bed.data"))
    open(os.path.join(tmp, "bed.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    r.cmd("run 0")
    t0 = time.perf_counter()
    r.cmd("run %d" % steps)
    dt = time.perf_counter() - t0
    json.dump({"seconds": dt, "n": len(c["tag"]), "steps": steps}, open(out, "w"))


def run_reference_procs(nproc, tiles, steps):
    outs, procs = [], []
    for k in range(nproc):
        out = tempfile.mktemp(suffix=".json")
        outs.append(out)
        procs.append(subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-worker", str(tiles), str(steps), out],
                                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL))
    for p in procs:
        p.wait()
    res = [json.load(open(o)) for o in outs if os.path.exists(o)]
    for o in outs:
        if os.path.exists(o):
            os.unlink(o)
    if not res:
        return None
    total = sum(r["n"] * r["steps"] for r in res)
    return total / max(r["seconds"] for r in res), len(res), res[0]["n"]


def ref_server(tiles):
    """persistent reference worker: loads the tile bed once, then serves `run S` requests on stdin"""
    import cases
    import ref_driver
    c = bed_case(tiles, tiles, name="cpu")
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "bed.data"))
    open(os.path.join(tmp, "bed.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    r.cmd("run 0")
    sys.stdout.write("ready %d\n" % len(c["tag"])); sys.stdout.flush()
    for line in sys.stdin:
        line = line.strip()
        if not line or line == "quit":
            break
        t0 = time.perf_counter()
        r.cmd("run %d pre no post no" % int(line))  # steady-state stepping: no Verlet::setup per request
        sys.stdout.write("done %.6f\n" % (time.perf_counter() - t0)); sys.stdout.flush()


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import ref_driver
    cfg = {"workload": "4,194,304-sphere settled polydisperse bed (16x16 replicas of bench_data/tile16k), hertz/history/cdt, floor + periodic xy",
           "inputs": "reference arm: every host core steps one 16,384-sphere tile of the same bed (serial reference build; the image has no MPI)"}
    if not ref_driver.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libliggghts_ref.so missing (build: make -C oracle ref)"}))
        return
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    sample_steps = 5  # reference timesteps per bench step and per core
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-server", "1"], stdin=subprocess.PIPE,
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1) for _ in range(cores)]
    n_tile = 0
    for p in procs:
        line = p.stdout.readline()
        while line and not line.startswith("ready"):
            line = p.stdout.readline()
        if not line:
            print(json.dumps({"impl": "reference", "unavailable": "reference worker failed to start"})); return
        n_tile = int(line.split()[1])

    def step():
        for p in procs:
            p.stdin.write("%d\n" % sample_steps); p.stdin.flush()
        for p in procs:
            line = p.stdout.readline()
            while line and not line.startswith("done"):
                line = p.stdout.readline()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    for p in procs:
        try:
            p.stdin.write("quit\n"); p.stdin.flush()
        except Exception:
            pass
    for p in procs:
        p.wait()
    value = cores * n_tile * sample_steps * args.steps / dt
    print(json.dumps({"impl": "reference", "metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": args.gpus,
                      "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
                      "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                      "config": cfg,
                      "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "reference",
                                       "sample": "%d concurrent serial reference processes, each one %d-sphere tile x %d timesteps per bench step"
                                                 % (cores, n_tile, sample_steps)},
                      "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


# --------------------------------------------------------------------------------------------
def own_arm(args):
    import torch
    import dem_b200
    import cases
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl")

    def new_engine():
        if world == 1:
            return dem_b200.Engine(device=local)
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(dem_b200.Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        return dem_b200.Engine(device=local, rank=rank, nranks=world, nccl_id=buf.cpu().numpy().tobytes())

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.MAX); return float(t.item())

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda"); dist.all_reduce(t, op=dist.ReduceOp.SUM); return float(t.item())

    tiles = args.tiles
    c = bed_case(tiles, tiles)
    n = len(c["tag"])
    eng = cases.apply(c, new_engine())
    eng.option("time_kernels", 1)
    if os.environ.get("DEM_DEBUG"):
        eng.option("debug", int(os.environ["DEM_DEBUG"]))  # profiling aid only (the numbers are not bench values)
    eng.setup()
    eng.run(max(args.warmup, 3))
    st0 = eng.stats()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(args.steps)
    e1.record(); torch.cuda.synchronize()
    if dist: dist.barrier()
    ms = allmax(e0.elapsed_time(e1))
    clocks = sampler.stop()
    st = eng.stats()
    value = n * args.steps / (ms * 1e-3)
    # roofline of the dominant (fused step) kernel
    nloc = max(int(st.nlocal), 1)
    K_half = st.npairs_full / 2.0 / nloc
    C_half = st.ncontacts_full / 2.0 / nloc
    dnum = st.dnum
    bytes_per_ps = 192.0 + 4.0 * K_half + (16.0 * dnum + 8.0) * C_half
    kms = st.step_kernel_ms / max(st.step_kernel_calls, 1)
    peak, peak_src = peaks()
    achieved = bytes_per_ps * nloc / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_bytes_per_launch.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("k_step_bytes_per_launch")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_step<hertz,cdt> (fused pair+wall+gravity+integrate step)", "kernel_ms": kms, "peak_source": peak_src,
                "bytes_per_particle_step": bytes_per_ps, "halflist_per_particle": K_half, "contacts_per_particle": C_half,
                "kernel_share_of_step": kms * st.step_kernel_calls / ms if ms > 0 else None}
    launches = st.kernel_launches - st0.kernel_launches
    nbuilds = st.nbuilds
    # upper bound of the cost of one neighbour rebuild, reported next to the step time: wall time of setup() on a running engine
    # (a full rebuild -- sort, gather, cell ranges, halo lists, list build + history remap; its stages sum to 3.2 ms for 4.19M
    # spheres under DEM_B200_TRACE -- plus Verlet::setup's force evaluation with force read-out and the host synchronisation)
    eng.setup()  # (the first rebuild after the initial one allocates the second list set: not steady state)
    torch.cuda.synchronize(); tr0 = time.perf_counter()
    eng.setup()
    torch.cuda.synchronize()
    rebuild_ms = allmax((time.perf_counter() - tr0) * 1e3) - (kms if kms > 0 else 0.0)

    # end-to-end through the public API with host buffers: upload -> setup -> run(K) -> download
    ke = args.steps
    eng.close()
    # the job's host buffers are page-locked (inputs and results), as a production caller's would be
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); v = t.numpy(); v[...] = a; return v
    import numpy as np
    cp = dict(c)
    for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density"):
        cp[k] = pinned(np.ascontiguousarray(c[k], np.int32 if k in ("tag", "type", "mask") else np.float64))
    xo = pinned(np.zeros((n, 3))) if world == 1 else None
    vo = pinned(np.zeros((n, 3))) if world == 1 else None
    # one short untimed pass of the same job first (warm-up, like the W steps of the kernel-level number)
    engw = cases.apply(cp, new_engine()); engw.setup(); engw.run(3); engw.download("x", out=xo); engw.close()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    t0 = time.perf_counter()
    eng2 = cases.apply(cp, new_engine())
    t1 = time.perf_counter()
    eng2.setup()
    t2 = time.perf_counter()
    eng2.run(ke)
    xo = eng2.download("x", out=xo); vo = eng2.download("v", out=vo)
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    t_e2e = allmax(t3 - t0)
    h2d = allsum(sum(int(cp[k].nbytes) for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density")))  # every rank receives the whole set and keeps its brick
    d2h = allsum(xo.nbytes + vo.nbytes + 2 * len(xo) * 4)
    e2e = {"value": n * ke / t_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d / ke, "d2h_bytes_per_step": d2h / ke,
           "job": "create + upload(page-locked host arrays) + setup + run(%d) + download x,v; %.3f s (create+upload %.3f, setup %.3f, run+download %.3f on rank 0)"
                  % (ke, t_e2e, t1 - t0, t2 - t1, t3 - t2)}
    eng2.close()

    # CPU baseline: the unmodified reference, 1 core, one tile of the same bed
    cpu = None
    if rank == 0 and not args.no_cpu:
        import ref_driver
        if ref_driver.available():
            steps_cpu = 1500
            r = run_reference_procs(1, 1, steps_cpu)
            if r:
                cpu = {"value": r[0], "unit": "particle-steps/s", "cores": 1, "kind": "reference",
                       "sample": "one 16,384-sphere tile of the bed x %d steps, serial reference build (oracle/_ref)" % steps_cpu}
    out = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "%d-sphere settled polydisperse bed (%dx%d replicas of bench_data/tile16k, radii U[1.5,3] mm), "
                                  "hertz/history/cdt, floor + periodic xy, dt 1e-5, skin 1 mm" % (n, tiles, tiles),
                      "particles": n, "l2_policy": "inputs (%.1f GB of state + lists) exceed the 126 MB L2" % (n * 600 / 1e9),
                      "rebuilds_in_timed_region": int(nbuilds), "resetup_ms": round(rebuild_ms, 3),
                      "parallelism": "1 GPU" if world == 1 else "%d GPUs: x-slab bricks, NCCL halo per step (roofline fields are rank 0's)" % world},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu}
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.barrier(); dist.destroy_process_group()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--ref-worker":
        ref_worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); return
    if len(sys.argv) > 1 and sys.argv[1] == "--ref-server":
        ref_server(int(sys.argv[2])); return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--tiles", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
