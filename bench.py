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

    def wait_first(self, timeout=5.0):
        """block until nvidia-smi has printed its first sample: its start-up (NVML initialisation over all GPUs of the box,
        tens to hundreds of milliseconds during which CUDA calls of running processes stall) must not fall into a timed region"""
        if not self.p:
            return
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.f.name) > 0:
                    return
            except OSError:
                return
            time.sleep(0.02)

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


def ref_state_worker(steps, out):
    """parity leg: the unmodified reference steps ONE periodic tile of the bed `steps` timesteps from the bench's initial state
    and writes x, v by tag -- the checker of the `parity` field, never timed"""
    import cases
    import ref_driver
    c = bed_case(1, 1, name="cpu")
    tmp = tempfile.mkdtemp()
    deck, data = cases.to_deck(c, os.path.join(tmp, "bed.data"))
    open(os.path.join(tmp, "bed.data"), "w").write(data)
    r = ref_driver.Ref()
    r.cmd(deck)
    r.cmd("run %d" % steps)
    a = r.atoms()
    np.savez(out, x=a["x"], v=a["v"], omega=a["omega"])


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
    # clocks / throttle reasons are sampled over the whole run of rank 0's GPU; nvidia-smi is started FIRST and its first
    # sample awaited, so that its start-up (which stalls CUDA calls of running processes) cannot fall into a timed region
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("DEM_BENCH_NO_SAMPLER")) else None
    have_comm = [False]

    def with_opts(e):  # developer aid: DEM_OPTS="chunk=65536,wave_skew=8" sets engine options on every engine of the run
        for kv in os.environ.get("DEM_OPTS", "").split(","):
            if "=" in kv:
                e.option(kv.split("=")[0].strip(), float(kv.split("=")[1]))
        return e

    def new_engine():
        return with_opts(new_engine_())

    def new_engine_():
        if world == 1:
            return dem_b200.Engine(device=local)
        if have_comm[0]:  # explicit reuse of the communicator the previous engine of this process released (dem_b200.h: dem_create)
            return dem_b200.Engine(device=local, rank=rank, nranks=world, nccl_id=None)
        buf = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf.copy_(torch.frombuffer(bytearray(dem_b200.Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(buf, 0)
        have_comm[0] = True
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
    if world > 1:
        # every rank holds (and uploads) its own brick of the bed, like a rank of the reference after read_data: the
        # engine's own decomposition code decides the ownership (dem_brick_layout == dem_upload_particles' test)
        lay = dem_b200.brick_layout(world, rank, c["lo"], c["hi"], c["periodic"], x=c["x"])
        mine = lay["mine"] != 0
        for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density"):
            c[k] = np.ascontiguousarray(c[k][mine])
    n_mine = len(c["tag"])
    eng = cases.apply(c, new_engine())
    eng.option("time_kernels", 1)
    if os.environ.get("DEM_DEBUG"):
        eng.option("debug", int(os.environ["DEM_DEBUG"]))  # profiling aid only (the numbers are not bench values)
    eng.setup()
    if sampler:
        sampler.wait_first()
    eng.run(max(args.warmup, 3))
    st0 = eng.stats()
    torch.cuda.synchronize()
    if dist: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    eng.run(args.steps)
    e1.record(); torch.cuda.synchronize()
    if dist: dist.barrier()
    ms = allmax(e0.elapsed_time(e1))
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

    # ---- neighbour rebuilds ("with and without amortised rebuilds" + a falling window) ------------------
    # (a) cost of ONE rebuild in the steady state of this bed: the same window again with forced rebuilds (`neigh_modify
    #     every R check no`); the difference to the rebuild-free window divided by the number of rebuilds
    R = max(2, min(10, args.steps // 2))
    nsteps_r = (args.steps // R) * R
    eng.neighbor(c["skin"], every=R, delay=0, check=False)
    eng.run(R)  # (the first rebuild after the initial one allocates the second list set: not steady state)
    torch.cuda.synchronize()
    if dist: dist.barrier()
    b0 = eng.stats().nbuilds
    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    r0.record(); eng.run(nsteps_r); r1.record(); torch.cuda.synchronize()
    ms_forced = allmax(r0.elapsed_time(r1))
    nb_forced = nsteps_r // R
    rebuild_ms = max(ms_forced - nsteps_r * ms / args.steps, 0.0) / max(nb_forced, 1)
    eng.neighbor(c["skin"], every=1, delay=0, check=True)
    eng.close()
    # (b) a FALLING window: the same bed stretched by 15 % in z (no new overlaps: distances only grow) and dropped; the bed
    #     collapses back onto itself, contacts re-form and the distance check trips rebuilds at its own pace.  Its
    #     particle-steps/s includes every rebuild -- the honest throughput of an unsettled bed.
    falling = None
    if not args.no_falling:
        cf = dict(c); cf["x"] = c["x"] * np.array([1.0, 1.0, 1.15]); cf["hi"] = [c["hi"][0], c["hi"][1], c["hi"][2] * 1.15 + 0.01]
        engf = cases.apply(cf, new_engine())
        engf.setup(); engf.run(20)
        torch.cuda.synchronize()
        if dist: dist.barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fsteps = max(args.falling_steps, 1)
        f0.record(); engf.run(fsteps); f1.record(); torch.cuda.synchronize()
        ms_f = allmax(f0.elapsed_time(f1))
        sf = engf.stats()
        falling = {"value": n * fsteps / (ms_f * 1e-3), "unit": "particle-steps/s", "steps": fsteps, "rebuilds": int(sf.nbuilds),
                   "ms_per_step": ms_f / fsteps, "workload": "the same bed stretched 1.15x in z and dropped (contacts re-form; rebuilds included)"}
        engf.close()
    amortised = None
    if falling and falling["rebuilds"] > 0:
        interval = falling["steps"] / falling["rebuilds"]
        amortised = {"value": n / ((ms / args.steps + rebuild_ms / interval) * 1e-3), "unit": "particle-steps/s",
                     "rebuild_interval_steps": interval, "note": "settled-bed step time + rebuild_ms spread over the falling window's measured rebuild interval"}

    # ---- end to end through the public API with host buffers: create -> upload -> setup -> run(K) -> download -----------
    ke = args.steps
    # the job's host buffers are page-locked (inputs and results), as a production caller's would be
    def pinned(a):
        t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); v = t.numpy(); v[...] = a; return v
    cp = dict(c)
    for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density"):
        cp[k] = pinned(np.ascontiguousarray(c[k], np.int32 if k in ("tag", "type", "mask") else np.float64))
    xo = pinned(np.zeros((n_mine, 3))); vo = pinned(np.zeros((n_mine, 3)))
    # one short untimed pass of the same job first (warm-up, like the W steps of the kernel-level number): the timed job
    # therefore runs in a WARM process -- device blocks, page-locked flag words and (multi-GPU) the NCCL communicator are
    # recycled from the engine before it, as in a service that runs job after job
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
    h2d = allsum(sum(int(cp[k].nbytes) for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density")))
    d2h = allsum(xo.nbytes + vo.nbytes)
    e2e = {"value": n * ke / t_e2e, "unit": "particle-steps/s", "h2d_bytes_per_step": h2d / ke, "d2h_bytes_per_step": d2h / ke,
           "job": "create + upload(page-locked host arrays%s) + setup + run(%d) + download x,v; %.3f s (create+upload %.3f, setup %.3f, run+download %.3f on rank 0); warm process (see bench.py)"
                  % ("" if world == 1 else ", every rank its own brick", ke, t_e2e, t1 - t0, t2 - t1, t3 - t2)}
    tags_mine = eng2.download("tag")
    eng2.close()
    # parity is checked on a short window: DEM trajectories are chaotic, after hundreds of steps rounding-level differences
    # have grown beyond any meaningful bound.  A long e2e job is therefore followed by a dedicated 20-step job.
    kp = ke
    if ke > 50:
        kp = 20
        engp = cases.apply(cp, new_engine()); engp.setup(); engp.run(kp)
        xo = engp.download("x", out=xo); vo = engp.download("v", out=vo); tags_mine = engp.download("tag")
        engp.close()

    # ---- parity inside the bench (all N): the e2e job's result against the UNMODIFIED reference stepping ONE periodic
    # tile of the bed for the same K steps from the same initial state (the bed is tiles x tiles replicas of that tile, so
    # every replica on every rank must reproduce it), plus the conservation of the particle set across ranks.  The
    # reference here is the checker, not the thing measured.
    parity = {"checked": False}
    try:
        ref = None
        if rank == 0:
            import ref_driver
            if ref_driver.available():
                out = tempfile.mktemp(suffix=".npz")
                subprocess.run([sys.executable, os.path.abspath(__file__), "--ref-state", str(kp), out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=600)
                if os.path.exists(out):
                    z = np.load(out); ref = np.concatenate([z["x"], z["v"]], axis=1); os.unlink(out)
        n0 = n // (tiles * tiles)
        if world > 1:
            flag = torch.tensor([1 if ref is not None else 0], device="cuda"); dist.broadcast(flag, 0)
            have_ref = bool(flag.item())
            rt = torch.zeros((n0, 6), dtype=torch.float64, device="cuda")
            if have_ref:
                if rank == 0: rt.copy_(torch.from_numpy(ref))
                dist.broadcast(rt, 0); ref = rt.cpu().numpy()
        else:
            have_ref = ref is not None
        tile0 = load_tile()
        Lx, Ly = tile0["hi"][0] - tile0["lo"][0], tile0["hi"][1] - tile0["lo"][1]
        t_idx = (tags_mine.astype(np.int64) - 1) // n0; k_idx = (tags_mine.astype(np.int64) - 1) % n0
        ntot = allsum(float(len(tags_mine))); tagsum = allsum(float(tags_mine.astype(np.float64).sum()))
        conserved = (int(ntot) == n) and (tagsum == n * (n + 1) / 2.0)
        if have_ref:
            off = np.stack([(t_idx // tiles) * Lx, (t_idx % tiles) * Ly, np.zeros(len(t_idx))], axis=1)
            dx = xo - off - ref[k_idx, 0:3]
            dx[:, 0] -= Lx * np.round(dx[:, 0] / Lx); dx[:, 1] -= Ly * np.round(dx[:, 1] / Ly)
            rmin = float(tile0["radius"].min())
            ex = allmax(float(np.abs(dx).max()) / rmin if len(dx) else 0.0)
            vscale = float(np.sqrt(9.81 * tile0["radius"].mean()))
            ev = allmax(float(np.abs(vo - ref[k_idx, 3:6]).max()) / vscale if len(dx) else 0.0)
            # bound: 1e-9 of the smallest radius / of sqrt(g r) after <= 50 steps (summation order, FMA contraction and the
            # replica offsets' own rounding: measured 7e-13 / 2e-11)
            tol = 1e-9
            parity = {"checked": True, "ok": bool(conserved and ex <= tol and ev <= tol), "particles_conserved": bool(conserved),
                      "max_dx_over_rmin": ex, "max_dv_over_sqrt_g_r": ev, "tol": tol, "steps": kp, "replicas_checked": tiles * tiles,
                      "against": "unmodified reference (oracle/_ref) stepping one periodic tile of the bed"}
        else:
            parity = {"checked": False, "particles_conserved": bool(conserved), "why": "oracle/_ref not available on this box"}
    except Exception as ex_:  # the bench line must survive a failing checker
        parity = {"checked": False, "error": repr(ex_)[:200]}

    # CPU baseline: the unmodified reference, 1 core, one tile of the same bed
    cpu = None
    if rank == 0 and not args.no_cpu and world == 1:
        import ref_driver
        if ref_driver.available():
            steps_cpu = 1500
            r = run_reference_procs(1, 1, steps_cpu)
            if r:
                cpu = {"value": r[0], "unit": "particle-steps/s", "cores": 1, "kind": "reference",
                       "sample": "one 16,384-sphere tile of the bed x %d steps, serial reference build (oracle/_ref)" % steps_cpu}
    clocks = sampler.stop() if sampler else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["sampler disabled"]}
    out = {"metric": METRIC, "value": value, "unit": "particle-steps/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
           "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": "%d-sphere settled polydisperse bed (%dx%d replicas of bench_data/tile16k, radii U[1.5,3] mm), "
                                  "hertz/history/cdt, floor + periodic xy, dt 1e-5, skin 1 mm" % (n, tiles, tiles),
                      "particles": n, "l2_policy": "inputs (%.1f GB of state + lists) exceed the 126 MB L2" % (n * 600 / 1e9),
                      "rebuilds_in_timed_region": int(nbuilds), "rebuild_ms": round(rebuild_ms, 3),
                      "rebuild_ms_how": "same window with %d forced rebuilds (neigh_modify every %d check no) minus the rebuild-free window" % (nb_forced, R),
                      "parallelism": "1 GPU" if world == 1 else "%d GPUs: x-slab bricks, peer-memory halo per step (roofline fields are rank 0's)" % world},
           "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu,
           "parity": parity, "falling": falling, "amortised": amortised}
    if rank == 0:
        print(json.dumps(out))
    if dist:
        dist.barrier(); dist.destroy_process_group()


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--ref-worker":
        ref_worker(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]); return
    if len(sys.argv) > 1 and sys.argv[1] == "--ref-state":
        ref_state_worker(int(sys.argv[2]), sys.argv[3]); return
    if len(sys.argv) > 1 and sys.argv[1] == "--ref-server":
        ref_server(int(sys.argv[2])); return
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="own")
    ap.add_argument("--tiles", type=int, default=16)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-falling", action="store_true")
    ap.add_argument("--falling-steps", type=int, default=2000)
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
