#!/usr/bin/env python
"""Phase timing of the end-to-end job bench.py measures (create, configure+upload, setup, run, download).
usage (GPU box): python tools/e2e_breakdown.py [tiles] [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "liggghts-inl_b200")); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, ROOT)
import torch
import dem_b200, cases, bench
tiles = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
torch.cuda.set_device(0)
c = bench.bed_case(tiles, tiles)
import numpy as np
def pinned(a):
    t = torch.empty(a.shape, dtype=torch.from_numpy(a[:0]).dtype, pin_memory=True); v = t.numpy(); v[...] = a; return v
for k in ("tag", "type", "mask", "x", "v", "omega", "radius", "density"):
    c[k] = pinned(np.ascontiguousarray(c[k], np.int32 if k in ("tag", "type", "mask") else np.float64))
n = len(c["tag"])
xo = pinned(np.zeros((n, 3))); vo = pinned(np.zeros((n, 3)))
for rep in range(3):
    torch.cuda.synchronize(); t = [time.perf_counter()]
    eng = dem_b200.Engine(device=0); torch.cuda.synchronize(); t.append(time.perf_counter())
    cases.apply(c, eng); torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.setup(); torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.run(steps); torch.cuda.synchronize(); t.append(time.perf_counter())
    x = eng.download("x", out=xo); v = eng.download("v", out=vo); torch.cuda.synchronize(); t.append(time.perf_counter())
    eng.close(); torch.cuda.synchronize(); t.append(time.perf_counter())
    names = ["create", "configure+upload", "setup", "run(%d)" % steps, "download x,v", "close"]
    print("rep %d: " % rep + "  ".join("%s %.1f ms" % (n, 1e3 * (b - a)) for n, a, b in zip(names, t[:-1], t[1:])) + "  total(no close) %.1f ms" % (1e3 * (t[-2] - t[0])))
