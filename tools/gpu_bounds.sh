#!/bin/bash
# Lower bounds of k_step on the bench bed (profiling aid, not bench values): option debug 1 = sweep only (no contact
# evaluation), 2 = no sweep either (own records + epilogue), 0 = the real kernel.  usage: tools/gpu_bounds.sh <tag>
TAG=$1
mkdir -p gpurun_out
for d in ${DBGS:-0 1 2}; do
  DEM_DEBUG=$d timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu --no-falling > gpurun_out/${TAG}_dbg$d.json 2>gpurun_out/${TAG}_dbg$d.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_dbg$d.json")); print("debug=$d  ms/step %.4f kernel_ms %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]))
except Exception as e:
    print("debug=$d failed", e); print(open("gpurun_out/${TAG}_dbg$d.err").read()[-800:])
PY
done
