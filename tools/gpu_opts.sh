#!/bin/bash
# bench (no CPU leg, no falling window) under several engine option sets.  usage: tools/gpu_opts.sh <tag> "opts1" "opts2" ...
TAG=$1; shift
mkdir -p gpurun_out
k=0
for o in "$@"; do
  k=$((k+1))
  DEM_OPTS="$o" timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu --no-falling > gpurun_out/${TAG}_opts$k.json 2>gpurun_out/${TAG}_opts$k.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_opts$k.json")); print("[$o]  %.3f G/s  kernel_ms %.4f  frac %.3f  e2e %.3f  parity %s rebuild_ms %s" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["parity"].get("ok"), d["config"]["rebuild_ms"]))
except Exception as e:
    print("[$o] failed", e); print(open("gpurun_out/${TAG}_opts$k.err").read()[-800:])
PY
done
