#!/bin/bash
# bench (no CPU leg, no falling window) for the default library and each variant library given.  usage: tools/gpu_variants.sh <tag> [lib.so ...]
TAG=$1; shift
mkdir -p gpurun_out
for lib in liggghts-inl_b200/libdem_b200.so "$@"; do
  DEM_B200_LIB=$PWD/$lib timeout 600 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu --no-falling > gpurun_out/${TAG}_bench_$(basename $lib .so).json 2>gpurun_out/${TAG}_bench_$(basename $lib .so).err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_bench_$(basename $lib .so).json")); print("$lib  %.3f G/s  kernel_ms %.4f  frac %.3f  e2e %.3f  parity %s" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e9, d["parity"].get("ok")))
except Exception as e:
    print("$lib failed", e); print(open("gpurun_out/${TAG}_bench_$(basename $lib .so).err").read()[-800:])
PY
done
