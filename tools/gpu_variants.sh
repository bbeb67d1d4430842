#!/bin/bash
# bench (no CPU leg) of library variants: usage tools/gpu_variants.sh <tag> lib1.so lib2.so ...
TAG=$1; shift
mkdir -p gpurun_out
for lib in "$@"; do
  DEM_B200_LIB=$PWD/$lib timeout 600 python bench.py --gpus 1 --no-cpu --steps 100 --warmup 10 > gpurun_out/${TAG}_$(basename $lib .so).json 2>gpurun_out/${TAG}_$(basename $lib .so).err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${TAG}_$(basename $lib .so).json")); print("$lib  %.3f G/s  kernel_ms %.4f  frac %.3f  e2e %.3f" % (d["value"]/1e9, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"]/1e9))
except Exception as ex:
    print("$lib FAILED", ex); print(open("gpurun_out/${TAG}_$(basename $lib .so).err").read()[-800:])
PY
done
