#!/bin/bash
# quick GPU visit: parity tests on the default library, then bench (no CPU leg) for each library variant given
TAG=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -15 gpurun_out/${TAG}_pytest.log
timeout 600 python bench.py --gpus 1 --no-cpu > gpurun_out/${TAG}_bench_default.json 2>gpurun_out/${TAG}_bench_default.err; python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_bench_default.json")); print("default", d["value"]/1e9, "G/s kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"]/1e9)
PY
for lib in "$@"; do
  DEM_B200_LIB=$PWD/$lib timeout 600 python bench.py --gpus 1 --no-cpu > gpurun_out/${TAG}_bench_$(basename $lib .so).json 2>&1; python - <<PY
import json; d=json.load(open("gpurun_out/${TAG}_bench_$(basename $lib .so).json")); print("$lib", d["value"]/1e9, "G/s kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"])
PY
done
