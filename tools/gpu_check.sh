#!/bin/bash
# one 1-GPU visit: parity tests + the driver's bench command (20 steps / 5 warm-up).  usage: tools/gpu_check.sh <tag> [bench args]
TAG=$1; shift
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/${TAG}_pytest.log | head -30
for f in tools/dbg/*.py; do [ -f "$f" ] && { echo "== $f"; timeout 300 python $f 2>&1 | tail -25; }; done
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 "$@" > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench.json"):
    if l.startswith("{"):
        d=json.loads(l); r=d["roofline"]
        print("value %.3f G/s ms/step %.4f kernel_ms %.4f frac %.3f e2e %.3f G (%s) rebuild_ms %s falling %s parity %s" % (d["value"]/1e9, d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["value"]/1e9, d["e2e"]["job"][-90:], d["config"].get("rebuild_ms"), d.get("falling"), d.get("parity")))
PY
tail -5 gpurun_out/${TAG}_bench.err
