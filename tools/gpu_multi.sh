#!/bin/bash
# multi-GPU visit: parity driver + the driver's bench command at N ranks (20 steps) + a long window.  usage: tools/gpu_multi.sh <N> <tag> [quick]
N=$1; TAG=$2; QUICK=$3
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$QUICK" ]; then
timeout 600 $TR --master-port 29533 tests/run_multi.py > gpurun_out/${TAG}_multi${N}.log 2>&1; grep -c "ok:" gpurun_out/${TAG}_multi${N}.log; grep "PARITY\|Error\|error" gpurun_out/${TAG}_multi${N}.log | head -5
fi
timeout 600 $TR --master-port 29534 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu > gpurun_out/${TAG}_bench${N}a.json 2>gpurun_out/${TAG}_bench${N}a.err
timeout 600 $TR --master-port 29536 bench.py --gpus $N --steps 1000 --warmup 20 --no-cpu --no-falling > gpurun_out/${TAG}_bench${N}long.json 2>gpurun_out/${TAG}_bench${N}long.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench${N}*.json")):
    for l in open(f):
        if l.startswith("{"):
            d=json.loads(l); p=d.get("parity") or {}
            print(f.split("_")[-1], "N=$N %.3f G/s ms/step %.4f kernel_ms %.4f share %.2f e2e %.3f G parity ok=%s dx %.1e dv %.1e rebuild_ms %s falling %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"], d["e2e"]["value"]/1e9, p.get("ok"), p.get("max_dx_over_rmin", -1), p.get("max_dv_over_sqrt_g_r", -1), d["config"].get("rebuild_ms"), (d.get("falling") or {}).get("value")))
PY
tail -3 gpurun_out/${TAG}_bench${N}a.err
