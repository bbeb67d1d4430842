#!/bin/bash
# multi-GPU: parity driver + bench at N ranks.  usage: tools/gpu_scale.sh <N> <tag>
N=$1; TAG=$2
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/run_multi.py > gpurun_out/${TAG}_multi${N}.log 2>&1; grep -c "ok:" gpurun_out/${TAG}_multi${N}.log; grep "PARITY\|Error\|error" gpurun_out/${TAG}_multi${N}.log | head -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 200 --warmup 20 --no-cpu > gpurun_out/${TAG}_bench${N}.json 2>gpurun_out/${TAG}_bench${N}.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_bench${N}.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N", d["value"]/1e9, "G/s ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "share", d["roofline"]["kernel_share_of_step"], "e2e", d["e2e"]["value"]/1e9)
PY
