#!/bin/bash
# multi-GPU step-time attribution with the engine's profiling switches (DEM_DEBUG bit 4: no flag all-reduce, bit 8: no
# halo traffic between rebuilds; timings only, the physics of such runs is not valid).  usage: tools/gpu_scale_dbg.sh <N> <tag>
N=$1; TAG=$2
mkdir -p gpurun_out
for dbg in 0 4 8 12; do
DEM_DEBUG=$dbg timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$((dbg % 10)) bench.py --gpus $N --steps 300 --warmup 20 --no-cpu > gpurun_out/${TAG}_dbg${dbg}.json 2>gpurun_out/${TAG}_dbg${dbg}.err
python - <<PY
import json
for l in open("gpurun_out/${TAG}_dbg${dbg}.json"):
    if l.startswith("{"):
        d=json.loads(l); print("N=$N debug=$dbg  %.3f G/s  ms/step %.4f  kernel_ms %.4f" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"]))
PY
done
