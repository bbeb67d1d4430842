#!/bin/bash
N=$1
for D in 0 4 8 12; do
DEM_DEBUG=$D timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$((D%10)) bench.py --gpus $N --steps 200 --warmup 20 --no-cpu 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('debug=$D N=$N', round(d['value']/1e9,3), 'G/s ms/step', round(d['ms_per_step'],4), 'kernel_ms', round(d['roofline']['kernel_ms'],4))
"
done
