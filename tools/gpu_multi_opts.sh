#!/bin/bash
# bench at N ranks under several engine option sets (200-step window, no CPU leg).  usage: tools/gpu_multi_opts.sh <N> <tag> "opts1" "opts2" ...
N=$1; TAG=$2; shift; shift
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
k=0
for o in "$@"; do
  k=$((k+1))
  DEM_OPTS="$o" timeout 600 $TR --master-port $((29540+k)) bench.py --gpus $N --steps ${STEPS:-200} --warmup 10 --no-cpu --no-falling > gpurun_out/${TAG}_n${N}_opts$k.json 2>gpurun_out/${TAG}_n${N}_opts$k.err
  python - <<PY
import json
try:
    for l in open("gpurun_out/${TAG}_n${N}_opts$k.json"):
        if l.startswith("{"):
            d=json.loads(l); p=d.get("parity") or {}
            print("[$o] N=$N %.3f G/s ms/step %.4f kernel_ms %.4f share %.2f e2e %.3f parity %s rebuild_ms %s" % (d["value"]/1e9, d["ms_per_step"], d["roofline"]["kernel_ms"], d["roofline"]["kernel_share_of_step"], d["e2e"]["value"]/1e9, p.get("ok"), d["config"].get("rebuild_ms")))
except Exception as e:
    print("[$o] failed", e); print(open("gpurun_out/${TAG}_n${N}_opts$k.err").read()[-1500:])
PY
done
