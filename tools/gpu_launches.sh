#!/bin/bash
# ncu launch list (durations per kernel) of a short bench run.  usage: tools/gpu_launches.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu --no-falling > gpurun_out/${TAG}_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
python tools/ncu_kernels.py gpurun_out/${TAG}_launches.csv
