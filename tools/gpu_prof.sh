#!/bin/bash
# ncu full capture of k_step on the default library
TAG=$1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 -f -o gpurun_out/${TAG}_k_step \
    python bench.py --steps 30 --warmup 3 --no-cpu > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
