#!/bin/bash
# ncu full capture of one kernel while tools/config_runs.py runs one config.  usage: tools/gpu_prof_cfg.sh <tag> <config> <kernel regex> [skip]
TAG=$1; CFG=$2; KRE=$3; SKIP=${4:-30}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
    python tools/config_runs.py $CFG > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu $KRE rc=$?"
