#!/bin/bash
# ncu full capture of one kernel of a short bench run.  usage: tools/gpu_prof.sh <tag> <kernel regex> [skip] [bench args...]
TAG=$1; KRE=$2; SKIP=${3:-12}; shift; shift; shift
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:$KRE -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
    python bench.py --steps 10 --warmup 3 --no-cpu --no-falling "$@" > gpurun_out/${TAG}_ncu_full.log 2>&1; echo "ncu full rc=$?"
