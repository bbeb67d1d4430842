#!/bin/bash
# fast GPU visit: the core parity file + bench under option sets.  usage: tools/gpu_fast.sh <tag> [opts ...]   ("" = defaults)
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
grep -E "^FAILED|^ERROR|passed|failed|rc=" gpurun_out/${TAG}_pytest.log | head -10
tools/gpu_opts.sh $TAG "$@"
