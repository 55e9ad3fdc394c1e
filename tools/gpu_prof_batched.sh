#!/bin/bash
mkdir -p gpurun_out
export VROD_BATCHED_DEBUG=${DBGMODE:-nocand}
timeout 300 python tests/tools/batched_check.py prof > gpurun_out/plain_b.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:batched_tile -s 9 -c 1 -f -o gpurun_out/prof_batched python tests/tools/batched_check.py prof > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log
