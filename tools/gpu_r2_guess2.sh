#!/bin/bash
mkdir -p gpurun_out
{
echo "=== d=1536 k=100 with and without guessing"
VROD_VERBOSE=1 timeout 300 python tests/tools/batched_check.py one 1000000 1536 0 100 256 2>&1 | grep -E "vrod\]|time " | tail -8
VROD_BATCHED_NO_GUESS=1 VROD_VERBOSE=1 timeout 300 python tests/tools/batched_check.py one 1000000 1536 0 100 256 2>&1 | grep -E "vrod\]|time " | tail -8
} > gpurun_out/guess2.log 2>&1
cat gpurun_out/guess2.log
python bench.py --workload cfg2 --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/plain_cfg2b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg2b.csv python bench.py --workload cfg2 --steps 4 --warmup 3 --no-extra --no-cpu-baseline > gpurun_out/ncu_launches_cfg2b.log 2>&1
tail -1 gpurun_out/plain_cfg2b.log | cut -c1-300
