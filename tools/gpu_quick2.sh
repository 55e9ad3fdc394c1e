#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q -m gpu 2>&1 | tail -2
CMD="python tools/batched_check.py prof10"
$CMD > gpurun_out/plain_p10.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_p10.csv $CMD > /dev/null 2>&1
grep time gpurun_out/plain_p10.log
