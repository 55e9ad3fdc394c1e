#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python tools/batched_check.py first 2>&1 | tail -8
timeout 900 python -m pytest tests/test_gpu_batched.py -x -q -m gpu 2>&1 | tail -15
for P in 3 4; do BPATH=$P timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "time "; done
VROD_BATCHED_DEBUG=1 timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "batched dbg\]|time " | tail -9 | cut -c1-330
} > gpurun_out/exp5.log 2>&1
cat gpurun_out/exp5.log
