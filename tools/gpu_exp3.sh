#!/bin/bash
mkdir -p gpurun_out
{
for ST in 4 5 6; do
  echo "=== stages $ST"
  VROD_BATCHED_STAGES=$ST timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "time "
done
echo "=== counters (6 stages)"
VROD_BATCHED_DEBUG=1 timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "batched dbg\]|time " | tail -4 | cut -c1-330
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q -m gpu 2>&1 | tail -2
} > gpurun_out/exp3.log 2>&1
cat gpurun_out/exp3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:batched_finish -s 10 -c 3 -f -o gpurun_out/prof_finish2 python tools/batched_check.py prof10 > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log
