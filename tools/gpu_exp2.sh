#!/bin/bash
# thread-local append + L2 prefetch: parity, counters, prefetch-distance sweep
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_batched.py -x -q -m gpu 2>&1 | tail -3
echo "=== counters (prefetch 4)"
VROD_BATCHED_DEBUG=1 timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "batched dbg\]|time " | tail -9 | cut -c1-330
for PF in 0 2 4 8 16; do
  echo "=== plain, prefetch $PF"
  VROD_BATCHED_PREFETCH=$PF timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "time "
done
} > gpurun_out/exp2.log 2>&1
cat gpurun_out/exp2.log
