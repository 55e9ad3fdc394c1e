#!/bin/bash
# guessed thresholds + early stage release: parity of the batched tests, then configs[2] timing over growth factors
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -x -q -m gpu --timeout 300 2>&1 | tail -15 > gpurun_out/guess_tests.log
cat gpurun_out/guess_tests.log
{
for G in 0 6 8 12 16 24; do
  echo "=== guess, VROD_BATCHED_GROWTH=$G"
  VROD_BATCHED_GROWTH=$G timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -2
done
for G in 2.5 3.5; do
  echo "=== VROD_BATCHED_NO_GUESS=1 VROD_BATCHED_GROWTH=$G"
  VROD_BATCHED_NO_GUESS=1 VROD_BATCHED_GROWTH=$G timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -2
done
echo "=== sweep shapes (default settings)"
timeout 300 python tests/tools/batched_check.py 2>&1 | grep -E "time |bad" | tail -12
} > gpurun_out/guess_timing.log 2>&1
cat gpurun_out/guess_timing.log
