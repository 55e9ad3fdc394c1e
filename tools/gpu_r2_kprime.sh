#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -x -q -m gpu --timeout 300 2>&1 | tail -5 > gpurun_out/kprime_tests.log
cat gpurun_out/kprime_tests.log
{
for KP in 0 192 160 128; do
  echo "=== k=100: KPRIME=$KP"
  VROD_BATCHED_KPRIME=$KP timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -1
  VROD_BATCHED_KPRIME=$KP timeout 200 python tests/tools/batched_check.py one 1000000 128 0 100 256 2>&1 | grep -E "time " | tail -1
done
for KP in 0 48 32 24; do
  echo "=== k=10: KPRIME=$KP"
  VROD_BATCHED_KPRIME=$KP timeout 200 python tests/tools/batched_check.py one 10000000 128 1 10 1024 2>&1 | grep -E "time " | tail -1
  VROD_BATCHED_KPRIME=$KP timeout 200 python tests/tools/batched_check.py one 1000000 128 0 10 256 2>&1 | grep -E "time " | tail -1
  VROD_BATCHED_KPRIME=$KP timeout 200 python tests/tools/batched_check.py one 1000000 64 0 10 256 2>&1 | grep -E "time " | tail -1
done
echo "=== high dims, default k'"
for shape in "1000000 1536 0 10 256" "1000000 1536 0 100 256" "1000000 1536 1 100 256" "1000000 768 0 100 256" "1000000 768 1 100 256"; do
  VROD_VERBOSE=1 timeout 300 python tests/tools/batched_check.py one $shape 2>&1 | grep -E "vrod\]|time " | tail -3
done
} > gpurun_out/kprime.log 2>&1
cat gpurun_out/kprime.log
