#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_gpu_batched.py tests/test_gpu_property.py -x -q -m gpu 2>&1 | tail -3
for G in 3.5 5 7 10; do
  echo "=== growth $G"
  VROD_BATCHED_GROWTH=$G timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "time "
done
python bench.py --workload cfg3b --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-2500
} > gpurun_out/exp8.log 2>&1
cat gpurun_out/exp8.log
