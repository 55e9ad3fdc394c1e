#!/bin/bash
mkdir -p gpurun_out
{
for R in 0 64 128; do
  echo "=== REST=$R"
  for shape in "1000000 128 0 10 256" "1000000 64 0 10 256" "1000000 384 1 10 256" "1000000 128 0 100 256" "4000000 128 0 10 1024"; do
    VROD_BATCHED_REST=$R timeout 200 python tests/tools/batched_check.py one $shape 2>&1 | grep -E "time " | tail -1
  done
done
} > gpurun_out/rest.log 2>&1
cat gpurun_out/rest.log
