#!/bin/bash
mkdir -p gpurun_out
{
for M in nocand noepi ldonly nomma noepi,nomma; do
  echo "=== bf16 VROD_BATCHED_DEBUG=$M"
  VROD_BATCHED_DEBUG=$M timeout 200 python tools/batched_check.py prof10 2>&1 | grep -E "tiles \[19330|time " | tail -2 | cut -c1-330
done
} > gpurun_out/exp6.log 2>&1
cat gpurun_out/exp6.log
