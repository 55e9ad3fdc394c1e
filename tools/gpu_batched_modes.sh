#!/bin/bash
# bottleneck isolation of the batched tile kernel (BPATH=3 bf16 mirror mode, BPATH=4 tf32): normal counters, no candidates,
# no epilogue, TMEM loads only, no MMAs, neither (VROD_BATCHED_DEBUG modes of knn_batched.cu)
mkdir -p gpurun_out
for M in 1 nocand noepi ldonly nomma noepi,nomma; do
  echo "=== VROD_BATCHED_DEBUG=$M" 
  VROD_BATCHED_DEBUG=$M timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "tiles \[(2368|9472|8288|[0-9]+),|time " | tail -9 | cut -c1-330
done > gpurun_out/exp1.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/exp1.log
cat gpurun_out/exp1.log
