#!/bin/bash
# bottleneck isolation of the batched tile kernel (BPATH=3 bf16 mirror mode, BPATH=4 tf32) with the DEBUG build of the library
# (make debug): normal counters, no candidates, no epilogue, TMEM loads only, no MMAs, neither (VROD_BATCHED_DEBUG modes)
mkdir -p gpurun_out
export VROD_LIB=$PWD/vrod_b200/libvrod_knn_dbg.so
for M in 1 nocand noepi ldonly nomma noepi,nomma; do
  echo "=== VROD_BATCHED_DEBUG=$M"
  VROD_BATCHED_DEBUG=$M timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "tiles \[|time " | tail -9 | cut -c1-330
done > gpurun_out/modes.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/modes.log
cat gpurun_out/modes.log
