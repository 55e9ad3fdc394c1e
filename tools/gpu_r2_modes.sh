#!/bin/bash
# where a candidate's cost comes from (debug build): production stamps, no candidates, parked-but-dropped, hits ignored
mkdir -p gpurun_out
export VROD_LIB=$PWD/vrod_b200/libvrod_knn_dbg.so
for M in 1 nocand noflush nopark; do
  echo "=== VROD_BATCHED_DEBUG=$M"
  VROD_BATCHED_DEBUG=$M timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "tiles \[|time " | tail -9 | cut -c1-400
done > gpurun_out/modes4.log 2>&1
unset VROD_LIB
for G in 2.5 3.5 5; do
  echo "=== release build, VROD_BATCHED_GROWTH=$G"
  VROD_BATCHED_GROWTH=$G timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -2
done >> gpurun_out/modes4.log 2>&1
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw --format=csv >> gpurun_out/modes4.log
cat gpurun_out/modes4.log
