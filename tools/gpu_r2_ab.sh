#!/bin/bash
# A/B on one box: the committed library (libvrod_knn_head.so) against the working tree's.
# Build the former first:  git stash && make && cp vrod_b200/libvrod_knn.so vrod_b200/libvrod_knn_head.so && git stash pop && make
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv > gpurun_out/ab.log
{
for rep in 1 2; do
for L in head new; do
  if [ $L = head ]; then export VROD_LIB=$PWD/vrod_b200/libvrod_knn_head.so; else unset VROD_LIB; fi
  echo "=== $L (rep $rep)"
  timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -2
  timeout 200 python tests/tools/batched_check.py one 1000000 128 0 10 256 2>&1 | grep -E "time " | tail -2
done
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu --format=csv
} >> gpurun_out/ab.log 2>&1
cat gpurun_out/ab.log
