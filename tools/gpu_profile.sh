#!/bin/bash
# One GPU call: ncu launch lists + full captures for the bench workloads (each after a plain run exited 0).
mkdir -p gpurun_out
for W in cfg3 cfg2; do
  CMD="python bench.py --workload $W --steps 4 --warmup 3 --no-cpu-baseline --no-extra"
  $CMD > gpurun_out/plain_$W.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_$W.csv $CMD > gpurun_out/ncu_launches_$W.log 2>&1
done
CMD="python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fast_scan -s 4 -c 1 -f -o gpurun_out/prof_scan_cfg3 $CMD > gpurun_out/ncu_full_cfg3.log 2>&1
# the shard sizes of N = 2 / 4 / 8 on one GPU: DRAM traffic of the scan per launch for roofline.traffic at N > 1
for R in 50000000 25000000 12500000; do
  CMD="python bench.py --workload cfg3 --rows $R --steps 4 --warmup 3 --no-cpu-baseline --no-extra"
  $CMD > gpurun_out/plain_r$R.log 2>&1 &&
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fast_scan -s 4 -c 1 --csv --log-file gpurun_out/traffic_r$R.csv $CMD > /dev/null 2>&1
done
# cfg2: all 4 phase launches of one batch (the 4th batch: 3 warm-up batches x 4 phases = 12 launches skipped)
CMD="python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:batched_tile -s 12 -c 4 -f -o gpurun_out/prof_tile_cfg2 $CMD > gpurun_out/ncu_full_cfg2.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:batched_finish -s 12 -c 4 -f -o gpurun_out/prof_finish_cfg2 $CMD > gpurun_out/ncu_full_finish.log 2>&1
tail -n 2 gpurun_out/ncu_full_cfg3.log; tail -n 2 gpurun_out/ncu_full_cfg2.log
