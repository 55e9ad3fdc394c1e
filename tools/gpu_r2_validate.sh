#!/bin/bash
# full GPU suite, then the batched profiles of the final code (4 phases per configs[2] batch: -s 12 skips 3 warm-up batches)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -8 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
CMD="python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu-baseline --no-extra"
$CMD > gpurun_out/plain_cfg2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_cfg2.csv $CMD > gpurun_out/ncu_launches_cfg2.log 2>&1
$CMD > gpurun_out/plain3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:batched_tile -s 12 -c 4 -f -o gpurun_out/prof_tile_cfg2 $CMD > gpurun_out/ncu_full_cfg2.log 2>&1
$CMD > gpurun_out/plain4.log 2>&1 &&
ncu --set full --clock-control none -k regex:batched_finish -s 12 -c 4 -f -o gpurun_out/prof_finish_cfg2 $CMD > gpurun_out/ncu_full_finish.log 2>&1
tail -1 gpurun_out/plain_cfg2.log | cut -c1-200
timeout 600 python tests/tools/sweep.py > gpurun_out/sweep.log 2>&1; tail -3 gpurun_out/sweep.log
