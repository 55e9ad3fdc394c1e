#!/bin/bash
# One GPU call: smoke, the bench lines, the ncu launch list and one full capture of the scan kernel.
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1 || { tail -20 gpurun_out/smoke.log; exit 1; }
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err || { tail -20 gpurun_out/bench_cfg3.err; exit 1; }
python bench.py --workload cfg1 --steps 200 --warmup 5 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err || { tail -20 gpurun_out/bench_cfg1.err; exit 1; }
CMD="python bench.py --workload ${PROF_WORKLOAD:-cfg1} --steps 5 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:fast_scan -s 4 -c 2 -f -o gpurun_out/prof_scan $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
cat gpurun_out/bench_cfg3.json gpurun_out/bench_cfg1.json
