#!/bin/bash
# full GPU suite + cfg2 bench line (bf16 and tf32 modes) + the cfg2 launch list
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -16
python bench.py --workload cfg2 --steps 50 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err || tail -5 gpurun_out/bench_cfg2.err
python bench.py --workload cfg2 --steps 50 --batched-kind tf32 --no-cpu-baseline > gpurun_out/bench_cfg2_tf32.json 2> gpurun_out/bench_cfg2_tf32.err || tail -5 gpurun_out/bench_cfg2_tf32.err
cat gpurun_out/bench_cfg2.json gpurun_out/bench_cfg2_tf32.json | cut -c1-1800
