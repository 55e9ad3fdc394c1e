#!/bin/bash
# round 2, call A (1 GPU): smoke, the whole GPU suite, the MMA issue-rate probe, the full bench line (headline + extra legs)
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -4
timeout 1500 python -m pytest tests -x -q -m gpu --durations=8 2>&1 | tail -18
timeout 120 tools/mma_probe 2000 > gpurun_out/mma_probe.log 2>&1; tail -32 gpurun_out/mma_probe.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err || tail -8 gpurun_out/bench_r2a.err
tail -c 6000 gpurun_out/bench_r2a.json
