#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -8 > gpurun_out/final_tests.log
cat gpurun_out/final_tests.log
timeout 900 python tests/tools/soak_batched.py > gpurun_out/soak.log 2>&1; tail -6 gpurun_out/soak.log
timeout 300 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -1
