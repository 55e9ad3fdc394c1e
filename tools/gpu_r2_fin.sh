#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_batched.py -x -q -m gpu --timeout 300 2>&1 | tail -3
tools/gpu_r2_b256.sh
timeout 300 python tests/tools/batched_check.py prof10 2>&1 | grep -E "time " | tail -1
