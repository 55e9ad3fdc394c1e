#!/bin/bash
mkdir -p gpurun_out
SOAK_CASES=48 timeout 600 python tests/tools/soak_batched.py > gpurun_out/soak.log 2>&1; grep -E "soak done|MISMATCH" -A3 gpurun_out/soak.log | cut -c1-420
