#!/bin/bash
# last check of the round on two GPUs: smoke (multi-GPU legs), the sharded and multi-context tests
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu --timeout 600 2>&1 | tail -3
