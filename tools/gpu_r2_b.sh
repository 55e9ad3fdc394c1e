#!/bin/bash
# round 2, call B (1 GPU): probe with the no-swizzle layout, many-queries diagnosis (debug build), rest of the GPU suite
mkdir -p gpurun_out
timeout 120 tools/mma_probe 2000 > gpurun_out/mma_probe2.log 2>&1; grep -E "NO-swizzle|grid 148" gpurun_out/mma_probe2.log | head -20
VROD_LIB=$PWD/vrod_b200/libvrod_knn_dbg.so timeout 300 python tools/diag_manyq.py 1 2>&1 | tail -12
timeout 300 python tools/diag_manyq.py 2>&1 | tail -12
timeout 1500 python -m pytest tests -q -m gpu --deselect tests/test_gpu_batched.py::test_many_queries_few_rows_keeps_the_tensor_core_answers 2>&1 | tail -8
