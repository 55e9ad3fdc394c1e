#!/bin/bash
# round 2 (1 GPU): full suite + soak after the merge fallback
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 150 --deselect tests/test_gpu_parity.py::test_config3_100m_x128_l2_top10_full_size 2>&1 | tail -8
SOAK_CASES=48 timeout 600 python tests/tools/soak_batched.py > gpurun_out/soak.log 2>&1; grep -E "soak done|MISMATCH" gpurun_out/soak.log | cut -c1-200; grep -E "clusters|lowrank" gpurun_out/soak.log | cut -c1-120 | head -20
