#!/bin/bash
# round 2 (1 GPU): whole GPU suite with a per-test time-out, scan tail stamps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 150 --deselect tests/test_gpu_parity.py::test_config3_100m_x128_l2_top10_full_size 2>&1 | tail -15
VROD_LIB=$PWD/vrod_b200/libvrod_knn_dbg.so VROD_SCAN_DEBUG=1 timeout 120 python tools/scan_probe.py 2>&1 | grep -E "==|scan dbg" | awk '/==/{print; n=0} /scan dbg/{n++; if(n>=4)print}' > gpurun_out/scan_tail.log; cat gpurun_out/scan_tail.log
