#!/bin/bash
# round 2, call E (1 GPU): MMA probe (layouts, lean issue), batched parity after the deferred-append stash, phase counters, cfg2 bench
mkdir -p gpurun_out
timeout 120 tools/mma_probe 2000 > gpurun_out/mma_probe3.log 2>&1; cat gpurun_out/mma_probe3.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_batched.py -q -m gpu -x 2>&1 | tail -5
VROD_LIB=$PWD/vrod_b200/libvrod_knn_dbg.so VROD_BATCHED_DEBUG=1 timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "tiles \[|time " | cut -c1-330
timeout 600 python bench.py --workload cfg2 --no-extra --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg2_e.json 2> gpurun_out/bench_cfg2_e.err || tail -8 gpurun_out/bench_cfg2_e.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_e.json").read().strip().splitlines()[-1])
print("cfg2: value %.0f qps, ms/step %.3f, e2e %.0f, roofline %.1f TF frac_burst %.3f, parity %s, clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac_of_burst"], d["parity"]["ok"], d["clocks"]))
PY
