#!/bin/bash
# round 2 (1 GPU): band mode -- batched tests, the stress-distribution soak, cfg2 bench unchanged?
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_batched.py -q -m gpu --timeout 150 2>&1 | tail -8
SOAK_CASES=24 timeout 600 python tests/tools/soak_batched.py > gpurun_out/soak.log 2>&1; grep -E "clusters|lowrank|grid|soak done|MISMATCH" gpurun_out/soak.log | cut -c1-200
timeout 600 python bench.py --workload cfg2 --no-extra --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg2_w.json 2> gpurun_out/bench_cfg2_w.err || tail -8 gpurun_out/bench_cfg2_w.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_w.json").read().strip().splitlines()[-1])
print("cfg2: value %.0f qps, ms/step %.3f, e2e %.0f, roofline %.1f TF frac_burst %.3f, parity %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac_of_burst"], d["parity"]["ok"]))
PY
