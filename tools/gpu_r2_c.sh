#!/bin/bash
# round 2, call C (1 GPU): TMEM load probe; batched path after the tiled-mirror / lean-issue rework: parity, many-queries, cfg2 + cfg3b bench
mkdir -p gpurun_out
timeout 120 tools/tmem_ld_probe > gpurun_out/tmem_ld_probe.log 2>&1; cat gpurun_out/tmem_ld_probe.log
timeout 900 python -m pytest tests/test_gpu_batched.py -q -m gpu -x 2>&1 | tail -8
timeout 300 python tools/diag_manyq.py 2>&1 | tail -9
timeout 600 python bench.py --workload cfg2 --no-extra --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg2_c.json 2> gpurun_out/bench_cfg2_c.err || tail -8 gpurun_out/bench_cfg2_c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_c.json").read().strip().splitlines()[-1])
print("cfg2: value %.0f qps, ms/step %.3f, e2e %.0f, roofline %.1f TF frac_burst %.3f, parity %s, clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac_of_burst"], d["parity"]["ok"], d["clocks"]))
PY
