#!/bin/bash
# round 2, call G (1 GPU): batched parity + phase counters + cfg2 bench (32-slot stash, dense start-phase kernel)
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_batched.py -q -m gpu -x --timeout 120 2>&1 | tail -5
export VROD_LIB=$PWD/vrod_b200/libvrod_knn_dbg.so
for M in 1 nocand; do
  echo "=== VROD_BATCHED_DEBUG=$M"
  VROD_BATCHED_DEBUG=$M timeout 200 python tests/tools/batched_check.py prof10 2>&1 | grep -E "tiles \[|time " | tail -9 | cut -c1-330
done > gpurun_out/modes3.log 2>&1
cat gpurun_out/modes3.log
unset VROD_LIB
timeout 600 python bench.py --workload cfg2 --no-extra --no-cpu-baseline --steps 30 > gpurun_out/bench_cfg2_g.json 2> gpurun_out/bench_cfg2_g.err || tail -8 gpurun_out/bench_cfg2_g.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_cfg2_g.json").read().strip().splitlines()[-1])
print("cfg2: value %.0f qps, ms/step %.3f, e2e %.0f, roofline %.1f TF frac_burst %.3f, parity %s, clocks %s" % (
    d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["achieved"], d["roofline"]["frac_of_burst"], d["parity"]["ok"], d["clocks"]))
PY
