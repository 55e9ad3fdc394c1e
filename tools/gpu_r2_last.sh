#!/bin/bash
# last check of the round on one GPU: smoke, the full GPU suite, the default bench line
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/bench_last_n1.json 2> gpurun_out/bench_last_n1.err; tail -c 300 gpurun_out/bench_last_n1.json | head -c 10 > /dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_last_n1.json").read().strip().splitlines()[-1])
print("N=1 cfg3: value %.1f e2e %.1f ms/step %.3f frac %.3f parity %s wall %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity"]["ok"], d.get("bench_wall_s")))
for e in d.get("extra", []):
    print(" extra", e["workload"], "value %.0f e2e %.0f ms/step %.3f roofline %s %.3f parity %s" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["roofline"]["bound"], e["roofline"]["frac"], e["parity"]["ok"]))
PY
