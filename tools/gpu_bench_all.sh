#!/bin/bash
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err || tail -5 gpurun_out/bench_cfg3.err
python bench.py --workload cfg1 --steps 400 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err || tail -5 gpurun_out/bench_cfg1.err
python bench.py --workload cfg2 --steps 50 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err || tail -5 gpurun_out/bench_cfg2.err
python bench.py --workload cfg0 --steps 400 > gpurun_out/bench_cfg0.json 2> gpurun_out/bench_cfg0.err || tail -5 gpurun_out/bench_cfg0.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err || tail -5 gpurun_out/bench_ref.err
for f in cfg3 cfg1 cfg2 cfg0 ref; do python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_$f.json"))
    r = d.get("roofline", {})
    print("$f", "value=%.1f" % d["value"], "e2e=%.1f" % d["e2e"]["value"], "ms/step=%.4f" % d["ms_per_step"], "roofline=%s %.1f %s frac=%.3f" % (r.get("bound"), r.get("achieved") or 0, r.get("unit"), r.get("frac") or 0) if r else "", "launches", d.get("gpu_launches"), "clocks", d.get("clocks"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
except Exception as e:
    print("$f FAILED", e)
PY
done
