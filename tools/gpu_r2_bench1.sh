#!/bin/bash
# N=1: the default bench line (cfg3 + extras + cpu baselines), the reference arm, cfg0/cfg1 lines
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err || tail -5 gpurun_out/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_n1.json 2> gpurun_out/bench_ref_n1.err || tail -5 gpurun_out/bench_ref_n1.err
timeout 300 python bench.py --workload cfg0 --no-extra --no-cpu-baseline > gpurun_out/bench_cfg0_n1.json 2>/dev/null
timeout 300 python bench.py --workload cfg2 --no-extra --no-cpu-baseline --steps 50 --warmup 5 > gpurun_out/bench_cfg2_n1.json 2>/dev/null
python - <<'PY'
import json
for f in ("bench_n1", "bench_ref_n1", "bench_cfg0_n1", "bench_cfg2_n1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "ERR", e); continue
    print(f, "value %.1f e2e %.1f ms/step %.4f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), "roofline", d.get("roofline", {}).get("frac"), "parity", d.get("parity", {}).get("ok"), "lat", d.get("latency_ms"))
    for e in d.get("extra", []):
        print("   extra", e["workload"], "value %.0f e2e %.0f ms/step %.3f roofline %s %.3f parity %s" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["roofline"]["bound"], e["roofline"]["frac"], e["parity"]["ok"]))
    if "cpu_baseline" in d: print("   cpu", json.dumps(d["cpu_baseline"])[:300])
PY
