#!/bin/bash
# round-end style validation: smoke, the whole GPU suite, the default bench line (both arms), the cfg2 line
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err || tail -5 gpurun_out/bench_cfg3.err
python bench.py --workload cfg2 --steps 50 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err || tail -5 gpurun_out/bench_cfg2.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err || tail -5 gpurun_out/bench_ref.err
for f in cfg3 cfg2 ref; do tail -1 gpurun_out/bench_$f.json | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d.get('roofline', {})
print('$f', 'value=%.1f' % d['value'], 'e2e=%.1f' % d['e2e']['value'], 'ms/step=%.4f' % d['ms_per_step'], 'roofline', r.get('bound'), r.get('achieved'), r.get('frac'), 'traffic', r.get('traffic'), 'launches', d.get('gpu_launches'), 'clocks', d.get('clocks'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))"; done
