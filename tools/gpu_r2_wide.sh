#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --timeout 600 2>&1 | tail -3
for shape in "1000000 1536 0 10 256" "1000000 1536 0 100 256" "1000000 1536 1 100 256" "1000000 768 0 100 256" "1000000 128 0 100 256"; do
  VROD_VERBOSE=1 timeout 300 python tests/tools/batched_check.py one $shape 2>&1 | grep -E "vrod\]|time " | tail -4 | cut -c1-220
done
timeout 300 python bench.py --workload cfg2 --no-extra --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 value %.0f e2e %.0f ms %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['latency'])"
timeout 900 python tests/tools/soak_batched.py 2>&1 | tail -2
