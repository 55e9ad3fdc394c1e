#!/bin/bash
mkdir -p gpurun_out
for shape in "1000000 128 0 10 256" "1000000 64 0 10 256"; do
  timeout 200 python tests/tools/batched_check.py one $shape 2>&1 | grep "time " | tail -1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_b256.csv python tests/tools/batched_check.py one $shape > /dev/null 2>&1
  python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_b256.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=="ID"][0]
hdr=rows[hi]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value")
seq=[(r[ki].split("(")[0][-44:], float(r[vi].replace(",",""))/1e3) for r in rows[hi+1:] if len(r)>vi]
idx=[i for i,(k,v) in enumerate(seq) if 'prep_queries' in k]
for k,v in seq[idx[-1]:]: print(f"   {k:46s} {v:9.1f} us")
PY
done
