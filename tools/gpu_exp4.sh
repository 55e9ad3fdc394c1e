#!/bin/bash
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_batched.py tests/test_gpu_property.py -x -q -m gpu 2>&1 | tail -3
CMD="python tools/batched_check.py prof10"
timeout 200 $CMD 2>&1 | grep -E "time "
timeout 200 $CMD 2>&1 | grep -E "time "
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_p10.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(l for l in open('gpurun_out/launches_p10.csv') if l.startswith('"')))
hdr=rows[0]; ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value')
out=[(r[ki].split('::')[-1][:24],int(r[vi].replace(',',''))) for r in rows[2:] if 'batched' in r[ki]]
print(out[-16:])
print('tile', sum(v for n,v in out[-16:] if 'Batched' in n), 'finish', sum(v for n,v in out[-16:] if 'Finish' in n))
PY
} > gpurun_out/exp4.log 2>&1
cat gpurun_out/exp4.log
