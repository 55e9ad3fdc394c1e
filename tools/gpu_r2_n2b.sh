#!/bin/bash
# round 2, 2-GPU call: fused exchange inside the scan kernels -- process-per-GPU tests (fused + NCCL), multi-GPU context tests, bench N=2
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu --timeout 600 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_n2b.json 2> gpurun_out/bench_n2b.err || tail -12 gpurun_out/bench_n2b.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2b.json").read().strip().splitlines()[-1])
print("N=2 cfg3: value %.1f qps e2e %.1f ms/step %.4f frac %.3f launches %d parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"]))
PY
