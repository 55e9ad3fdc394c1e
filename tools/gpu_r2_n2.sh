#!/bin/bash
# round 2, 2-GPU call: smoke (multi-GPU legs), single-process multi-GPU context tests, process-per-GPU tests (fused + NCCL exchange), bench at N=2
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python __graft_entry__.py smoke 2>&1 | tail -6
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_gpu_sharded.py -q -m gpu 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err || tail -12 gpurun_out/bench_n2.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2.json").read().strip().splitlines()[-1])
print("N=2 cfg3: value %.1f qps e2e %.1f ms/step %.3f frac %.3f parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["parity"]))
for e in d.get("extra", []):
    print(" extra", e["workload"], "value %.0f e2e %.0f ms/step %.3f roofline %s %.3f parity %s" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["roofline"]["bound"], e["roofline"]["frac"], e["parity"]["ok"]))
print("wall", d.get("bench_wall_s"))
PY
