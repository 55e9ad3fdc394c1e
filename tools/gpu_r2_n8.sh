#!/bin/bash
# round 2, 8-GPU call: process-per-GPU tests at 8 ranks (fused exchange + NCCL fallback), multi-GPU context tests at 8 devices,
# bench N=8 (headline + extras + parity), the single-process context on 100M x 128
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 900 python -m pytest tests/test_gpu_sharded.py -q -m gpu -k "8" --timeout 600 > gpurun_out/n8_sharded_tests.log 2>&1; tail -4 gpurun_out/n8_sharded_tests.log
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu -k "8] or cli" --timeout 600 > gpurun_out/n8_multi_tests.log 2>&1; tail -4 gpurun_out/n8_multi_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err || tail -12 gpurun_out/bench_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n8.json").read().strip().splitlines()[-1])
print("N=8 cfg3: value %.1f qps e2e %.1f ms/step %.4f frac %.3f launches %d parity %s lat %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"], d["latency"]["resident_ms"]))
for e in d.get("extra", []):
    print(" extra", e["workload"], "value %.0f e2e %.0f ms/step %.3f roofline %s %.3f parity %s" % (e["value"], e["e2e"]["value"], e["ms_per_step"], e["roofline"]["bound"], e["roofline"]["frac"], e["parity"]["ok"]))
print("wall", d.get("bench_wall_s"))
PY
timeout 600 python tests/tools/multi_ctx_bench.py > gpurun_out/multi_ctx_n8.json 2> gpurun_out/multi_ctx_n8.err; tail -1 gpurun_out/multi_ctx_n8.json | cut -c1-900; tail -3 gpurun_out/multi_ctx_n8.err
