#!/bin/bash
# round 2, 2-GPU call: block-cyclic sharding -- single-GPU suite (sanity), process-per-GPU tests (fused + NCCL), multi-GPU context tests, bench N=2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --timeout 300 -x --deselect tests/test_gpu_parity.py::test_config3_100m_x128_l2_top10_full_size 2>&1 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-extra --no-cpu-baseline > gpurun_out/bench_n2c.json 2> gpurun_out/bench_n2c.err || tail -12 gpurun_out/bench_n2c.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_n2c.json").read().strip().splitlines()[-1])
print("N=2 cfg3: value %.1f qps e2e %.1f ms/step %.4f frac %.3f launches %d parity %s" % (d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["gpu_launches"], d["parity"]["ok"]))
PY
