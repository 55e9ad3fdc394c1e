#!/bin/bash
# multi-GPU box: the batched workloads row-sharded over N GPUs (one process per GPU); W="cfg3b cfg2" N=8 by default
mkdir -p gpurun_out
N=${N:-8}
for W in ${W:-cfg3b cfg2}; do
  timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus $N --workload $W --steps 12 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${W}_n$N.json 2> gpurun_out/bench_${W}_n$N.err || tail -5 gpurun_out/bench_${W}_n$N.err
  tail -1 gpurun_out/bench_${W}_n$N.json | cut -c1-220
done
