#!/usr/bin/env python
"""bench.py -- throughput of the exact top-k kNN scan (vRod SEARCH hot path) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg1|cfg0|cfg2] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is ONE pass of the hot path over one batch of synthetic queries (batch = 1 query unless the
workload says otherwise): f32 scan of this rank's row shard + exact f64 rerank + guard (+ NCCL
all-gather and merge of the per-rank top-k lists when N > 1).  The default workload is BASELINE.json
configs[3] -- 100M x 128 f32, Euclidean, top-10 -- the collection the metric "queries/sec for exact
top-10 kNN at 1/2/4/8 B200" is quoted on; it fits one GPU (51.2 GB), and is row-sharded over the N
ranks (strong scaling: the collection is fixed, per-GPU rows shrink as N grows).

Prints ONE JSON line on rank 0 (see the task contract): value = queries/s with queries and results
resident in HBM, timed with CUDA events on the library's stream, max over ranks; e2e = the same metric
through the host-buffer C-ABI call (vrod_collection_search: H2D of the query and D2H of the ids and
distances inside the timed region); roofline = the scan kernel against the measured HBM peak;
cpu_baseline = the CPU oracle (test infrastructure, oracle/) timed on this box's host cores on a
bounded sample of the same collection.  --impl reference times that CPU oracle as the reference arm
(the reference itself, sekulas/vRod, has an empty SEARCH body and no toolchain here: DESIGN.md).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows, dim, metric, k, batch, BASELINE.json config)
    "cfg3": (100_000_000, 128, 0, 10, 1, "configs[3]: 100Mx128 f32 L2, exact top-10, row-sharded over the GPUs"),
    "cfg1": (1_000_000, 768, 1, 10, 1, "configs[1]: 1Mx768 f32 cosine, single-query exact top-10"),
    "cfg0": (10_000, 128, 0, 10, 1, "configs[0]: 10kx128 f32 Euclidean, single-query exact top-10"),
    "cfg2": (10_000_000, 128, 0, 100, 1024, "configs[2]: 10Mx128 f32 L2, 1024 queries, exact top-100"),
    "cfg3b": (100_000_000, 128, 0, 10, 1024, "configs[3] collection with the north_star's batched load: 100Mx128 f32 L2, "
                                             "1024 queries, exact top-10, row-sharded over the GPUs"),
}
DATA_SEED, QUERY_SEED = 0x5EED0001, 0x5EED0002
METRIC_NAME = "queries/sec, exact top-k kNN"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def measured_bf16_peak():
    """Dense bf16 tensor peak from MEASURED_PEAKS.json: the SUSTAINED figure (the batched kernels are timed inside
    a long run of back-to-back batches), with the burst figure beside it; else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if "bf16_tflops_sustained" in d:
            return float(d["bf16_tflops_sustained"]), float(d.get("bf16_tflops", d["bf16_tflops_sustained"])), \
                "measured (MEASURED_PEAKS.json bf16_tflops_sustained: torch.matmul bf16 8192^3 back to back)"
        if "bf16_tflops" in d:
            return float(d["bf16_tflops"]), float(d["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
    return 2250.0, 2250.0, "fallback (B200_PROFILING.md: 2.25 PFLOP/s dense bf16 nominal)"


def ncu_traffic(workload):
    """dram bytes per scan launch from the committed ncu capture (profiles/roofline_traffic.json), else None."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(workload)
    return None


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line), sampled every
    50 ms from a background thread through NVML in-process.  (An `nvidia-smi -lms` child process was measured
    to slow the sampled GPU by ~50 % on an 8-GPU box, which then held back every other rank.)"""

    def __init__(self, index):
        self.index, self.samples, self._stop, self._thr = index, [], threading.Event(), None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: match by UUID of the torch device when possible
            import torch
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            handle = None
            for i in range(pynvml.nvmlDeviceGetCount()):
                h = pynvml.nvmlDeviceGetHandleByIndex(i)
                u = pynvml.nvmlDeviceGetUUID(h)
                u = u.decode() if isinstance(u, bytes) else u
                if uuid in u or u.replace("GPU-", "") == uuid:
                    handle = h
                    break
            if handle is None:
                handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._nv, self._h = pynvml, handle
        except Exception:
            self._nv = None
            return

        def loop():
            nv, h = self._nv, self._h
            while not self._stop.is_set():
                try:
                    self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                         nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM),
                                         (getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons)(h)))
                except Exception:
                    pass
                self._stop.wait(0.05)

        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def clear(self):
        self.samples = []

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        nv = getattr(self, "_nv", None)
        sm = [s[0] for s in self.samples]
        mx = [s[1] for s in self.samples]
        reasons = set()
        if nv:
            bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                    "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
            for s in self.samples:
                for name, bit in bits.items():
                    if s[2] & bit:
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvml in-process, 50 ms period"}


def cpu_oracle_leg(rows, dim, metric, k, steps, warmup, budget_s=20.0, batch=1):
    """Time the CPU oracle on a bounded sample of the workload: the first `sample` rows of the same
    seeded collection, min(batch, 4) queries per step, all host threads.  qps is scaled to the full row count."""
    from oracle import oracle as O
    O.build()
    # all the host cores this process may run on, stated explicitly: torchrun exports OMP_NUM_THREADS=1 to its
    # workers, which would otherwise make the N>1 reference arm a single-threaded run
    try:
        nthreads = len(os.sched_getaffinity(0))
    except AttributeError:
        nthreads = os.cpu_count() or 1
    sample = min(rows, max(10_000, (1 << 30) // (dim * 4)))     # <= 1 GiB of rows on the host
    qb = min(batch, 4)
    X = O.fill(sample, dim, DATA_SEED)
    Q = O.fill(max(steps + warmup, 1) * qb, dim, QUERY_SEED)
    for i in range(warmup):
        O.search(X, Q[i * qb:(i + 1) * qb], k, metric, nthreads=nthreads)
    t_used, times = 0.0, []
    for i in range(steps):
        t0 = time.perf_counter()
        O.search(X, Q[(warmup + i) * qb:(warmup + i + 1) * qb], k, metric, nthreads=nthreads)
        dt = time.perf_counter() - t0
        times.append(dt)
        t_used += dt
        if t_used > budget_s and len(times) >= 3:
            break
    per_query_full = statistics.mean(times) / qb * (rows / sample)
    return {"value": 1.0 / per_query_full, "unit": "queries/s", "cores": nthreads, "kind": "port",
            "sample": f"first {sample} of {rows} rows x {dim} (same Philox stream), {len(times)} steps of {qb} top-{k} "
                      f"queries by oracle/knn_oracle.c (canonical f64), time scaled by {rows / sample:g} to the full collection",
            "ms_per_step_sample": statistics.mean(times) * 1e3, "steps": len(times)}


def measure_tf32_peak():
    """Dense TF32 tensor peak of this GPU, measured the way MEASURED_PEAKS.json measures bf16: torch.matmul on
    8192^3 with TF32 allowed, best of 10 (burst).  Library GEMM used ONLY as the roofline denominator."""
    import torch
    torch.backends.cuda.matmul.allow_tf32 = True
    a = torch.randn(8192, 8192, device="cuda")
    b = torch.randn(8192, 8192, device="cuda")
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        a @ b
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    return 2 * 8192 ** 3 / (best / 1e3) / 1e12


def run_reference(args, rank):
    rows, dim, metric, k, batch, label = WORKLOADS[args.workload]
    if rank != 0:
        return
    leg = cpu_oracle_leg(rows, dim, metric, k, args.steps, args.warmup, budget_s=60.0, batch=batch)
    line = {"impl": "reference", "metric": METRIC_NAME, "value": leg["value"], "unit": "queries/s", "n_gpus": args.gpus,
            "steps": leg["steps"], "warmup": args.warmup, "ms_per_step": 1e3 / leg["value"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": label, "rows": rows, "dim": dim, "k": k, "batch": batch,
                       "note": "sekulas/vRod's SEARCH body is empty (src/command/types.rs:114-119) and rustc is absent: "
                               "the reference arm is the CPU oracle port of the written semantics, all host threads"},
            "cpu_baseline": {k2: leg[k2] for k2 in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vrod_b200", choices=["vrod_b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batched-kind", default="bf16", choices=["bf16", "tf32"],
                    help="operand mode of the batched (tensor-core) path: bf16 mirror (default) or the f32 rows as tf32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.rows:
        w = list(WORKLOADS[args.workload])
        w[0] = args.rows
        WORKLOADS[args.workload] = tuple(w)
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world != args.gpus:
        if args.gpus != 1 or world != 1:
            raise SystemExit(f"--gpus {args.gpus} needs WORLD_SIZE={args.gpus} (launch with torch.distributed.run)")

    import torch
    import torch.distributed as dist
    from vrod_b200 import ffi
    from vrod_b200.dist import share_comm_id

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm_id = share_comm_id(ffi.comm_unique_id, rank, world) if world > 1 else None
    ctx = ffi.Context(local, rank, world, comm_id)
    stream = torch.cuda.ExternalStream(ctx.stream())

    rows, dim, metric, k, batch, label = WORKLOADS[args.workload]
    coll = ctx.create("bench", dim, metric, rows)
    coll.fill_synthetic(rows, DATA_SEED)
    if batch > 1 and args.batched_kind == "tf32":
        coll.set_path(4)
    base, local_rows = coll.shard()

    # queries: one fresh Philox draw per step (never a row of X), generated by the library's own
    # device generator in a private unsharded context so that every rank holds the same full set
    nq = (args.steps + args.warmup) * batch
    qctx = ffi.Context(local)
    qcoll = qctx.create("queries", dim, 0, nq)
    qcoll.fill_synthetic(nq, QUERY_SEED)
    from_host = qcoll.read_rows(0, nq)
    qctx.close()
    q_dev = torch.from_numpy(from_host).cuda()
    ids_dev = torch.empty((batch, k), dtype=torch.int64, device="cuda")
    dist_dev = torch.empty((batch, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def step_resident(i):
        coll.search_device(q_dev[i * batch:(i + 1) * batch].data_ptr(), batch, k, ids_dev.data_ptr(), dist_dev.data_ptr())

    # ---- resident leg: `value` ----
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    if rank == 0:
        sampler.clear()          # keep only what is sampled from here on (the timed region)
    s0 = ctx.stats()
    ctx.profile(True)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        step_resident(args.warmup + i)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms_total = e0.elapsed_time(e1)
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile(False)
    s1 = ctx.stats()
    clocks = sampler.stop() if rank == 0 else None
    last_ids = ids_dev.cpu().numpy().astype(np.uint64)
    last_dist = dist_dev.cpu().numpy()

    # ---- e2e leg: host buffers through vrod_collection_search ----
    q_pinned = torch.from_numpy(from_host).pin_memory()
    q_np = q_pinned.numpy()
    for i in range(args.warmup):
        coll.search(q_np[i * batch:(i + 1) * batch], k)
    barrier()
    se0 = ctx.stats()
    t0 = time.perf_counter()
    for i in range(args.steps):
        j = args.warmup + i
        h_ids, h_dist = coll.search(q_np[j * batch:(j + 1) * batch], k)
    barrier()
    e2e_s = time.perf_counter() - t0
    se1 = ctx.stats()
    assert np.array_equal(h_ids, last_ids) and np.array_equal(h_dist.view(np.uint32), last_dist.view(np.uint32)), \
        "host-buffer and resident legs disagree"

    # max over ranks
    t = torch.tensor([ms_total, e2e_s * 1e3, kern_ms / max(kern_n, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kern_avg_ms = [float(x) for x in t.tolist()]

    if rank == 0:
        peak, peak_src = measured_peaks()
        qps = args.steps * batch / (ms_total / 1e3)
        algo_bytes = 4.0 * local_rows * dim                 # 4*N_local*d per pass, once per batch
        achieved = algo_bytes / (kern_avg_ms / 1e3) / 1e9 if kern_avg_ms > 0 else None
        traffic = ncu_traffic(args.workload if world == 1 else f"{args.workload}@{world}")
        used_batched = (s1["batched_tiles"] - s0["batched_tiles"]) > 0
        line = {
            "metric": METRIC_NAME, "value": qps, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": label, "rows": rows, "rows_per_gpu": local_rows, "dim": dim, "k": k, "batch": batch,
                       "metric": "cosine" if metric else "euclidean", "parallelism": f"row-shard x{world}",
                       "data_seed": hex(DATA_SEED), "query_seed": hex(QUERY_SEED),
                       "l2_policy": f"inputs larger than L2: {algo_bytes / 1e9:.2f} GB scanned per step per GPU vs 126 MB L2",
                       "arithmetic": "f32 scan, exact f64 rerank + guard (bit-identical to the oracle)"},
            "e2e": {"value": args.steps * batch / (e2e_ms / 1e3), "unit": "queries/s",
                    "h2d_bytes_per_step": int((se1["h2d_bytes"] - se0["h2d_bytes"]) // max(args.steps, 1)),
                    "d2h_bytes_per_step": int((se1["d2h_bytes"] - se0["d2h_bytes"]) // max(args.steps, 1)),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(s1["kernel_launches"] - s0["kernel_launches"]),
            "exact_rescans": int(s1["exact_rescans"] - s0["exact_rescans"]),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if achieved else None, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "fast_scan_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                         "kernel_ms": kern_avg_ms, "launches_timed": int(kern_n)},
            "clocks": clocks,
        }
        if used_batched:
            # batched path: a dense contraction, 2*B*N_local*d flops per step; the bracketed time is
            # the whole phased tile-kernel sequence (tiles + inter-phase merges) of one batch
            flops = 2.0 * batch * local_rows * dim
            ach = flops / (kern_avg_ms / 1e3) / 1e12
            if args.batched_kind == "tf32":
                tpeak, tburst = measure_tf32_peak(), None
                tsrc = "measured here: torch.matmul 8192^3 TF32, best of 10 (burst); the kernel's MMA kind is tf32"
                kname = "batched_tile_kernel (tcgen05.mma kind::tf32 on the stored f32 rows) + inter-phase batched_finish_kernel"
            else:
                tpeak, tburst, tsrc = measured_bf16_peak()
                kname = ("batched_tile_kernel (tcgen05.mma kind::f16, bf16 mirror of the rows, thresholds folded into the "
                         "contraction) + inter-phase batched_finish_kernel")
            line["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                                "frac": ach / tpeak, "traffic": traffic, "peak_source": tsrc, "peak_burst": tburst,
                                "kernel": kname,
                                "algorithmic_flops_per_step": flops, "kernel_ms": kern_avg_ms, "launches_timed": int(kern_n),
                                "hbm_floor": {"algorithmic_bytes": algo_bytes, "peak_gbs": peak}}
            line["config"]["arithmetic"] = (f"{args.batched_kind} tensor-core pass, exact f64 rerank + guard "
                                            "(bit-identical to the oracle)")
            line["dtype"] = args.batched_kind      # the type the dominant kernel computes in (f32 accumulate)
        # cheap live check of the last answer: sorted, and the claimed rows sit at the claimed distances
        assert np.all(np.diff(last_dist, axis=1) >= 0)
        if world == 1 and not args.no_cpu_baseline:
            from oracle import oracle as O
            O.build()
            qlast = from_host[(args.warmup + args.steps - 1) * batch]   # query 0 of the last batch
            for j in (0, k - 1):
                row = O.fill(1, dim, DATA_SEED, row0=int(last_ids[0, j]))[0]
                assert np.float32(O.distance(row, qlast, metric)) == last_dist[0, j], "distance check against the oracle failed"
            line["cpu_baseline"] = {k2: v for k2, v in cpu_oracle_leg(rows, dim, metric, k, 12, 1, batch=batch).items()
                                    if k2 in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    barrier()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
