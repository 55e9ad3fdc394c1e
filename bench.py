#!/usr/bin/env python
"""bench.py -- throughput of the exact top-k kNN scan (vRod SEARCH hot path) on 1..8 B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg3|cfg1|cfg0|cfg2|cfg3b] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A step is ONE pass of the hot path over one batch of synthetic queries (batch = 1 query unless the
workload says otherwise): f32 scan of this rank's row shard + exact f64 rerank + guard (+ the fused
NVLink exchange and merge of the per-rank top-k lists when N > 1).  The headline workload is
BASELINE.json configs[3] -- 100M x 128 f32, Euclidean, top-10 -- the collection the metric "queries/sec
for exact top-10 kNN at 1/2/4/8 B200" is quoted on; it fits one GPU (51.2 GB) and is row-sharded over
the N ranks (strong scaling: the collection is fixed, per-GPU rows shrink as N grows).

Prints ONE JSON line on rank 0 (see the task contract):
  value         queries/s with queries and results resident in HBM, CUDA events on the library's stream, max over ranks
  e2e           the same metric through the host-buffer C-ABI call (vrod_collection_search: H2D of the queries and
                D2H of ids + distances inside the timed region)
  roofline      the dominant kernel against the measured peak (HBM for the scan, bf16 tensor for the batched path)
  parity        every rank's answers of ALL timed steps hashed and compared across ranks, and queries of the last
                step replayed by the CPU oracle over the WHOLE collection (chunked Philox replay): a missed
                neighbour cannot pass
  latency       median / p99 per step of both legs
  cpu_baseline  the CPU oracle (test infrastructure, oracle/) on this box's host cores on a bounded sample of the
                same collection; cpu_baseline_variants lists single-thread naive f32, all-cores naive f32 and
                all-cores canonical f64
  extra         the other BASELINE configs measured in the same process after the headline legs -- configs[3] with
                the north_star's 1024-query batches (tensor cores, the collection is already resident), configs[1]
                (1M x 768 cosine) and configs[2] (10M x 128, 1024 queries, top-100) -- each with value, e2e, roofline,
                parity, clocks
--impl reference times the CPU oracle as the reference arm (sekulas/vRod's SEARCH body is empty and there is no
rustc here: DESIGN.md); its `config` is the vrod arm's, key for key.
"""
import argparse
import hashlib
import json
import os
import statistics
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
# extra legs after the headline one: (workload, timed steps, queries of the last step replayed by the oracle)
EXTRA = {"cfg3": [("cfg3b", 10, 2), ("cfg1", 100, 2), ("cfg2", 30, 2)]}
DATA_SEED, QUERY_SEED = 0x5EED0001, 0x5EED0002
METRIC_NAME = "queries/sec, exact top-k kNN"
ARITH_SCAN = "f32 scan, exact f64 rerank + guard (bit-identical to the oracle)"


def workload_config(name, world):
    """The `config` object of the JSON line: the SAME dict, key for key, in the vrod arm and the reference arm."""
    rows, dim, metric, k, batch, label = WORKLOADS[name]
    per = (rows + world - 1) // world
    return {"workload": label, "rows": rows, "rows_per_gpu": per, "dim": dim, "k": k, "batch": batch,
            "metric": "cosine" if metric else "euclidean", "parallelism": f"row-shard x{world}",
            "data_seed": hex(DATA_SEED), "query_seed": hex(QUERY_SEED),
            "l2_policy": f"inputs larger than L2: {4.0 * per * dim / 1e9:.2f} GB scanned per step per GPU vs 126 MB L2"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def measured_bf16_peak():
    """Dense bf16 tensor peak from MEASURED_PEAKS.json: (sustained, burst, source); else the profiling recipe's fallback."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        if "bf16_tflops" in d:
            burst = float(d["bf16_tflops"])
            return float(d.get("bf16_tflops_sustained", burst)), burst, "measured (MEASURED_PEAKS.json: torch.matmul bf16 8192^3)"
    return 1400.0, 1590.0, "fallback (B200_PROFILING.md: 1.59 PFLOP/s burst, ~1.4 sustained)"


def ncu_traffic(key):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/roofline_traffic.json)."""
    p = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get(key)
    return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class ClockSampler:
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md clocks line), sampled every
    50 ms from a background thread through NVML in-process.  (An `nvidia-smi -lms` child process was measured
    to slow the sampled GPU by ~50 % on an 8-GPU box, which then held back every other rank.)"""

    def __init__(self, index):
        self.index, self.samples, self._stop, self._thr, self._nv = index, [], threading.Event(), None, None

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may renumber devices: match by UUID of the torch device when possible
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
            return self

        def loop():
            nv, h = self._nv, self._h
            reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            while not self._stop.is_set():
                try:
                    self.samples.append((nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                         nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM), reasons(h)))
                except Exception:
                    pass
                self._stop.wait(0.05)

        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()
        return self

    def clear(self):
        self.samples = []

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=1.0)
        nv = self._nv
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


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(round(q * (len(xs) - 1))))] if xs else None


# ------------------------------------------------------------------------------------------------------------
# CPU legs (the oracle is test infrastructure: it is only ever the checker or the reported baseline here)
# ------------------------------------------------------------------------------------------------------------
def cpu_oracle_leg(name, steps, warmup, mode, nthreads, budget_s=20.0, sample_bytes=1 << 30, X=None):
    """Time the CPU oracle on a bounded sample of the workload: the first `sample` rows of the same seeded
    collection, min(batch, 4) queries per step.  mode 0 = canonical f64 (the graded arithmetic), 1 = naive f32
    (what an unoptimised Rust loop over Vec<f32> would compute).  `value` is scaled to the full row count;
    ms_per_step_sample is what was measured."""
    from oracle import oracle as O
    rows, dim, metric, k, batch, _ = WORKLOADS[name]
    O.build()
    O.set_threads(host_threads())
    sample = min(rows, max(10_000, sample_bytes // (dim * 4)))
    qb = min(batch, 4)
    if X is None or X.shape[0] != sample:
        X = O.fill(sample, dim, DATA_SEED)
    Q = O.fill(max(steps + warmup, 1) * qb, dim, QUERY_SEED)
    for i in range(warmup):
        O.search(X, Q[i * qb:(i + 1) * qb], k, metric, mode=mode, nthreads=nthreads)
    t_used, times = 0.0, []
    for i in range(steps):
        t0 = time.perf_counter()
        O.search(X, Q[(warmup + i) * qb:(warmup + i + 1) * qb], k, metric, mode=mode, nthreads=nthreads)
        dt = time.perf_counter() - t0
        times.append(dt)
        t_used += dt
        if t_used > budget_s and len(times) >= 3:
            break
    scale = rows / sample
    per_query_full = statistics.mean(times) / qb * scale
    arith = "canonical f64 (graded arithmetic)" if mode == 0 else "naive sequential f32 (plain Rust loop arithmetic)"
    return {"value": 1.0 / per_query_full, "unit": "queries/s", "cores": nthreads, "kind": "port", "arithmetic": arith,
            "sample": f"first {sample} of {rows} rows x {dim} (same Philox stream), {len(times)} steps of {qb} top-{k} "
                      f"queries by oracle/knn_oracle.c, {arith}, {nthreads} thread(s); qps scaled by 1/{scale:g} to the full collection",
            "ms_per_step_sample": statistics.mean(times) * 1e3, "sample_rows": sample, "sample_scale": scale,
            "queries_per_step": qb, "steps": len(times)}, X


def cpu_baseline_block(name, steps=8, warmup=1):
    """The three CPU variants of BASELINE.md section 4; the headline cpu_baseline is the FASTEST all-cores one."""
    nt = host_threads()
    a, X = cpu_oracle_leg(name, steps, warmup, 1, 1, budget_s=8.0)
    b, X = cpu_oracle_leg(name, steps, warmup, 1, nt, budget_s=6.0, X=X)
    c, X = cpu_oracle_leg(name, steps, warmup, 0, nt, budget_s=8.0, X=X)
    keys = ("value", "unit", "cores", "kind", "sample")
    best = b if b["value"] >= c["value"] else c
    variants = [dict({k2: v[k2] for k2 in keys}, variant=nm) for nm, v in
                (("A: 1 thread, naive f32", a), ("B1: all cores, naive f32", b), ("B2: all cores, canonical f64", c))]
    return {k2: best[k2] for k2 in keys}, variants


def run_reference(args, rank):
    """Reference arm: the CPU oracle, all host threads, naive-f32 arithmetic (the closest stand-in for what a Rust loop
    over Vec<Vec<f32>> computes, and the faster of the two CPU arithmetics: the conservative baseline)."""
    if rank != 0:
        return
    name = args.workload
    nt = host_threads()
    leg, _ = cpu_oracle_leg(name, args.steps, args.warmup, 1, nt, budget_s=60.0, sample_bytes=4 << 30)
    line = {"impl": "reference", "metric": METRIC_NAME, "value": leg["value"], "unit": "queries/s", "n_gpus": args.gpus,
            "steps": leg["steps"], "warmup": args.warmup,
            # measured: one step = one bounded sample (sample_rows of the collection); `value` scales it to all rows
            "ms_per_step": leg["ms_per_step_sample"], "sample_scale": leg["sample_scale"], "sample_rows": leg["sample_rows"],
            "ms_per_step_full_collection_scaled": leg["ms_per_step_sample"] * leg["sample_scale"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(name, args.gpus),
            "arithmetic": leg["arithmetic"],
            "note": "sekulas/vRod's SEARCH body is empty (src/command/types.rs:114-119) and rustc is absent: the reference "
                    "arm is the CPU oracle port of the written semantics on all host threads",
            "cpu_baseline": {k2: leg[k2] for k2 in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": leg["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


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


# ------------------------------------------------------------------------------------------------------------
# one workload, both legs, parity, roofline
# ------------------------------------------------------------------------------------------------------------
class Env:
    pass


def make_queries(env, dim, nq):
    """Fresh Philox draws (never rows of X), generated by the library's own device generator in a private
    unsharded context so that every rank holds the same full set."""
    from vrod_b200 import ffi
    qctx = ffi.Context(env.local)
    qcoll = qctx.create("queries", dim, 0, nq)
    qcoll.fill_synthetic(nq, QUERY_SEED)
    q = qcoll.read_rows(0, nq)
    qctx.close()
    return q


def run_workload(env, name, steps, warmup, coll=None, check_queries=1, batched_kind="bf16"):
    """Both legs of one workload on the context env.ctx.  Returns (result dict on rank 0 / None, collection)."""
    import torch
    ctx, dist, rank, world = env.ctx, env.dist, env.rank, env.world
    rows, dim, metric, k, batch, label = WORKLOADS[name]
    if coll is None:
        coll = ctx.create("bench_" + name, dim, metric, rows)
        coll.fill_synthetic(rows, DATA_SEED)
    coll.set_path(4 if (batch > 1 and batched_kind == "tf32") else 0)
    base, local_rows = coll.shard()
    nq = (steps + warmup) * batch
    from_host = make_queries(env, dim, nq)
    q_dev = torch.from_numpy(from_host).cuda()
    ids_dev = torch.empty((steps, batch, k), dtype=torch.int64, device="cuda")     # every timed step keeps its answer
    dist_dev = torch.empty((steps, batch, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    stream = env.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def step_resident(i, slot):
        coll.search_device(q_dev[i * batch:(i + 1) * batch].data_ptr(), batch, k, ids_dev[slot].data_ptr(), dist_dev[slot].data_ptr())

    # ---- resident leg: `value` ----
    sampler = ClockSampler(env.local).start() if rank == 0 else None
    for i in range(warmup):
        step_resident(i, 0)
    barrier()
    if sampler:
        sampler.clear()          # keep only what is sampled from here on (the timed region)
    s0 = ctx.stats()
    ctx.profile(True)
    ctx.profile_read()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record(stream)
    for i in range(steps):
        step_resident(warmup + i, i)
        evs[i + 1].record(stream)
    evs[-1].synchronize()
    barrier()
    ms_total = evs[0].elapsed_time(evs[-1])
    step_ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(steps)]
    kern_ms, kern_n = ctx.profile_read()
    ctx.profile(False)
    s1 = ctx.stats()
    clocks = sampler.stop() if sampler else None
    all_ids = ids_dev.cpu().numpy().astype(np.uint64)
    all_dist = dist_dev.cpu().numpy()

    # ---- e2e leg: host buffers through vrod_collection_search ----
    q_pinned = torch.from_numpy(from_host).pin_memory()
    q_np = q_pinned.numpy()
    for i in range(warmup):
        coll.search(q_np[i * batch:(i + 1) * batch], k)
    barrier()
    se0 = ctx.stats()
    call_ms = []
    t0 = time.perf_counter()
    for i in range(steps):
        j = warmup + i
        tc = time.perf_counter()
        h_ids, h_dist = coll.search(q_np[j * batch:(j + 1) * batch], k)
        call_ms.append((time.perf_counter() - tc) * 1e3)
    barrier()
    e2e_s = time.perf_counter() - t0
    se1 = ctx.stats()
    legs_agree = bool(np.array_equal(h_ids, all_ids[-1]) and np.array_equal(h_dist.view(np.uint32), all_dist[-1].view(np.uint32)))

    # ---- parity, part 1: every rank hashes ALL its timed answers; the hashes must agree ----
    digest = (hashlib.sha256(all_ids.tobytes() + all_dist.tobytes()).hexdigest(), legs_agree)
    digests = [digest]
    if world > 1:
        digests = [None] * world
        dist.all_gather_object(digests, digest)
    ranks_agree = all(d[0] == digests[0][0] for d in digests)
    legs_agree = all(d[1] for d in digests)

    # max over ranks
    t = torch.tensor([ms_total, e2e_s * 1e3, kern_ms / max(kern_n, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, e2e_ms, kern_avg_ms = [float(x) for x in t.tolist()]
    if rank != 0:
        barrier()     # rank 0 replays the oracle now; nobody starts the next (collective) workload before it is back
        return None, coll

    # ---- parity, part 2 (rank 0, outside every timed region): replay queries of the LAST step over the WHOLE
    # collection with the CPU oracle (chunked Philox replay) -- ids and f32 distance bits must be identical ----
    from oracle import oracle as O
    O.build()
    O.set_threads(host_threads())
    nchk = min(check_queries, batch)
    qsel = sorted(set([0, batch - 1][:nchk])) if nchk <= 2 else list(range(nchk))
    last0 = (warmup + steps - 1) * batch
    t_or = time.perf_counter()
    if nchk:
        rid, rdist = O.search_chunked(rows, dim, DATA_SEED, from_host[[last0 + j for j in qsel]], k, metric, nthreads=host_threads())
        oracle_ok = bool(np.array_equal(all_ids[-1][qsel], rid) and
                         np.array_equal(all_dist[-1][qsel].view(np.uint32), rdist.view(np.uint32)))
    else:
        oracle_ok = None
    sorted_ok = bool(np.all(np.diff(all_dist, axis=2) >= 0))
    parity = {"checked_queries": len(qsel) if nchk else 0, "ok": bool(ranks_agree and legs_agree and sorted_ok and oracle_ok is not False),
              "ranks_agree": bool(ranks_agree), "ranks": world, "hashed_steps": steps, "host_and_resident_legs_agree": legs_agree,
              "oracle_replay_ok": oracle_ok, "oracle_replay_s": round(time.perf_counter() - t_or, 2),
              "how": f"sha256 over (ids, dist bits) of all {steps} timed steps equal on all {world} rank(s); queries {qsel} of the last "
                     f"step replayed by oracle/knn_oracle.c over all {rows} rows (4M-row Philox chunks, (dist,id) merge): "
                     "identical ids and identical f32 distance bits required"}

    peak, peak_src = measured_peaks()
    qps = steps * batch / (ms_total / 1e3)
    algo_bytes = 4.0 * local_rows * dim                 # 4*N_local*d per pass, once per batch
    used_batched = (s1["batched_tiles"] - s0["batched_tiles"]) > 0
    tkey = name if world == 1 else f"{name}@{world}"
    traffic = ncu_traffic(tkey)
    res = {
        "workload": name, "value": qps, "unit": "queries/s", "steps": steps, "warmup": warmup, "ms_per_step": ms_total / steps,
        "config": workload_config(name, world), "arithmetic": ARITH_SCAN, "dtype": "f32",
        "e2e": {"value": steps * batch / (e2e_ms / 1e3), "unit": "queries/s",
                "h2d_bytes_per_step": int((se1["h2d_bytes"] - se0["h2d_bytes"]) // max(steps, 1)),
                "d2h_bytes_per_step": int((se1["d2h_bytes"] - se0["d2h_bytes"]) // max(steps, 1)),
                "ms_per_step": e2e_ms / steps},
        "latency": {"resident_ms": {"median": pct(step_ms, 0.5), "p99": pct(step_ms, 0.99), "max": max(step_ms)},
                    "e2e_ms": {"median": pct(call_ms, 0.5), "p99": pct(call_ms, 0.99), "max": max(call_ms)},
                    "note": "per step on rank 0: CUDA events between consecutive resident steps; wall clock around each host-buffer call"},
        "gpu_launches": int(s1["kernel_launches"] - s0["kernel_launches"]),
        "exact_rescans": int(s1["exact_rescans"] - s0["exact_rescans"]),
        "parity": parity, "clocks": clocks,
    }
    if not used_batched:
        achieved = algo_bytes / (kern_avg_ms / 1e3) / 1e9 if kern_avg_ms > 0 else None
        res["roofline"] = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                           "frac": achieved / peak if achieved else None, "traffic": traffic,
                           "traffic_source": ("from_profile: profiles/roofline_traffic.json (ncu --set full, dram__bytes_read.sum + "
                                              "dram__bytes_write.sum of one launch with this shard size; not measured in this run)")
                           if traffic is not None else None,
                           "peak_source": peak_src, "kernel": "fast_scan_kernel", "algorithmic_bytes_per_launch": algo_bytes,
                           "kernel_ms": kern_avg_ms, "launches_timed": int(kern_n)}
    else:
        # batched path: a dense contraction, 2*B*N_local*d flops per step; the bracketed time is the whole
        # tile-kernel sequence (tiles + inter-phase merges) of one batch
        flops = 2.0 * batch * local_rows * dim
        ach = flops / (kern_avg_ms / 1e3) / 1e12
        if batched_kind == "tf32":
            tpeak = measure_tf32_peak()
            tsust, tburst = tpeak, tpeak
            tsrc = "measured here: torch.matmul 8192^3 TF32, best of 10 (burst); the kernel's MMA kind is tf32"
            kname = "batched_tile_kernel (tcgen05.mma kind::tf32 on the stored f32 rows) + inter-phase batched_finish_kernel"
        else:
            tsust, tburst, tsrc = measured_bf16_peak()
            kname = ("batched_tile_kernel (tcgen05.mma kind::f16, bf16 mirror of the rows, thresholds folded into the "
                     "contraction) + inter-phase batched_finish_kernel")
        # denominator: the burst figure when the timed region is short (the GPU has not settled under the power cap),
        # the sustained one for multi-second regions; both fractions are printed
        region_s = ms_total / 1e3
        use_burst = region_s < 2.0
        tpeak = tburst if use_burst else tsust
        res["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s", "frac": ach / tpeak,
                           "peak_kind": "burst" if use_burst else "sustained", "timed_region_s": region_s,
                           "frac_of_burst": ach / tburst, "frac_of_sustained": ach / tsust, "peak_burst": tburst, "peak_sustained": tsust,
                           "traffic": traffic, "traffic_source": "from_profile: profiles/roofline_traffic.json" if traffic is not None else None,
                           "peak_source": tsrc, "kernel": kname, "algorithmic_flops_per_step": flops, "kernel_ms": kern_avg_ms,
                           "launches_timed": int(kern_n), "hbm_floor": {"algorithmic_bytes": algo_bytes, "peak_gbs": peak}}
        res["arithmetic"] = f"{batched_kind} tensor-core pass, exact f64 rerank + guard (bit-identical to the oracle)"
        res["dtype"] = batched_kind      # the type the dominant kernel computes in (f32 accumulate)
    barrier()
    return res, coll


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="vrod_b200", choices=["vrod_b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--rows", type=int, default=0, help="override the row count (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra workloads measured after the headline one")
    ap.add_argument("--batched-kind", default="bf16", choices=["bf16", "tf32"],
                    help="operand mode of the batched (tensor-core) path: bf16 mirror (default) or the f32 rows as tf32")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.rows:
        for nm in (args.workload, "cfg3b" if args.workload == "cfg3" else args.workload):
            w = list(WORKLOADS[nm])
            w[0] = args.rows
            WORKLOADS[nm] = tuple(w)
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
    env = Env()
    env.ctx = ffi.Context(local, rank, world, comm_id)
    env.stream = torch.cuda.ExternalStream(env.ctx.stream())
    env.dist, env.rank, env.world, env.local = dist, rank, world, local

    t_start = time.perf_counter()
    head, coll = run_workload(env, args.workload, args.steps, args.warmup, check_queries=1, batched_kind=args.batched_kind)
    extras = []
    if not args.no_extra:
        for (nm, st, chk) in EXTRA.get(args.workload, []):
            reuse = coll if WORKLOADS[nm][:3] == WORKLOADS[args.workload][:3] else None
            if reuse is None and coll is not None:
                env.ctx.drop(coll.name)
                coll = None
            r, c2 = run_workload(env, nm, st, 3, coll=reuse, check_queries=chk, batched_kind=args.batched_kind)
            if reuse is None:
                env.ctx.drop(c2.name)
            if r is not None:
                extras.append(r)

    if rank == 0:
        line = {"metric": METRIC_NAME, "value": head["value"], "unit": "queries/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic", "config": head["config"],
                "arithmetic": head["arithmetic"], "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
                "exact_rescans": head["exact_rescans"], "roofline": head["roofline"], "parity": head["parity"],
                "latency": head["latency"], "clocks": head["clocks"]}
        if not args.no_cpu_baseline:
            line["cpu_baseline"], line["cpu_baseline_variants"] = cpu_baseline_block(args.workload)
        if extras:
            line["extra"] = extras
        line["bench_wall_s"] = round(time.perf_counter() - t_start, 1)
        if not head["parity"]["ok"] or any(not e["parity"]["ok"] for e in extras):
            print(json.dumps(line), flush=True)
            raise SystemExit("bench.py: PARITY FAILED (see the parity objects of the line above)")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    env.ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
