"""BASELINE configs[4]: d in {64,128,384,768,1536} x k in {1,10,100} at 1M rows, batch 1 and 256.
Every cell is parity-checked against the oracle on a few queries, then timed with queries and results resident
(CUDA events on the library's stream).  Writes gpurun_out/sweep.json and prints a table."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from vrod_b200 import ffi
from oracle import oracle as O

N = int(os.environ.get("SWEEP_ROWS", 1_000_000))
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "..", "MEASURED_PEAKS.json")) else 6650.0
ctx = ffi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
rows = []
for d in (64, 128, 384, 768, 1536):
    for metric in (0, 1):
        c = ctx.create("sw", d, metric, N)
        c.fill_synthetic(N, 0x5EED0001)
        X = O.fill(N, d, 0x5EED0001)
        Qh = O.fill(256, d, 0x5EED0002)
        Q = torch.from_numpy(Qh).cuda()
        for k in (1, 10, 100):
            for b in (1, 256):
                # parity on a sample
                ids, dist = c.search(Qh[:b], k)
                sample = list(range(min(b, 3)))
                rid, rdist = O.search(X, Qh[sample], k, metric)
                ok = bool(np.array_equal(ids[sample], rid) and np.array_equal(dist[sample].view(np.uint32), rdist.view(np.uint32)))
                oi = torch.empty((b, k), dtype=torch.int64, device="cuda"); od = torch.empty((b, k), dtype=torch.float32, device="cuda")
                torch.cuda.synchronize()
                iters = 30 if b == 1 else 10
                for i in range(3): c.search_device(Q[(i % 256) if b == 1 else 0:].data_ptr(), b, k, oi.data_ptr(), od.data_ptr())
                ctx.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s0 = ctx.stats()
                e0.record(stream)
                for i in range(iters): c.search_device(Q[(i % 256) if b == 1 else 0:].data_ptr(), b, k, oi.data_ptr(), od.data_ptr())
                e1.record(stream); e1.synchronize()
                s1 = ctx.stats()
                ms = e0.elapsed_time(e1) / iters
                path = "batched" if s1["batched_tiles"] > s0["batched_tiles"] else "scan"
                r = dict(rows=N, dim=d, metric="cosine" if metric else "euclidean", k=k, batch=b, path=path, ms_per_batch=ms,
                         qps=b / ms * 1e3, parity=ok)
                if path == "scan":
                    r["scan_gbs"] = N * d * 4 / ms / 1e6 * b; r["frac_of_measured_hbm"] = r["scan_gbs"] / PEAK / b
                else:
                    r["tflops"] = 2.0 * b * N * d / ms / 1e9
                rows.append(r)
                print(f"d={d:5d} {r['metric'][:3]} k={k:3d} b={b:3d} {path:7s} {ms*1e3:9.1f} us/batch {r['qps']:10.0f} qps "
                      + (f"{r['frac_of_measured_hbm']*100:5.1f}% of HBM peak" if path == 'scan' else f"{r['tflops']:6.0f} TFLOP/s")
                      + ("" if ok else "  PARITY MISMATCH"), flush=True)
        ctx.drop("sw")
        del X
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/sweep.json", "w"), indent=1)
print("all parity ok:", all(r["parity"] for r in rows))
sys.exit(0 if all(r["parity"] for r in rows) else 1)
