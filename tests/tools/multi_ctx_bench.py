"""Single-process multi-GPU context (vrod_ctx_create_multi) on all visible GPUs: configs[3] -- 100M x 128 f32 L2 top-10 -- through
the host-buffer call a vRod SearchCommand would make (one thread, host query in, host ids / distances out), parity-checked
against a full oracle replay; then the same collection with 1024-query batches.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from oracle import oracle as O
from vrod_b200 import ffi

ndev = torch.cuda.device_count()
rows, dim, k = int(os.environ.get("ROWS", 100_000_000)), 128, 10
O.build()
O.set_threads(len(os.sched_getaffinity(0)))
ctx = ffi.Context(list(range(ndev)))
c = ctx.create("cfg3", dim, 0, rows)
t0 = time.perf_counter()
c.fill_synthetic(rows, 0x5EED0001)
fill_s = time.perf_counter() - t0
Q = O.fill(400, dim, 0x5EED0002)
for i in range(10):
    c.search(Q[i], k)
lat = []
for i in range(10, 310):
    t = time.perf_counter()
    ids, dist = c.search(Q[i], k)
    lat.append((time.perf_counter() - t) * 1e3)
lat.sort()
rid, rdist = O.search_chunked(rows, dim, 0x5EED0001, Q[309], k, 0, nthreads=len(os.sched_getaffinity(0)))
ok1 = bool(np.array_equal(ids, rid) and np.array_equal(dist.view(np.uint32), rdist.view(np.uint32)))
# batched: 1024 queries per call
QB = O.fill(1024 * 6, dim, 0x5EED0003).reshape(6, 1024, dim)
c.search(QB[0], k)
tb = []
for i in range(1, 6):
    t = time.perf_counter()
    bids, bdist = c.search(QB[i], k)
    tb.append((time.perf_counter() - t) * 1e3)
sel = [0, 1023]
rid2, rdist2 = O.search_chunked(rows, dim, 0x5EED0001, QB[5][sel], k, 0, nthreads=len(os.sched_getaffinity(0)))
ok2 = bool(np.array_equal(bids[sel], rid2) and np.array_equal(bdist[sel].view(np.uint32), rdist2.view(np.uint32)))
st = ctx.stats()
print(json.dumps({"what": "single-process multi-GPU context, host-buffer search", "devices": ndev, "rows": rows, "dim": dim, "k": k,
                  "fill_s": round(fill_s, 2),
                  "single_query_ms": {"mean": sum(lat) / len(lat), "median": lat[len(lat) // 2], "p99": lat[int(len(lat) * 0.99)], "calls": len(lat)},
                  "single_query_qps": 1e3 / (sum(lat) / len(lat)),
                  "batch1024_ms": {"mean": sum(tb) / len(tb), "min": min(tb)}, "batch1024_qps": 1024e3 / (sum(tb) / len(tb)),
                  "parity": {"single_query_vs_oracle_full_replay": ok1, "batched_two_queries_vs_oracle_full_replay": ok2},
                  "kernel_launches": st["kernel_launches"], "batched_tiles": st["batched_tiles"]}), flush=True)
ctx.close()
sys.exit(0 if ok1 and ok2 else 1)
