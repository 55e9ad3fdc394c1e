"""Quick parity + timing sweep on one GPU (development aid; the graded tests are tests/test_gpu_*.py)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from vrod_b200 import ffi
from oracle import oracle as O

ctx = ffi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
bad = 0
def check(n, d, metric, k, b=3, seed=1, path=0):
    global bad
    name = f"c_{n}_{d}_{metric}_{k}_{path}"
    c = ctx.create(name, d, metric, max(n, 1))
    c.fill_synthetic(n, seed)
    c.set_path(path)
    X = O.fill(n, d, seed)
    if n: assert np.array_equal(c.read_rows(0, min(n, 100)), X[:100]), "fill mismatch"
    q = O.fill(b, d, seed + 1000)
    ids, dist = c.search(q, k)
    rid, rdist = O.search(X, q, k, metric)
    ok = np.array_equal(ids, rid) and np.array_equal(dist.view(np.uint32), rdist.view(np.uint32))
    st = ctx.stats()
    print(f"n={n} d={d} metric={metric} k={k} path={path}: {'OK' if ok else 'MISMATCH'} rescans={st['exact_rescans']}", flush=True)
    if not ok:
        bad += 1
        print(ids[0][:10], rid[0][:10]); print(dist[0][:10], rdist[0][:10])
    ctx.drop(name)

for path in (0, 2):
    for (n, d) in [(10000, 128), (1000, 64), (5000, 32), (3000, 768), (2000, 1536), (1234, 100), (777, 3), (50, 128), (5, 128), (0, 16), (20000, 384), (100000, 256)]:
        for metric in (0, 1):
            for k in (1, 10, 100):
                check(n, d, metric, k, path=path)

def timeit(n, d, metric, k, iters=20):
    c = ctx.create("t", d, metric, n)
    c.fill_synthetic(n, 7)
    q = torch.from_numpy(O.fill(iters, d, 8)).cuda()
    ids = torch.empty((1, k), dtype=torch.int64, device="cuda"); dist = torch.empty((1, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for i in range(3): c.search_device(q[i].data_ptr(), 1, k, ids.data_ptr(), dist.data_ptr())
    ctx.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(iters): c.search_device(q[i].data_ptr(), 1, k, ids.data_ptr(), dist.data_ptr())
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / iters
    gbs = n * d * 4 / ms / 1e6
    print(f"time n={n} d={d} metric={metric} k={k}: {ms*1000:.1f} us/query  {gbs:.0f} GB/s  ({gbs/6548.8*100:.1f}% of measured copy peak) rescans={ctx.stats()['exact_rescans']}", flush=True)
    ctx.drop("t")

for (n, d, m) in [(1000000, 768, 1), (1000000, 128, 0), (1000000, 64, 0), (1000000, 384, 1), (1000000, 1536, 0), (10000000, 128, 0)]:
    for k in (10, 100):
        timeit(n, d, m, k)
print("bad =", bad)
sys.exit(1 if bad else 0)
