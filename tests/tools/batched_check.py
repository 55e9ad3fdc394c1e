"""Development aid: batched (tensor-core) path parity + timing on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import torch
from vrod_b200 import ffi
from oracle import oracle as O

ctx = ffi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
bad = 0
BPATH = int(os.environ.get("BPATH", "3"))   # 3 = batched (bf16 mirror), 4 = batched tf32
def check(n, d, metric, k, b, seed=1):
    global bad
    name = f"c_{n}_{d}_{metric}"
    c = ctx.create(name, d, metric, max(n, 1)); c.fill_synthetic(n, seed); c.set_path(BPATH)
    X = O.fill(n, d, seed); Q = O.fill(b, d, seed + 1000)
    s0 = ctx.stats()
    ids, dist = c.search(Q, k)
    s1 = ctx.stats()
    rid, rdist = O.search(X, Q, k, metric)
    ok = np.array_equal(ids, rid) and np.array_equal(dist.view(np.uint32), rdist.view(np.uint32))
    print(f"n={n} d={d} metric={metric} k={k} b={b}: {'OK' if ok else 'MISMATCH'} tiles={s1['batched_tiles']-s0['batched_tiles']} rescanned={s1['fast_scans']-s0['fast_scans']}", flush=True)
    if not ok:
        bad += 1
        wrong = np.where((ids != rid).any(axis=1))[0]
        print("  wrong queries:", wrong[:10], "of", len(wrong))
        if len(wrong): print(ids[wrong[0]][:10], rid[wrong[0]][:10]); print(dist[wrong[0]][:10], rdist[wrong[0]][:10])
    ctx.drop(name)

cases = [(1000, 128, 0, 10, 4), (50000, 128, 0, 10, 300), (50000, 128, 1, 100, 256), (100, 128, 0, 10, 7), (20000, 64, 0, 10, 513),
         (20000, 100, 1, 10, 64), (30000, 32, 0, 100, 100), (200000, 128, 0, 100, 1024), (300000, 96, 1, 10, 1000)]
if len(sys.argv) > 1 and sys.argv[1] == "first":
    cases = cases[:1]
if len(sys.argv) > 1 and (sys.argv[1].startswith("prof") or sys.argv[1] == "one"):
    cases = []
for cs in cases:
    check(*cs)

def timeit(n, d, metric, k, b, iters=3):
    c = ctx.create("t", d, metric, n); c.fill_synthetic(n, 7); c.set_path(BPATH)
    q = torch.from_numpy(O.fill(b, d, 8)).cuda()
    ids = torch.empty((b, k), dtype=torch.int64, device="cuda"); dist = torch.empty((b, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    c.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr()); ctx.synchronize()
    ctx.profile(True); ctx.profile_read()
    s0 = ctx.stats()
    t0 = time.perf_counter()
    for i in range(iters): c.search_device(q.data_ptr(), b, k, ids.data_ptr(), dist.data_ptr())
    ctx.synchronize()
    dt = (time.perf_counter() - t0) / iters
    kms, kn = ctx.profile_read(); ctx.profile(False)
    s1 = ctx.stats()
    tf = 2.0 * b * n * d / (kms / kn / 1e3) / 1e12
    print(f"time n={n} d={d} metric={metric} k={k} b={b}: {dt*1e3:.2f} ms/batch {b/dt:.0f} qps; tile kernel {kms/kn:.3f} ms = {tf:.0f} TFLOP/s (path {BPATH}); rescanned={s1['fast_scans']-s0['fast_scans']}", flush=True)
    ctx.drop("t")
if len(sys.argv) > 1 and sys.argv[1] == "prof":
    timeit(1000000, 128, 0, 100, 1024, iters=2)
elif len(sys.argv) > 1 and sys.argv[1] == "one":      # one n d metric k b
    timeit(*[int(a) for a in sys.argv[2:7]])
elif len(sys.argv) > 1 and sys.argv[1] == "prof10":
    timeit(10000000, 128, 0, 100, 1024, iters=1)
elif bad == 0 and not (len(sys.argv) > 1 and sys.argv[1] == "first"):
    timeit(1000000, 128, 0, 10, 256)
    timeit(1000000, 64, 0, 10, 256)
    timeit(1000000, 384, 1, 10, 256)
    timeit(1000000, 768, 1, 10, 256)
    timeit(1000000, 1536, 0, 100, 256)
    timeit(10000000, 128, 0, 100, 1024)
    timeit(10000000, 128, 1, 10, 1024)
print("bad =", bad)
sys.exit(1 if bad else 0)
