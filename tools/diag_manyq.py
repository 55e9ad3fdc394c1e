"""Development aid: how many queries of a large batch fall back to single scans (guard failures) on the batched path."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vrod_b200 import ffi

ctx = ffi.Context(0)
cases = [(100_000, 64, 8192, 100, 0), (100_000, 64, 1024, 100, 0), (100_000, 64, 8192, 10, 0), (100_000, 128, 8192, 100, 0),
         (1_000_000, 64, 8192, 100, 0), (100_000, 64, 8192, 100, 1), (100_000, 64, 4096, 100, 0), (100_000, 64, 2048, 100, 0)]
if len(sys.argv) > 1:
    cases = cases[:int(sys.argv[1])]
for (n, d, b, k, metric) in cases:
    c = ctx.create("diag", d, metric, n)
    c.fill_synthetic(n, 77)
    c.set_path(3)
    qc = ctx.create("diagq", d, 0, b)
    qc.fill_synthetic(b, 78)
    Q = qc.read_rows(0, b)
    ctx.drop("diagq")
    s0 = ctx.stats()
    t0 = time.time()
    c.search(Q, k)
    s1 = ctx.stats()
    print(f"n={n} d={d} b={b} k={k} metric={metric}: rescanned {s1['fast_scans'] - s0['fast_scans']} of {b}, "
          f"tiles {s1['batched_tiles'] - s0['batched_tiles']}, {time.time() - t0:.2f} s", flush=True)
    ctx.drop("diag")
ctx.close()
