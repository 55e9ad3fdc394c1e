"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck): single-query scans (both metrics, the
exact scan, a generic dimension), the batched tcgen05 path in bf16 and tf32 operand modes (dense start phase, stash,
finish kernels, band mode), inserts with growth.  Answers are checked against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import oracle as O
from vrod_b200 import ffi

def same(a, b):
    return np.array_equal(a[0], b[0]) and np.array_equal(a[1].view(np.uint32), b[1].view(np.uint32))

O.build()
with ffi.Context(0) as ctx:
    for (n, d, metric, k) in [(3000, 128, 0, 10), (3000, 768, 1, 10), (2000, 100, 1, 7), (2500, 64, 0, 100)]:
        c = ctx.create(f"s{d}", d, metric, n)
        c.fill_synthetic(n, 5)
        X, Q = O.fill(n, d, 5), O.fill(2, d, 6)
        assert same(c.search(Q, k), O.search(X, Q, k, metric)), ("scan", n, d)
        c.set_path(2)
        assert same(c.search(Q[:1], k), O.search(X, Q[:1], k, metric)), ("exact", n, d)
        ctx.drop(c.name)
    for (n, d, metric, k, b, path) in [(6000, 128, 0, 10, 64, 3), (5000, 96, 1, 100, 300, 3), (4000, 128, 0, 10, 40, 4), (3000, 400, 1, 5, 33, 3)]:
        c = ctx.create(f"b{d}_{path}", d, metric, n)
        c.fill_synthetic(n, 7)
        c.set_path(path)
        X, Q = O.fill(n, d, 7), O.fill(b, d, 8)
        assert same(c.search(Q, k), O.search(X, Q, k, metric)), ("batched", n, d, path)
        ctx.drop(c.name)
    rng = np.random.default_rng(3)
    cen = rng.standard_normal((20, 64)).astype(np.float32)
    X = (cen[rng.integers(0, 20, 4000)] + rng.standard_normal((4000, 64)).astype(np.float32) * 0.01).astype(np.float32)
    c = ctx.create("clu", 64, 0, 1000)             # grows twice
    c.insert(X[:1500]); c.insert(X[1500:])
    c.set_path(3)
    Q = X[:80] + np.float32(1e-4)
    assert same(c.search(Q, 10), O.search(X, Q, 10, 0)), "band mode"
    print("sanitize case ok:", ctx.stats())
