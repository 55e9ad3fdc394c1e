"""Randomised parity soak of the batched (tensor-core) path on data that stresses the guard: clustered rows,
normalised embeddings, integer grids (exact distance ties), sparse and low-rank rows.  Every case is compared
bit for bit with the CPU oracle; prints how many queries each case had to rescan."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from vrod_b200 import ffi
from oracle import oracle as O

def make(kind, n, d, rng):
    if kind == "clusters":
        c = rng.standard_normal((max(n // 200, 2), d)).astype(np.float32)
        X = c[rng.integers(0, len(c), n)] + (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    elif kind == "normalized":
        X = rng.standard_normal((n, d)).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True).astype(np.float32)
    elif kind == "grid":
        X = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    elif kind == "sparse":
        X = (rng.standard_normal((n, d)) * (rng.random((n, d)) < 0.05)).astype(np.float32)
    elif kind == "lowrank":
        X = (rng.standard_normal((n, 4)) @ rng.standard_normal((4, d))).astype(np.float32)
    else:
        X = rng.uniform(-1, 1, size=(n, d)).astype(np.float32)
    return np.ascontiguousarray(X, dtype=np.float32)

def main():
    seed = int(os.environ.get("SOAK_SEED", "1"))
    ncase = int(os.environ.get("SOAK_CASES", "36"))
    rng = np.random.default_rng(seed)
    ctx = ffi.Context(0)
    bad = 0
    t0 = time.time()
    for i in range(ncase):
        kind = ["clusters", "normalized", "grid", "sparse", "lowrank", "uniform"][i % 6]
        d = int(rng.choice([16, 48, 64, 100, 128, 256, 384]))
        n = int(rng.integers(2000, 120000))
        b = int(rng.choice([64, 130, 256, 700]))
        k = int(rng.choice([1, 10, 50, 100, 120]))
        metric = int(rng.integers(0, 2))
        path = 3 if i % 4 else 4
        X = make(kind, n, d, rng)
        Q = make(kind, b, d, rng)
        if kind in ("clusters", "grid"):
            Q[: b // 2] = X[rng.integers(0, n, b // 2)]          # queries sitting on rows
        c = ctx.create(f"soak{i}", d, metric, n)
        c.insert(X)
        c.set_path(path)
        s0 = ctx.stats()
        ids, dist = c.search(Q, k)
        s1 = ctx.stats()
        rid, rdist = O.search(X, Q, k, metric)
        ok = np.array_equal(ids, rid) and np.array_equal(dist.view(np.uint32), rdist.view(np.uint32))
        wrong = int((ids != rid).any(axis=1).sum()) if not ok else 0
        print(f"case {i:2d} {kind:10s} n={n:6d} d={d:3d} b={b:3d} k={k:3d} metric={metric} path={path}: "
              f"{'OK' if ok else 'MISMATCH in %d queries' % wrong} batched={s1['batched_tiles'] > s0['batched_tiles']} "
              f"rescanned={s1['fast_scans'] - s0['fast_scans']}", flush=True)
        if not ok:
            wq = np.where((ids != rid).any(axis=1) | (dist.view(np.uint32) != rdist.view(np.uint32)).any(axis=1))[0]
            q0 = int(wq[0])
            j0 = int(np.where((ids[q0] != rid[q0]) | (dist[q0].view(np.uint32) != rdist[q0].view(np.uint32)))[0][0])
            print(f"   wrong queries {wq[:12].tolist()}; query {q0} first differs at rank {j0}: got id {ids[q0, j0]} dist {dist[q0, j0]!r}, "
                  f"want id {rid[q0, j0]} dist {rdist[q0, j0]!r}; |q|^2 = {float((Q[q0].astype(np.float64) ** 2).sum())!r}; "
                  f"got-id in wanted list: {int(ids[q0, j0]) in set(rid[q0].tolist())}, wanted-id in got list: {int(rid[q0, j0]) in set(ids[q0].tolist())}", flush=True)
            lo, hi = max(0, j0 - 2), min(k, j0 + 4)
            print("   got ", ids[q0, lo:hi].tolist(), dist[q0, lo:hi].tolist())
            print("   want", rid[q0, lo:hi].tolist(), rdist[q0, lo:hi].tolist(), flush=True)
        bad += 0 if ok else 1
        ctx.drop(c.name)
    print(f"soak done: {ncase} cases, {bad} mismatching, {time.time() - t0:.0f} s")
    sys.exit(1 if bad else 0)

main()
