"""Generates tests/golden/kat.json.

The reference (sekulas/vRod) holds no test, golden vector or fixture for SEARCH (its body is empty,
reference src/command/types.rs:114-119), so these vectors are AUTHORED here:
  * `hand`: small cases whose answers are checkable by hand (3-4-5 triangles, unit vectors, ties).
  * `numpy`: random small cases answered by an independent NumPy float64 brute force (no code
    shared with oracle/knn_oracle.c); distances stored as the f32 bit pattern,
    rows and queries as base64 of little-endian f32.
Run:  python tests/golden/make_golden.py   (deterministic; rewrites kat.json)
"""
import base64
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
INF = float("inf")
PAD = 0xFFFFFFFFFFFFFFFF


def f32bits(a):
    return [int(x) for x in np.asarray(a, dtype=np.float32).view(np.uint32).ravel()]


def f32b64(a):
    return base64.b64encode(np.ascontiguousarray(a, dtype="<f4").tobytes()).decode()


def brute(rows, q, k, metric):
    X = np.asarray(rows, dtype=np.float32).astype(np.float64)
    v = np.asarray(q, dtype=np.float32).astype(np.float64)
    if metric == "euclidean":
        d = np.sqrt(((X - v[None, :]) ** 2).sum(axis=1))
    else:
        nx = np.sqrt((X * X).sum(axis=1))
        nq = np.sqrt((v * v).sum())
        with np.errstate(divide="ignore", invalid="ignore"):
            sim = (X @ v) / (nx * nq)
        sim = np.where((nx == 0) | (nq == 0), 0.0, sim)
        d = 1.0 - sim
    d32 = d.astype(np.float32)
    order = np.lexsort((np.arange(len(d32)), d32))[:k]
    ids = [int(i) for i in order] + [PAD] * (k - len(order))
    dist = [float(d32[i]) for i in order] + [INF] * (k - len(order))
    return ids, dist


def main():
    hand = [
        dict(name="l2_345", metric="euclidean", rows=[[0, 0], [3, 4], [6, 8], [0, 5]], query=[0, 0], k=3,
             ids=[0, 1, 3], dist=[0.0, 5.0, 5.0]),
        dict(name="l2_345_all", metric="euclidean", rows=[[0, 0], [3, 4], [6, 8], [0, 5]], query=[0, 0], k=4,
             ids=[0, 1, 3, 2], dist=[0.0, 5.0, 5.0, 10.0]),
        dict(name="l2_3d_pythagorean", metric="euclidean", rows=[[2, 3, 6], [1, 4, 8], [4, 4, 7], [0, 0, 0]],
             query=[0, 0, 0], k=4, ids=[3, 0, 1, 2], dist=[0.0, 7.0, 9.0, 9.0]),
        dict(name="l2_self_match", metric="euclidean", rows=[[1.5, -2.25], [0.5, 0.5], [1.5, -2.25]],
             query=[1.5, -2.25], k=2, ids=[0, 2], dist=[0.0, 0.0]),
        dict(name="cos_units", metric="cosine", rows=[[1, 0], [0, 1], [-1, 0], [0, 0], [2, 0]], query=[1, 0], k=5,
             ids=[0, 4, 1, 3, 2], dist=[0.0, 0.0, 1.0, 1.0, 2.0]),
        dict(name="cos_zero_query", metric="cosine", rows=[[1, 0], [0, 1], [3, 4]], query=[0, 0], k=2,
             ids=[0, 1], dist=[1.0, 1.0]),
        dict(name="ties_straddle_k", metric="euclidean", rows=[[1, 1]] * 6 + [[0, 0]], query=[1, 1], k=4,
             ids=[0, 1, 2, 3], dist=[0.0, 0.0, 0.0, 0.0]),
        dict(name="k_gt_n_pads", metric="euclidean", rows=[[3, 4], [0, 0]], query=[0, 0], k=4,
             ids=[1, 0, PAD, PAD], dist=[0.0, 5.0, INF, INF]),
        dict(name="empty_collection", metric="euclidean", rows=[], query=[0, 0], k=2, dim=2,
             ids=[PAD, PAD], dist=[INF, INF]),
        dict(name="k1", metric="cosine", rows=[[0, 2], [5, 5], [1, 0]], query=[0, 1], k=1, ids=[0], dist=[0.0]),
    ]
    for h in hand:
        h["dist_bits"] = f32bits(h["dist"])
        h["dist"] = [("inf" if d == INF else d) for d in h["dist"]]

    rng = np.random.default_rng(20261018)
    cases = []
    for (n, d, k) in [(200, 17, 7), (64, 128, 10), (300, 3, 25), (33, 64, 40), (60, 260, 5), (128, 1, 4)]:
        for metric in ("euclidean", "cosine"):
            rows = rng.standard_normal((n, d)).astype(np.float32)
            if n > 8:
                rows[5] = rows[2]          # duplicate rows: exact ties
                rows[7] = 0.0              # a zero row (cosine rule)
            q = rng.standard_normal(d).astype(np.float32)
            ids, dist = brute(rows, q, k, metric)
            cases.append(dict(name=f"np_{metric}_{n}x{d}_k{k}", metric=metric, n=n, dim=d, k=k,
                              rows_b64=f32b64(rows), query_b64=f32b64(q), ids=ids, dist_bits=f32bits(dist)))
    out = dict(about="authored KATs for vRod SEARCH semantics; see make_golden.py", hand=hand, numpy=cases)
    with open(os.path.join(HERE, "kat.json"), "w") as f:
        json.dump(out, f)
    print("wrote kat.json:", len(hand), "hand cases,", len(cases), "numpy cases")


if __name__ == "__main__":
    main()
