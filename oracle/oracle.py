"""ctypes binding of the CPU oracle (oracle/knn_oracle.c).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under vrod_b200/ imports this module.
Parity unpinned: see the header of knn_oracle.c (the reference's SearchCommand::execute,
reference src/command/types.rs:114-119, is an empty body).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvrod_oracle.so")

EUCLIDEAN, COSINE = 0, 1
CANONICAL, NAIVE_F32 = 0, 1
PAD_ID = np.uint64(0xFFFFFFFFFFFFFFFF)


def build(force=False):
    src = os.path.join(_HERE, "knn_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        L.vrod_oracle_philox.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.vrod_oracle_fill.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
        L.vrod_oracle_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int]
        L.vrod_oracle_distance.restype = C.c_float
        L.vrod_oracle_search.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_int, C.c_void_p,
                                         C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, C.c_int, C.c_void_p,
                                         C.c_void_p]
        L.vrod_oracle_search.restype = C.c_int
        L.vrod_oracle_merge.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                        C.c_void_p]
        L.vrod_oracle_max_threads.restype = C.c_int
        L.vrod_oracle_set_threads.argtypes = [C.c_int]
        L.vrod_oracle_set_threads.restype = None
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().vrod_oracle_philox(_p(c), _p(k), _p(out))
    return out


def fill(n, d, seed, row0=0):
    """Synthetic rows row0..row0+n of the collection seeded `seed` (uniform [-1, 1), Philox4x32-10)."""
    rows = np.empty((n, d), dtype=np.float32)
    lib().vrod_oracle_fill(_p(rows), row0, n, d, seed)
    return rows


def distance(x, q, metric):
    x = np.ascontiguousarray(x, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    return float(lib().vrod_oracle_distance(_p(x), _p(q), x.shape[0], metric))


def search(rows, queries, k, metric, id_base=0, ids=None, mode=CANONICAL, nthreads=0):
    """Exact top-k: returns (ids [b,k] uint64, dist [b,k] float32) ordered by (dist, id)."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries[None, :]
    n, d = rows.shape if rows.ndim == 2 else (0, queries.shape[1])
    b = queries.shape[0]
    assert queries.shape[1] == d
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.uint64)
    out_ids = np.empty((b, k), dtype=np.uint64)
    out_dist = np.empty((b, k), dtype=np.float32)
    rc = lib().vrod_oracle_search(_p(rows) if n else None, _p(ids), n, d, metric, _p(queries), b, k, id_base, mode,
                                  nthreads, _p(out_ids), _p(out_dist))
    if rc != 0:
        raise ValueError("vrod_oracle_search: invalid arguments")
    return out_ids, out_dist


def merge(ids, dist):
    """Merge per-shard lists [g,b,k] -> [b,k] under the same (dist, id) order."""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    dist = np.ascontiguousarray(dist, dtype=np.float32)
    g, b, k = ids.shape
    out_ids = np.empty((b, k), dtype=np.uint64)
    out_dist = np.empty((b, k), dtype=np.float32)
    lib().vrod_oracle_merge(_p(ids), _p(dist), g, b, k, _p(out_ids), _p(out_dist))
    return out_ids, out_dist


def max_threads():
    return int(lib().vrod_oracle_max_threads())


def set_threads(n):
    """OpenMP team size of fill() and of search(nthreads=0); see vrod_oracle_set_threads."""
    lib().vrod_oracle_set_threads(int(n))


def search_chunked(n, d, seed, queries, k, metric, chunk=4_000_000, nthreads=0):
    """Exact top-k over the synthetic collection (n x d, `seed`) WITHOUT holding it: the Philox stream is replayed
    in `chunk`-row pieces and the per-chunk lists are merged under the same (dist, id) order."""
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries[None, :]
    parts_i, parts_d = [], []
    for lo in range(0, n, chunk):
        X = fill(min(chunk, n - lo), d, seed, row0=lo)
        pi, pd = search(X, queries, k, metric, id_base=lo, nthreads=nthreads)
        parts_i.append(pi)
        parts_d.append(pd)
    return merge(np.stack(parts_i), np.stack(parts_d))
