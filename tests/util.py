import base64
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
METRIC = {"euclidean": 0, "cosine": 1}


def load_kat():
    with open(os.path.join(HERE, "golden", "kat.json")) as f:
        kat = json.load(f)
    cases = []
    for h in kat["hand"]:
        dim = h.get("dim", len(h["query"]))
        rows = np.asarray(h["rows"], dtype=np.float32).reshape(-1, dim)
        cases.append(dict(name=h["name"], metric=METRIC[h["metric"]], rows=rows,
                          query=np.asarray(h["query"], dtype=np.float32), k=h["k"],
                          ids=np.asarray(h["ids"], dtype=np.uint64),
                          dist=np.asarray(h["dist_bits"], dtype=np.uint32).view(np.float32)))
    for c in kat["numpy"]:
        rows = np.frombuffer(base64.b64decode(c["rows_b64"]), dtype="<f4").reshape(c["n"], c["dim"]).copy()
        q = np.frombuffer(base64.b64decode(c["query_b64"]), dtype="<f4").copy()
        cases.append(dict(name=c["name"], metric=METRIC[c["metric"]], rows=rows, query=q, k=c["k"],
                          ids=np.asarray(c["ids"], dtype=np.uint64),
                          dist=np.asarray(c["dist_bits"], dtype=np.uint32).view(np.float32)))
    return cases


def assert_same(ids, dist, rids, rdist, what=""):
    """Bit-exact: identical ids and identical f32 distance bit patterns."""
    ids = np.asarray(ids, dtype=np.uint64)
    rids = np.asarray(rids, dtype=np.uint64)
    d = np.asarray(dist, dtype=np.float32)
    rd = np.asarray(rdist, dtype=np.float32)
    assert ids.shape == rids.shape, what
    assert np.array_equal(ids, rids), f"{what}: ids differ\n{ids}\n{rids}"
    assert np.array_equal(d.view(np.uint32), rd.view(np.uint32)), f"{what}: distances differ\n{d}\n{rd}"
