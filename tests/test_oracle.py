"""CPU tests of the oracle (oracle/knn_oracle.c) -- the checker every GPU parity test relies on.

The reference has no golden vectors for SEARCH: the oracle is pinned against the
published Philox4x32-10 known-answer vectors (Random123 kat_vectors), the authored hand KATs and an
independent NumPy float64 brute force (tests/golden/kat.json, made by tests/golden/make_golden.py).
"""
import numpy as np
import pytest

from tests.util import assert_same, load_kat

KAT = load_kat()


def test_philox_known_answers(oracle):
    # Random123 kat_vectors, philox4x32 10 rounds
    assert [hex(x) for x in oracle.philox([0, 0, 0, 0], [0, 0])] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    assert [hex(x) for x in oracle.philox([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2)] == \
        ["0x408f276d", "0x41c83b0e", "0xa20bc7c6", "0x6d5451fd"]
    assert [hex(x) for x in oracle.philox([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0])] == \
        ["0xd16cfe09", "0x94fdcceb", "0x5001e420", "0x24126ea1"]


def test_fill_is_positional_and_uniform(oracle):
    X = oracle.fill(1000, 128, 0x5EED0001)
    assert X.dtype == np.float32 and X.min() >= -1.0 and X.max() < 1.0
    assert abs(float(X.mean())) < 0.01 and abs(float(X.var()) - 1.0 / 3.0) < 0.01
    # any slice regenerates bit-identically (what lets the CPU replay a 51 GB collection in chunks)
    assert np.array_equal(oracle.fill(10, 128, 0x5EED0001, row0=500), X[500:510])
    # dims that are not a multiple of the 4-word Philox block still address elements by i*d + j
    Y = oracle.fill(50, 7, 3)
    assert np.array_equal(oracle.fill(3, 7, 3, row0=20), Y[20:23])
    # element (i, j) is word (i*d+j)&3 of block (i*d+j)>>2
    e = 13 * 7 + 5
    w = oracle.philox([e >> 2, 0, 0, 0], [3, 0])[e & 3]
    assert Y[13, 5] == np.float32((int(w >> 8) - (1 << 23)) * 2.0 ** -23)


@pytest.mark.parametrize("case", KAT, ids=[c["name"] for c in KAT])
def test_golden(oracle, case):
    ids, dist = oracle.search(case["rows"], case["query"], case["k"], case["metric"])
    assert_same(ids[0], dist[0], case["ids"], case["dist"], case["name"])


@pytest.mark.parametrize("metric", [0, 1])
def test_matches_numpy_f64_bruteforce(oracle, metric):
    X = oracle.fill(4000, 96, 11)
    Q = oracle.fill(5, 96, 12)
    ids, dist = oracle.search(X, Q, 20, metric)
    Xd, Qd = X.astype(np.float64), Q.astype(np.float64)
    if metric == 0:
        D = np.sqrt(((Xd[None] - Qd[:, None]) ** 2).sum(-1))
    else:
        D = 1.0 - (Qd @ Xd.T) / np.linalg.norm(Qd, axis=1)[:, None] / np.linalg.norm(Xd, axis=1)[None]
    D32 = D.astype(np.float32)
    for i in range(5):
        order = np.lexsort((np.arange(4000), D32[i]))[:20]
        assert np.array_equal(order.astype(np.uint64), ids[i])
        assert np.allclose(D[i, order], dist[i], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("metric", [0, 1])
def test_naive_f32_mode_within_tolerance(oracle, metric):
    """What a plain single-threaded f32 Rust loop would return stays within 1e-5 relative (north_star)."""
    X = oracle.fill(3000, 768, 21)
    q = oracle.fill(1, 768, 22)
    ids, dist = oracle.search(X, q, 10, metric)
    nid, ndist = oracle.search(X, q, 10, metric, mode=oracle.NAIVE_F32, nthreads=1)
    assert np.allclose(ndist, dist, rtol=1e-5)
    assert len(set(ids[0]) & set(nid[0])) >= 9


def test_threads_and_shards_do_not_change_the_answer(oracle):
    X = oracle.fill(5000, 64, 31)
    X[100] = X[4000]                      # a tie across shards
    Q = np.vstack([oracle.fill(3, 64, 32), X[4000][None]])
    for metric in (0, 1):
        one = oracle.search(X, Q, 16, metric, nthreads=1)
        many = oracle.search(X, Q, 16, metric, nthreads=0)
        assert_same(*one, *many)
        for g in (2, 3, 8):
            per = (5000 + g - 1) // g
            parts = [oracle.search(X[s * per:(s + 1) * per], Q, 16, metric, id_base=s * per) for s in range(g)]
            merged = oracle.merge(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]))
            assert_same(*merged, *one, f"g={g}")


def test_padding_and_empty(oracle):
    X = oracle.fill(3, 8, 1)
    ids, dist = oracle.search(X, X[1], 5, 0)
    assert ids[0, 0] == 1 and dist[0, 0] == 0.0
    assert list(ids[0, 3:]) == [oracle.PAD_ID] * 2 and np.all(np.isinf(dist[0, 3:]))
    ids, dist = oracle.search(np.zeros((0, 8), np.float32), X[:2], 3, 1)
    assert np.all(ids == oracle.PAD_ID) and np.all(np.isinf(dist))
