"""GPU parity tests of the batched-query path (tcgen05 tiles + exact f64 rerank + guard) against the CPU oracle:
bit-exact ids and distances, like the scan path.  set_path(3) forces the batched path in its default operand mode
(bf16 mirror of the rows with the thresholds folded into the contraction), set_path(4) the tf32 mode that feeds the
stored f32 rows."""
import numpy as np
import pytest

from tests.util import assert_same

pytestmark = pytest.mark.gpu

CASES = [  # n, d, metric, k, b
    (1000, 128, 0, 10, 4), (50000, 128, 0, 10, 300), (50000, 128, 1, 100, 256), (100, 128, 0, 10, 7),
    (20000, 64, 0, 10, 513), (20000, 100, 1, 10, 64), (30000, 32, 0, 100, 100), (200000, 128, 0, 100, 1024),
    (300000, 96, 1, 10, 1000), (129, 8, 0, 1, 3), (4000, 128, 1, 120, 33),
    # fewer rows than k (padded answers), a single row
    (5, 16, 1, 10, 70), (1, 128, 0, 3, 65),
    # dims above 128: the query slabs are streamed with the row slabs
    (30000, 384, 1, 10, 256), (20000, 768, 1, 10, 70), (9000, 1536, 0, 100, 300), (15000, 200, 0, 10, 40), (5000, 260, 1, 5, 9),
]


@pytest.mark.parametrize("path", [3, 4], ids=["bf16", "tf32"])
@pytest.mark.parametrize("n,d,metric,k,b", CASES)
def test_batched_matches_oracle(ctx, oracle, n, d, metric, k, b, path):
    c = ctx.create(f"b{n}_{d}_{metric}_{k}_{path}", d, metric, n)
    c.fill_synthetic(n, 300 + d)
    c.set_path(path)
    X = oracle.fill(n, d, 300 + d)
    Q = oracle.fill(b, d, 400 + d)
    s0 = ctx.stats()
    ids, dist = c.search(Q, k)
    s1 = ctx.stats()
    assert s1["batched_tiles"] > s0["batched_tiles"], "the tensor-core path did not run"
    assert_same(ids, dist, *oracle.search(X, Q, k, metric), f"n={n} d={d} metric={metric} k={k} b={b}")
    ctx.drop(c.name)


def test_batched_is_chosen_automatically_for_large_batches(ctx, oracle):
    n, d = 60000, 128
    c = ctx.create("auto_b", d, 0, n)
    c.fill_synthetic(n, 5)
    Q = oracle.fill(128, d, 6)
    s0 = ctx.stats()
    ids, dist = c.search(Q, 10)
    s1 = ctx.stats()
    assert s1["batched_tiles"] > s0["batched_tiles"] and s1["fast_scans"] == s0["fast_scans"]
    assert_same(ids, dist, *oracle.search(oracle.fill(n, d, 5), Q, 10, 0))
    # on a small collection a handful of queries is cheaper as single-query scans (cost model in vrod_capi.cu)
    c2 = ctx.create("auto_b2", 768, 1, 3000)
    c2.fill_synthetic(3000, 7)
    Q2 = oracle.fill(6, 768, 8)
    s0 = ctx.stats()
    ids, dist = c2.search(Q2, 5)
    s1 = ctx.stats()
    assert s1["batched_tiles"] == s0["batched_tiles"] and s1["fast_scans"] - s0["fast_scans"] == 6
    assert_same(ids, dist, *oracle.search(oracle.fill(3000, 768, 7), Q2, 5, 1))
    # ... while on a large one even 8 queries go through the tensor cores
    c3 = ctx.create("auto_b3", 128, 0, 2_000_000)
    c3.fill_synthetic(2_000_000, 9)
    Q3 = oracle.fill(8, 128, 10)
    s0 = ctx.stats()
    ids, dist = c3.search(Q3, 10)
    s1 = ctx.stats()
    assert s1["batched_tiles"] > s0["batched_tiles"]
    assert_same(ids, dist, *oracle.search(oracle.fill(2_000_000, 128, 9), Q3, 10, 0))
    ctx.drop("auto_b3")
    ctx.drop("auto_b")
    ctx.drop("auto_b2")


def test_mirror_follows_appends_and_growth(ctx, oracle):
    """The bf16 mirror is built by the first batched search; rows inserted afterwards (including an insert that
    makes the collection grow and reallocate) must be in it at the next search."""
    d, k = 96, 10
    X = oracle.fill(9000, d, 31)
    Q = oracle.fill(130, d, 32)
    for metric in (0, 1):
        c = ctx.create(f"mir{metric}", d, metric, 4000)
        c.set_path(3)
        c.insert(X[:3000])
        assert_same(*c.search(Q, k), *oracle.search(X[:3000], Q, k, metric), "first build")
        c.insert(X[3000:3900])
        assert_same(*c.search(Q, k), *oracle.search(X[:3900], Q, k, metric), "appended rows")
        c.insert(X[3900:])          # beyond the capacity: the collection grows, the mirror is rebuilt
        assert_same(*c.search(Q, k), *oracle.search(X, Q, k, metric), "after growth")
        ctx.drop(c.name)


@pytest.mark.parametrize("path", [3, 4], ids=["bf16", "tf32"])
@pytest.mark.parametrize("metric", [0, 1], ids=["euclidean", "cosine"])
def test_batched_awkward_inputs(ctx, oracle, metric, path):
    """Zero rows, a zero query, a query outside the fast paths' range, an exact hit, and row norms spread over six
    orders of magnitude (the bf16 mode's folded threshold loses its resolution for the small rows: those queries
    must fail the guard and be rescanned, not answered wrongly).  The two out-of-range queries are answered by the
    exact scan on their own: the rest of the batch still goes through the tensor cores."""
    n, d, k, b = 20000, 64, 10, 130
    rng = np.random.default_rng(7)
    X = oracle.fill(n, d, 51)
    X *= (10.0 ** rng.uniform(-3, 3, size=(n, 1))).astype(np.float32)
    X[17] = 0.0
    X[4000:4003] = 0.0
    Q = oracle.fill(b, d, 52)
    Q[0] = 0.0
    Q[1] = X[123]
    Q[2:40] *= np.float32(1e-3)          # queries among the small rows
    Q[40:60] *= np.float32(300.0)
    Q[61] *= np.float32(1e25)
    c = ctx.create(f"awk{metric}_{path}", d, metric, n)
    c.insert(X)
    c.set_path(path)
    s0 = ctx.stats()
    ids, dist = c.search(Q, k)
    assert ctx.stats()["batched_tiles"] > s0["batched_tiles"]
    assert_same(ids, dist, *oracle.search(X, Q, k, metric), f"metric={metric} path={path}")
    ctx.drop(c.name)


def test_batched_guard_failures_are_rescanned(ctx, oracle):
    """Blocks of duplicate rows make the tf32 pass unable to prove its candidates for some queries; those
    queries are re-answered by the scan path and the result is still exact."""
    n, d, k = 40000, 128, 10
    X = oracle.fill(n, d, 21)
    X[1000:1400] = X[7]            # 401 copies of one vector
    Q = oracle.fill(80, d, 22)
    Q[3] = X[7]                    # a query sitting exactly on the duplicate block
    Q[5] = X[7] + np.float32(1e-3)
    c = ctx.create("dups_b", d, 0, n)
    c.insert(X)
    c.set_path(3)
    s0 = ctx.stats()
    ids, dist = c.search(Q, k)
    s1 = ctx.stats()
    assert_same(ids, dist, *oracle.search(X, Q, k, 0))
    assert list(ids[3][:2]) == [7, 1000] and dist[3][0] == 0.0
    assert s1["fast_scans"] > s0["fast_scans"], "the duplicate-block query should have failed the tf32 guard"
    ctx.drop("dups_b")


def test_batched_config2_shape_spot_check(ctx, oracle):
    """BASELINE configs[2] (10M x 128 L2, 1024 queries, top-100) at full size; the oracle answers a sample of
    the 1024 queries (~0.5 s each on the host), the rest are checked by size-independent properties."""
    n, d, k, b = 10_000_000, 128, 100, 1024
    c = ctx.create("cfg2b", d, 0, n)
    c.fill_synthetic(n, 0x5EED0001)
    Q = oracle.fill(b, d, 0x5EED0002)
    c.set_path(3)
    ids, dist = c.search(Q, k)
    assert np.all(np.diff(dist, axis=1) >= 0) and np.all(ids < n)
    assert all(len(set(r)) == k for r in ids.tolist())
    X = oracle.fill(n, d, 0x5EED0001)
    sample = [0, 1, 255, 256, 511, 777, 1023]
    assert_same(ids[sample], dist[sample], *oracle.search(X, Q[sample], k, 0))
    # the scan path gives the same rows for other queries (two independent GPU implementations)
    c.set_path(1)
    other = [100, 600, 900]
    sid, sdist = c.search(Q[other], k)
    assert_same(ids[other], dist[other], sid, sdist)
    ctx.drop("cfg2b")


def test_batched_falls_back_to_tf32_without_a_mirror(tmp_path):
    """When the bf16 mirror cannot be allocated (forced here with VROD_NO_MIRROR=1, read once per process) path 3
    still answers through the tensor cores -- with the f32 rows as tf32 operands -- and the results do not change."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = (
        "import numpy as np\n"
        "from vrod_b200 import ffi\n"
        "from oracle import oracle as O\n"
        "ctx = ffi.Context(0)\n"
        "c = ctx.create('nm', 128, 0, 50000); c.fill_synthetic(50000, 3); c.set_path(3)\n"
        "Q = O.fill(200, 128, 4)\n"
        "s0 = ctx.stats(); ids, dist = c.search(Q, 10); s1 = ctx.stats()\n"
        "assert s1['batched_tiles'] > s0['batched_tiles']\n"
        "rid, rdist = O.search(O.fill(50000, 128, 3), Q, 10, 0)\n"
        "assert np.array_equal(ids, rid) and np.array_equal(dist.view(np.uint32), rdist.view(np.uint32))\n"
        "print('NO_MIRROR_OK', s1['kernel_launches'] - s0['kernel_launches'])\n"
    )
    outs = {}
    for name, env in (("mirror", {}), ("nomirror", {"VROD_NO_MIRROR": "1"})):
        r = subprocess.run([sys.executable, "-c", code], cwd=root, env={**os.environ, **env}, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0 and "NO_MIRROR_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
        outs[name] = int(r.stdout.split("NO_MIRROR_OK")[1].split()[0])
    # the mirror mode launches two kernels more (build_mirror, prep_queries): proof that the modes differed
    assert outs["mirror"] == outs["nomirror"] + 2, outs


def _stress_rows(kind, n, d, rng):
    if kind == "clusters":      # tight clusters: neighbours closer than the tensor-core pass can resolve
        c = rng.standard_normal((max(n // 200, 2), d)).astype(np.float32)
        X = c[rng.integers(0, len(c), n)] + (rng.standard_normal((n, d)) * 0.01).astype(np.float32)
    elif kind == "normalized":  # unit vectors, as an embedding model emits them
        X = rng.standard_normal((n, d)).astype(np.float32)
        X /= np.linalg.norm(X, axis=1, keepdims=True).astype(np.float32)
    elif kind == "grid":        # small integers: many exactly equal distances, the (dist, id) tie-break decides
        X = rng.integers(-3, 4, size=(n, d)).astype(np.float32)
    elif kind == "sparse":
        X = (rng.standard_normal((n, d)) * (rng.random((n, d)) < 0.05)).astype(np.float32)
    else:                       # rank 4
        X = (rng.standard_normal((n, 4)) @ rng.standard_normal((4, d))).astype(np.float32)
    return np.ascontiguousarray(X, dtype=np.float32)


@pytest.mark.parametrize("kind", ["clusters", "normalized", "grid", "sparse", "lowrank"])
def test_batched_on_stress_distributions(ctx, oracle, kind):
    """Data that stresses the guard (tests/tools/soak_batched.py runs the long version): whatever the tensor-core pass
    cannot prove is rescanned, the answer is the oracle's bit for bit on both metrics and both operand modes."""
    rng = np.random.default_rng(sum(kind.encode()))
    n, d, b, k = 30000, 96, 140, 20
    X = _stress_rows(kind, n, d, rng)
    Q = _stress_rows(kind, b, d, rng)
    Q[: b // 2] = X[rng.integers(0, n, b // 2)]
    for metric, path in ((0, 3), (1, 3), (0, 4)):
        c = ctx.create(f"st_{kind}_{metric}_{path}", d, metric, n)
        c.insert(X)
        c.set_path(path)
        s0 = ctx.stats()
        ids, dist = c.search(Q, k)
        assert ctx.stats()["batched_tiles"] > s0["batched_tiles"]
        assert_same(ids, dist, *oracle.search(X, Q, k, metric), f"{kind} metric={metric} path={path}")
        ctx.drop(c.name)


def test_clustered_data_stays_on_the_tensor_cores(ctx, oracle):
    """Tight clusters under the Euclidean metric: neighbour distances are far below the row norms, so the bf16 contraction
    cannot tell the rows of a cluster apart and the fixed-k' proof fails for every query (round 1 then answered the batch
    with one single-query scan per query and stopped using the tensor cores for the collection).  Now the first such batch
    switches the collection to BAND mode -- every candidate within the contraction's error band above the k-th best is
    kept and re-evaluated exactly -- and is answered by one more tensor-core pass; later batches stay batched."""
    rng = np.random.default_rng(5)
    n, d, b, k = 60000, 64, 96, 10
    X = _stress_rows("clusters", n, d, rng)
    Q = X[rng.integers(0, n, b)] + np.float32(1e-4)
    c = ctx.create("learn", d, 0, n)
    c.insert(X)
    want = oracle.search(X, Q, k, 0)
    tiles, rescans = [], []
    for _ in range(3):
        s0 = ctx.stats()
        assert_same(*c.search(Q, k), *want, "clustered, automatic path")
        s1 = ctx.stats()
        tiles.append(s1["batched_tiles"] - s0["batched_tiles"])
        rescans.append(s1["fast_scans"] - s0["fast_scans"])
    assert all(t > 0 for t in tiles), f"every batch should go through the tensor cores: {tiles}"
    assert tiles[0] > tiles[1], "the first batch runs the pass again with the wide margin and then in band mode"
    assert max(rescans) <= b // 20, f"at most 5 % of a batch may fall back to single-query scans: {rescans}"
    ctx.drop("learn")


def test_rows_sorted_by_distance_defeat_the_guessed_thresholds_once(ctx, oracle):
    """The batched pass filters each phase at a GUESSED threshold: the rank among the keys seen so far below which the phase
    should find its k' keys if the rows to come resemble the rows already seen.  Rows inserted nearest-first break that
    assumption: every guess is too tight, the finish kernels notice (fewer than k' keys under the threshold the rows were
    filtered with) and flag the queries.  The first such batch is answered again at the k'-th best key (no assumption),
    the collection stops guessing, later batches run one plain pass -- and every answer is the oracle's throughout."""
    rng = np.random.default_rng(11)
    n, d, b, k = 400_000, 64, 64, 10
    X = oracle.fill(n, d, 93)
    q0 = oracle.fill(1, d, 94)[0]
    X = X[np.argsort(((X.astype(np.float64) - q0) ** 2).sum(axis=1), kind="stable")]   # nearest to q0 first
    Q = (q0 + 0.01 * rng.standard_normal((b, d))).astype(np.float32)
    c = ctx.create("sorted", d, 0, n)
    c.insert(X)
    c.set_path(3)
    want = oracle.search(X, Q, k, 0)
    tiles, rescans = [], []
    for _ in range(3):
        s0 = ctx.stats()
        assert_same(*c.search(Q, k), *want, "rows sorted by distance")
        s1 = ctx.stats()
        tiles.append(s1["batched_tiles"] - s0["batched_tiles"])
        rescans.append(s1["fast_scans"] - s0["fast_scans"])
    assert tiles[0] == 2 * tiles[1] == 2 * tiles[2], f"the first batch runs twice (guessed, then plain), later ones once: {tiles}"
    assert rescans[1] == rescans[2] == 0, f"no single-query rescans once the collection has stopped guessing: {rescans}"
    ctx.drop("sorted")


def test_guessed_thresholds_hold_on_exchangeable_rows(ctx, oracle):
    """Rows in random order (the synthetic fill): the guesses hold, no query is rescanned, one pass per batch."""
    n, d, b, k = 1_000_000, 64, 300, 100
    c = ctx.create("guess_ok", d, 0, n)
    c.fill_synthetic(n, 91)
    c.set_path(3)
    Q = oracle.fill(b, d, 92)
    s0 = ctx.stats()
    ids, dist = c.search(Q, k)
    s1 = ctx.stats()
    assert s1["fast_scans"] == s0["fast_scans"], "a guessed threshold failed on rows in random order"
    assert s1["batched_tiles"] - s0["batched_tiles"] == ((n + 127) // 128) * ((b + 255) // 256)
    assert_same(ids[:16], dist[:16], *oracle.search(oracle.fill(n, d, 91), Q[:16], k, 0), "guessed thresholds, random order")
    ctx.drop("guess_ok")


@pytest.mark.parametrize("metric", [0, 1])
def test_many_queries_few_rows_keeps_the_tensor_core_answers(ctx, oracle, metric):
    """b >= 8192 with k ~ 100 on ~100k rows: few row tiles per CTA and 8x phase growth make the per-(CTA, query) lists
    pass their prune mark INSIDE a phase.  In bf16 mode the prune must not touch the threshold that is folded into the
    contraction (the surrogate of a later candidate is recovered as thr_folded - D): round 1 overwrote it, every later
    candidate was understated, and most of the batch failed its proof and was rescanned one query at a time."""
    n, d, b, k = 100_000, 64, 8192, 100
    c = ctx.create(f"manyq{metric}", d, metric, n)
    c.fill_synthetic(n, 77)
    c.set_path(3)
    X = oracle.fill(n, d, 77)
    Q = oracle.fill(b, d, 78)
    s0 = ctx.stats()
    ids, dist = c.search(Q, k)
    s1 = ctx.stats()
    assert s1["batched_tiles"] > s0["batched_tiles"]
    rescanned = s1["fast_scans"] - s0["fast_scans"]
    assert rescanned <= b // 100, f"{rescanned} of {b} queries fell back to single scans"
    sel = np.r_[0:64, b - 64:b]                     # the oracle checks a slice (k = 100 over 100k rows x 8192 queries is slow)
    assert_same(ids[sel], dist[sel], *oracle.search(X, Q[sel], k, metric), "many queries, few rows")
    ctx.drop(c.name)


def test_band_overflow_in_an_early_phase_is_not_forgotten(ctx, oracle):
    """Sparse rows (5 % non-zero components; many all-zero and near-duplicate rows) under the Euclidean metric with k = 120:
    most fixed-k' proofs fail, the collection switches to band mode, and for many queries the error band holds more rows
    than a list can keep.  A merge BETWEEN phases that overran its list has dropped keys: the flag must reach the last
    phase (round 2's first band-mode build kept it in shared memory only, the last phase happened to fit, its guard passed,
    and true neighbours were missing from 210 of 700 answers)."""
    rng = np.random.default_rng(45)
    n, d, b, k = 99549, 48, 700, 120
    X = _stress_rows("sparse", n, d, rng)
    Q = _stress_rows("sparse", b, d, rng)
    c = ctx.create("bandovf", d, 0, n)
    c.insert(X)
    c.set_path(3)
    want = oracle.search(X, Q, k, 0)
    for _ in range(2):      # the switching batch, then a batch that starts in band mode
        assert_same(*c.search(Q, k), *want, "sparse rows, band mode")
    ctx.drop("bandovf")
