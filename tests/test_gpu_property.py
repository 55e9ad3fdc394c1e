"""Property tests (hypothesis): for random shapes, metrics, k, batch sizes, paths and data with planted
duplicates / zero rows / exact query hits, the CUDA answer through the C ABI equals the oracle's bit for bit.
Sizes are small so that the oracle answers each example in milliseconds."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from tests.util import assert_same

pytestmark = pytest.mark.gpu
_n = [0]


@st.composite
def cases(draw):
    d = draw(st.sampled_from([1, 3, 8, 17, 32, 64, 100, 128, 130, 256, 384, 513, 768]))
    n = draw(st.integers(min_value=1, max_value=6000))
    k = draw(st.sampled_from([1, 2, 10, 33, 100, 120]))
    b = draw(st.sampled_from([1, 2, 5, 64, 70]))
    metric = draw(st.sampled_from([0, 1]))
    path = draw(st.sampled_from([0, 1, 2, 3, 3, 4]))   # auto, scan, exact, batched (bf16 mirror), batched tf32
    seed = draw(st.integers(min_value=1, max_value=2**31))
    dup = draw(st.sampled_from([0, 0, 3, 40]))      # copies of one row planted elsewhere
    zero = draw(st.booleans())
    hit = draw(st.booleans())                       # one query equals a stored row
    return dict(d=d, n=n, k=k, b=b, metric=metric, path=path, seed=seed, dup=dup, zero=zero, hit=hit)


@settings(max_examples=200, deadline=None, suppress_health_check=[HealthCheck.function_scoped_fixture], derandomize=True)
@given(c=cases())
def test_random_cases_match_oracle(ctx, oracle, c):
    n, d, k, b = c["n"], c["d"], c["k"], c["b"]
    X = oracle.fill(n, d, c["seed"])
    rng = np.random.default_rng(c["seed"])
    if c["dup"] and n > c["dup"] + 2:
        src = int(rng.integers(0, n))
        X[rng.choice(n, size=c["dup"], replace=False)] = X[src]
    if c["zero"]:
        X[int(rng.integers(0, n))] = 0.0
    Q = oracle.fill(b, d, c["seed"] ^ 0x5555)
    if c["hit"]:
        Q[0] = X[int(rng.integers(0, n))]
    _n[0] += 1
    coll = ctx.create(f"prop{_n[0]}", d, c["metric"], n)
    try:
        coll.insert(X)
        coll.set_path(c["path"])
        ids, dist = coll.search(Q, k)
        assert_same(ids, dist, *oracle.search(X, Q, k, c["metric"]), str(c))
    finally:
        ctx.drop(coll.name)
