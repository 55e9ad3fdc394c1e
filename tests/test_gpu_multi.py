"""Single-process multi-GPU context (vrod_ctx_create_multi): one process, one host thread, every collection
row-sharded over the devices -- the process model SURVEY.md section 8(b)/(e) asks for because the reference's
caller is one single-threaded process (Rc<RefCell<Database>>, src/command/types.rs:10; fn main, src/main.rs:42).
The answers must equal the oracle's (and therefore the single-GPU and the process-per-GPU answers) bit for bit.
Needs >= 2 GPUs (`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`); skipped on a 1-GPU box."""
import os
import subprocess

import numpy as np
import pytest

from tests.util import assert_same

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "vrod_b200", "host", "vrod")


def ndev():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module", params=[2, 4, 8])
def mctx(request):
    from vrod_b200 import ffi
    if ndev() < request.param:
        pytest.skip(f"needs {request.param} GPUs, have {ndev()}")
    c = ffi.Context(list(range(request.param)))
    assert c.devices == request.param and c.world == 1 and c.rank == 0
    yield c
    c.close()


def test_one_device_list_is_a_plain_context(oracle):
    from vrod_b200 import ffi
    with ffi.Context([0]) as c:
        assert c.devices == 1
        col = c.create("one", 64, 0, 1000)
        col.fill_synthetic(1000, 3)
        Q = oracle.fill(2, 64, 4)
        assert_same(*col.search(Q, 5), *oracle.search(oracle.fill(1000, 64, 3), Q, 5, 0))


def test_scan_exact_and_padding_match_the_oracle(mctx, oracle):
    for (n, d, metric, k) in [(100_003, 128, 0, 10), (50_000, 768, 1, 10), (20_001, 64, 0, 100), (7, 128, 0, 10), (1, 32, 1, 3),
                              (30_000, 100, 1, 7)]:
        c = mctx.create(f"m{n}_{d}", d, metric, n)
        c.fill_synthetic(n, 41)
        assert c.shard() == (0, n)                                  # the parent holds everything
        X = oracle.fill(n, d, 41)
        assert np.array_equal(c.read_rows(0, n), X)                 # global row indices across the devices
        if n > 10:
            assert np.array_equal(c.read_rows(n // 3, n // 2), X[n // 3:n // 3 + n // 2])
        Q = oracle.fill(5, d, 42)
        Q[4] = X[n - 1]                                             # an exact hit on the last device's shard
        c.set_path(1)
        ids, dd = c.search(Q, k)
        assert_same(ids, dd, *oracle.search(X, Q, k, metric), f"multi n={n} d={d}")
        c.set_path(2)
        assert_same(*c.search(Q[:2], k), ids[:2], dd[:2], "exact path, multi")
        c.set_path(0)
        assert_same(*c.search(Q, k), ids, dd, "automatic path, multi")
        mctx.drop(c.name)


def test_batched_path_fused_and_gathered_exchange(mctx, oracle):
    # b*k <= 4096 with b <= 256: the fused NVLink push + merge kernel; beyond: peer copies to device 0 + merge
    for (n, d, metric, k, b, path) in [(300_000, 128, 0, 10, 200, 3), (300_000, 128, 0, 10, 1024, 3), (120_000, 96, 1, 100, 257, 3),
                                       (150_000, 128, 0, 10, 64, 4)]:
        c = mctx.create(f"mb{n}_{b}_{path}", d, metric, n)
        c.fill_synthetic(n, 46)
        c.set_path(path)
        Q = oracle.fill(b, d, 47)
        s0 = mctx.stats()
        ids, dd = c.search(Q, k)
        assert mctx.stats()["batched_tiles"] > s0["batched_tiles"], "the tensor-core path did not run"
        assert_same(ids, dd, *oracle.search(oracle.fill(n, d, 46), Q, k, metric), f"multi batched n={n} b={b} path={path}")
        mctx.drop(c.name)


def test_insert_ties_across_devices_and_rejects_bad_rows(mctx, oracle):
    from vrod_b200 import ffi
    from vrod_b200.dist import SHARD_BLOCK
    W = mctx.devices
    n, d = 9 * SHARD_BLOCK + 777, 96
    X = oracle.fill(n, d, 43)
    last_lo = (W - 1) * SHARD_BLOCK + 5                             # a row of the last device's first block
    X[last_lo] = X[3]                                               # the same vector on the first and the last device
    c = mctx.create("mins", d, 0, n // 3)                           # the shards grow as the rows arrive
    assert c.insert(X[:12_345]) == 0
    bad = X[12_345:12_400].copy()
    bad[-1, 5] = np.nan                                             # lands on ONE device: all of them must reject the batch
    with pytest.raises(ffi.VrodError) as e:
        c.insert(bad)
    assert e.value.status == ffi.EINVAL and c.info()["count"] == 12_345
    assert c.insert(X[12_345:]) == 12_345
    ids, dd = c.search(X[3], 5)
    assert ids[0, 0] == 3 and ids[0, 1] == last_lo and dd[0, 0] == 0 and dd[0, 1] == 0
    assert_same(ids, dd, *oracle.search(X, X[3], 5, 0))
    assert c.info()["count"] == n and c.info()["capacity"] >= n     # grown in place: every device continues its own deal
    assert np.array_equal(c.read_rows(0, n), X)
    mctx.drop("mins")


def test_save_and_load_across_process_models(mctx, oracle, tmp_path):
    """A collection saved by the multi-GPU context loads into a single-GPU context (and back) with the same answers."""
    from vrod_b200 import ffi
    n, d = 40_001, 64
    X = oracle.fill(n, d, 51)
    Q = oracle.fill(3, d, 52)
    c = mctx.create("msave", d, 1, n)
    c.insert(X)
    want = oracle.search(X, Q, 10, 1)
    assert_same(*c.search(Q, 10), *want)
    f = tmp_path / "msave.vrc"
    c.save(f)
    mctx.drop("msave")
    with ffi.Context(0) as one:
        assert_same(*one.load("back", f).search(Q, 10), *want)
    c2 = mctx.load("again", f)
    assert c2.info()["count"] == n
    assert_same(*c2.search(Q, 10), *want)
    mctx.drop("again")


def test_device_pointer_api_is_refused(mctx):
    from vrod_b200 import ffi
    c = mctx.create("mdev", 32, 0, 100)
    c.fill_synthetic(100, 1)
    with pytest.raises(ffi.VrodError) as e:
        c.search_device(0x1000, 1, 1, 0x1000, 0x1000)
    assert e.value.status == ffi.EINVAL
    mctx.drop("mdev")


def test_cli_devices_flag(oracle, tmp_path):
    """`vrod --devices 0,1 -e SEARCH` answers like the oracle: the drop-in CLI can use every GPU of the box."""
    if ndev() < 2:
        pytest.skip("needs 2 GPUs")
    n, d = 5000, 48
    X = oracle.fill(n, d, 61)
    rec = tmp_path / "rows.txt"
    with open(rec, "w") as f:
        for i, row in enumerate(X):
            f.write(",".join(repr(float(v)) for v in row) + f";w{i}\n")
    q = oracle.fill(1, d, 62)[0]
    script = tmp_path / "s.txt"
    script.write_text(f"CREATE - t;{d};euclidean;{n}\nBULKINSERT t {rec}\nSEARCH t 5;" + ",".join(repr(float(v)) for v in q) + "\n")
    devs = ",".join(str(i) for i in range(min(ndev(), 8)))
    r = subprocess.run([CLI, "--devices", devs, "--script", str(script)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    rid, rdist = oracle.search(X, q, 5, 0)
    hits = [ln.split("\t") for ln in r.stdout.splitlines() if ln[:1].isdigit() and "\t" in ln]
    assert [int(h[1]) for h in hits] == [int(v) for v in rid[0]], r.stdout
    assert [h[3] for h in hits] == [f"w{int(v)}" for v in rid[0]]
    assert np.array_equal(np.array([float(h[2]) for h in hits], dtype=np.float32).view(np.uint32), rdist[0].view(np.uint32))
