"""Worker of tests/test_gpu_sharded.py: one process per GPU (torchrun), row-sharded collections,
NCCL all-gather + merge inside the library; rank 0 checks the global answer against the oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import oracle as O
    from tests.util import assert_same
    from vrod_b200 import ffi
    from vrod_b200.dist import share_comm_id, shard_range

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    cid = share_comm_id(ffi.comm_unique_id, rank, world)
    ctx = ffi.Context(local, rank, world, cid)
    assert ctx.rank == rank and ctx.world == world

    for (n, d, metric, k) in [(100_003, 128, 0, 10), (50_000, 768, 1, 10), (20_001, 64, 0, 100), (7, 128, 0, 10), (1, 32, 1, 3)]:
        c = ctx.create(f"s{n}_{d}", d, metric, n)
        c.fill_synthetic(n, 41)
        base, rows_here = c.shard()
        lo, hi = shard_range(n, rank, world)
        assert (base, rows_here) == (lo if hi > lo else base, hi - lo), (base, rows_here, lo, hi)
        X = O.fill(n, d, 41)
        if rows_here:
            assert np.array_equal(c.read_rows(0, rows_here), X[lo:hi])
        Q = O.fill(5, d, 42)
        Q[4] = X[n - 1]                                   # an exact hit on the last shard
        ids, dd = c.search(Q, k)                          # collective; every rank gets the global answer
        assert_same(ids, dd, *O.search(X, Q, k, metric), f"rank {rank} n={n} d={d}")
        c.set_path(2)
        assert_same(*c.search(Q[:2], k), ids[:2], dd[:2], "exact path, sharded")
        ctx.drop(c.name)

    # batched (tensor-core) path on every shard, both operand modes, then the exchange + merge of b x k hits
    for (n, d, metric, k, b, path) in [(300_000, 128, 0, 10, 300, 3), (120_000, 96, 1, 100, 257, 3), (150_000, 128, 0, 10, 64, 4)]:
        c = ctx.create(f"sb{n}_{path}", d, metric, n)
        c.fill_synthetic(n, 46)
        c.set_path(path)
        Q = O.fill(b, d, 47)
        s0 = ctx.stats()
        ids, dd = c.search(Q, k)
        assert ctx.stats()["batched_tiles"] > s0["batched_tiles"], "the tensor-core path did not run"
        if rank == 0:                                     # (every rank holds the same global answer; one oracle run is enough)
            assert_same(ids, dd, *O.search(O.fill(n, d, 46), Q, k, metric), f"sharded batched n={n} path={path}")
        gathered = [None] * world
        dist.all_gather_object(gathered, (ids.tobytes(), dd.tobytes()))
        assert all(g == gathered[0] for g in gathered), "ranks disagree on the global answer"
        ctx.drop(c.name)

    # INSERT: every rank passes the same rows; ties across the shard boundary break by id
    n, d = 30_000, 96
    X = O.fill(n, d, 43)
    lo1, _ = shard_range(n, world - 1, world)
    X[lo1] = X[3]                                         # same vector in the first and the last shard
    c = ctx.create("ins", d, 0, n)
    c.insert(X[:12_345])
    c.insert(X[12_345:])
    ids, dd = c.search(X[3], 5)
    assert ids[0, 0] == 3 and ids[0, 1] == lo1 and dd[0, 0] == 0 and dd[0, 1] == 0
    assert_same(ids, dd, *O.search(X, X[3], 5, 0))
    ctx.drop("ins")

    # a batch with ONE non-finite value lands on one rank only: every rank must reject it (and keep its count), or the
    # ranks' ids drift apart for every later insert
    c = ctx.create("bad", d, 0, n)
    c.insert(X[:1000])
    bad = X[1000:1100].copy()
    bad[0, 0] = np.inf
    try:
        c.insert(bad)
        raise AssertionError("a batch with an infinity was accepted")
    except ffi.VrodError as e:
        assert e.status == ffi.EINVAL
    assert c.info()["count"] == 1000
    assert c.insert(X[1000:]) == 1000
    assert_same(*c.search(X[n - 1], 3), *O.search(X, X[n - 1], 3, 0))
    ctx.drop("bad")

    # large batches leave the fused exchange for ncclAllGather + merge; alternating the two paths with different (b, k)
    # shifts the packed result layout over stale words -- no call may report a spurious exchange time-out
    c = ctx.create("alt", 128, 0, 60_000)
    c.fill_synthetic(60_000, 48)
    Xa = O.fill(60_000, 128, 48)
    for (b, k) in [(3, 10), (1024, 10), (1, 1), (300, 100), (7, 3), (1024, 1), (2, 100)]:
        Q = O.fill(b, 128, 49 + b)
        ids, dd = c.search(Q, k)
        if rank == 0:
            assert_same(ids, dd, *O.search(Xa, Q, k, 0), f"alternating exchange paths b={b} k={k}")
    ctx.synchronize()
    ctx.drop("alt")

    # resident (device-pointer) entry point, also collective
    c = ctx.create("dev", 128, 0, 200_000)
    c.fill_synthetic(200_000, 44)
    Q = O.fill(3, 128, 45)
    q = torch.from_numpy(Q).cuda()
    ids_t = torch.empty((3, 10), dtype=torch.int64, device="cuda")
    dd_t = torch.empty((3, 10), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    c.search_device(q.data_ptr(), 3, 10, ids_t.data_ptr(), dd_t.data_ptr())
    ctx.synchronize()
    assert_same(ids_t.cpu().numpy().astype(np.uint64), dd_t.cpu().numpy(), *O.search(O.fill(200_000, 128, 44), Q, 10, 0))
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    if rank == 0:
        print("SHARDED_OK world", world)


if __name__ == "__main__":
    main()
