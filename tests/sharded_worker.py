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
    from vrod_b200.dist import SHARD_BLOCK, share_comm_id, shard_ids

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
        mine = shard_ids(n, rank, world)                  # block-cyclic deal: blocks of SHARD_BLOCK ids, round robin
        assert (base, rows_here) == (rank * SHARD_BLOCK, len(mine)), (base, rows_here, len(mine))
        X = O.fill(n, d, 41)
        if rows_here:
            assert np.array_equal(c.read_rows(0, rows_here), X[mine.astype(np.int64)])
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
    n, d = 9 * SHARD_BLOCK + 777, 96
    X = O.fill(n, d, 43)
    lo1 = (world - 1) * SHARD_BLOCK + 5                   # a row of the last rank's first block
    X[lo1] = X[3]                                         # same vector on the first and the last rank
    X[world * SHARD_BLOCK + 9] = X[3]                     # ... and again on the first rank, one deal later
    c = ctx.create("ins", d, 0, n // 3)                   # a third of what will arrive: the shards grow, twice
    c.insert(X[:12_345])
    c.insert(X[12_345:])
    assert c.info()["count"] == n and c.info()["capacity"] >= n
    assert c.shard()[1] == len(shard_ids(n, rank, world))
    ids, dd = c.search(X[3], 5)
    assert ids[0].tolist()[:3] == [3, lo1, world * SHARD_BLOCK + 9] and not dd[0, :3].any()   # ties: by id, across interleaved shards
    assert_same(ids, dd, *O.search(X, X[3], 5, 0))
    ctx.drop("ins")

    # persistence, collectively: every rank writes / reads only the blocks it holds (one file, block offsets)
    path = f"/tmp/vrod_sharded_{os.environ.get('MASTER_PORT', '0')}.vrc"
    c = ctx.create("sv", d, 1, n)
    c.insert(X)
    Qs = O.fill(3, d, 60)
    want = O.search(X, Qs, 7, 1)
    assert_same(*c.search(Qs, 7), *want, "before save")
    c.save(path)
    ctx.drop("sv")
    c2 = ctx.load("sv2", path)
    assert c2.info()["count"] == n and c2.shard()[1] == len(shard_ids(n, rank, world))
    assert_same(*c2.search(Qs, 7), *want, "after the collective load")
    ctx.drop("sv2")
    if rank == 0:                                         # the same file in a plain single-GPU context
        with ffi.Context(local) as one:
            assert_same(*one.load("sv3", path).search(Qs, 7), *want, "sharded file, single-GPU load")
    dist.barrier()
    if rank == 0:
        os.remove(path)

    # a batch with ONE non-finite value lands on one rank only: every rank must reject it (and keep its count), or the
    # ranks' ids drift apart for every later insert
    n = 30_000
    X = X[:n]
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
