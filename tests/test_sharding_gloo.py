"""World-size-2 (and 3) gloo tests of the N>1 host logic on CPU: the (block-cyclic) partition rule, shipping the
communicator id, and that per-rank top-k lists exchanged by an all-gather and merged under
(dist, id) equal the unsharded answer.  The per-rank scan here is the oracle (there is no GPU in
this container); on the GPU box the same exchange is the library's ncclAllGather + merge kernel
(tests/test_gpu_sharded.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from vrod_b200.dist import SHARD_BLOCK, row_id, shard_ids, shard_rows_at


def test_block_cyclic_deal_partitions_everything():
    """The partition rule of the library (knn_scan.cuh: row_id / shard_rows_at), restated in vrod_b200/dist.py."""
    for count, world in [(10, 1), (10, 2), (3 * SHARD_BLOCK + 5, 2), (7 * SHARD_BLOCK, 3), (100_000_000, 8), (1, 4), (0, 2)]:
        sizes = [shard_rows_at(count, r, world) for r in range(world)]
        assert sum(sizes) == count
        assert max(sizes) - min(sizes) <= SHARD_BLOCK                 # balanced at ANY fill level
        if count <= 8 * SHARD_BLOCK:
            seen = []
            for r in range(world):
                ids = shard_ids(count, r, world)
                assert len(ids) == sizes[r]
                assert np.all(np.diff(ids.astype(np.int64)) > 0)       # local order is id order
                assert [row_id(i, r, world) for i in (0, len(ids) - 1) if len(ids)] == [int(ids[i]) for i in (0, len(ids) - 1) if len(ids)]
                seen += ids.tolist()
            assert sorted(seen) == list(range(count))
    # growth continues the deal: what a rank holds at a smaller count is a prefix of what it holds later
    for r in range(3):
        small, big = shard_ids(2 * SHARD_BLOCK + 17, r, 3), shard_ids(9 * SHARD_BLOCK + 1, r, 3)
        assert np.array_equal(big[:len(small)], small)


def _worker(rank, world, port, n, d, k, metric, out):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from vrod_b200.dist import share_comm_id
    cid = share_comm_id(lambda: bytes(range(128)), rank, world)
    assert cid == bytes(range(128))
    mine = shard_ids(n, rank, world)
    X = O.fill(n, d, 77)[mine.astype(np.int64)]         # this rank's rows only (block-cyclic deal)
    Q = O.fill(4, d, 78)
    ids, dd = O.search(X, Q, k, metric, ids=mine)
    t_ids = torch.from_numpy(ids.astype(np.int64))
    t_dd = torch.from_numpy(dd)
    g_ids = [torch.empty_like(t_ids) for _ in range(world)]
    g_dd = [torch.empty_like(t_dd) for _ in range(world)]
    dist.all_gather(g_ids, t_ids)
    dist.all_gather(g_dd, t_dd)
    m_ids, m_dd = O.merge(np.stack([g.numpy().astype(np.uint64) for g in g_ids]), np.stack([g.numpy() for g in g_dd]))
    if rank == 0:
        np.save(out + ".ids.npy", m_ids)
        np.save(out + ".dist.npy", m_dd)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,metric", [(2, 0), (2, 1), (3, 0), (4, 1)])
def test_gloo_sharded_equals_unsharded(tmp_path, oracle, world, metric):
    n, d, k = 3 * SHARD_BLOCK + 1001, 48, 12
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "merged")
    mp.spawn(_worker, args=(world, port, n, d, k, metric, out), nprocs=world, join=True)
    X = oracle.fill(n, d, 77)
    Q = oracle.fill(4, d, 78)
    rid, rdd = oracle.search(X, Q, k, metric)
    assert np.array_equal(np.load(out + ".ids.npy"), rid)
    assert np.array_equal(np.load(out + ".dist.npy").view(np.uint32), rdd.view(np.uint32))
