"""World-size-2 (and 3) gloo tests of the N>1 host logic on CPU: the partition rule, shipping the
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

from vrod_b200.dist import shard_range


def test_shard_range_partitions_everything():
    for cap, world in [(10, 1), (10, 2), (10, 3), (7, 8), (100_000_000, 8), (1, 4)]:
        covered = []
        for r in range(world):
            lo, hi = shard_range(cap, r, world)
            assert 0 <= lo <= hi <= cap
            covered += list(range(lo, hi)) if cap <= 100 else []
            if r:
                assert lo == shard_range(cap, r - 1, world)[1]
        assert shard_range(cap, world - 1, world)[1] == cap
        if cap <= 100:
            assert covered == list(range(cap))


def _worker(rank, world, port, n, d, k, metric, out):
    import torch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from vrod_b200.dist import share_comm_id
    cid = share_comm_id(lambda: bytes(range(128)), rank, world)
    assert cid == bytes(range(128))
    lo, hi = shard_range(n, rank, world)
    X = O.fill(hi - lo, d, 77, row0=lo)                 # this rank's rows only
    Q = O.fill(4, d, 78)
    ids, dd = O.search(X, Q, k, metric, id_base=lo)
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


@pytest.mark.parametrize("world,metric", [(2, 0), (2, 1), (3, 0)])
def test_gloo_sharded_equals_unsharded(tmp_path, oracle, world, metric):
    n, d, k = 3001, 48, 12
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
