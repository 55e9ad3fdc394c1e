"""Multi-rank plumbing around the C ABI: one process per GPU (torchrun), torch.distributed only for
the rendezvous (shipping the communicator id, barriers, max-over-ranks timing).  The data path -- the
per-rank scan and the exchange + merge of the top-k lists -- is inside libvrod_knn.so.

shard_ids() / shard_rows_at() restate the library's partition rule (knn_scan.cuh: row_id, block-cyclic
sharding) so host code and tests can reason about which rank owns which ids:
global ids are dealt out in blocks of SHARD_BLOCK consecutive rows, block j to rank j mod world; a rank
stores its blocks back to back, so its local order is id order.
"""
import os

import numpy as np

SHARD_BLOCK = 4096


def shard_rows_at(count, rank, world, block=SHARD_BLOCK):
    """Rows rank `rank` holds when the collection holds `count` rows."""
    cycle = block * world
    full, rem = divmod(count, cycle)
    return full * block + min(max(rem - rank * block, 0), block)


def shard_ids(count, rank, world, block=SHARD_BLOCK):
    """Global ids of rank `rank`'s rows, in local order (ascending)."""
    ids = np.arange(count, dtype=np.uint64)
    return ids[(ids // block) % world == rank]


def row_id(local_row, rank, world, block=SHARD_BLOCK):
    return ((local_row // block) * world + rank) * block + local_row % block


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def share_comm_id(make_id, rank, world):
    """Rank 0 creates the communicator id (bytes); everyone receives it through torch.distributed."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]
