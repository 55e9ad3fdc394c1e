"""Multi-rank plumbing around the C ABI: one process per GPU (torchrun), torch.distributed only for
the rendezvous (shipping the communicator id, barriers, max-over-ranks timing).  The data path -- the
per-rank scan and the ncclAllGather + merge of the top-k lists -- is inside libvrod_knn.so.

shard_range() restates the library's partition rule (vrod_capi.cu: vrod_collection_create) so host
code and tests can reason about which rank owns which ids.
"""
import os


def shard_range(capacity, rank, world):
    """Global id range [lo, hi) owned by `rank`: contiguous blocks of ceil(capacity/world) rows."""
    per = (capacity + world - 1) // world
    lo = min(capacity, per * rank)
    hi = min(capacity, per * (rank + 1))
    return lo, hi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def share_comm_id(make_id, rank, world):
    """Rank 0 creates the communicator id (bytes); everyone receives it through torch.distributed."""
    import torch.distributed as dist
    box = [make_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]
