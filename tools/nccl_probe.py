"""Development aid: why is the async resident loop slower than the per-step-sync loop at N >= 4?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from vrod_b200 import ffi
from vrod_b200.dist import share_comm_id
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = ffi.Context(local, rank, world, share_comm_id(ffi.comm_unique_id, rank, world))
stream = torch.cuda.ExternalStream(ctx.stream())
n, d, k = int(os.environ.get("ROWS", 100_000_000)), 128, 10
c = ctx.create("x", d, 0, n); c.fill_synthetic(n, 1)
q = torch.randn(512, d, device="cuda"); ids = torch.empty((1, k), dtype=torch.int64, device="cuda"); dd = torch.empty((1, k), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
def run(name, steps=200, sync_every=0, profile=False):
    for i in range(5): c.search_device(q[i].data_ptr(), 1, k, ids.data_ptr(), dd.data_ptr())
    ctx.synchronize(); dist.barrier(); torch.cuda.synchronize()
    ctx.profile(profile); ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for i in range(steps):
        c.search_device(q[i % 512].data_ptr(), 1, k, ids.data_ptr(), dd.data_ptr())
        if sync_every and (i + 1) % sync_every == 0: ctx.synchronize()
    e1.record(stream); e1.synchronize(); t1 = time.perf_counter()
    kms, kn = ctx.profile_read(); ctx.profile(False)
    t = torch.tensor([e0.elapsed_time(e1) / steps, (t1 - t0) * 1e3 / steps, kms / max(kn, 1)], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{name:32s} event {t[0].item():.3f} ms/step  wall {t[1].item():.3f} ms/step  scan {t[2].item():.3f} ms", flush=True)
run("async")
run("async+profile", profile=True)
run("sync every step", sync_every=1)
run("sync every 4", sync_every=4)
run("sync every 16", sync_every=16)
run("async again")
ctx.close(); dist.destroy_process_group()
