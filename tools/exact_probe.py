import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrod_b200 import ffi
ctx = ffi.Context(0)
stream = torch.cuda.ExternalStream(ctx.stream())
for (n, d, m, k) in [(1000000, 128, 0, 10), (1000000, 768, 1, 10), (10000000, 128, 0, 10), (1000000, 1536, 0, 100), (1000000, 64, 1, 10)]:
    c = ctx.create("t", d, m, n); c.fill_synthetic(n, 7); c.set_path(2)
    q = torch.randn(16, d, device="cuda"); ids = torch.empty((1, k), dtype=torch.int64, device="cuda"); dd = torch.empty((1, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for i in range(3): c.search_device(q[i].data_ptr(), 1, k, ids.data_ptr(), dd.data_ptr())
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(10): c.search_device(q[i].data_ptr(), 1, k, ids.data_ptr(), dd.data_ptr())
    e1.record(stream); e1.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"exact scan n={n} d={d} metric={m} k={k}: {ms*1e3:.1f} us/query {n*d*4/ms/1e6:.0f} GB/s ({n*d*4/ms/1e6/6548.8*100:.0f}% of measured copy peak)", flush=True)
    ctx.drop("t")
