import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrod_b200 import ffi
ctx = ffi.Context(0); stream = torch.cuda.ExternalStream(ctx.stream())
for (n, d, m, k, b) in [(1000000, 64, 0, 10, 256), (1000000, 128, 0, 10, 256), (1000000, 128, 0, 100, 256), (1000000, 768, 1, 10, 256), (2000000, 128, 0, 10, 1024), (200000, 128, 0, 10, 256)]:
    c = ctx.create("t", d, m, n); c.fill_synthetic(n, 7)
    q = torch.randn(b, d, device="cuda"); ids = torch.empty((b, k), dtype=torch.int64, device="cuda"); dd = torch.empty((b, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    for i in range(3): c.search_device(q.data_ptr(), b, k, ids.data_ptr(), dd.data_ptr())
    ctx.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(10): c.search_device(q.data_ptr(), b, k, ids.data_ptr(), dd.data_ptr())
    e1.record(stream); e1.synchronize()
    print(f"n={n} d={d} m={m} k={k} b={b}: {e0.elapsed_time(e1)/10*1e3:.1f} us/batch", flush=True)
    ctx.drop("t")
