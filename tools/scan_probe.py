import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrod_b200 import ffi
ctx = ffi.Context(0)
for (n, d, m, k) in [(1000000, 128, 0, 10), (1000000, 128, 0, 100), (1000000, 64, 0, 10), (10000, 128, 0, 10), (1000000, 768, 1, 10)]:
    c = ctx.create("t", d, m, n); c.fill_synthetic(n, 7)
    q = torch.randn(8, d, device="cuda"); ids = torch.empty((1, k), dtype=torch.int64, device="cuda"); dd = torch.empty((1, k), dtype=torch.float32, device="cuda")
    torch.cuda.synchronize()
    print(f"== n={n} d={d} metric={m} k={k}", file=sys.stderr, flush=True)
    for i in range(5): c.search_device(q[i].data_ptr(), 1, k, ids.data_ptr(), dd.data_ptr())
    ctx.synchronize(); ctx.drop("t")
