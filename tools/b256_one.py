import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from vrod_b200 import ffi
ctx = ffi.Context(0)
n, d, m, k, b = 1000000, 128, 0, 10, 256
c = ctx.create("t", d, m, n); c.fill_synthetic(n, 7)
q = torch.randn(b, d, device="cuda"); ids = torch.empty((b, k), dtype=torch.int64, device="cuda"); dd = torch.empty((b, k), dtype=torch.float32, device="cuda")
torch.cuda.synchronize()
for i in range(3): c.search_device(q.data_ptr(), b, k, ids.data_ptr(), dd.data_ptr())
ctx.synchronize()
