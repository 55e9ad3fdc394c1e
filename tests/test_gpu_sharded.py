"""Row-sharded search over 2+ GPUs (one process per GPU, NCCL all-gather + merge inside the library)
equals the unsharded oracle answer.  Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest
tests -m gpu`; skipped on a 1-GPU box (the CPU-side sharding logic is covered by test_sharding_gloo)."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("exchange", ["fused", "nccl"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_search_matches_oracle(world, exchange):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, have {torch.cuda.device_count()}")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "sharded_worker.py")]
    env = dict(os.environ, VROD_NO_P2P_EXCHANGE="1") if exchange == "nccl" else dict(os.environ)
    env.pop("VROD_NO_P2P_EXCHANGE", None) if exchange == "fused" else None
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0 and f"SHARDED_OK world {world}" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
