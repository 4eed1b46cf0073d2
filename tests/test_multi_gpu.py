"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a 1-GPU box): the sharded path through the
C ABI + NCCL all-gather reproduces the oracle bit for bit on every rank."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("comm", ["native", "torch"])
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_update_matches_oracle(world, comm, oracle_mod):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    for attempt in range(4):        # a port handed out by bind(0) can still be taken by the time torchrun listens on it: try another
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            port = s.getsockname()[1]
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
               "--master-port", str(port), os.path.join(HERE, "_nccl_worker.py"), comm]
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        if r.returncode == 0 or "EADDRINUSE" not in r.stdout:
            break
    assert r.returncode == 0, r.stdout[-4000:]
    for k in range(world):
        assert f"rank {k}/{world} ok" in r.stdout, r.stdout[-4000:]
