"""Row-sharded operator on >= 2 GPUs: one process per GPU under torchrun; skipped on a 1-GPU box.
The host-side sharding logic is covered on CPU by tests/test_shard_dist_cpu.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.parametrize("p2p", ["1", "0"])
def test_sharded_operator_and_solver(p2p):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    env = dict(os.environ, TB_P2P=p2p)
    port = 29500 + (os.getpid() % 400) + (0 if p2p == "1" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "dist_worker ok" in r.stdout
