"""Hardware multi-GPU correctness: runs only where >= 2 CUDA devices are visible (gpurun --gpus 2+,
the driver's scaling box); the protocol itself is covered on CPU/gloo by tests/test_distributed_cpu.py."""
import os
import socket
import subprocess
import sys

import pytest

from tests.conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_nccl_sharded_stats_equal_single_gpu_and_numpy(world):
    import torch
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip("needs %d CUDA devices" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "multigpu_stats_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTIGPU_STATS_OK world=%d" % world in r.stdout
