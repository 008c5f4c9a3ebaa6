"""Data-parallel FM fit on real GPUs under torchrun (NCCL), world = min(2, visible GPUs)."""
import os
import socket
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("exchange", ["nvlink", "nccl"])
def test_data_parallel_fit_matches_reference_golden(exchange):
    """exchange = nvlink: the library's fused reduce + apply kernel over CUDA-IPC peer memory;
    nccl: torch.distributed all-reduce + apply. Both must reproduce the reference trajectory."""
    import torch
    world = min(2, torch.cuda.device_count())
    assert world >= 1
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dp_worker.py")]
    env = dict(os.environ, RFM_DP_EXCHANGE=exchange)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DP_OK world=%d exchange=%s" % (world, exchange) in res.stdout
