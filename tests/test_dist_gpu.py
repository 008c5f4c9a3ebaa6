"""Data-parallel FM fit, sharded upload, sharded scoring and sharded metrics on real GPUs under torchrun (NCCL).

With two or more GPUs the worker runs at world = 2. On a single-GPU box the same worker runs at world = 1 -- every
collective and the peer-memory exchange degenerate to one rank, which still exercises the code paths -- and the test
is reported as XFAIL (not as a pass): multi-rank parity was NOT checked there. bench.py repeats the data-parallel
parity check (dp_parity) inside every N > 1 run, where the driver sees it."""
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
    if world < 2:
        pytest.xfail("only one GPU visible: the worker ran at world = 1, multi-rank parity was not checked on this box "
                     "(bench.py's dp_parity covers it in every N > 1 run)")
