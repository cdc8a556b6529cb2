"""Row-sharded table on 2 GPUs (skipped on a single-GPU box): tools/check_sharded.py under torchrun compares the
P2P lookup bit for bit, the probabilities and the reduce-scattered table gradient with an unsharded replica."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_sharded_table_two_gpus():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "check_sharded.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "SHARDED OK world=2" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
