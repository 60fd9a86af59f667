"""Multi-GPU parity (needs >= 2 visible GPUs; skipped on a single-GPU box): the assembly
tree is partitioned over 2 ranks, contribution blocks cross GPUs by NCCL send/recv, and the
distributed solve must reach the same backward error as the single-GPU path."""
import json
import os
import socket
import subprocess
import sys

import pytest

import sylver_b200 as sb

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("kind,k", [("lap27", 20), ("lap7", 30)])
def test_two_gpu_factor_and_solve(lib, kind, k):
    sb.require_gpu()
    if sb.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "scripts", "multi_gpu_worker.py"), kind, str(k), "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    rec = json.loads(line)
    assert rec["world"] == 2 and rec["bwderr"] <= 1e-14
