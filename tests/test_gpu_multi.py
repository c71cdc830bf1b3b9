"""Multi-GPU parity (needs ≥ 2 GPUs, skipped otherwise): a sharded run with NCCL boundary hand-over equals the
single-GPU run bit for bit.  The host-side plan itself is covered on CPU by tests/test_shard.py (gloo)."""
import os
import subprocess
import sys

import pytest

from hydpy_b200 import device_count

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.skipif(device_count() < 2, reason="needs at least two GPUs")
def test_sharded_run_over_nccl_is_bit_identical():
    world = min(device_count(), 4)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(HERE, "multi_gpu_worker.py")]
    proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout[-2000:] + proc.stderr[-2000:]
    assert "bit_identical=True" in proc.stdout
