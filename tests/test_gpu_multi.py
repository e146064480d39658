"""Multi-GPU parity under pytest: spawns tests/mgpu_check.py under torchrun on 2 (and, if present, 4) GPUs of this box.
Skipped on boxes with fewer than 2 devices (bench.py --gpus N carries the same check in its JSON line as `parity`)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("world", [2, 4])
def test_multi_gpu_parity_against_single_domain_oracle(world):
    if _device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    port = 29500 + (os.getpid() % 400) + world
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert "mgpu_check OK" in r.stdout


def test_no_p2p_environment_selects_the_nccl_transport():
    """B200FEM_NO_P2P at communicator attach: same results through ncclSend/Recv + ncclAllReduce (the fallback path)"""
    if _device_count() < 2:
        pytest.skip("needs 2 GPUs")
    port = 29900 + (os.getpid() % 90)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "mgpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=ROOT, env=dict(os.environ, B200FEM_NO_P2P="1"))
    assert r.returncode == 0, r.stdout[-3000:] + "\n" + r.stderr[-6000:]
    assert "mgpu_check OK" in r.stdout and "transport=nccl" in r.stdout
