"""The C++ host mirror (dune_fem_b200/host/b200fem.hh) through its self test executable."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dune_fem_b200", "lib", "host_selftest")


def _run():
    if not os.path.exists(EXE):
        import __graft_entry__
        __graft_entry__.build()
    return subprocess.run([EXE], capture_output=True, text=True, timeout=300)


def test_host_mirror_reports_missing_device_cleanly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run()
    assert r.returncode == 0 and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_host_mirror_apply_and_cg_on_gpu():
    r = _run()
    assert r.returncode == 0, r.stdout + r.stderr
    assert "host selftest OK" in r.stdout
