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


# ---- the reference-side binding dune/fem/schemes/b200galerkin.hh, compiled against stand-in DUNE headers (tests/dune_stub) ----
EXE2 = os.path.join(ROOT, "dune_fem_b200", "lib", "dune_binding_selftest")


def _run2():
    if not os.path.exists(EXE2):
        import __graft_entry__
        __graft_entry__.build()
    return subprocess.run([EXE2], capture_output=True, text=True, timeout=300)


def test_dune_binding_compiles_and_reports_missing_device_as_dune_exception():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = _run2()
    assert r.returncode == 1 and "InvalidStateException" in r.stdout and "no CPU fallback" in r.stdout


@pytest.mark.gpu
def test_dune_binding_apply_and_gmres_match_the_oracle():
    """B200GalerkinOperator< GeneratedIntegrands, DF > on a dgonb P2 space over an 8x8 mesh, used through Dune::Fem::Operator, and
    B200KrylovInverseOperator (GMRES): the apply and the solution against the oracle integrating the same source."""
    import re
    import sys
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    r = _run2()
    assert r.returncode == 0, r.stdout + r.stderr
    src = open(os.path.join(ROOT, "tests", "dune_binding_selftest.cpp")).read()
    literals = lambda text: "".join(eval(m) for m in re.findall(r'^\s*(?:return\s+)?("(?:[^"\\]|\\.)*")\s*;?\s*$', text, flags=re.M))
    body = literals(src[src.index("struct GeneratedIntegrands"):src.index("struct GeneratedInteriorIntegrands")])
    sp = ol.Space([8, 8], [-1.0, -1.0], [1.0, 1.0], ol.DG_ONB, 2)
    op = ol.UserOperator(sp, body, [0.5, 0.3, 80.0])
    u = np.sin(0.37 * np.arange(sp.size))
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    ref = float(np.sum(op.apply(u) ** 2))
    assert abs(float(out["apply_norm2"]) - ref) < 1e-11 * ref
    # the solution of L[x] = 0, i.e. A x = -L[0]: the oracle's operator probed into a dense matrix (384 dofs) and solved directly
    b = -op.apply(np.zeros(sp.size))
    A = np.stack([op.apply(e, linear=True) for e in np.eye(sp.size)], axis=1)
    x = np.linalg.solve(A, b)
    for i in range(6):
        assert abs(float(out[f"x{i}"]) - x[i]) < 1e-7 * np.abs(x).max()
    # the unstructured path of the binding: the (stub) ALUGrid-like grid part is walked once, the arrays go to b200fem_mesh_unstructured;
    # P2 on a 3 x 3 patch of distorted quadrilaterals, interior-only generated integrands, against the oracle on the same arrays
    body2 = literals(src[src.index("struct GeneratedInteriorIntegrands"):src.index("static int unstructuredPart")])
    n = 3
    vx = np.array([[i / n + 0.05 * np.sin(5.0 * j / n) * (i / n) * (1 - i / n), j / n + 0.04 * np.sin(4.0 * i / n) * (j / n) * (1 - j / n)]
                   for j in range(n + 1) for i in range(n + 1)])
    cubes = np.array([[i + (n + 1) * j, i + (n + 1) * j + 1, i + (n + 1) * j + n + 1, i + (n + 1) * j + n + 2] for j in range(n) for i in range(n)], dtype=np.int64)
    uop = ol.UnstructuredOperator(vx, cubes, 2, user_source=body2, constants=[0.7, 0.4])
    assert uop.size == 49
    wu = uop.apply(np.sin(0.37 * np.arange(uop.size)))
    assert abs(float(out["unstructured_apply_norm2"]) - float(np.sum(wu ** 2))) < 1e-11 * float(np.sum(wu ** 2))
    for i in range(4):
        assert abs(float(out[f"uw{i}"]) - wu[i]) < 1e-12 * np.abs(wu).max()
