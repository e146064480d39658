"""Run-time compiled integrands, the part that needs no GPU: NVRTC turns the user's source + the library's own quadrature
kernel header into sm_100a code (b200fem_jit_compile_check), compile errors come back with the log, and the same source text
compiles for the host into the callbacks the oracle integrates."""
import ctypes as C
import os

import numpy as np
import pytest

from dune_fem_b200 import _capi
import oracle_lib as ol

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = open(os.path.join(HERE, "integrands", "adr_variable.cuh")).read()


def _nvrtc_present():
    for n in ("libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12"):
        try:
            C.CDLL(n)
            return True
        except OSError:
            pass
    return False


pytestmark = pytest.mark.skipif(not _nvrtc_present(), reason="libnvrtc.so.12 not installed")


@pytest.mark.parametrize("order", [1, 2, 4])
def test_user_integrands_compile_into_the_quadrature_kernel(order):
    log = C.create_string_buffer(1 << 16)
    rc = _capi.lib().b200fem_jit_compile_check(SOURCE.encode(), order, log, len(log))
    assert rc == 0, log.value.decode()


def test_compile_errors_are_reported_with_the_log():
    log = C.create_string_buffer(1 << 16)
    rc = _capi.lib().b200fem_jit_compile_check((SOURCE + "\n__device__ void broken() { undefined_symbol(); }\n").encode(), 2, log, len(log))
    assert rc == -1                                           # B200FEM_ERR_INVALID
    assert "undefined_symbol" in log.value.decode()
    assert "undefined_symbol" in _capi.lib().b200fem_last_error().decode()


def test_oracle_integrates_the_same_source_on_the_host():
    """the callbacks reproduce the built-in family when they are given the same form (constant coefficients)"""
    src = """
__device__ void interior(const double* x, const PointValue& u, PointRange& r, const double* c, int dim) {
  r.s = c[2] * u.u; for (int d = 0; d < dim; ++d) r.F[d] = c[0] * u.du[d] - (d == 0 ? c[1] : 0.0) * u.u; }
__device__ void skeleton(const double* x, int axis, double sign, double ihe, const PointValue& in, const PointValue& out,
                         PointRange& rin, PointRange& rout, const double* c, int dim) {
  const double jump = in.u - out.u, bn = (axis == 0 ? c[1] : 0.0) * sign;
  const double cj = c[0] * c[3] * ihe * jump - c[0] * 0.5 * (in.du[axis] + out.du[axis]) * sign + 0.5 * (bn + fabs(bn)) * in.u - 0.5 * (-bn + fabs(bn)) * out.u;
  rin.s = cj; rout.s = -cj; rin.F[axis] = rout.F[axis] = -c[0] * jump * 0.5 * sign; }
__device__ void boundary(const double* x, int axis, int side, double ihbnd, const PointValue& u, PointRange& r, const double* c, int dim) {
  if (axis != 0) return;
  const double sign = side ? 1.0 : -1.0, bn = c[1] * sign;
  r.s = c[0] * c[3] * ihbnd * u.u + 0.5 * (bn + fabs(bn)) * u.u; }
"""
    sp = ol.Space([4, 3, 2], [-1.0] * 3, [1.0] * 3, ol.DG_LEGENDRE_HIER, 2)
    u = np.random.default_rng(0).uniform(-1, 1, sp.size)
    builtin = ol.Operator(sp, eps=0.1, b=(1.0, 0.0, 0.0), c=0.3, beta=80.0, dirichlet_mask=0b000011, data=0, skeleton=True, boundary=True)
    user = ol.UserOperator(sp, src, constants=[0.1, 1.0, 0.3, 80.0])
    ref = builtin.apply(u)
    assert np.abs(user.apply(u) - ref).max() < 1e-13 * np.abs(ref).max()
    assert np.abs(user.apply(u, linear=True) - ref).max() < 1e-13 * np.abs(ref).max()
