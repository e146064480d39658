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


# ---- vector-valued spaces (dimRange > 1) and continuous Lagrange spaces ----
VSRC = {name: open(os.path.join(HERE, "integrands", name + ".cuh")).read() for name in ("testoperator_vector", "testoperator_vector_lin", "system_dg")}


@pytest.mark.parametrize("name,kind,dim,order,R,skel,bnd", [("testoperator_vector", _capi.LAGRANGE, 2, 2, 2, 0, 0), ("testoperator_vector", _capi.LAGRANGE, 3, 2, 3, 0, 0),
                                                            ("testoperator_vector", _capi.LAGRANGE, 3, 1, 2, 0, 0), ("system_dg", _capi.DG_LEGENDRE_HIER, 3, 2, 2, 1, 1),
                                                            ("system_dg", _capi.DG_ONB, 2, 3, 3, 1, 1), ("system_dg", _capi.DG_LEGENDRE_HIER, 3, 1, 4, 1, 1)])
def test_vector_integrands_compile_into_the_kernels(name, kind, dim, order, R, skel, bnd):
    log = C.create_string_buffer(1 << 16)
    rc = _capi.lib().b200fem_jit_compile_check_space(VSRC[name].encode(), kind, dim, order, R, skel, bnd, log, len(log))
    assert rc == 0, log.value.decode()


def test_scalar_integrands_compile_into_the_lagrange_kernels():
    log = C.create_string_buffer(1 << 16)
    for dim in (2, 3):
        for order in (2, 3):
            rc = _capi.lib().b200fem_jit_compile_check_space(SOURCE.encode(), _capi.LAGRANGE, dim, order, 1, 0, 1, log, len(log))
            assert rc == 0, log.value.decode()
    rc = _capi.lib().b200fem_jit_compile_check_space(SOURCE.encode(), _capi.LAGRANGE, 2, 4, 1, 0, 1, log, len(log))
    assert rc == _capi.ERR_NOT_IMPLEMENTED                      # Lagrange orders 1..3


def test_vector_oracle_reduces_to_the_scalar_one_for_uncoupled_components():
    """VectorOperator (fem_oracle.cpp) against the scalar element loop: R independent copies of the scalar form"""
    scalar = open(os.path.join(HERE, "integrands", "adr_variable.cuh")).read()
    # the scalar source applied per component: wrap it
    wrap = "namespace sc {\nstruct PointValue { double u; double du[3]; };\nstruct PointRange { double s; double F[3]; };\n" + scalar + "\n}\n" + """
__device__ void interior(const double* x, const VectorValue& u, VectorRange& r, const double* c, int dim) {
  for (int k = 0; k < dimRange; ++k) { sc::PointValue v; sc::PointRange q; v.u = u.u[k]; q.s = 0; for (int d = 0; d < 3; ++d) { v.du[d] = u.du[k][d]; q.F[d] = 0; }
    sc::interior(x, v, q, c, dim); r.s[k] = q.s; for (int d = 0; d < 3; ++d) r.F[k][d] = q.F[d]; } }
__device__ void skeleton(const double* x, int axis, double sign, double ihe, const VectorValue& in, const VectorValue& out, VectorRange& rin, VectorRange& rout, const double* c, int dim) {
  for (int k = 0; k < dimRange; ++k) { sc::PointValue a, b; sc::PointRange p, q; a.u = in.u[k]; b.u = out.u[k]; p.s = q.s = 0; for (int d = 0; d < 3; ++d) { a.du[d] = in.du[k][d]; b.du[d] = out.du[k][d]; p.F[d] = q.F[d] = 0; }
    sc::skeleton(x, axis, sign, ihe, a, b, p, q, c, dim); rin.s[k] = p.s; rout.s[k] = q.s; for (int d = 0; d < 3; ++d) { rin.F[k][d] = p.F[d]; rout.F[k][d] = q.F[d]; } } }
__device__ void boundary(const double* x, int axis, int side, double ihbnd, const VectorValue& u, VectorRange& r, const double* c, int dim) {
  for (int k = 0; k < dimRange; ++k) { sc::PointValue v; sc::PointRange q; v.u = u.u[k]; q.s = 0; for (int d = 0; d < 3; ++d) { v.du[d] = u.du[k][d]; q.F[d] = 0; }
    sc::boundary(x, axis, side, ihbnd, v, q, c, dim); r.s[k] = q.s; for (int d = 0; d < 3; ++d) r.F[k][d] = q.F[d]; } }
"""
    const = [0.05, 1.0, -0.5, 0.25, 80.0, 0.3, 0.7]
    for kind, n, skel in ((ol.DG_LEGENDRE_HIER, [4, 3, 2], True), (ol.LAGRANGE, [4, 3], False), (ol.DG_ONB, [4, 3], True)):
        lo, hi = [-1.0] * len(n), [1.0, 0.5, 2.0][:len(n)]
        sp = ol.Space(n, lo, hi, kind, 2)
        R = 3
        u = np.random.default_rng(5).uniform(-1, 1, sp.size * R)
        vop = ol.VectorUserOperator(sp, R, wrap, const, skeleton=skel, boundary=True)
        sop = ol.UserOperator(sp, scalar, const, skeleton=skel, boundary=True)
        w = vop.apply(u).reshape(-1, R)
        for k in range(R):
            ref = sop.apply(np.ascontiguousarray(u.reshape(-1, R)[:, k]))
            assert np.abs(w[:, k] - ref).max() <= 1e-14 * np.abs(ref).max()


def test_reference_vector_operator_check_holds_on_the_oracle():
    """dune/fempy/test/testoperator.py:55-66 on its own configuration (Lagrange order 2, dimRange 2, 40 x 40): op(ubar) == linop(ubar)"""
    n, R = [40, 40], 2
    sp = ol.Space(n, [0.0, 0.0], [1.0, 1.0], ol.LAGRANGE, 2)
    ubar = np.repeat((sp.node_positions() ** 2).sum(axis=1), R)        # interpolate(as_vector([dot(x,x),]*dimR))
    op = ol.VectorUserOperator(sp, R, VSRC["testoperator_vector"], skeleton=False, boundary=False)
    linop = ol.VectorUserOperator(sp, R, VSRC["testoperator_vector_lin"], skeleton=False, boundary=False)
    a, d = op.apply(ubar), linop.apply(ubar)
    assert np.abs(a).max() > 1e-3
    assert np.abs(a - d).max() < 1e-15 + 1e-13 * np.abs(a).max()


def test_scalar_integrands_compile_into_the_unstructured_kernel():
    log = C.create_string_buffer(1 << 16)
    for dim in (2, 3):
        for order in (1, 2):
            rc = _capi.lib().b200fem_jit_compile_check_unstructured(SOURCE.encode(), dim, order, log, len(log))
            assert rc == 0, log.value.decode()
