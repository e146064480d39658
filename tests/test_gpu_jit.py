"""GPU parity of run-time compiled integrands (b200fem_operator_create_jit): a variable-coefficient, non-linear
advection-diffusion-reaction form that the built-in family cannot express, compiled by NVRTC into the generic quadrature
kernel -- against the CPU oracle integrating the SAME source text compiled for the host.  Tolerance 1e-12 of max|w|."""
import os

import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE = open(os.path.join(HERE, "integrands", "adr_variable.cuh")).read()
CONST = [0.05, 1.0, -0.5, 0.25, 80.0, 0.3, 0.7]       # k0, b0, b1, b2, penalty, c0, c1 (cubic reaction)


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def spaces(kind, dim, order, n):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    g = fem.structuredGrid(lo, hi, n)
    if kind == "onb":
        return fem.space.dgonb(g, order=order), ol.Space(n, lo, hi, ol.DG_ONB, order)
    return fem.space.dglegendre(g, order=order, hierarchical=True), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, order)


@pytest.mark.parametrize("kind,dim,order,n", [("hier", 3, 1, [5, 4, 3]), ("hier", 3, 2, [5, 4, 3]), ("hier", 3, 3, [3, 3, 2]), ("hier", 3, 4, [3, 2, 2]),
                                              ("hier", 2, 2, [7, 5]), ("onb", 2, 2, [7, 5]), ("onb", 3, 2, [4, 3, 3])])
def test_compiled_integrands_match_the_oracle(kind, dim, order, n):
    space, osp = spaces(kind, dim, order, n)
    const = list(CONST)
    const[4] = 20.0 * order ** 2
    op = fem.operator.galerkinJit(space, SOURCE, const)
    oop = ol.UserOperator(osp, SOURCE, const)
    u = np.random.default_rng(order + 10 * dim).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    assert op.timing()["kernel"] == _capi.KERNEL_QUADRATURE
    # b = -L[0] and the difference L[u] - L[0]
    assert rel(op.loadVector(), -oop.apply(np.zeros(space.size))) < TOL
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < 1e-11
    # constants change without recompilation (dune.ufl.Constant)
    const[0], const[6] = 0.2, 0.0
    op.setConstants(const)
    oop2 = ol.UserOperator(osp, SOURCE, const)
    op(u, w)
    assert rel(w, oop2.apply(u)) < TOL
    # non-default quadrature orders select another compiled kernel
    if order <= 3 and kind == "hier":
        op.setQuadratureOrders(2 * order + 2, 2 * order + 2)
        osp_q = ol.Space(n, [-1.0] * dim, [1.0, 0.5, 2.0][:dim], ol.DG_LEGENDRE_HIER, order, interior_order=2 * order + 2, surface_order=2 * order + 2)
        op(u, w)
        assert rel(w, ol.UserOperator(osp_q, SOURCE, const).apply(u)) < TOL


def test_compiled_linear_integrands_are_solved_with_gmres():
    """linear variable-coefficient problem (cubic term off): GMRES on A u = L[u] - L[0] converges to the root of L"""
    space, osp = spaces("hier", 2, 2, [8, 8])
    const = list(CONST)
    const[0], const[4], const[6] = 0.5, 80.0, 0.0
    op = fem.operator.galerkinJit(space, SOURCE, const)
    inv = fem.solver.GmresInverseOperator({"tolerance": 1e-11, "maxiterations": 4000, "gmres.restart": 50})
    inv.bind(op)
    x = np.zeros(space.size)
    b = op.loadVector()
    inv(b, x)
    assert inv.iterations > 0
    r = ol.UserOperator(osp, SOURCE, const).apply(x)
    assert np.abs(r).max() < 1e-8 * np.abs(b).max()


def test_compile_error_raises_with_the_log():
    space, _ = spaces("hier", 3, 1, [2, 2, 2])
    op = fem.operator.galerkinJit(space, SOURCE + "\n__device__ void broken() { undefined_symbol(); }\n", CONST)
    with pytest.raises(_capi.B200FemError) as ei:
        op(np.zeros(space.size), np.empty(space.size))
    assert "undefined_symbol" in str(ei.value)


def _oracle_newton(oop, w0, tol, maxit, lin_tol, lin_maxit, restart, tolcrit=2):
    """NewtonInverseOperator::operator() (newtoninverseoperator.hh:690-803) restated on the oracle's pieces (u = 0, no line search)"""
    w = w0.copy()
    res = oop.apply(w)
    delta = np.sqrt(res @ res)
    it = lit = 0
    while True:
        oop.linearize(w)
        if lin_maxit - lit <= 0:
            break
        li, dw, _ = oop.gmres_jacobian(res, np.zeros_like(w), lin_tol, lin_maxit - lit, tolcrit, restart)
        if li < 0:
            lit = li
            break
        lit += li
        w -= dw
        res = oop.apply(w)
        delta = np.sqrt(res @ res)
        it += 1
        if delta < tol or it >= maxit or lit >= lin_maxit:
            break
    return it, lit, delta, w


@pytest.mark.parametrize("builtin", [True, False])
def test_newton_inverse_operator(builtin):
    """the non-linear reaction-diffusion problem solved in ONE call: same Newton iterates as the restatement on the oracle (iteration
    counts, final residual and solution), built-in model and run-time compiled integrands"""
    n, lo, hi = [4, 4, 3], [-1.0] * 3, [1.0] * 3
    space, osp = spaces("hier", 3, 2, n)
    space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=2)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    if builtin:
        kw = dict(eps=0.5, b=(1.0, 0.0, 0.0), c=1.0, gamma=2.0, beta=80.0, dirichlet_mask=0b000011, data=1)
        op, oop = fem.operator.galerkin(space, **kw), ol.Operator(osp, skeleton=True, boundary=True, **kw)
    else:
        const = [0.5, 1.0, -0.5, 0.25, 80.0, 1.0, 2.0]
        op, oop = fem.operator.galerkinJit(space, SOURCE, const), ol.UserOperator(osp, SOURCE, const)
    # (the difference quotient carries ~1e-8 of rounding noise: the linear solves are asked for a residual REDUCTION and the Newton
    # tolerance sits above that floor, as in test_difference_quotient_jacobian_and_newton_krylov)
    newton = fem.solver.NewtonInverseOperator({"tolerance": 1e-7, "linear.method": "gmres", "linear.tolerance": 1e-7, "linear.errormeasure": "residualreduction",
                                               "linear.maxiterations": 4000, "linear.gmres.restart": 30})
    newton.bind(op)
    w = np.zeros(space.size)
    newton(None, w)
    it, lit, delta, w_ref = _oracle_newton(oop, np.zeros(space.size), 1e-7, 2 ** 31 - 1, 1e-7, 4000, 30)
    assert newton.converged and newton.iterations == it and 2 <= it <= 12
    assert abs(newton.linearIterations - lit) <= max(5, lit // 10)
    assert newton.residual < 1e-7 and np.linalg.norm(oop.apply(w)) < 2e-7
    assert rel(w, w_ref) < 1e-7
    # a linear iteration budget that is too small: NewtonFailure::LinearSolverFailed (the Krylov solver reports a negative count)
    newton = fem.solver.NewtonInverseOperator({"tolerance": 1e-9, "linear.tolerance": 1e-13, "linear.maxiterations": 5, "linear.gmres.restart": 5})
    newton.bind(op)
    w = np.zeros(space.size)
    newton(None, w)
    assert not newton.converged and newton.failure in (6, 7)


def test_newton_line_search():
    """NewtonInverseOperator::lineSearch ("simple", newtoninverseoperator.hh:588-629) on the device against the restatement that
    tests/test_reference_pieces.py pins to the reference's own class: an indefinite reaction term (c < 0) makes full Newton steps overshoot,
    the search halves the step in iterations 4 and 6 (the case is stable under 1e-7 perturbations of the initial guess: same iteration
    count and halvings) and the iteration count differs from plain Newton's."""
    n, lo, hi = [3, 3, 2], [-1.0] * 3, [1.0] * 3
    space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=1)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 1)
    kw = dict(eps=0.5, b=(1.0, 0.0, 0.0), c=-8.0, gamma=5.0, beta=40.0, dirichlet_mask=0b000011, data=1)
    op, oop = fem.operator.galerkin(space, **kw), ol.Operator(osp, skeleton=True, boundary=True, **kw)
    w0 = 2.0 * np.random.default_rng(4).uniform(-1, 1, space.size)
    counts = {}
    for search in ("simple", "none"):
        trace = []
        it, lit, fail, delta, w_ref = ol.newton(oop, w0, 1e-7, 40, 1e-8, 20000, 48, 2, search == "simple", trace=trace)
        assert fail == 0 and (sum(trace) >= 1) == (search == "simple")
        newton = fem.solver.NewtonInverseOperator({"tolerance": 1e-7, "maxiterations": 40, "linesearch": search, "linear.method": "gmres", "linear.tolerance": 1e-8,
                                                   "linear.errormeasure": "residualreduction", "linear.maxiterations": 20000, "linear.gmres.restart": 48})
        newton.bind(op)
        w = w0.copy()
        newton(None, w)
        assert newton.converged and newton.iterations == it
        assert abs(newton.linearIterations - lit) <= max(5, lit // 10)
        assert newton.residual < 1e-7 and rel(w, w_ref) < 1e-6
        counts[search] = it
    assert counts["simple"] != counts["none"]
