"""GPU parity on vector-valued spaces (dimRange > 1, b200fem_space_create_vector) and of run-time compiled integrands on continuous
Lagrange spaces: the device kernels (dg_quadrature.cuh / lagrange_quadrature.cuh with NVRTC-compiled integrands) against the CPU
oracle integrating the SAME source text (fem_oracle.cpp: VectorOperator).  Tolerance 1e-12 of max|w|.  The first test is the
reference's own matrix-free check on a vector-valued space (dune/fempy/test/testoperator.py) on its own configuration."""
import os

import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))
SRC = {name: open(os.path.join(HERE, "integrands", name + ".cuh")).read() for name in ("testoperator_vector", "testoperator_vector_lin", "system_dg", "adr_variable")}


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def test_reference_testoperator_configuration():
    """testoperator.py:14-66: Lagrange order 2, dimRange 2, structuredGrid([0,0],[1,1],[40,40]); op(ubar) == linop(ubar), and both
    equal the oracle's"""
    n, R = [40, 40], 2
    grid = fem.structuredGrid([0.0, 0.0], [1.0, 1.0], n)
    space = fem.space.lagrange(grid, order=2, dimRange=R)
    osp = ol.Space(n, [0.0, 0.0], [1.0, 1.0], ol.LAGRANGE, 2)
    assert space.size == osp.size * R == 81 * 81 * 2
    ubar = np.repeat((osp.node_positions() ** 2).sum(axis=1), R)
    op = fem.operator.galerkinJit(space, SRC["testoperator_vector"], skeleton=False, boundary=False)
    linop = fem.operator.galerkinJit(space, SRC["testoperator_vector_lin"], skeleton=False, boundary=False)
    a, d = np.empty(space.size), np.empty(space.size)
    op(ubar, a)
    linop(ubar, d)
    ref = ol.VectorUserOperator(osp, R, SRC["testoperator_vector"], skeleton=False, boundary=False).apply(ubar)
    assert rel(a, ref) < TOL and rel(d, ref) < TOL
    # err = integrate((destA - destD)**2) < 1e-15 (testoperator.py:64-66): bounded by |Omega| * max|a - d|^2 * max|phi|^2
    assert np.abs(a - d).max() ** 2 < 1e-15
    # random argument (the quadratic ubar exercises few modes)
    u = np.random.default_rng(1).uniform(-1, 1, space.size)
    op(u, a)
    assert rel(a, ol.VectorUserOperator(osp, R, SRC["testoperator_vector"], skeleton=False, boundary=False).apply(u)) < TOL


@pytest.mark.parametrize("dim,order,R,n", [(2, 1, 2, [9, 7]), (2, 2, 3, [6, 5]), (3, 1, 2, [5, 4, 3]), (3, 2, 2, [4, 3, 3]), (3, 2, 3, [3, 3, 2]), (3, 1, 4, [4, 4, 3])])
def test_vector_lagrange_spaces(dim, order, R, n):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=order, dimRange=R)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, order)
    assert space.size == osp.size * R
    u = np.random.default_rng(dim + 10 * order + 100 * R).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    # the system form without its skeleton terms (continuous space), boundary terms on
    op = fem.operator.galerkinJit(space, SRC["system_dg"], [0.05, 0.02, 0.7, 20.0 * order ** 2], skeleton=False, boundary=True)
    oop = ol.VectorUserOperator(osp, R, SRC["system_dg"], [0.05, 0.02, 0.7, 20.0 * order ** 2], skeleton=False, boundary=True)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    assert rel(op.loadVector(), -oop.apply(np.zeros(space.size))) < TOL
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < 1e-11


@pytest.mark.parametrize("kind,dim,order,R,n", [("hier", 3, 1, 2, [5, 4, 3]), ("hier", 3, 2, 2, [4, 3, 3]), ("hier", 3, 2, 3, [3, 3, 2]), ("hier", 3, 3, 2, [3, 2, 2]),
                                                ("hier", 3, 1, 4, [4, 3, 3]), ("hier", 2, 2, 2, [7, 5]), ("onb", 2, 2, 3, [6, 5]), ("onb", 3, 2, 2, [4, 3, 3])])
def test_vector_dg_spaces(kind, dim, order, R, n):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    g = fem.structuredGrid(lo, hi, n)
    if kind == "onb":
        space, osp = fem.space.dgonb(g, order=order, dimRange=R), ol.Space(n, lo, hi, ol.DG_ONB, order)
    else:
        space, osp = fem.space.dglegendre(g, order=order, hierarchical=True, dimRange=R), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, order)
    assert space.size == osp.size * R
    const = [0.05, 0.02, 0.7, 20.0 * order ** 2]
    u = np.random.default_rng(dim + 10 * order + 100 * R).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    op = fem.operator.galerkinJit(space, SRC["system_dg"], const)
    oop = ol.VectorUserOperator(osp, R, SRC["system_dg"], const)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    assert op.timing()["kernel"] == _capi.KERNEL_QUADRATURE
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < 1e-11
    const[0], const[2] = 0.2, 0.0
    op.setConstants(const)
    op(u, w)
    assert rel(w, ol.VectorUserOperator(osp, R, SRC["system_dg"], const).apply(u)) < TOL


@pytest.mark.parametrize("dim,order,n", [(2, 1, [9, 7]), (2, 2, [8, 5]), (3, 1, [5, 4, 3]), (3, 2, [4, 3, 3])])
def test_compiled_scalar_integrands_on_lagrange_spaces(dim, order, n):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=order)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, order)
    const = [0.05, 1.0, -0.5, 0.25, 20.0 * order ** 2, 0.3, 0.7]
    u = np.random.default_rng(dim + 10 * order).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    op = fem.operator.galerkinJit(space, SRC["adr_variable"], const, skeleton=False, boundary=True)
    oop = ol.UserOperator(osp, SRC["adr_variable"], const, skeleton=False, boundary=True)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL


def test_vector_space_errors():
    g = fem.structuredGrid([0.0] * 3, [1.0] * 3, [2, 2, 2])
    space = fem.space.dglegendre(g, order=1, dimRange=2)
    with pytest.raises(_capi.B200FemError):      # the built-in family is scalar
        fem.operator.galerkin(space)
    with pytest.raises(_capi.B200FemError):
        fem.space.dglegendre(g, order=1, dimRange=5)
    with pytest.raises(_capi.B200FemError):      # skeleton terms on a continuous space
        fem.operator.galerkinJit(fem.space.lagrange(g, order=1), SRC["adr_variable"], [0.0] * 7, skeleton=True, boundary=True)


def test_vector_space_krylov_solve():
    """a linear vector-valued problem (reaction coupling off) solved with GMRES through the vector operator: root of L"""
    n, R = [6, 6], 2
    lo, hi = [-1.0] * 2, [1.0, 0.5]
    space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=2, dimRange=R)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    const = [0.5, 0.1, 0.0, 80.0]
    op = fem.operator.galerkinJit(space, SRC["system_dg"], const)
    inv = fem.solver.GmresInverseOperator({"tolerance": 1e-11, "maxiterations": 4000, "gmres.restart": 50})
    inv.bind(op)
    x = np.zeros(space.size)
    b = op.loadVector()
    inv(b, x)
    assert inv.iterations > 0
    r = ol.VectorUserOperator(osp, R, SRC["system_dg"], const).apply(x)
    assert np.abs(r).max() < 1e-8 * np.abs(b).max()
