"""GPU parity for continuous Lagrange spaces of order 3 (quadrature kernels with the generalised closed-form dof map: several nodes
inside an edge / face / cell) against the oracle: numbering bit-exact, values to 1e-12, CG iterates, compiled integrands and
vector-valued P3."""
import os

import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12
HERE = os.path.dirname(os.path.abspath(__file__))


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("dim,n", [(2, [7, 5]), (3, [4, 3, 3])])
def test_order3_dofmap_apply_dirichlet_and_diagonal(dim, n):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=3)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, 3)
    assert space.size == osp.size
    for e in range(osp.elements):
        assert (space.mapper(e) == osp.dofmap(e)).all()
    full = (1 << (2 * dim)) - 1
    kw = dict(eps=0.7, b=(1.0, -0.5, 0.25)[:dim], c=0.3, gamma=0.5, data=2, dirichlet_mask=full & 0b011011, strong_dirichlet=True)
    op, oop = fem.operator.galerkin(space, **kw), ol.Operator(osp, **kw)
    assert op.nonlinear
    mask, vals = op.dirichlet()
    omask, ovals = oop.dirichlet()
    assert (mask == omask).all() and np.abs(vals - ovals).max() < 1e-14
    u = np.random.default_rng(dim).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    assert op.timing()["kernel"] == _capi.KERNEL_QUADRATURE
    # linear model: affine structure, matrix-free diagonal
    kwl = dict(kw, gamma=0.0)
    opl, oopl = fem.operator.galerkin(space, **kwl), ol.Operator(osp, **kwl)
    opl(u, w)
    assert rel(w, oopl.apply(u)) < TOL
    opl.applyLinear(u, w)
    assert rel(w, oopl.apply(u, linear=True)) < TOL
    assert rel(opl.loadVector(), -oopl.apply(np.zeros(space.size))) < TOL
    assert rel(opl.diagonal(), oopl.diagonal()) < TOL
    with pytest.raises(_capi.B200FemError):             # the lattice kernels carry orders 1 and 2
        fem.operator.galerkin(space, kernel=_capi.KERNEL_KRONECKER, **kwl)(u, w)


def test_order3_cg_iterates_match_the_oracle():
    n, lo, hi = [6, 5], [0.0, 0.0], [1.0, 1.0]
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=3)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, 3)
    kw = dict(eps=1.0, c=0.5, data=2, dirichlet_mask=0b1111, strong_dirichlet=True)
    op, oop = fem.operator.galerkin(space, **kw), ol.Operator(osp, **kw)
    b = op.loadVector()
    inv = fem.solver.CgInverseOperator({"tolerance": 1e-12, "maxiterations": 400})
    inv.bind(op)
    x = np.zeros(space.size)
    inv(b, x)
    it, xo, hist = oop.cg(b, np.zeros(space.size), 1e-12, 400)
    assert inv.iterations == it > 3           # (the data are close to an eigenfunction: CG needs few steps)
    assert rel(x, xo) < 1e-10 and np.allclose(inv.residuals[:it - 1], hist[:it - 1], rtol=1e-8)   # (the last residual is rounding noise: the Krylov space is exhausted)
    assert osp.l2error(x, 2) < 2e-4            # fourth-order space on a 6 x 5 mesh


@pytest.mark.parametrize("dim,n,R", [(2, [5, 4], 1), (3, [3, 3, 2], 1), (2, [4, 4], 2)])
def test_order3_compiled_integrands(dim, n, R):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=3, dimRange=R)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, 3)
    u = np.random.default_rng(7).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    if R == 1:
        src = open(os.path.join(HERE, "integrands", "adr_variable.cuh")).read()
        const = [0.05, 1.0, -0.5, 0.25, 180.0, 0.3, 0.7]
        op = fem.operator.galerkinJit(space, src, const, skeleton=False, boundary=True)
        ref = ol.UserOperator(osp, src, const, skeleton=False, boundary=True).apply(u)
    else:
        src = open(os.path.join(HERE, "integrands", "system_dg.cuh")).read()
        const = [0.05, 0.02, 0.7, 180.0]
        op = fem.operator.galerkinJit(space, src, const, skeleton=False, boundary=True)
        ref = ol.VectorUserOperator(osp, R, src, const, skeleton=False, boundary=True).apply(u)
    op(u, w)
    assert rel(w, ref) < TOL
