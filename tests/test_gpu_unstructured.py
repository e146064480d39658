"""GPU parity on UNSTRUCTURED conforming cube meshes (b200fem_mesh_unstructured; lagrange_unstructured.cuh: element -> dof index
arrays, per-element multilinear geometry, colour-ordered scatter) against the oracle's restatement (fem_oracle.cpp:
UnstructuredLagrange): numbering bit-exact, values to 1e-12 of max|w|; CG / Jacobi-PCG on a distorted mesh against a dense solve of
the oracle's operator."""
import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol
from test_unstructured_cpu import distorted

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("dim,n", [(2, [9, 7]), (3, [5, 4, 3])])
@pytest.mark.parametrize("order", [1, 2])
def test_cartesian_mesh_as_unstructured(dim, n, order):
    """the same mesh through both paths of the library: index arrays + geometry kernel == closed-form Cartesian kernels
    (adaptive-leaf numbering), and both == the oracle"""
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    coords, elems = ol.cartesian_as_unstructured(n, lo, hi)
    kw = dict(eps=0.7, b=(1.0, -0.5, 0.25)[:dim], c=0.3, gamma=0.5, data=2, strong_dirichlet=True)
    uspace = fem.space.lagrange(fem.unstructuredGrid(coords, elems), order=order)
    sspace = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=order, numbering=_capi.NUMBERING_ADAPTIVE_LEAF)
    oop = ol.UnstructuredOperator(coords, elems, order, **kw)
    assert uspace.size == sspace.size == oop.size
    for e in range(len(elems)):
        assert (uspace.mapper(e) == oop.dofmap(e)).all() and (uspace.mapper(e) == sspace.mapper(e)).all()
    uop = fem.operator.galerkin(uspace, dirichlet_mask=1, **kw)
    sop = fem.operator.galerkin(sspace, dirichlet_mask=(1 << (2 * dim)) - 1, kernel=_capi.KERNEL_QUADRATURE, **kw)
    u = np.random.default_rng(dim + order).uniform(-1, 1, uspace.size)
    wu, ws = np.empty(uspace.size), np.empty(uspace.size)
    uop(u, wu)
    sop(u, ws)
    ref = oop.apply(u)
    assert rel(wu, ref) < TOL and rel(ws, ref) < TOL
    assert uop.timing()["kernel"] == _capi.KERNEL_QUADRATURE


@pytest.mark.parametrize("dim,n", [(2, [11, 8]), (3, [6, 5, 4])])
@pytest.mark.parametrize("order", [1, 2])
def test_distorted_shuffled_mesh(dim, n, order):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    coords, elems = distorted(n, lo, hi, seed=3 * dim + order)
    kw = dict(eps=0.7, b=(1.0, -0.5, 0.25)[:dim], c=0.3, gamma=0.5, data=2, strong_dirichlet=True)
    space = fem.space.lagrange(fem.unstructuredGrid(coords, elems), order=order)
    oop = ol.UnstructuredOperator(coords, elems, order, **kw)
    assert space.size == oop.size
    for e in range(len(elems)):
        assert (space.mapper(e) == oop.dofmap(e)).all()                  # first-touch numbering: bit-exact
    op = fem.operator.galerkin(space, dirichlet_mask=1, **kw)
    mask, vals = op.dirichlet()
    x, bnd = oop.nodes()
    assert (mask == bnd).all()
    g = np.prod(np.sin(np.pi * x), axis=1)
    assert np.abs(vals[bnd == 1] - g[bnd == 1]).max() < 1e-14
    u = np.random.default_rng(5).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < TOL
    assert rel(op.loadVector(), -oop.apply(np.zeros(space.size))) < TOL
    # without constraints (natural boundary conditions)
    kw2 = dict(kw, strong_dirichlet=False)
    op2 = fem.operator.galerkin(space, **kw2)
    op2(u, w)
    assert rel(w, ol.UnstructuredOperator(coords, elems, order, **kw2).apply(u)) < TOL


@pytest.mark.parametrize("dim,n,order", [(2, [8, 7], 2), (3, [4, 4, 3], 1), (3, [3, 3, 2], 2)])
def test_krylov_solvers_on_a_distorted_mesh(dim, n, order):
    """Poisson + reaction with strong Dirichlet data: CG and Jacobi-PCG against a dense solve of the ORACLE's operator"""
    lo, hi = [0.0] * dim, [1.0] * dim
    coords, elems = distorted(n, lo, hi, seed=11 * dim + order, amplitude=0.15)
    kw = dict(eps=1.0, c=0.5, data=2, strong_dirichlet=True)
    space = fem.space.lagrange(fem.unstructuredGrid(coords, elems), order=order)
    op = fem.operator.galerkin(space, dirichlet_mask=1, **kw)
    oop = ol.UnstructuredOperator(coords, elems, order, **kw)
    N = space.size
    A = np.column_stack([oop.apply(np.eye(N)[:, j], linear=True) for j in range(N)])
    b = -oop.apply(np.zeros(N))
    x_ref = np.linalg.solve(A, b)
    assert rel(op.diagonal(), np.diag(A)) < 1e-12
    mask, vals = op.dirichlet()
    for cls in (fem.solver.CgInverseOperator, fem.solver.JacobiCgInverseOperator):
        inv = cls({"tolerance": 1e-12, "maxiterations": 2000})
        inv.bind(op)
        x = np.where(mask == 1, vals, 0.0)          # the constrained dofs start at their values (DirichletConstraints::operator()(u)):
        inv(op.loadVector(), x)                     # their residual is zero and stays zero, CG sees the symmetric interior block
        assert inv.iterations > 0
        assert rel(x, x_ref) < 1e-9
    # the discrete solution approximates g = prod sin(pi x_k) (it solves the PDE with these data)
    xs, _ = oop.nodes()
    assert np.abs(x_ref - np.prod(np.sin(np.pi * xs), axis=1)).max() < (0.2 if order == 1 else 0.05)


def test_unstructured_mesh_errors():
    coords, elems = ol.cartesian_as_unstructured([2, 2], [0.0, 0.0], [1.0, 1.0])
    grid = fem.unstructuredGrid(coords, elems)
    with pytest.raises(_capi.B200FemError):                    # DG spaces need face connectivity
        fem.space.dglegendre(grid, order=1)
    with pytest.raises(_capi.B200FemError):                    # order 3
        fem.space.lagrange(grid, order=3)
    with pytest.raises(_capi.B200FemError):                    # weak boundary terms
        fem.operator.galerkin(fem.space.lagrange(grid, order=1), boundary=True)
    with pytest.raises(_capi.B200FemError):                    # the Kronecker form needs a Cartesian mesh
        op = fem.operator.galerkin(fem.space.lagrange(grid, order=1), kernel=_capi.KERNEL_KRONECKER)
        op(np.zeros(9), np.zeros(9))
    bad = elems.copy()
    bad[0] = bad[0][[1, 0, 3, 2]]                              # mirrored element: negative Jacobian determinant
    with pytest.raises(_capi.B200FemError):
        fem.unstructuredGrid(coords, bad)
    with pytest.raises(_capi.B200FemError):
        fem.unstructuredGrid(coords, elems + 100)


@pytest.mark.parametrize("dim,n,order", [(2, [9, 6], 2), (3, [4, 4, 3], 1), (3, [4, 3, 3], 2)])
def test_compiled_integrands_on_a_distorted_mesh(dim, n, order):
    """run-time compiled interior() (variable coefficients, cubic reaction) on an unstructured mesh against the oracle integrating
    the same source through the same multilinear geometry"""
    import os
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "integrands", "adr_variable.cuh")).read()
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    coords, elems = distorted(n, lo, hi, seed=17 * dim + order)
    const = [0.05, 1.0, -0.5, 0.25, 0.0, 0.3, 0.7]
    space = fem.space.lagrange(fem.unstructuredGrid(coords, elems), order=order)
    op = fem.operator.galerkinJit(space, src, const, skeleton=False, boundary=False)
    u = np.random.default_rng(9).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    op(u, w)
    oop = ol.UnstructuredOperator(coords, elems, order, user_source=src, constants=const)
    assert rel(w, oop.apply(u)) < TOL
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u) - oop.apply(np.zeros(space.size))) < 1e-11
