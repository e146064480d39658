"""dune.fem.scheme.galerkin mirrored on the device (dune_fem_b200/scheme.py -> FemScheme::solve, schemes/femscheme.hh:194-254): the
constraints are set on the target, then the (non-linear) inverse operator solves L[uh] = rhs; results against the oracle."""
import os

import numpy as np
import pytest

import dune_fem_b200 as fem
import oracle_lib as ol

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_linear_scheme_with_dirichlet_constraints():
    """the reference's solver acceptance problem (solver/test/inverseoperatortest.cc): P2 Poisson + reaction, data prod sin(pi x)"""
    n, lo, hi = [6, 6, 5], [0.0] * 3, [1.0] * 3
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=2)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, 2)
    kw = dict(eps=1.0, c=0.5, data=2, dirichlet_mask=0b111111, strong_dirichlet=True)
    scheme = fem.scheme.galerkin(space, solver="cg", parameters={"newton.linear.tolerance": 1e-13, "newton.linear.maxiterations": 2000}, **kw)
    uh = np.full(space.size, 0.3)                       # any initial guess: solve() sets the constraints first
    info = scheme.solve(target=uh)
    assert info["converged"] and info["iterations"] == 1 and info["linear_iterations"] > 5
    oop = ol.Operator(osp, **kw)
    assert np.abs(oop.apply(uh)).max() < 1e-9           # L[uh] = 0, constrained rows included (uh_d = g_d)
    assert osp.l2error(uh, 2) < 2e-3
    # right-hand side: L[uh] = f
    f = np.random.default_rng(0).uniform(-1, 1, space.size) * 1e-2
    uh2 = np.zeros(space.size)
    scheme.solve(target=uh2, rhs=f)
    assert np.abs(oop.apply(uh2) - f).max() < 1e-9
    # the scheme is the operator
    w = np.empty(space.size)
    scheme(uh, w)
    assert np.abs(w - oop.apply(uh)).max() < 1e-12


def test_nonlinear_scheme_runs_newton():
    n, lo, hi = [4, 4, 3], [-1.0] * 3, [1.0] * 3
    space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=2)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    kw = dict(eps=0.5, b=(1.0, 0.0, 0.0), c=1.0, gamma=2.0, beta=80.0, dirichlet_mask=0b000011, data=1)
    scheme = fem.scheme.galerkin(space, solver="gmres", parameters={"newton.tolerance": 1e-7, "newton.linear.tolerance": 1e-7, "newton.linear.errormeasure": "residualreduction",
                                                                    "newton.linear.maxiterations": 4000, "newton.linear.gmres.restart": 30}, **kw)
    uh = np.zeros(space.size)
    info = scheme.solve(target=uh)
    assert info["converged"] and 2 <= info["iterations"] <= 12 and info["linear_iterations"] > info["iterations"]
    r = ol.Operator(osp, skeleton=True, boundary=True, **kw).apply(uh)
    assert np.linalg.norm(r) < 2e-7


def test_scheme_over_compiled_integrands():
    src = open(os.path.join(HERE, "integrands", "adr_variable.cuh")).read()
    const = [0.5, 1.0, -0.5, 0.25, 80.0, 1.0, 2.0]
    n, lo, hi = [6, 5], [-1.0] * 2, [1.0, 0.5]
    space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=2)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    scheme = fem.scheme.galerkin(space, solver="gmres", integrands=src, constants=const,
                                 parameters={"nonlinear.tolerance": 1e-7, "nonlinear.linear.tolerance": 1e-7, "nonlinear.linear.errormeasure": "residualreduction",
                                             "nonlinear.linear.maxiterations": 4000, "nonlinear.linear.gmres.restart": 40})
    uh = np.zeros(space.size)
    info = scheme.solve(target=uh)
    assert info["converged"]
    assert np.linalg.norm(ol.UserOperator(osp, src, const).apply(uh)) < 2e-7


def test_mol_scheme_is_the_inverse_mass_operator():
    n, lo, hi = [4, 3, 3], [-1.0] * 3, [1.0] * 3
    space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=1)
    kw = dict(eps=0.1, b=(1.0, 0.0, 0.0), beta=20.0, dirichlet_mask=0b000011, data=1)
    mol, plain = fem.scheme.molGalerkin(space, **kw), fem.scheme.galerkin(space, **kw)
    u = np.random.default_rng(2).uniform(-1, 1, space.size)
    a, b = np.empty(space.size), np.empty(space.size)
    mol(u, a)
    plain(u, b)
    vol = np.prod([(hi[d] - lo[d]) / n[d] for d in range(3)])
    assert np.abs(a * vol - b).max() < 1e-12 * np.abs(b).max()
