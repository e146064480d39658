"""The oracle's solver restatements against golden vectors produced by RUNNING the reference's own solver code
(tests/golden/reference_solvers.json, made by tests/golden/make_golden_solvers.py from oracle/_ref in the build container):
LinearSolver::cg / bicgstab / gmres, the legacy ConjugateGradientSolver, AutomaticDifferenceOperator and NewtonInverseOperator.
Same comparisons as the live ones in tests/test_reference_pieces.py, available where neither oracle/_ref nor /root/reference exists."""
import json
import os

import numpy as np
import pytest

import oracle_lib as ol
import solver_cases as sc

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_solvers.json")))


@pytest.mark.parametrize("crit", [0, 1, 2])
def test_cg_golden(crit):
    sp, op, b = sc.poisson()
    g = GOLD["cg"][str(crit)]
    it, x, h = op.cg(b, np.zeros(sp.size), 1e-9, 60, crit)
    assert it == g["iterations"] and len(h) == abs(it)
    np.testing.assert_array_equal(h, g["history"])           # same operations in the same order as cg.hh: bit-identical
    np.testing.assert_array_equal(x, g["x"])


def test_preconditioned_and_legacy_cg_golden():
    sp, op, b = sc.poisson()
    g = GOLD["pcg"]["0"]
    it, x, h = op.pcg(op.diagonal(), b, np.zeros(sp.size), 1e-10, 80, 0)
    assert it == g["iterations"]
    np.testing.assert_allclose(h, g["history"], rtol=1e-12)
    np.testing.assert_allclose(x, g["x"], rtol=1e-11, atol=1e-14)
    for measure in (0, 1):
        g = GOLD["legacy_cg"][str(measure)]
        it, x, _ = op.cg(b, np.zeros(sp.size), 1e-9, 400, measure)
        assert it == g["iterations"]
        np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-13 * np.abs(g["x"]).max())


@pytest.mark.parametrize("crit", [0, 1, 2])
def test_bicgstab_golden(crit):
    sp, op, b = sc.advdiff(1, eps=1.0)
    g = GOLD["bicgstab"][str(crit)]
    it, x, h = op.bicgstab(b, np.zeros(sp.size), 1e-12, 12, crit)
    assert it == g["iterations"] == -12 and len(g["history"]) == 11     # the reference logs `res` only for iterations that continue
    np.testing.assert_allclose(h[:11], g["history"], rtol=1e-9)
    np.testing.assert_allclose(h[:4], g["history"][:4], rtol=1e-13)
    np.testing.assert_allclose(x, g["x"], rtol=0, atol=1e-9 * np.abs(g["x"]).max())


@pytest.mark.parametrize("crit", [0, 1, 2])
def test_gmres_golden(crit):
    sp, op, b = sc.advdiff(1)
    g = GOLD["gmres"][str(crit)]
    it, x, h = op.gmres(b, np.zeros(sp.size), 1e-7, 600, crit, 5)
    assert it == g["iterations"] and it > 0 and len(h) == len(g["history"])
    np.testing.assert_allclose(h, g["history"], rtol=1e-9)
    np.testing.assert_allclose(h[:10], g["history"][:10], rtol=1e-13)
    np.testing.assert_allclose(x, g["x"], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("eps", [0.0, 1e-6])
def test_difference_quotient_golden(eps):
    sp, op, u, args = sc.cubic_lagrange()
    op.linearize(u, eps)
    for a, r in zip(args, GOLD["difference_quotient"][repr(eps)]):
        np.testing.assert_array_equal(op.applyJacobian(a)[0], r)


@pytest.mark.parametrize("name", sorted(sc.NEWTON_CASES))
def test_newton_golden(name):
    gamma, amp, c, seed, line_search, maxit = sc.NEWTON_CASES[name]
    g = GOLD["newton"][name]
    sp, op = sc.reaction_diffusion(gamma, c)
    w0 = amp * np.random.default_rng(seed).uniform(-1, 1, sp.size)
    it, lit, fail, delta, w = ol.newton(op, w0, 1e-7, maxit, 1e-8, 20000, 48, 2, line_search)
    assert (it, fail) == (g["iterations"], g["failure"])
    assert fail == {"plain": 0, "line_search": 0, "too_many_iterations": 5, "linear_solver_failed": 7}[name]
    # the reference GMRES reports -(maxIterations + 1) where oracle and device report -maxIterations
    assert lit == g["linear_iterations"] or (fail == 7 and lit == g["linear_iterations"] + 1)
    np.testing.assert_allclose(delta, g["residual"], rtol=1e-12)
    np.testing.assert_allclose(w, g["w"], rtol=0, atol=1e-12 * np.abs(g["w"]).max())
