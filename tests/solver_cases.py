"""The small problems the solver pins run on (tests/golden/make_golden_solvers.py generates the reference's outputs for them,
tests/test_oracle_solver_goldens.py checks the oracle against those).  Everything is seeded; right-hand sides come from numpy's PCG64."""
import numpy as np

import oracle_lib as ol


def poisson(dim=3, order=2, n=(4, 4, 3)):
    sp = ol.Space(list(n), [0.0] * dim, [1.0] * dim, ol.LAGRANGE, order)
    op = ol.Operator(sp, eps=1.0, c=0.5, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
    mask, _ = op.dirichlet()
    return sp, op, np.random.default_rng(11).uniform(-1, 1, sp.size) * (1 - mask)


def advdiff(order, n=(4, 3, 3), eps=1e-2):
    sp = ol.Space(list(n), [-1.0] * 3, [1.0] * 3, ol.DG_LEGENDRE_HIER, order)
    op = ol.Operator(sp, eps=eps, b=(1.0, 0.3, 0.0), beta=20.0 * order * order, dirichlet_mask=0b000011, data=1, skeleton=True, boundary=True)
    return sp, op, -op.apply(np.zeros(sp.size)) + np.random.default_rng(12).uniform(-1, 1, sp.size)


def cubic_lagrange():
    sp = ol.Space([5, 4, 3], [0.0] * 3, [1.0] * 3, ol.LAGRANGE, 2)
    op = ol.Operator(sp, eps=1.0, c=0.5, gamma=2.0, data=2, dirichlet_mask=0b111111, strong_dirichlet=True)
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, sp.size)
    args = np.stack([rng.uniform(-1, 1, sp.size), 1e-9 * rng.uniform(-1, 1, sp.size), 1e4 * rng.uniform(-1, 1, sp.size), np.zeros(sp.size)])
    return sp, op, u, args


def reaction_diffusion(gamma, c):
    sp = ol.Space([3, 3, 2], [-1.0] * 3, [1.0] * 3, ol.DG_LEGENDRE_HIER, 1)
    return sp, ol.Operator(sp, skeleton=True, boundary=True, eps=0.5, b=(1.0, 0.0, 0.0), c=c, gamma=gamma, beta=40.0, dirichlet_mask=0b000011, data=1)


# name -> gamma, amplitude of the initial guess, c, seed, line search, maxiterations
NEWTON_CASES = {"plain": (10.0, 2.0, -8.0, 4, False, 40), "line_search": (10.0, 2.0, -8.0, 4, True, 40),
                "too_many_iterations": (5.0, 4.0, -8.0, 4, True, 3), "linear_solver_failed": (10.0, 2.0, -12.0, 3, True, 40)}


def newton_keys(tol, maxit, lin_tol, lin_maxit, restart, line_search, errormeasure="residualreduction"):
    """the keys the reference reads (newtoninverseoperator.hh:163-170, 206, 234, 285; solver/parameter.hh:96-186)"""
    return {"fem.solver.nonlinear.tolerance": tol, "fem.solver.nonlinear.maxiterations": maxit, "fem.solver.nonlinear.linesearch": "simple" if line_search else "none",
            "fem.solver.linear.method": "gmres", "fem.solver.linear.tolerance": lin_tol, "fem.solver.linear.errormeasure": errormeasure,
            "fem.solver.linear.maxiterations": lin_maxit, "fem.solver.linear.gmres.restart": restart}
