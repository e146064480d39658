#!/usr/bin/env python
"""How far do the device CG iterates drift from the oracle's (sequential summation) over 50 iterations?  (contract: 1e-12 rel)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import dune_fem_b200 as fem
import oracle_lib as ol
for dim, order, n in ((2, 1, [256, 256]), (3, 2, [24, 24, 24])):
    lo, hi = [0.0] * dim, [1.0] * dim
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=order)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, order)
    kw = dict(eps=1.0, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
    op = fem.operator.galerkin(space, **kw); oop = ol.Operator(osp, threads=8, **kw)
    mask, g = op.dirichlet(); x0 = np.zeros(space.size)
    b = np.random.default_rng(5).uniform(-1, 1, space.size) * (1 - mask)      # rough right-hand side: 50 genuine CG iterations
    inv = fem.solver.CgInverseOperator({"tolerance": 1e-30, "maxiterations": 50}); inv.bind(op)
    x = x0.copy(); it = inv(b, x)
    it_ref, x_ref, h_ref = oop.cg(b, x0, 1e-30, 50)
    dh = np.abs(inv.residuals - h_ref) / h_ref
    print(dim, order, n, "it", it, it_ref, "hist max rel dev", dh.max(), "first10", dh[:10].max(), "x rel", np.abs(x - x_ref).max() / np.abs(x_ref).max(), flush=True)
