#!/usr/bin/env python
"""Golden vectors produced by RUNNING the reference's own solver code (oracle/_ref: dune/fem/solver/linear/{cg,bicgstab,gmres}.hh,
solver/cginverseoperator.hh, operator/common/automaticdifferenceoperator.hh, solver/newtoninverseoperator.hh compiled from
/root/reference, oracle/ref_bind.cpp) on the oracle's operators through callbacks.  Run in the build container only:
    make -C oracle ref && python tests/golden/make_golden_solvers.py
Writes tests/golden/reference_solvers.json; tests/test_oracle_solver_goldens.py checks the oracle's restatements against it where
neither oracle/_ref nor the reference tree exists (the live comparison is tests/test_reference_pieces.py).  The problems are built by
tests/solver_cases.py, shared with that test."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reference_lib as rl  # noqa: E402
import solver_cases as sc  # noqa: E402

out = {"cg": {}, "pcg": {}, "legacy_cg": {}, "bicgstab": {}, "gmres": {}, "difference_quotient": {}, "newton": {}}
sp, op, b = sc.poisson()
A = lambda u: op.apply(u, linear=True)  # noqa: E731
x0 = np.zeros(sp.size)
for crit in (0, 1, 2):
    it, x, h = rl.cg(A, b, x0, 1e-9, 60, crit)
    out["cg"][str(crit)] = {"iterations": it, "history": h.tolist(), "x": x.tolist()}
d = op.diagonal()
it, x, h = rl.cg(A, b, x0, 1e-10, 80, 0, precon=lambda r: r / d)
out["pcg"]["0"] = {"iterations": it, "history": h.tolist(), "x": x.tolist()}
for measure in (0, 1):
    it, x = rl.legacy_cg(A, b, x0, 1e-9, 400, measure)
    out["legacy_cg"][str(measure)] = {"iterations": it, "x": x.tolist()}
sp, op, b = sc.advdiff(1, eps=1.0)
A = lambda u: op.apply(u, linear=True)  # noqa: E731
x0 = np.zeros(sp.size)
for crit in (0, 1, 2):
    it, x, h = rl.bicgstab(A, b, x0, 1e-12, 12, crit)
    out["bicgstab"][str(crit)] = {"iterations": it, "history": h.tolist(), "x": x.tolist()}
sp, op, b = sc.advdiff(1)
A = lambda u: op.apply(u, linear=True)  # noqa: E731
for crit in (0, 1, 2):
    it, x, h = rl.gmres(A, b, x0, 1e-7, 600, crit, 5)
    out["gmres"][str(crit)] = {"iterations": it, "history": h.tolist(), "x": x.tolist()}
sp, op, u, args = sc.cubic_lagrange()
for eps in (0.0, 1e-6):
    out["difference_quotient"][repr(eps)] = rl.difference_quotient(lambda v: op.apply(v), u, args, eps=eps).tolist()
for name, (gamma, amp, c, seed, line_search, maxit) in sc.NEWTON_CASES.items():
    sp, op = sc.reaction_diffusion(gamma, c)
    w0 = amp * np.random.default_rng(seed).uniform(-1, 1, sp.size)
    it, lit, fail, delta, w = rl.newton(lambda v: op.apply(v), w0, sc.newton_keys(1e-7, maxit, 1e-8, 20000, 48, line_search))
    out["newton"][name] = {"iterations": it, "linear_iterations": lit, "failure": fail, "residual": delta, "w": w.tolist()}
json.dump(out, open(os.path.join(HERE, "reference_solvers.json"), "w"))
print("written:", {k: len(v) for k, v in out.items()}, os.path.getsize(os.path.join(HERE, "reference_solvers.json")), "bytes")
