#!/usr/bin/env python
"""Golden vectors produced by RUNNING pieces of the reference itself (oracle/_ref: the parts of DUNE-FEM that compile from their own
source files under /root/reference, oracle/ref_bind.cpp) on fixed inputs.  Run in the build container only:
    make -C oracle ref && python tests/golden/make_golden_ref.py
Writes tests/golden/reference_pieces.json:
  lagrange_points   GenericLagrangePoint of the cube: per (dim, order) the local coordinates, (codim, subEntity, dofNumber) of every node
  lagrange_basis    GenericLagrangeBaseFunction: values and reference gradients of all basis functions at fixed points
  legendre_sets     LegendreShapeFunctionSet (plain / hierarchical): values and gradients of all functions, in the set's order
  cube_quadrature   CubeQuadrature: points, weights, order of the selected rule
  onb               OrthonormalBase_{2,3}D (the `dgonb` cube bases): values and gradients of the first functions at the fixed points
tests/test_oracle_tables.py checks the oracle against this file (works where neither oracle/_ref nor the reference tree exists)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reference_lib as rl  # noqa: E402

POINTS = {2: [[0.3, 0.55], [0.81, 0.12]], 3: [[0.3, 0.55, 0.7], [0.81, 0.12, 0.43]]}
out = {"points": POINTS, "lagrange_points": {}, "lagrange_basis": {}, "legendre_sets": {}, "cube_quadrature": {}, "onb": {}}
for dim in (2, 3):
    for order in (1, 2, 3):
        x, codim, sub, num = rl.lagrange_cube_points(dim, order)
        out["lagrange_points"][f"{dim},{order}"] = {"x": x.tolist(), "codim": codim.tolist(), "sub": sub.tolist(), "dof": num.tolist()}
        vals = []
        for xp in POINTS[dim]:
            pv = [rl.lagrange_cube_evaluate(dim, order, b, xp) for b in range((order + 1) ** dim)]
            vals.append({"phi": [float(v) for v, _ in pv], "dphi": [d.tolist() for _, d in pv]})
        out["lagrange_basis"][f"{dim},{order}"] = vals
    for order in (1, 2, 3, 4):
        for hier in (0, 1):
            vals = []
            for xp in POINTS[dim]:
                phi, dphi = rl.legendre_set(dim, order, hier, xp)
                vals.append({"phi": phi.tolist(), "dphi": dphi.tolist()})
            out["legendre_sets"][f"{dim},{order},{hier}"] = vals
for dim in (2, 3):
    order = 4                                                       # P_k bases are nested: the first functions of P_4 are P_1 .. P_3
    nb = (order + 1) * (order + 2) // 2 if dim == 2 else (order + 1) * (order + 2) * (order + 3) // 6
    vals = []
    for xp in POINTS[dim]:
        x = np.zeros(3)
        x[:dim] = xp
        pv = [rl.onb_cube(dim, i, x, grad=True) for i in range(nb)]
        vals.append({"phi": [float(v) for v, _ in pv], "dphi": [g.tolist() for _, g in pv]})
    out["onb"][f"{dim},{order}"] = vals
for dim in (1, 2, 3):
    for order in (0, 1, 3, 4, 5, 7, 10):
        x, w, exact = rl.cube_quadrature(dim, order)
        out["cube_quadrature"][f"{dim},{order}"] = {"x": x.tolist(), "w": w.tolist(), "exact": exact}
json.dump(out, open(os.path.join(HERE, "reference_pieces.json"), "w"))
print("written:", {k: len(v) for k, v in out.items()})
