"""Developer check (run on the GPU box): z-marching Kronecker kernel against the CPU oracle and the tile kernel."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


worst = 0.0
for n in ([8, 4, 4], [16, 16, 1], [18, 4, 9], [16, 16, 7], [32, 33, 5], [34, 20, 21], [2, 1, 2], [40, 40, 40]):
    for hier in (False, True):
        lo, hi = [-1, -1, -1], [1, 1.5, 1]
        space = fem.space.dglegendre(fem.structuredGrid(lo, hi, n), order=2, hierarchical=hier)
        kw = dict(eps=0.3, b=(1.0, -0.5, 0.25), c=0.7, dirichlet_mask=0b011011, data=1)
        u = np.random.default_rng(7).uniform(-1, 1, space.size)
        res = {}
        for variant in ("tensor", "march"):
            os.environ["B200FEM_KRON_VARIANT"] = variant
            op = fem.operator.galerkin(space, beta=80.0, kernel=_capi.KERNEL_KRONECKER, **kw)
            w = np.full(space.size, np.nan); wl = np.full(space.size, np.nan)
            op(u, w); op.applyLinear(u, wl)
            res[variant] = (w, wl)
        d = max(rel(res["march"][0], res["tensor"][0]), rel(res["march"][1], res["tensor"][1]))
        msg = f"n={n} hier={hier}: march vs tensor {d:.2e}"
        if np.prod(n) <= 16000:
            osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER if hier else ol.DG_LEGENDRE, 2)
            oop = ol.Operator(osp, beta=80.0, skeleton=True, boundary=True, **kw)
            do = max(rel(res["march"][0], oop.apply(u)), rel(res["march"][1], oop.apply(u, linear=True)))
            msg += f"  vs oracle {do:.2e}"; d = max(d, do)
        print(msg, flush=True)
        worst = max(worst, d)
print("worst", worst)
assert worst < 1e-12
