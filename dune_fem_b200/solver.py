"""Krylov solvers.  Mirrors Dune::Fem::CgInverseOperator / KrylovInverseOperator<DF, cg>
(dune/fem/solver/krylovinverseoperators.hh:46-281): bind(op), __call__(rhs, x), iterations(); parameters follow
fem.solver.{tolerance, errormeasure, maxiterations, verbose} (dune/fem/solver/parameter.hh:21-295)."""
import ctypes as C

import numpy as np

from . import _capi as capi

_ERRORMEASURE = {"absolute": capi.TOL_ABSOLUTE, "relative": capi.TOL_RELATIVE, "residualreduction": capi.TOL_RESIDUAL_REDUCTION}


class CgInverseOperator:
    def __init__(self, parameters=None):
        p = {"tolerance": 1e-8, "errormeasure": "absolute", "maxiterations": 1000, "verbose": False}
        for k, v in (parameters or {}).items():
            p[k.replace("fem.solver.", "")] = v
        self.parameters = p
        self._op = None
        self._iterations = 0
        self.residuals = np.zeros(0)

    def bind(self, op):
        self._op = op

    def unbind(self):
        self._op = None

    def __call__(self, rhs, x):
        if self._op is None:
            raise RuntimeError("CgInverseOperator: no operator bound")      # DUNE_THROW(InvalidStateException) analogue
        p = self.parameters
        it = C.c_int()
        hist = np.zeros(max(int(p["maxiterations"]), 1))
        capi.check(capi.lib().b200fem_cg_solve(self._op.handle, capi.ptr(rhs), capi.ptr(x), float(p["tolerance"]),
                                               int(p["maxiterations"]), _ERRORMEASURE[p["errormeasure"]], C.byref(it),
                                               capi.ptr(hist)))
        self._iterations = it.value
        self.residuals = hist[:abs(it.value)]
        if p["verbose"]:
            for i, r in enumerate(self.residuals):
                print(f"Fem::CG it: {i} : residual {r}")                   # solver/linear/cg.hh:110-113
        return it.value

    @property
    def iterations(self):
        return self._iterations

    @property
    def converged(self):
        return self._iterations >= 0


KrylovInverseOperator = CgInverseOperator
