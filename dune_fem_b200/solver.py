"""Krylov solvers.  Mirrors Dune::Fem::CgInverseOperator / KrylovInverseOperator<DF, cg>
(dune/fem/solver/krylovinverseoperators.hh:46-281): bind(op), __call__(rhs, x), iterations(); parameters follow
fem.solver.{tolerance, errormeasure, maxiterations, verbose} (dune/fem/solver/parameter.hh:21-295)."""
import ctypes as C

import numpy as np

from . import _capi as capi

_ERRORMEASURE = {"absolute": capi.TOL_ABSOLUTE, "relative": capi.TOL_RELATIVE, "residualreduction": capi.TOL_RESIDUAL_REDUCTION}


class CgInverseOperator:
    _solve = "b200fem_cg_solve"          # LinearSolver::cg (solver/linear/cg.hh)
    _label = "Fem::CG it: {} : residual {}"

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
        capi.check(getattr(capi.lib(), self._solve)(self._op.handle, capi.ptr(rhs), capi.ptr(x), float(p["tolerance"]),
                                                    int(p["maxiterations"]), _ERRORMEASURE[p["errormeasure"]], C.byref(it),
                                                    capi.ptr(hist)))
        self._iterations = it.value
        self.residuals = hist[:abs(it.value)]
        if p["verbose"]:
            for i, r in enumerate(self.residuals):
                print(self._label.format(i, r))                            # solver/linear/cg.hh:110-113, bicgstab.hh:196
        return it.value

    @property
    def iterations(self):
        return self._iterations

    @property
    def converged(self):
        return self._iterations >= 0


class JacobiCgInverseOperator(CgInverseOperator):
    """CG with "fem.solver.preconditioning.method: jacobi": the preconditioned branch of LinearSolver::cg
    (solver/linear/cg.hh:52-56, 72-107) with B = diag(A)^-1 built matrix-free (the reference's DiagonalPreconditioner needs an
    assembled operator, solver/diagonalpreconditioner.hh:37-57)."""
    _solve = "b200fem_pcg_solve"


class BicgstabInverseOperator(CgInverseOperator):
    """KrylovInverseOperator< DF, SolverParameter::bicgstab > (solver/krylovinverseoperators.hh:288 ->
    solver/linear/bicgstab.hh:64-214): for non-symmetric operators (advection-diffusion)."""
    _solve = "b200fem_bicgstab_solve"
    _label = "Fem::BiCGstab it: {} : {}"


class GmresInverseOperator(CgInverseOperator):
    """KrylovInverseOperator< DF, SolverParameter::gmres > (solver/krylovinverseoperators.hh:295 ->
    solver/linear/gmres.hh:117-301): restarted GMRES, "gmres.restart" (default 20, solver/parameter.hh:197-201)."""
    _label = "Fem::GMRES it: {} : {}"

    def __call__(self, rhs, x):
        if self._op is None:
            raise RuntimeError("GmresInverseOperator: no operator bound")
        p = self.parameters
        it = C.c_int()
        hist = np.zeros(max(int(p["maxiterations"]), 1))
        capi.check(capi.lib().b200fem_gmres_solve(self._op.handle, capi.ptr(rhs), capi.ptr(x), int(p.get("gmres.restart", 20)),
                                                  float(p["tolerance"]), int(p["maxiterations"]), _ERRORMEASURE[p["errormeasure"]],
                                                  C.byref(it), capi.ptr(hist)))
        self._iterations = it.value
        self.residuals = hist[:abs(it.value)]
        if p["verbose"]:
            for i, r in enumerate(self.residuals):
                print(self._label.format(i, r))
        return it.value


def KrylovInverseOperator(parameters=None):
    """fem.solver.method selects the Krylov method (solver/parameter.hh; krylovinverseoperators.hh:83,126-131)."""
    method = {k.replace("fem.solver.", ""): v for k, v in (parameters or {}).items()}.get("method", "cg")
    precon = {k.replace("fem.solver.", ""): v for k, v in (parameters or {}).items()}.get("preconditioning.method", "none")
    if method == "cg":
        return JacobiCgInverseOperator(parameters) if precon == "jacobi" else CgInverseOperator(parameters)
    if method == "bicgstab":
        return BicgstabInverseOperator(parameters)
    if method == "gmres":
        return GmresInverseOperator(parameters)
    raise NotImplementedError(f"KrylovInverseOperator: method {method!r} (cg, bicgstab and gmres are available)")


class NewtonInverseOperator:
    """Dune::Fem::NewtonInverseOperator (solver/newtoninverseoperator.hh:423-803): bind(op); __call__(u, w) solves L[w] = u from the
    initial guess in w (u = None: L[w] = 0).  Parameters as fem.solver.nonlinear.*: tolerance (1e-6), maxiterations, linesearch
    ("none" | "simple"; "linesearch.method" is accepted too), and for the linear solves fem.solver.linear.* (the prefix the reference derives,
    :163-165; "nonlinear.linear.*" is accepted too): method, tolerance (1e-8), errormeasure ("absolute"), maxiterations, gmres.restart (20).
    Deviation: the total linear-iteration budget defaults to 1000 here, to `int` max in the reference (solver/parameter.hh:150).  The Jacobian is the difference
    quotient of AutomaticDifferenceLinearOperator; the whole iteration runs on the device (b200fem_newton_solve)."""
    _METHODS = {"cg": 0, "bicgstab": 1, "gmres": 2}
    FAILURES = {0: "Success", 1: "InvalidResidual", 4: "LineSearchFailed", 5: "TooManyIterations", 6: "TooManyLinearIterations", 7: "LinearSolverFailed"}

    def __init__(self, parameters=None):
        p = {"tolerance": 1e-6, "maxiterations": 2 ** 31 - 1, "linesearch.method": "none", "verbose": False,
             "linear.method": "gmres", "linear.tolerance": 1e-8, "linear.errormeasure": "absolute", "linear.maxiterations": 1000, "linear.gmres.restart": 20}
        for k, v in (parameters or {}).items():
            k = k.replace("fem.solver.", "").replace("nonlinear.", "")
            if k in ("linesearch", "lineSearch"):       # the key the reference reads (newtoninverseoperator.hh:279-285): values "none" | "simple"
                k = "linesearch.method"
            p[k] = v
        self.parameters = p
        self._op = None
        self.iterations = self.linearIterations = 0
        self.residual = float("nan")
        self.failure = 0

    def bind(self, op):
        self._op = op

    def unbind(self):
        self._op = None

    def __call__(self, u, w):
        if self._op is None:
            raise RuntimeError("NewtonInverseOperator: no operator bound")
        p = self.parameters
        it, lit, fail, res = C.c_int(), C.c_int(), C.c_int(), C.c_double()
        capi.check(capi.lib().b200fem_newton_solve(self._op.handle, None if u is None else capi.ptr(u), capi.ptr(w), float(p["tolerance"]), int(min(p["maxiterations"], 2 ** 31 - 1)),
                                                   self._METHODS[p["linear.method"]], float(p["linear.tolerance"]), int(p["linear.maxiterations"]),
                                                   _ERRORMEASURE[p["linear.errormeasure"]], int(p["linear.gmres.restart"]), int(p["linesearch.method"] == "simple"),
                                                   C.byref(it), C.byref(lit), C.byref(res), C.byref(fail)))
        self.iterations, self.linearIterations, self.residual, self.failure = it.value, lit.value, res.value, fail.value
        if p["verbose"]:
            print(f"Newton iterations: {it.value}, linear iterations: {lit.value}, |residual| = {res.value} ({self.FAILURES.get(fail.value, fail.value)})")
        return it.value

    @property
    def converged(self):
        return self.failure == 0
