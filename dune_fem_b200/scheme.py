"""Schemes.  Mirrors dune.fem.scheme.galerkin / molGalerkin (python/dune/fem/scheme/_schemes.py:617-637) and the C++ class behind them,
Dune::Fem::FemScheme (dune/fem/schemes/femscheme.hh:60-260): a Galerkin operator together with the (non-linear) inverse operator,
`scheme(u, w)` = the operator, `scheme.solve(target=uh [, rhs=f])` = set the constraints on uh, then solve L[uh] = rhs
(femscheme.hh:194-221, 248-254) and return the solver info."""
import numpy as np

from . import _capi as capi
from . import operator as _operator
from . import solver as _solver

_KRYLOV = {"cg": _solver.CgInverseOperator, "bicgstab": _solver.BicgstabInverseOperator, "gmres": _solver.GmresInverseOperator}


def _strip(parameters):
    """accepts "newton.linear.tolerance" / "nonlinear.linear.tolerance" / "fem.solver.newton.linear.tolerance" alike"""
    out = {}
    for k, v in (parameters or {}).items():
        k = k.replace("fem.solver.", "")
        for pre in ("newton.", "nonlinear."):
            if k.startswith(pre):
                k = k[len(pre):]
        out[k] = v
    return out


class GalerkinScheme:
    def __init__(self, space, solver=None, parameters=None, integrands=None, constants=(), mol=False, **model):
        """integrands: CUDA C++ source of interior / skeleton / boundary (run-time compiled, operator.galerkinJit); otherwise the
        keyword arguments describe the built-in advection-diffusion-reaction family (operator.galerkin)"""
        self.space = space
        if integrands is not None:
            self.operator = _operator.galerkinJit(space, integrands, constants, skeleton=model.pop("skeleton", space.kind != capi.LAGRANGE),
                                                  boundary=model.pop("boundary", space.kind != capi.LAGRANGE))
            self._linear = False                       # (compiled forms are not known to be linear: Newton)
        else:
            self.operator = _operator.galerkin(space, **model)
            self._linear = not self.operator.nonlinear
        if mol:
            self.operator.setInverseMass(True)
        self.model = self.operator.model
        p = _strip(parameters)
        # default linear solver as _schemes.py: cg for symmetric problems is the caller's choice; gmres is always safe
        self._method = solver if isinstance(solver, str) else p.get("linear.method", "gmres")
        if self._method not in _KRYLOV:
            raise ValueError(f"solver must be one of {sorted(_KRYLOV)}")
        self.parameters = p
        self._mask = self._vals = None
        if integrands is None and model.get("strong_dirichlet"):
            self._mask, self._vals = self.operator.dirichlet()

    # --- FemScheme::operator() (femscheme.hh:181-190) ---
    def __call__(self, u, w):
        self.operator(u, w)

    def setQuadratureOrders(self, interior, surface):
        self.operator.setQuadratureOrders(interior, surface)

    # --- constraints (femscheme.hh:145-178 -> dirichletconstraints.hh:237-262) ---
    def setConstraints(self, u):
        if self._mask is not None:
            u[self._mask == 1] = self._vals[self._mask == 1]

    def solve(self, target, rhs=None):
        """FemScheme::solve(rhs, solution): solution = g on the Dirichlet boundary (+ rhs there), then invOp(rhs, solution).
        Returns the reference's info dictionary: converged, iterations (Newton), linear_iterations."""
        p = self.parameters
        self.setConstraints(target)
        if rhs is not None and self._mask is not None:
            target[self._mask == 1] += rhs[self._mask == 1]
        if self._linear:
            # a linear operator: the one Newton step with the exact Jacobian, A x = b + rhs (newtoninverseoperator.hh:761, 791-792)
            inv = _KRYLOV[self._method]({"tolerance": p.get("linear.tolerance", 1e-8), "errormeasure": p.get("linear.errormeasure", "absolute"),
                                         "maxiterations": p.get("linear.maxiterations", 1000), "gmres.restart": p.get("linear.gmres.restart", 20),
                                         "verbose": p.get("linear.verbose", False)})
            inv.bind(self.operator)
            b = self.operator.loadVector()
            if rhs is not None:
                b = b + rhs
                if self._mask is not None:                 # constrained rows: w_d = u_d - g_d - rhs_d = 0  <=>  u_d = g_d + rhs_d
                    b[self._mask == 1] = target[self._mask == 1]
            it = inv(np.ascontiguousarray(b), target)
            return {"converged": it >= 0, "iterations": 1, "linear_iterations": abs(it)}
        newton = _solver.NewtonInverseOperator({"tolerance": p.get("tolerance", 1e-6), "maxiterations": p.get("maxiterations", 2 ** 31 - 1),
                                                "linesearch.method": p.get("linesearch", p.get("linesearch.method", "none")), "verbose": p.get("verbose", False),
                                                "linear.method": self._method, "linear.tolerance": p.get("linear.tolerance", 1e-8),
                                                "linear.errormeasure": p.get("linear.errormeasure", "absolute"),
                                                "linear.maxiterations": p.get("linear.maxiterations", 1000), "linear.gmres.restart": p.get("linear.gmres.restart", 20)})
        newton.bind(self.operator)
        newton(rhs, target)
        return {"converged": newton.converged, "iterations": newton.iterations, "linear_iterations": newton.linearIterations}


def galerkin(space, solver=None, parameters=None, **kwargs):
    """dune.fem.scheme.galerkin(integrands, space, solver, parameters)"""
    return GalerkinScheme(space, solver=solver, parameters=parameters, **kwargs)


def molGalerkin(space, solver=None, parameters=None, **kwargs):
    """dune.fem.scheme.molGalerkin: MethodOfLinesScheme (w = M^-1 L[u])"""
    return GalerkinScheme(space, solver=solver, parameters=parameters, mol=True, **kwargs)
