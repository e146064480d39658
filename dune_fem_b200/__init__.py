"""dune_fem_b200 -- B200-native matrix-free Galerkin operator + CG behind DUNE-FEM's operator interface.

Host-side mirror of the reference's Python entry points for this path (python/dune/fem/...):
    grid.structuredGrid, space.lagrange / space.dglegendre, operator.galerkin, solver.CgInverseOperator.
Everything computes on the GPU through the C ABI in include/b200fem.h; there is no CPU fallback.
"""
from . import _capi  # noqa: F401
from .grid import structuredGrid, unstructuredGrid  # noqa: F401
from . import space, operator, solver, scheme  # noqa: F401

__all__ = ["structuredGrid", "unstructuredGrid", "space", "operator", "solver", "scheme"]
