"""Host logic of the Python mirror that needs no device: parameter keys of the inverse operators as the reference reads them
(solver/parameter.hh:96-186, solver/newtoninverseoperator.hh:163-170, 206, 234, 285 -- the key names are checked against the
reference's own classes in tests/test_reference_pieces.py) and the errors the mirror raises before any C call."""
import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
from dune_fem_b200.scheme import _strip


def test_newton_parameters_accept_the_reference_keys():
    ref = {"fem.solver.nonlinear.tolerance": 1e-5, "fem.solver.nonlinear.maxiterations": 7, "fem.solver.nonlinear.linesearch": "simple",
           "fem.solver.linear.method": "bicgstab", "fem.solver.linear.tolerance": 1e-9, "fem.solver.linear.errormeasure": "residualreduction",
           "fem.solver.linear.maxiterations": 123, "fem.solver.linear.gmres.restart": 11}
    p = fem.solver.NewtonInverseOperator(ref).parameters
    assert (p["tolerance"], p["maxiterations"], p["linesearch.method"]) == (1e-5, 7, "simple")
    assert (p["linear.method"], p["linear.tolerance"], p["linear.errormeasure"], p["linear.maxiterations"], p["linear.gmres.restart"]) == \
        ("bicgstab", 1e-9, "residualreduction", 123, 11)
    # the nested spelling ("nonlinear.linear.*") and the short one give the same table
    nested = {k.replace("fem.solver.linear.", "fem.solver.nonlinear.linear."): v for k, v in ref.items()}
    short = {k.replace("fem.solver.nonlinear.", "").replace("fem.solver.", ""): v for k, v in ref.items()}
    assert fem.solver.NewtonInverseOperator(nested).parameters == p == fem.solver.NewtonInverseOperator(short).parameters
    # the reference's defaults (newtoninverseoperator.hh:206, 285; solver/parameter.hh:118, 104)
    d = fem.solver.NewtonInverseOperator().parameters
    assert d["tolerance"] == 1e-6 and d["linesearch.method"] == "none" and d["linear.tolerance"] == 1e-8 and d["linear.errormeasure"] == "absolute"
    assert d["linear.gmres.restart"] == 20 and d["maxiterations"] == 2 ** 31 - 1


def test_scheme_parameter_prefixes():
    assert _strip({"fem.solver.newton.linear.tolerance": 1e-9, "nonlinear.tolerance": 1e-7, "newton.linesearch": "simple", "verbose": True}) == \
        {"linear.tolerance": 1e-9, "tolerance": 1e-7, "linesearch": "simple", "verbose": True}


def test_unbound_inverse_operators_refuse_to_run():
    x = np.zeros(4)
    with pytest.raises(RuntimeError):
        fem.solver.NewtonInverseOperator()(None, x)
    for cls in (fem.solver.CgInverseOperator, fem.solver.BicgstabInverseOperator, fem.solver.GmresInverseOperator):
        with pytest.raises(RuntimeError):
            cls()(x, x)


def test_raw_pointer_helper_rejects_reinterpretation():
    # the C ABI reads raw memory: float32 / integer / non-contiguous arrays must not be passed on silently
    _capi.ptr(np.zeros(3))
    for bad in (np.zeros(3, dtype=np.float32), np.zeros(3, dtype=np.int64), np.zeros((4, 4))[:, 1], [0.0, 1.0]):
        with pytest.raises(TypeError):
            _capi.ptr(bad)
    _capi.ptr(np.zeros(3, dtype=np.uint8), np.uint8)
