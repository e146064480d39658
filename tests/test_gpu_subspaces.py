"""GPU parity of the DG spaces that are SUB-BASES of the 3-D tensor Legendre basis the device kernels work on: Q_k Legendre on
2-D meshes (the setting of pydemo/advectiondiffusion.py:113) and `dgonb` P_k in 2-D and 3-D (the space that demo imports,
:9).  Everything through the C ABI against the CPU oracle; tolerance 1e-12 of max|w|."""
import math

import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def make(kind, dim, order, n, lo, hi):
    g = fem.structuredGrid(lo, hi, n)
    if kind == "onb":
        return fem.space.dgonb(g, order=order), ol.Space(n, lo, hi, ol.DG_ONB, order)
    hier = kind == "hier"
    return fem.space.dglegendre(g, order=order, hierarchical=hier), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER if hier else ol.DG_LEGENDRE, order)


CASES = [("lex", 2, 1), ("hier", 2, 2), ("lex", 2, 3), ("hier", 2, 4), ("hier", 2, 5),
         ("onb", 2, 1), ("onb", 2, 2), ("onb", 2, 3), ("onb", 2, 4),
         ("onb", 3, 1), ("onb", 3, 2), ("onb", 3, 3), ("onb", 3, 4)]


@pytest.mark.parametrize("kind,dim,order", CASES)
def test_subspace_apply_affine_linear_and_load_vector(kind, dim, order):
    n = [7, 5] if dim == 2 else ([5, 4, 3] if order <= 2 else [3, 3, 2])
    lo, hi = [-1.0] * dim, ([1.0, 0.5] if dim == 2 else [1.0, 0.5, 2.0])
    space, osp = make(kind, dim, order, n, lo, hi)
    assert space.size == osp.size and space.localBlockSize == osp.local_size
    b = (1.0, -0.5) if dim == 2 else (1.0, -0.5, 0.25)
    kw = dict(eps=0.05, b=b, c=0.3, beta=20.0 * order ** 2, dirichlet_mask=0b0111 if dim == 2 else 0b010011, data=1)
    u = np.random.default_rng(10 * order + dim).uniform(-1, 1, space.size)
    oop = ol.Operator(osp, skeleton=True, boundary=True, **kw)
    op = fem.operator.galerkin(space, **kw)
    w = np.empty(space.size)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    assert op.timing()["kernel"] == _capi.KERNEL_QUADRATURE
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < TOL
    assert rel(op.loadVector(), -oop.apply(np.zeros(space.size))) < TOL


@pytest.mark.parametrize("kind,dim,order", [("hier", 2, 2), ("onb", 2, 2), ("onb", 3, 2)])
def test_subspace_nonlinear_model_and_inverse_mass(kind, dim, order):
    n = [6, 5] if dim == 2 else [4, 3, 3]
    lo, hi = [-1.0] * dim, [1.0] * dim
    space, osp = make(kind, dim, order, n, lo, hi)
    kw = dict(eps=0.1, b=(1.0, 0.0, 0.0)[:dim], gamma=0.7, beta=80.0, dirichlet_mask=0b0011, data=1)
    u = np.random.default_rng(3).uniform(-1, 1, space.size)
    oop = ol.Operator(osp, skeleton=True, boundary=True, **kw)
    op = fem.operator.galerkin(space, **kw)
    w = np.empty(space.size)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    # MOLGalerkinOperator: the bases are orthonormal, the inverse mass is the scalar 1 / volume (molgalerkin.hh:100-197)
    mop = fem.operator.molGalerkin(space, **kw)
    oop.setInverseMass(True)
    mop(u, w)
    assert rel(w, oop.apply(u)) < TOL


def test_kronecker_kernel_is_refused_for_subspaces():
    space, _ = make("onb", 3, 2, [4, 4, 4], [-1.0] * 3, [1.0] * 3)
    op = fem.operator.galerkin(space, eps=0.1, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1, kernel=_capi.KERNEL_KRONECKER)
    with pytest.raises(Exception):
        op(np.zeros(space.size), np.empty(space.size))


@pytest.mark.parametrize("kind", ["onb", "hier"])
def test_pydemo_advection_diffusion_2d_eoc_on_the_device(kind):
    """pydemo/advectiondiffusion.py:93-147 as written: 2-D, order 2, SIPG + upwind with weak Dirichlet data on x0 = +-1, solved with
    BiCGStab on the device; the L2 error against sin(x0 x1) must converge with EOC >= order + 1 - 0.1 (:121-145) and agree
    with the oracle's solution of the same system."""
    order, errs = 2, []
    for n in (4, 8, 16):
        space, osp = make(kind, 2, order, [n, n], [-1.0, -1.0], [1.0, 1.0])
        kw = dict(eps=0.1, b=(1.0, 0.0), beta=20.0 * order ** 2, dirichlet_mask=0b0011, data=1)
        op = fem.operator.galerkin(space, **kw)
        inv = fem.solver.BicgstabInverseOperator({"tolerance": 1e-12, "maxiterations": 5000})
        inv.bind(op)
        x = np.zeros(space.size)
        inv(op.loadVector(), x)
        assert inv.iterations > 0
        errs.append(osp.l2error(x, 1))
        oop = ol.Operator(osp, skeleton=True, boundary=True, **kw)
        assert rel(op.loadVector(), -oop.apply(np.zeros(space.size))) < TOL
        r = oop.apply(x)                               # L[x] = A x - b = 0 at the solution
        assert np.abs(r).max() < 1e-9 * np.abs(op.loadVector()).max()
    eoc = [math.log(errs[i + 1] / errs[i]) / math.log(0.5) for i in range(2)]
    assert eoc[-1] > order + 1 - 0.1, (errs, eoc)


@pytest.mark.parametrize("kind,dim,order,periodic", [("hier", 3, 2, (1, 0, 1)), ("lex", 3, 1, (1, 1, 1)), ("onb", 2, 2, (0, 1)), ("hier", 2, 3, (1, 1)), ("hier", 3, 3, (0, 1, 0))])
def test_periodic_grids(kind, dim, order, periodic):
    """Periodic sides are skeleton faces whose neighbour is the element on the far side (galerkin.hh:859-861); incl. the case of two
    cells along an axis, where both faces of an element meet the same neighbour."""
    n = ([5, 2] if dim == 2 else [4, 2, 3])
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    g = fem.structuredGrid(lo, hi, n, periodic=periodic)
    if kind == "onb":
        space, osp = fem.space.dgonb(g, order=order), ol.Space(n, lo, hi, ol.DG_ONB, order)
    else:
        space, osp = fem.space.dglegendre(g, order=order, hierarchical=kind == "hier"), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER if kind == "hier" else ol.DG_LEGENDRE, order)
    osp.set_periodic(g.periodic)
    kw = dict(eps=0.05, b=(1.0, -0.5, 0.25)[:dim], c=0.3, beta=20.0 * order ** 2, dirichlet_mask=0b11 if not periodic[0] else 0, data=1)
    u = np.random.default_rng(order).uniform(-1, 1, space.size)
    oop = ol.Operator(osp, skeleton=True, boundary=True, **kw)
    ref = oop.apply(u)
    osp_np = ol.Space(n, lo, hi, osp.kind, order)
    assert rel(ol.Operator(osp_np, skeleton=True, boundary=True, **kw).apply(u), ref) > 1e-3     # the wrap-around matters
    op = fem.operator.galerkin(space, **kw)
    w = np.empty(space.size)
    op(u, w)
    assert rel(w, ref) < TOL
    assert op.timing()["kernel"] == _capi.KERNEL_QUADRATURE       # the Kronecker kernels do not wrap: AUTO steps aside
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < TOL


def test_periodic_with_compiled_x_dependent_integrands():
    """on a periodic face the integrand sees the INSIDE element's point (the two sides differ by the domain length)"""
    import os
    src = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "integrands", "adr_variable.cuh")).read()
    n, lo, hi = [4, 3, 2], [-1.0] * 3, [1.0, 0.5, 2.0]
    g = fem.structuredGrid(lo, hi, n, periodic=(1, 1, 0))
    space, osp = fem.space.dglegendre(g, order=2), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    osp.set_periodic(g.periodic)
    const = [0.05, 1.0, -0.5, 0.25, 80.0, 0.3, 0.7]
    u = np.random.default_rng(1).uniform(-1, 1, space.size)
    w = np.empty(space.size)
    fem.operator.galerkinJit(space, src, const)(u, w)
    assert rel(w, ol.UserOperator(osp, src, const).apply(u)) < TOL
