"""GPU parity of the Lagrange Kronecker (sum-factorised lattice stencil) kernel, lagrange_kronecker.cuh, through the C ABI.

Small meshes against the CPU oracle (both dof numberings, 2-D and 3-D, P1 and P2, Poisson with strong Dirichlet data and an
advection-diffusion-reaction model with boundary terms); a larger mesh against the generic quadrature kernel (which is pinned
to the oracle in test_gpu_parity.py).  Mesh extents are chosen to give partial tiles (tile = 28 x 12 / 30 x 14 lattice nodes),
several z-segments and single-element axes.  Tolerance 1e-12 relative to max|w| (north_star).
"""
import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


MODELS = {
    "poisson_strong_dirichlet": lambda dim: dict(eps=1.0, c=0.25, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True),
    "adr_boundary_terms": lambda dim: dict(eps=0.7, b=(1.0, 0.5, -0.25), c=0.3, beta=20.0, data=1, dirichlet_mask=0b100111 if dim == 3 else 0b0111, boundary=True),
    "homogeneous": lambda dim: dict(eps=0.4, b=(-0.3, 0.2, 0.1), c=1.5),
}


@pytest.mark.parametrize("model", sorted(MODELS))
@pytest.mark.parametrize("dim,order,numbering,n", [
    (2, 1, 0, [7, 6]), (2, 2, 0, [17, 9]), (2, 2, 1, [5, 4]), (3, 1, 0, [7, 6, 5]), (3, 1, 0, [33, 1, 18]),
    (3, 2, 0, [7, 6, 5]), (3, 2, 1, [4, 5, 3]), (3, 2, 0, [15, 8, 9]), (3, 2, 0, [1, 1, 1])])
def test_lagrange_kronecker_against_oracle(dim, order, numbering, n, model):
    lo, hi = [0.0] * dim, [1.0, 2.0, 1.5][:dim]
    kw = MODELS[model](dim)
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=order, numbering=numbering)
    osp = ol.Space(n, lo, hi, ol.LAGRANGE, order, numbering=numbering)
    u = np.random.default_rng(3).uniform(-1, 1, space.size)
    oop = ol.Operator(osp, **kw)
    op = fem.operator.galerkin(space, kernel=_capi.KERNEL_KRONECKER, **kw)
    w = np.full(space.size, np.nan)
    op(u, w)
    assert rel(w, oop.apply(u)) < TOL
    assert op.timing()["kernel"] == _capi.KERNEL_KRONECKER
    op.applyLinear(u, w)
    assert rel(w, oop.apply(u, linear=True)) < TOL


@pytest.mark.parametrize("order,numbering,n", [(2, 0, [40, 37, 29]), (1, 0, [70, 45, 50]), (2, 1, [20, 21, 22])])
def test_lagrange_kronecker_against_quadrature_kernel_large(order, numbering, n):
    lo, hi = [0.0, 0.0, 0.0], [1.0, 2.0, 1.5]
    kw = dict(eps=1.0, b=(0.5, -1.0, 0.25), c=0.1, data=2, dirichlet_mask=0b111111, strong_dirichlet=True)
    space = fem.space.lagrange(fem.structuredGrid(lo, hi, n), order=order, numbering=numbering)
    u = np.random.default_rng(5).uniform(-1, 1, space.size)
    res = {}
    for kernel in (_capi.KERNEL_QUADRATURE, _capi.KERNEL_KRONECKER):
        op = fem.operator.galerkin(space, kernel=kernel, **kw)
        w, wl = np.full(space.size, np.nan), np.full(space.size, np.nan)
        op(u, w)
        op.applyLinear(u, wl)
        assert op.timing()["kernel"] == kernel
        res[kernel] = (w, wl)
    assert rel(res[2][0], res[1][0]) < TOL and rel(res[2][1], res[1][1]) < TOL
