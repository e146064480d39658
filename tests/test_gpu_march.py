"""GPU parity of the z-marching Kronecker DG kernel (dg_kronecker_march.cuh) through the C ABI.

Oracle comparisons on meshes the CPU restatement finishes in seconds; on a larger mesh the marching kernel is compared
with the plain tile kernel (dg_kronecker.cuh, B200FEM_KERNEL_KRONECKER_TILE), which is itself pinned to the oracle in test_gpu_parity.py.
Tolerance 1e-12 relative to max|w| (north_star).  Mesh sizes cover: partial tiles in x and y (tile = 16 x 16 elements),
a single plane, runs that cross column boundaries, one-element extents, both local dof orderings, and both the dense and
the checkerboard (no advection along y, z) self-matrix code paths.
"""
import numpy as np
import pytest

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
import oracle_lib as ol

pytestmark = pytest.mark.gpu
TOL = 1e-12
LO, HI = [-1, -1, -1], [1, 1.5, 1]


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


def run(space, variant, u, **kw):
    kernel = {"march": _capi.KERNEL_KRONECKER, "tile": _capi.KERNEL_KRONECKER_TILE}[variant]
    op = fem.operator.galerkin(space, beta=80.0, kernel=kernel, **kw)
    w, wl = np.full(space.size, np.nan), np.full(space.size, np.nan)
    op(u, w)
    op.applyLinear(u, wl)
    return w, wl


@pytest.mark.parametrize("bvel", [(1.0, -0.5, 0.25), (1.0, 0.0, 0.0)])
@pytest.mark.parametrize("hier", [False, True])
@pytest.mark.parametrize("n", [[8, 4, 4], [16, 16, 1], [18, 4, 9], [16, 16, 7], [32, 33, 5], [34, 20, 21], [2, 1, 2]])
def test_march_kernel_against_oracle(n, hier, bvel):
    space = fem.space.dglegendre(fem.structuredGrid(LO, HI, n), order=2, hierarchical=hier)
    kw = dict(eps=0.3, b=bvel, c=0.7, dirichlet_mask=0b011011, data=1)
    u = np.random.default_rng(7).uniform(-1, 1, space.size)
    w, wl = run(space, "march", u, **kw)
    osp = ol.Space(n, LO, HI, ol.DG_LEGENDRE_HIER if hier else ol.DG_LEGENDRE, 2)
    oop = ol.Operator(osp, beta=80.0, skeleton=True, boundary=True, **kw)
    assert rel(w, oop.apply(u)) < TOL
    assert rel(wl, oop.apply(u, linear=True)) < TOL


@pytest.mark.parametrize("bvel", [(1.0, -0.5, 0.25), (1.0, 0.0, 0.0)])
def test_march_kernel_against_tile_kernel_large(bvel):
    n = [48, 40, 56]
    space = fem.space.dglegendre(fem.structuredGrid(LO, HI, n), order=2, hierarchical=True)
    kw = dict(eps=1e-5, b=bvel, dirichlet_mask=0b000011, data=1)
    u = np.random.default_rng(11).uniform(-1, 1, space.size)
    wm, wlm = run(space, "march", u, **kw)
    wt, wlt = run(space, "tile", u, **kw)
    assert rel(wm, wt) < TOL and rel(wlm, wlt) < TOL
