"""The oracle's restatement for unstructured cube meshes (fem_oracle.cpp: UnstructuredLagrange), pinned on what can be pinned
without a GPU: on a Cartesian mesh handed over as vertex / element arrays it reproduces the structured oracle bit for bit in the
numbering (AdaptiveLeafIndexSet first-touch order) and to rounding in the values; on distorted meshes it passes the patch test
(a size-independent property of the multilinear geometry + Gauss quadrature) and is invariant under element renumbering."""
import numpy as np
import pytest

import oracle_lib as ol


def distorted(n, lo, hi, seed, amplitude=0.2, shuffle=True):
    """Cartesian vertices displaced by up to `amplitude` cell widths (boundary vertices slide along the boundary only in the
    patch tests' sense: they move too -- the domain changes, the mesh stays conforming), elements in random order"""
    coords, elems = ol.cartesian_as_unstructured(n, lo, hi)
    rng = np.random.default_rng(seed)
    h = (np.array(hi) - np.array(lo)) / np.array(n)
    coords = coords + rng.uniform(-amplitude, amplitude, coords.shape) * h
    if shuffle:
        elems = elems[rng.permutation(len(elems))]
    return coords, elems


@pytest.mark.parametrize("dim,n", [(2, [5, 4]), (3, [4, 3, 2])])
@pytest.mark.parametrize("order", [1, 2])
def test_cartesian_mesh_as_unstructured_reproduces_the_structured_oracle(dim, n, order):
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    coords, elems = ol.cartesian_as_unstructured(n, lo, hi)
    kw = dict(eps=0.7, b=(1.0, -0.5, 0.25)[:dim], c=0.3, gamma=0.5, data=2, strong_dirichlet=True)
    uo = ol.UnstructuredOperator(coords, elems, order, **kw)
    sp = ol.Space(n, lo, hi, ol.LAGRANGE, order, numbering=ol.NUMBERING_ADAPTIVE_LEAF)
    so = ol.Operator(sp, dirichlet_mask=(1 << (2 * dim)) - 1, **kw)
    assert uo.size == sp.size
    for e in range(sp.elements):
        assert (uo.dofmap(e) == sp.dofmap(e)).all()                      # numbering: bit-exact
    u = np.random.default_rng(1).uniform(-1, 1, sp.size)
    ref = so.apply(u)
    assert np.abs(uo.apply(u) - ref).max() < 1e-14 * np.abs(ref).max()
    ref = so.apply(u, linear=True)
    assert np.abs(uo.apply(u, linear=True) - ref).max() < 1e-14 * np.abs(ref).max()
    x, bnd = uo.nodes()
    assert np.allclose(x, sp.node_positions(), atol=1e-14)
    on = np.zeros(sp.size, dtype=bool)
    for d in range(dim):
        on |= np.isclose(x[:, d], lo[d]) | np.isclose(x[:, d], hi[d])
    assert (bnd.astype(bool) == on).all()


@pytest.mark.parametrize("dim,n", [(2, [6, 5]), (3, [4, 3, 3])])
@pytest.mark.parametrize("order", [1, 2])
def test_patch_test_and_renumbering_invariance_on_distorted_meshes(dim, n, order):
    lo, hi = [0.0] * dim, [1.0] * dim
    coords, elems = distorted(n, lo, hi, seed=dim * 10 + order)
    op = ol.UnstructuredOperator(coords, elems, order, eps=1.0)
    x, bnd = op.nodes()
    u = 0.3 + x @ np.array([1.0, -2.0, 0.5][:dim])                       # a linear function: in the space on any multilinear mesh
    w = op.apply(u)
    # Laplace of a linear function: zero residual at every interior node (grad phi_i detJ is polynomial -> exact quadrature)
    scale = np.abs(w[bnd == 1]).max()
    assert scale > 1e-3 and np.abs(w[bnd == 0]).max() < 1e-13 * max(scale, 1.0)
    # mass: sum_i (M 1)_i = area of the (distorted) domain, whatever the element order
    m = ol.UnstructuredOperator(coords, elems, order, eps=0.0, c=1.0)
    vol = m.apply(np.ones(m.size)).sum()
    rng = np.random.default_rng(7)
    elems2 = elems[rng.permutation(len(elems))]
    m2 = ol.UnstructuredOperator(coords, elems2, order, eps=0.0, c=1.0)
    assert abs(m2.apply(np.ones(m2.size)).sum() - vol) < 1e-13 * vol
    # the operator itself is the same up to the permutation of the dofs induced by the new first-touch order
    op2 = ol.UnstructuredOperator(coords, elems2, order, eps=1.0)
    x2, _ = op2.nodes()
    key = lambda a: [tuple(np.round(r, 9)) for r in a]
    pos = {k: i for i, k in enumerate(key(x))}
    perm = np.array([pos[k] for k in key(x2)])                           # dof i of op2 is dof perm[i] of op
    v = rng.uniform(-1, 1, op.size)
    assert np.abs(op2.apply(v[perm]) - op.apply(v)[perm]).max() < 1e-12 * np.abs(op.apply(v)).max()


@pytest.mark.parametrize("dim,n", [(2, [7, 5]), (3, [4, 3, 3])])
@pytest.mark.parametrize("order", [1, 2])
def test_library_numbering_and_colouring_without_a_device(dim, n, order):
    """host logic of the product (b200fem_unstructured_numbering, no GPU): the dof map equals the oracle's bit for bit on a distorted,
    shuffled mesh; the boundary marks agree; the colouring is valid (no dof shared inside a colour) and uses few colours"""
    import ctypes as C
    from dune_fem_b200 import _capi
    lo, hi = [-1.0] * dim, [1.0, 0.5, 2.0][:dim]
    coords, elems = distorted(n, lo, hi, seed=5 * dim + order)
    coords, elems = np.ascontiguousarray(coords), np.ascontiguousarray(elems, dtype=np.int64)
    oop = ol.UnstructuredOperator(coords, elems, order)
    nb = (order + 1) ** dim
    size = C.c_int64()
    dofs = np.empty((len(elems), nb), dtype=np.int32)
    colour = np.empty(len(elems), dtype=np.int32)
    bnd = np.empty(oop.size, dtype=np.uint8)
    rc = _capi.lib().b200fem_unstructured_numbering(dim, len(coords), _capi.ptr(coords), len(elems), _capi.ptr(elems, np.int64), order, C.byref(size),
                                                   _capi.ptr(dofs, np.int32), _capi.ptr(colour, np.int32), _capi.ptr(bnd, np.uint8))
    assert rc == 0 and size.value == oop.size
    for e in range(len(elems)):
        assert (dofs[e] == oop.dofmap(e)).all()
    assert (bnd == oop.nodes()[1]).all()
    ncol = colour.max() + 1
    assert colour.min() == 0 and ncol <= 2 ** dim + 4               # a structured-like hexahedral mesh needs 2^dim colours; greedy stays close
    for c in range(ncol):
        d = dofs[colour == c].ravel()
        assert len(np.unique(d)) == len(d)
