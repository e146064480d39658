"""Lagrange order 3 in the oracle (several nodes inside an edge / face / cell: space/mapper/indexsetdofmapper.hh:414-427 with the
generic Lagrange point set, space/lagrange/genericlagrangepoints.hh:862-876): the dof map is a bijection onto 0..N-1 laid out by
geometry type, shared nodes of neighbouring elements coincide geometrically, and the discretisation reproduces cubics exactly."""
import numpy as np
import pytest

import oracle_lib as ol


@pytest.mark.parametrize("dim,n", [(2, [4, 3]), (3, [3, 2, 2])])
def test_order3_dofmap_and_cubic_exactness(dim, n):
    lo, hi = [0.0] * dim, [1.0, 0.5, 2.0][:dim]
    sp = ol.Space(n, lo, hi, ol.LAGRANGE, 3)
    assert sp.size == int(np.prod([3 * k + 1 for k in n])) and sp.local_size == 4 ** dim
    seen = set()
    for e in range(sp.elements):
        g = sp.dofmap(e)
        assert len(set(g.tolist())) == sp.local_size
        seen.update(g.tolist())
    assert seen == set(range(sp.size))
    # blocks by geometry type: the vertex dofs come first and are numbered lexicographically
    nv = int(np.prod([k + 1 for k in n]))
    x = sp.node_positions()                  # (asserts implicitly that shared nodes get ONE position: later elements overwrite equal values)
    h = (np.array(hi) - np.array(lo)) / np.array(n)
    on_vertex = np.all(np.abs((x - lo) / h - np.round((x - lo) / h)) < 1e-12, axis=1)
    assert on_vertex[:nv].all() and not on_vertex[nv:].any()
    # every element sees its nodes where the shape functions put them
    mi = sp.multiindex()[:, :dim]
    for e in range(sp.elements):
        ec, r = [], e
        for d in range(dim):
            ec.append(r % n[d])
            r //= n[d]
        assert np.allclose(x[sp.dofmap(e)], np.array(lo) + h * (np.array(ec) + mi / 3.0), atol=1e-14)
    # u cubic => A u = M (-lap u) on interior nodes (Galerkin orthogonality with u in the space, -lap u in the space)
    u = 0.3 + x @ np.array([1.0, -2.0, 0.5][:dim]) + (x ** 3).sum(axis=1)
    f = -6.0 * x.sum(axis=1)
    w = ol.Operator(sp, eps=1.0).apply(u)
    mf = ol.Operator(sp, eps=0.0, c=1.0).apply(f)
    interior = np.all((x > np.array(lo) + 1e-12) & (x < np.array(hi) - 1e-12), axis=1)
    assert np.abs(w - mf)[interior].max() < 1e-13
