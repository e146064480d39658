"""Oracle vs the numeric tables the reference pins (tests/golden/*.json, made by make_golden.py)."""
import json
import os

import numpy as np

import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_gauss_points_match_reference_tables():
    table = json.load(open(os.path.join(GOLD, "gauss_points.json")))
    for m, rule in table.items():
        m = int(m)
        x, w = ol.quadrature(1, rule["order"])
        assert len(w) == m
        np.testing.assert_allclose(x[:, 0], rule["x"], rtol=0, atol=2.3e-16)
        np.testing.assert_allclose(w, rule["w"], rtol=0, atol=2.3e-16)


def test_rule_selection_smallest_rule_covering_order():
    # femquadratures_inline.hh:59-70: first m with order(m) = 2m-1 >= requested; order <= 0 -> 1
    for order, m in [(-1, 1), (0, 1), (1, 1), (2, 2), (3, 2), (4, 3), (5, 3), (6, 4), (7, 4), (10, 6), (11, 6), (19, 10)]:
        assert len(ol.quadrature(1, order)[1]) == m


def test_tensor_rule_first_coordinate_fastest():
    x, w = ol.quadrature(3, 4)
    x1, w1 = ol.quadrature(1, 4)
    m = len(w1)
    assert len(w) == m ** 3
    for i in range(m ** 3):
        d = (i % m, (i // m) % m, i // (m * m))
        np.testing.assert_allclose(x[i], [x1[d[0], 0], x1[d[1], 0], x1[d[2], 0]])
        np.testing.assert_allclose(w[i], w1[d[0]] * w1[d[1]] * w1[d[2]])
    np.testing.assert_allclose(w.sum(), 1.0, rtol=1e-15)


def test_legendre_factors_match_reference_table():
    tab = json.load(open(os.path.join(GOLD, "legendre_table.json")))
    fac = np.array(tab["factor"])
    wgt = np.array(tab["weight"])
    # evaluate at a few points by Horner with the reference table, compare with the oracle's generated table
    # Horner sums cancel heavily; compare relative to the coefficient magnitudes (gcc fuses a*x+b)
    for num in range(11):
        scale = 1e-15 * wgt[num] * np.abs(fac[num]).sum() * max(num, 1)
        for x in [0.0, 0.1127016653792583, 0.5, 0.77, 1.0]:
            phi = fac[num][num]
            for i in range(num - 1, -1, -1):
                phi = phi * x + fac[num][i]
            ref = wgt[num] * phi
            got = ol.lib().fo_legendre(num, x, 0)
            assert abs(got - ref) <= scale, (num, x, got, ref)
            dphi = 0.0
            if num >= 1:
                dphi = fac[num][num] * num
                for i in range(num - 1, 0, -1):
                    dphi = dphi * x + fac[num][i] * i
            refd = wgt[num] * dphi
            gotd = ol.lib().fo_legendre(num, x, 1)
            assert abs(gotd - refd) <= scale * max(num, 1), (num, x, gotd, refd)


def test_legendre_orthonormal_on_unit_interval_up_to_order_9():
    # order 10 is excluded: the reference table has a typo there (see fem_oracle.cpp)
    x, w = ol.quadrature(1, 19)
    P = np.array([[ol.lib().fo_legendre(n, xi, 0) for xi in x[:, 0]] for n in range(10)])
    M = (P * w) @ P.T
    # monomial Horner evaluation (as in the reference) loses digits at high order
    np.testing.assert_allclose(M[:6, :6], np.eye(6), atol=1e-13)
    np.testing.assert_allclose(M, np.eye(10), atol=1e-9)


def test_legendre_multiindex_orderings():
    sp = ol.Space([2, 2, 2], [0, 0, 0], [1, 1, 1], ol.DG_LEGENDRE, 2)
    mi = sp.multiindex()
    # last coordinate fastest (legendre.hh:169-194)
    assert mi[0].tolist() == [0, 0, 0] and mi[1].tolist() == [0, 0, 1] and mi[3].tolist() == [0, 1, 0] and mi[9].tolist() == [1, 0, 0]
    hp = ol.Space([2, 2, 2], [0, 0, 0], [1, 1, 1], ol.DG_LEGENDRE_HIER, 2)
    mh = hp.multiindex()
    orders = mh.max(axis=1)
    assert (np.diff(orders) >= 0).all()           # sorted by max order (legendre.hh:236-250)
    assert mh[0].tolist() == [0, 0, 0]
    assert mh[1:8].tolist() == [[0, 0, 1], [0, 1, 0], [0, 1, 1], [1, 0, 0], [1, 0, 1], [1, 1, 0], [1, 1, 1]]
    assert sorted(map(tuple, mh.tolist())) == sorted(map(tuple, mi.tolist()))


def test_lagrange_shape_functions_are_nodal():
    for dim, order in [(2, 1), (2, 2), (3, 2)]:
        sp = ol.Space([1] * dim, [0] * dim, [1] * dim, ol.LAGRANGE, order)
        mi = sp.multiindex()
        for l in range(sp.local_size):
            phi, dphi = sp.shape(mi[l, :dim] / order)
            e = np.zeros(sp.local_size)
            e[l] = 1
            np.testing.assert_allclose(phi, e, atol=1e-14)
        # partition of unity and zero gradient sum
        phi, dphi = sp.shape(np.array([0.3, 0.6, 0.2])[:dim])
        assert abs(phi.sum() - 1) < 1e-14 and np.abs(dphi.sum(axis=0)).max() < 1e-13


# ---- golden vectors produced by running the compiled reference pieces (tests/golden/make_golden_ref.py -> reference_pieces.json):
# the same pins as tests/test_reference_pieces.py, available where neither oracle/_ref nor the reference tree exists ----
def _ref_golden():
    return json.load(open(os.path.join(GOLD, "reference_pieces.json")))


def test_golden_cube_quadratures():
    g = _ref_golden()
    for key, q in g["cube_quadrature"].items():
        dim, order = map(int, key.split(","))
        x, w = ol.quadrature(dim, order)
        assert len(w) == len(q["w"]) and q["exact"] >= order
        assert np.abs(x[:, :dim] - np.array(q["x"])).max() == 0.0 and np.abs(w - np.array(q["w"])).max() <= 4e-16


def test_golden_legendre_shape_function_sets():
    g = _ref_golden()
    for key, vals in g["legendre_sets"].items():
        dim, order, hier = map(int, key.split(","))
        sp = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.DG_LEGENDRE_HIER if hier else ol.DG_LEGENDRE, order)
        for xp, v in zip(g["points"][str(dim)], vals):
            phi, dphi = sp.shape(xp)
            assert np.abs(phi - np.array(v["phi"])).max() < 1e-12 * max(1.0, np.abs(v["phi"]).max())
            assert np.abs(dphi[:, :dim] - np.array(v["dphi"])).max() < 1e-11 * max(1.0, np.abs(v["dphi"]).max())


def test_golden_lagrange_points_basis_and_numbering():
    g = _ref_golden()
    for key, pts in g["lagrange_points"].items():
        dim, order = map(int, key.split(","))
        sp = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.LAGRANGE, order)
        assert np.abs(np.array(pts["x"]) - sp.multiindex()[:, :dim] / order).max() < 1e-15
        for xp, v in zip(g["points"][str(dim)], g["lagrange_basis"][key]):
            phi, dphi = sp.shape(xp)
            assert np.abs(phi - np.array(v["phi"])).max() < 1e-13 and np.abs(dphi[:, :dim] - np.array(v["dphi"])).max() < 1e-12
        if order == 2:          # one-element mesh: first-touch dof = offset[entity dimension] + reference sub-entity number
            codim, sub = np.array(pts["codim"]), np.array(pts["sub"])
            spa = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.LAGRANGE, 2, numbering=ol.NUMBERING_ADAPTIVE_LEAF)
            offset = np.concatenate([[0], np.cumsum([int((codim == dim - p).sum()) for p in range(dim + 1)])])
            assert (spa.dofmap(0) == offset[dim - codim] + sub).all()
        if order == 3:          # several dofs inside an entity: numbered in the order of the reference's dofNumber, contiguously
            codim, sub, num = np.array(pts["codim"]), np.array(pts["sub"]), np.array(pts["dof"])
            gl = sp.dofmap(0)
            for c in range(dim + 1):
                for s_ in np.unique(sub[codim == c]):
                    sel = np.where((codim == c) & (sub == s_))[0]
                    d = gl[sel][np.argsort(num[sel])]
                    assert (np.diff(d) == 1).all()


def test_golden_dgonb_functions():
    g = _ref_golden()
    for key, vals in g["onb"].items():
        dim, kmax = map(int, key.split(","))
        for order in range(1, kmax + 1):        # P_k bases are nested: the first functions of the P_4 vectors
            sp = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.DG_ONB, order)
            for xp, v in zip(g["points"][str(dim)], vals):
                phi, dphi = sp.shape(xp)
                nb = sp.local_size
                assert np.abs(phi - np.array(v["phi"][:nb])).max() < 1e-12 * max(1.0, np.abs(v["phi"][:nb]).max())
                assert np.abs(dphi[:, :dim] - np.array(v["dphi"][:nb])[:, :dim]).max() < 1e-11 * max(1.0, np.abs(v["dphi"][:nb]).max())
