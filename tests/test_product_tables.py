"""The PRODUCT's own host tables (dune_fem_b200/csrc/tables.hpp -- Gauss rules, rule selection, 1-D Legendre / Lagrange bases, local
numbering of the DG spaces; the header the device tables are built from) against the vectors the compiled reference produced
(tests/golden/reference_pieces.json, made by tests/golden/make_golden_ref.py from oracle/_ref) and against the reference's Gauss table
(tests/golden/gauss_points.json).  CPU only, no oracle in between: tests/native/product_tables_bind.cpp is compiled with g++ here."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLD = os.path.join(HERE, "golden")
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


@pytest.fixture(scope="module")
def pt():
    src = os.path.join(HERE, "native", "product_tables_bind.cpp")
    out = os.path.join(ROOT, "build", "product_tables.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    deps = [src, os.path.join(ROOT, "dune_fem_b200", "csrc", "tables.hpp")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        tmp = f"{out}.{os.getpid()}"
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-ffp-contract=off", "-shared", "-o", tmp, src])
        os.replace(tmp, out)
    L = C.CDLL(out)
    L.pt_gauss_rule.argtypes = [C.c_int, _dp, _dp]
    L.pt_tabulate_1d.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp]
    L.pt_basis_1d.restype = C.c_double
    L.pt_basis_1d.argtypes = [C.c_int, C.c_int, C.c_int, C.c_double, C.c_int]
    L.pt_dg_tensor_map.argtypes = [C.c_int, C.c_int, C.c_int, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")]
    return L


@pytest.fixture(scope="module")
def gold():
    return json.load(open(os.path.join(GOLD, "reference_pieces.json")))


def _rule(pt, m):
    x, w = np.zeros(m), np.zeros(m)
    pt.pt_gauss_rule(m, x, w)
    return x, w


def test_gauss_rules_are_the_reference_table(pt):
    table = json.load(open(os.path.join(GOLD, "gauss_points.json")))
    for m, rule in table.items():
        x, w = _rule(pt, int(m))
        np.testing.assert_allclose(x, rule["x"], rtol=0, atol=2.3e-16)
        np.testing.assert_allclose(w, rule["w"], rtol=0, atol=2.3e-16)


def test_rule_selection_and_tensor_rules_are_the_reference_cube_quadratures(pt, gold):
    # CubeQuadrature (femquadratures_inline.hh:59-110): smallest rule with 2m-1 >= order, tensor product with x0 fastest -- the kernels
    # build exactly this from the 1-D rule (quadrature point q = (q2*m + q1)*m + q0)
    for key, q in gold["cube_quadrature"].items():
        dim, order = map(int, key.split(","))
        m = pt.pt_gauss_points_for_order(order)
        x1, w1 = _rule(pt, m)
        assert m ** dim == len(q["w"]) and 2 * m - 1 == q["exact"]
        idx = np.indices((m,) * dim).reshape(dim, -1)[::-1].T          # column d = index along axis d, axis 0 fastest
        np.testing.assert_array_equal(x1[idx], np.array(q["x"]).reshape(-1, dim))
        np.testing.assert_allclose(np.prod(w1[idx], axis=1), q["w"], rtol=0, atol=4e-16)
    assert pt.pt_gauss_points_for_order(20) == -1                      # beyond the reference's table (MAXP = 10)


@pytest.mark.parametrize("hier", [0, 1])
def test_dg_legendre_spaces_hold_the_reference_functions_in_the_reference_order(pt, gold, hier):
    # LegendreShapeFunctionSet< FunctionSpace, hierarchical > (shapefunctionset/legendre.hh): function l of the set at a point ==
    # product of the product's 1-D polynomials for the multi-index dg_tensor_map stores at l
    for key, vals in gold["legendre_sets"].items():
        dim, order, h = map(int, key.split(","))
        if h != hier:
            continue
        n = order + 1
        tmap = np.full(n ** 3, -2, dtype=np.int32)
        nb = pt.pt_dg_tensor_map(dim, order, 2 if hier else 1, tmap)
        assert nb == n ** dim == len(vals[0]["phi"]) and sorted(tmap[tmap >= 0]) == list(range(nb))
        for xp, v in zip(gold["points"][str(dim)], vals):
            phi, dphi = np.zeros(nb), np.zeros((nb, dim))
            for t in np.where(tmap >= 0)[0]:
                mi = [t // (n * n), (t // n) % n, t % n]
                assert all(mi[d] == 0 for d in range(dim, 3))          # 2-D spaces: constant along the third axis
                f = [pt.pt_basis_1d(1, order, mi[d], xp[d], 0) for d in range(dim)]
                g = [pt.pt_basis_1d(1, order, mi[d], xp[d], 1) for d in range(dim)]
                phi[tmap[t]] = np.prod(f)
                for d in range(dim):
                    dphi[tmap[t], d] = np.prod([g[e] if e == d else f[e] for e in range(dim)])
            assert np.abs(phi - np.array(v["phi"])).max() <= 1e-13 * max(1.0, np.abs(v["phi"]).max())
            assert np.abs(dphi - np.array(v["dphi"])).max() <= 1e-12 * max(1.0, np.abs(v["dphi"]).max())


def test_dgonb_spaces_hold_the_reference_functions_in_the_reference_order(pt, gold):
    # OrthonormalBase_{2,3}D (orthonormal/orthonormalbase_{2,3}d.hh): on cubes the graded P_k basis is a sub-basis of the tensor Legendre
    # basis; P_k bases are nested, so the P_4 vectors cover orders 1..4
    for key, vals in gold["onb"].items():
        dim, kmax = map(int, key.split(","))
        for order in range(1, kmax + 1):
            n = order + 1
            tmap = np.full(n ** 3, -2, dtype=np.int32)
            nb = pt.pt_dg_tensor_map(dim, order, 3, tmap)
            assert nb == ((order + 1) * (order + 2) // 2 if dim == 2 else (order + 1) * (order + 2) * (order + 3) // 6)
            for xp, v in zip(gold["points"][str(dim)], vals):
                for t in np.where(tmap >= 0)[0]:
                    mi = [t // (n * n), (t // n) % n, t % n]
                    assert sum(mi) <= order
                    f = [pt.pt_basis_1d(1, order, mi[d], xp[d], 0) for d in range(dim)]
                    g = [pt.pt_basis_1d(1, order, mi[d], xp[d], 1) for d in range(dim)]
                    ref_phi, ref_dphi = v["phi"][tmap[t]], np.array(v["dphi"][tmap[t]])[:dim]
                    assert abs(np.prod(f) - ref_phi) < 1e-12 * max(1.0, abs(ref_phi))      # the reference evaluates expanded monomials
                    for d in range(dim):
                        assert abs(np.prod([g[e] if e == d else f[e] for e in range(dim)]) - ref_dphi[d]) < 1e-11 * max(1.0, np.abs(ref_dphi).max())


def test_lagrange_1d_bases_give_the_reference_cube_basis(pt, gold):
    # GenericLagrangeBaseFunction of the cube (space/lagrange/genericbasefunctions.hh): local function b has the multi-index with
    # coordinate 0 fastest; values and reference gradients are products of the product's 1-D equidistant Lagrange polynomials
    for key, vals in gold["lagrange_basis"].items():
        dim, order = map(int, key.split(","))
        n = order + 1
        for xp, v in zip(gold["points"][str(dim)], vals):
            for b in range(n ** dim):
                a = [(b // n ** d) % n for d in range(dim)]
                f = [pt.pt_basis_1d(0, order, a[d], xp[d], 0) for d in range(dim)]
                g = [pt.pt_basis_1d(0, order, a[d], xp[d], 1) for d in range(dim)]
                assert abs(np.prod(f) - v["phi"][b]) < 1e-14
                for d in range(dim):
                    assert abs(np.prod([g[e] if e == d else f[e] for e in range(dim)]) - v["dphi"][b][d]) < 1e-13
        # nodal at the reference's Lagrange points
        pts = np.array(gold["lagrange_points"][key]["x"])
        for b in range(n ** dim):
            a = [(b // n ** d) % n for d in range(dim)]
            for c, xc in enumerate(pts):
                val = np.prod([pt.pt_basis_1d(0, order, a[d], xc[d], 0) for d in range(dim)])
                assert abs(val - (1.0 if c == b else 0.0)) < 1e-13


def test_tabulation_the_kernels_read_is_the_basis_at_the_rule_points(pt):
    for legendre, order in [(1, 1), (1, 2), (1, 3), (1, 5), (0, 1), (0, 2), (0, 3)]:
        for m in (order + 1, order + 2):
            n = order + 1
            B, G = np.zeros(m * n), np.zeros(m * n)
            pt.pt_tabulate_1d(legendre, order, m, B, G)
            x, _ = _rule(pt, m)
            for q in range(m):
                for i in range(n):
                    assert B[q * n + i] == pt.pt_basis_1d(legendre, order, i, x[q], 0)
                    assert G[q * n + i] == pt.pt_basis_1d(legendre, order, i, x[q], 1)


def _kron_apply(K, n1, cells, u):
    """w_K = sum_d [ (S_d + lo-boundary D_d + hi-boundary D_d) u_K + L_d u_{K-e_d} + R_d u_{K+e_d} ] along axis d of the element's
    n1 x n1 x n1 coefficient tensor -- the apply the Kronecker kernels perform (dg_kronecker.cuh:80-110), in numpy"""
    S, Dlo, Dhi, L, R = K
    nx, ny, nz = cells
    U = u.reshape(nz, ny, nx, n1, n1, n1)                 # element (x fastest), local index (i0*n1 + i1)*n1 + i2
    W = np.zeros_like(U)
    for d, (eaxis, laxis, ne) in enumerate([(2, 3, nx), (1, 4, ny), (0, 5, nz)]):
        def along(M, V):
            return np.moveaxis(np.tensordot(M, V, axes=([1], [laxis])), 0, laxis)
        W += along(S[d], U)
        lo = [slice(None)] * 6
        hi = [slice(None)] * 6
        lo[eaxis], hi[eaxis] = slice(0, 1), slice(ne - 1, ne)
        W[tuple(lo)] += along(Dlo[d], U[tuple(lo)])
        W[tuple(hi)] += along(Dhi[d], U[tuple(hi)])
        if ne > 1:
            inner, left, right = [slice(None)] * 6, [slice(None)] * 6, [slice(None)] * 6
            inner[eaxis], left[eaxis] = slice(1, ne), slice(0, ne - 1)
            W[tuple(inner)] += along(L[d], U[tuple(left)])           # coupling to the element below along axis d
            W[tuple(left)] += along(R[d], U[tuple(inner)])           # ... and above
    return W.reshape(-1)


@pytest.mark.parametrize("order,cells,mask", [(1, [3, 2, 2], 0b000011), (2, [3, 3, 2], 0b111111), (3, [2, 2, 3], 0b010010)])
def test_kronecker_operator_matrices_reproduce_the_dense_reference_loop(pt, order, cells, mask):
    """The 1-D matrices the product folds the integrands into (kron_tables.hpp) applied in Kronecker form == the oracle's restatement
    of the reference's dense quadrature loop (galerkin.hh:332-537 with the integrands of pydemo/advectiondiffusion.py:33-60), for the
    homogeneous part of the SIPG / upwind advection-diffusion-reaction operator, boundary elements included."""
    import oracle_lib as ol
    lo, hi = [-1.0, 0.0, 0.5], [1.0, 0.5, 2.0]
    h = np.array([(hi[d] - lo[d]) / cells[d] for d in range(3)])
    eps, b, c, beta = 0.3, (1.0, -0.4, 0.25), 0.7, 10.0 * (order + 1) ** 2
    n1 = order + 1
    out = np.zeros(5 * 3 * n1 * n1)
    pt.pt_kron_tables.argtypes = [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_double, _dp]
    assert pt.pt_kron_tables(3, order, h, np.array([eps, *b, c, beta]), mask, 1, 1, 1.0, out) == n1
    K = out.reshape(5, 3, n1, n1)
    sp = ol.Space(cells, lo, hi, ol.DG_LEGENDRE, order)
    op = ol.Operator(sp, eps=eps, b=b, c=c, beta=beta, dirichlet_mask=mask, data=1, skeleton=True, boundary=True)
    u = np.random.default_rng(order).uniform(-1, 1, sp.size)
    w_ref = op.apply(u, linear=True)
    w = _kron_apply(K, n1, cells, u)
    assert np.abs(w - w_ref).max() < 1e-13 * np.abs(w_ref).max()


@pytest.mark.parametrize("dim,order,cells,mask", [(2, 1, [5, 4], 0b0110), (2, 2, [4, 3], 0b1111), (3, 1, [3, 3, 2], 0b100001), (3, 2, [3, 2, 2], 0b111111)])
def test_lagrange_row_tables_reproduce_the_dense_reference_loop(pt, dim, order, cells, mask):
    """A = T_0 x M_1 x M_2 + M_0 x T_1 x M_2 + M_0 x M_1 x T_2 on the node lattice, with the banded 1-D tables the product builds
    (kron_tables.hpp: build_lagrange_rows) == the oracle's restatement of the reference's dense element loop on the continuous
    Lagrange space (weak boundary terms on the masked sides, no strong constraints), through the oracle's dof map."""
    import oracle_lib as ol
    lo, hi = [-1.0, 0.0, 0.5][:dim], [1.0, 0.5, 2.0][:dim]
    h = np.array([(hi[d] - lo[d]) / cells[d] for d in range(dim)] + [1.0] * (3 - dim))
    eps, b, c, beta = 0.3, (1.0, -0.4, 0.25 if dim == 3 else 0.0), 0.7, 12.0
    k, W = order, 2 * order + 1
    n3 = np.array(list(cells) + [1] * (3 - dim), dtype=np.int32)
    pt.pt_lagrange_rows.restype = C.c_longlong
    pt.pt_lagrange_rows.argtypes = [C.c_int, C.c_int, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"), _dp, _dp, C.c_int, C.c_int, C.c_void_p]
    par = np.array([eps, *b, c, beta])
    total = pt.pt_lagrange_rows(dim, order, n3, h, par, mask, 1, None)
    out = np.zeros(total)
    pt.pt_lagrange_rows(dim, order, n3, h, par, mask, 1, out.ctypes.data_as(C.c_void_p))
    L = [k * cells[d] + 1 if d < dim else 1 for d in range(3)]
    mats, pos = [], 0
    for _ in range(2):                     # M then T: banded rows -> dense L_d x L_d
        per_axis = []
        for d in range(3):
            rows = out[pos:pos + L[d] * W].reshape(L[d], W)
            pos += L[d] * W
            D = np.zeros((L[d], L[d]))
            for g in range(L[d]):
                for j in range(W):
                    col = g - k + j
                    if 0 <= col < L[d]:
                        D[g, col] = rows[g, j]
                    else:
                        assert rows[g, j] == 0.0
            per_axis.append(D)
        mats.append(per_axis)
    M, T = mats
    sp = ol.Space(cells, lo, hi, ol.LAGRANGE, order)
    op = ol.Operator(sp, eps=eps, b=b, c=c, beta=beta, dirichlet_mask=mask, data=1, boundary=True)
    # lattice coordinate of every dof from the oracle's dof map: k * element coordinate + local multi-index
    mi = sp.multiindex()[:, :dim]
    lat = np.zeros((sp.size, 3), dtype=np.int64)
    ne = int(np.prod(cells))
    for e in range(ne):
        ec = [(e // int(np.prod(cells[:d]))) % cells[d] for d in range(dim)]
        lat[sp.dofmap(e), :dim] = k * np.array(ec) + mi
    u = np.random.default_rng(7 * dim + order).uniform(-1, 1, sp.size)
    U = np.zeros(L)
    U[lat[:, 0], lat[:, 1], lat[:, 2]] = u
    Wl = np.zeros(L)
    for d in range(dim):
        f = [T[a] if a == d else M[a] for a in range(3)]
        Wl += np.einsum("ia,jb,kc,abc->ijk", f[0], f[1], f[2], U)
    w = Wl[lat[:, 0], lat[:, 1], lat[:, 2]]
    w_ref = op.apply(u, linear=True)
    assert np.abs(w - w_ref).max() < 1e-13 * np.abs(w_ref).max()


@pytest.mark.parametrize("dim,order,cells,mask", [(2, 1, [5, 4], 0b0110), (3, 2, [3, 2, 2], 0b111111), (3, 1, [1, 2, 3], 0b000011)])
def test_lattice_stencil_is_the_row_table_by_node_type(pt, dim, order, cells, mask):
    """build_lagrange_stencil (what lagrange_lattice.cuh reads: one row per node type + corrections of the diagonal on the first / last
    lattice plane) expands to exactly the assembled rows of build_lagrange_rows (checked against the dense loop above)."""
    h = np.array([0.4, 0.25, 0.75])
    par = np.array([0.3, 1.0, -0.4, 0.25, 0.7, 12.0])
    k, W = order, 2 * order + 1
    n3 = np.array(list(cells) + [1] * (3 - dim), dtype=np.int32)
    _ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    pt.pt_lagrange_rows.restype = C.c_longlong
    pt.pt_lagrange_rows.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, C.c_int, C.c_void_p]
    pt.pt_lagrange_stencil.argtypes = [C.c_int, C.c_int, _ip, _dp, _dp, C.c_int, C.c_int, _dp]
    rows = np.zeros(pt.pt_lagrange_rows(dim, order, n3, h, par, mask, 1, None))
    pt.pt_lagrange_rows(dim, order, n3, h, par, mask, 1, rows.ctypes.data_as(C.c_void_p))
    st = np.zeros(72)
    pt.pt_lagrange_stencil(dim, order, n3, h, par, mask, 1, st)
    sM, sT = st[:30].reshape(3, 2, 5), st[30:60].reshape(3, 2, 5)
    Mlo, Mhi, Tlo, Thi = st[60:63], st[63:66], st[66:69], st[69:72]
    L = [k * cells[d] + 1 if d < dim else 1 for d in range(3)]
    pos = 0
    for S, lo_c, hi_c in ((sM, Mlo, Mhi), (sT, Tlo, Thi)):
        for d in range(3):
            R = rows[pos:pos + L[d] * W].reshape(L[d], W)
            pos += L[d] * W
            for g in range(L[d]):
                expect = S[d][0 if g % k == 0 else 1][:W].copy()
                if d < dim:
                    if g == 0:
                        expect[k] += lo_c[d]
                    if g == L[d] - 1:
                        expect[k] += hi_c[d]
                for j in range(W):
                    if 0 <= g - k + j < L[d]:             # nodes outside the box read as zero
                        assert abs(expect[j] - R[g, j]) <= 1e-15 * max(1.0, abs(R[g, j]))


def test_legendre_polynomials_are_the_reference_table_by_horner(pt):
    # LegendrePolynomials::evaluate / jacobian (shapefunctionset/legendrepolynomials.hh:24-46) over the reference's coefficient table
    # (legendrepolynomials.cc, extracted into tests/golden/legendre_table.json incl. its order-10 typo -34920): the product evaluates
    # the same Horner scheme in the same order (built with -ffp-contract=off like the library's host code) -> bit-identical
    tab = json.load(open(os.path.join(GOLD, "legendre_table.json")))
    fac, wgt = np.array(tab["factor"]), np.array(tab["weight"])
    assert fac[10][3] == -34920.0
    for num in range(11):
        for x in [0.0, 0.1127016653792583, 0.5, 0.77, 1.0]:
            phi = fac[num][num]
            for i in range(num - 1, -1, -1):
                phi = phi * x + fac[num][i]
            assert pt.pt_basis_1d(1, 10, num, x, 0) == wgt[num] * phi
            dphi = 0.0
            if num >= 1:
                dphi = fac[num][num] * num
                for i in range(num - 1, 0, -1):
                    dphi = dphi * x + fac[num][i] * i
            assert pt.pt_basis_1d(1, 10, num, x, 1) == wgt[num] * dphi
