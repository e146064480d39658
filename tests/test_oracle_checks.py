"""The reference's own four acceptance checks, run against the CPU oracle.

The reference pins this path by (1) an analytic L2-error threshold, (2) matrix-free == assembled,
(3) EOC >= k+1-0.1 and (4) invariance under communication; there are no golden vectors.
"""
import math

import numpy as np
import pytest
import scipy.sparse.linalg as spla

import oracle_lib as ol

ALL2 = 0b1111
ALL3 = 0b111111


def test_check1_mass_system_l2_error_inverseoperatortest():
    # dune/fem/solver/test/inverseoperatortest.cc: YaspGrid 2D 16x16 refined once, P2 Lagrange, mass system
    # for prod sin(pi x_k), fem.solver.tolerance 1e-15 absolute; pass iff L2 distance < 5e-6.
    sp = ol.Space([32, 32], [0, 0], [1, 1], ol.LAGRANGE, 2)
    op = ol.Operator(sp, eps=0.0, c=1.0, data=2)
    b = -op.apply(np.zeros(sp.size))                  # affine shift b = -L[0]
    it, x, hist = op.cg(b, np.zeros(sp.size), 1e-15, 1000, tolcrit=0)
    assert it > 0, "CG did not converge"
    err = sp.l2error(x, 2)
    assert err < 5e-6, err


@pytest.mark.parametrize("dim,kind,order", [(2, ol.DG_LEGENDRE, 2), (2, ol.DG_LEGENDRE_HIER, 1), (3, ol.DG_LEGENDRE, 1),
                                            (3, ol.DG_LEGENDRE_HIER, 2)])
def test_check2_matrix_free_equals_assembled_dg(dim, kind, order):
    n = [3, 2, 2][:dim]
    sp = ol.Space(n, [-1.0] * dim, [1.0, 0.5, 2.0][:dim], kind, order)
    op = ol.Operator(sp, eps=0.3, b=(1.0, -0.5, 0.25), c=0.7, beta=20.0 * order * order, dirichlet_mask=0b000011,
                     data=1, skeleton=True, boundary=True)
    A = op.assemble_dense()
    rng = np.random.default_rng(1)
    u = rng.uniform(-1, 1, sp.size)
    w = op.apply(u, linear=True)
    np.testing.assert_allclose(w, A @ u, rtol=0, atol=1e-12 * np.abs(A).sum(axis=1).max())
    # affine structure: L[u] - L[0] == A u
    w_aff = op.apply(u) - op.apply(np.zeros(sp.size))
    np.testing.assert_allclose(w_aff, w, rtol=0, atol=1e-12 * np.abs(A).sum(axis=1).max())


@pytest.mark.parametrize("dim,order,numbering", [(2, 1, 0), (2, 2, 0), (3, 2, 0), (3, 2, 1), (2, 2, 1)])
def test_check2_matrix_free_equals_assembled_lagrange_dirichlet(dim, order, numbering):
    n = [3, 2, 2][:dim]
    sp = ol.Space(n, [0.0] * dim, [1.0] * dim, ol.LAGRANGE, order, numbering=numbering)
    op = ol.Operator(sp, eps=1.0, data=2, dirichlet_mask=ALL3 if dim == 3 else ALL2, strong_dirichlet=True)
    A = op.assemble_dense()
    rng = np.random.default_rng(2)
    u = rng.uniform(-1, 1, sp.size)
    np.testing.assert_allclose(op.apply(u, linear=True), A @ u, rtol=0, atol=1e-12 * np.abs(A).sum(axis=1).max())
    # the dof map is a bijection onto [0, size)
    seen = np.zeros(sp.size, dtype=int)
    for e in range(sp.elements):
        seen[sp.dofmap(e)] += 1
    assert (seen > 0).all()


def _solve_nonsymmetric(op, sp):
    b = -op.apply(np.zeros(sp.size))
    A = spla.LinearOperator((sp.size, sp.size), matvec=lambda v: op.apply(np.ascontiguousarray(v), linear=True))
    x, info = spla.gmres(A, b, rtol=1e-12, atol=0.0, restart=200, maxiter=50)
    assert info == 0
    return x


@pytest.mark.parametrize("eps", [1.0, 1e-5])
def test_check3_eoc_advection_diffusion_dg(eps):
    # pydemo/advectiondiffusion.py:93-147: [-1,1]^2, 4x4 start, 3 refinements, order 2, beta = 20 k^2,
    # weak Dirichlet on x0 = +-1, EOC of the last refinement >= order+1-0.1
    order = 2
    errs = []
    for n in [4, 8, 16, 32]:
        sp = ol.Space([n, n], [-1, -1], [1, 1], ol.DG_LEGENDRE_HIER, order)
        op = ol.Operator(sp, eps=eps, b=(1.0, 0.0), beta=20.0 * order ** 2, dirichlet_mask=0b0011, data=1,
                         skeleton=True, boundary=True)
        errs.append(sp.l2error(_solve_nonsymmetric(op, sp), 1))
    eoc = [math.log(errs[i + 1] / errs[i]) / math.log(0.5) for i in range(3)]
    assert eoc[-1] - (order + 1) > -0.1, (errs, eoc)


def test_check3_eoc_poisson_lagrange_p2_cg():
    errs = []
    for n in [4, 8, 16]:
        sp = ol.Space([n, n], [0, 0], [1, 1], ol.LAGRANGE, 2)
        op = ol.Operator(sp, eps=1.0, data=2, dirichlet_mask=ALL2, strong_dirichlet=True)
        mask, g = op.dirichlet()
        b = -op.apply(np.zeros(sp.size))
        x0 = np.where(mask, g, 0.0)                   # FemScheme::solve sets constraints first (femscheme.hh:247-250)
        it, x, hist = op.cg(b, x0, 1e-12, 2000)
        assert it > 0
        errs.append(sp.l2error(x, 2))
    eoc = [math.log(errs[i + 1] / errs[i]) / math.log(0.5) for i in range(2)]
    assert eoc[-1] > 3 - 0.1, (errs, eoc)


def test_check4_invariance_under_domain_decomposition_dg():
    # rank-local apply with ghost neighbours (one-sided face integrals, galerkin.hh:866-878) must reproduce
    # the single-domain result on the owned elements (dgcomm.cc:183-229 checks the same invariance)
    sp = ol.Space([4, 3, 2], [-1, -1, -1], [1, 1, 1], ol.DG_LEGENDRE, 2)
    op = ol.Operator(sp, eps=0.1, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1, skeleton=True, boundary=True)
    u = np.random.default_rng(3).uniform(-1, 1, sp.size)
    w = op.apply(u)
    w0 = op.apply_box(u, [0, 0, 0], [2, 3, 2])
    w1 = op.apply_box(u, [2, 0, 0], [4, 3, 2])
    nb = sp.local_size
    for e in range(sp.elements):
        ex = e % 4
        mine = w0 if ex < 2 else w1
        np.testing.assert_allclose(mine[e * nb:(e + 1) * nb], w[e * nb:(e + 1) * nb], rtol=0, atol=1e-12 * np.abs(w).max())


def test_check4_invariance_under_domain_decomposition_lagrange_add():
    # Lagrange spaces communicate with Add on shared dofs (lagrange/space.hh:92)
    sp = ol.Space([4, 4, 2], [0, 0, 0], [1, 1, 1], ol.LAGRANGE, 2)
    op = ol.Operator(sp, eps=1.0, c=0.5, data=2)
    u = np.random.default_rng(4).uniform(-1, 1, sp.size)
    w = op.apply(u)
    w0 = op.apply_box(u, [0, 0, 0], [2, 4, 2])
    w1 = op.apply_box(u, [2, 0, 0], [4, 4, 2])
    np.testing.assert_allclose(w0 + w1, w, rtol=0, atol=1e-12 * np.abs(w).max())


def test_threaded_apply_matches_serial():
    sp = ol.Space([5, 4, 3], [-1, -1, -1], [1, 1, 1], ol.DG_LEGENDRE, 2)
    kw = dict(eps=0.1, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1, skeleton=True, boundary=True)
    u = np.random.default_rng(5).uniform(-1, 1, sp.size)
    w1 = ol.Operator(sp, **kw).apply(u)
    w4 = ol.Operator(sp, threads=4, **kw).apply(u)
    np.testing.assert_allclose(w4, w1, rtol=0, atol=1e-13 * np.abs(w1).max())


def test_cg_sign_conventions_and_negative_count():
    # r = Ax - b, p = b - Ax, iterations negative when maxIterations is hit (cg.hh:116)
    sp = ol.Space([8, 6], [0, 0], [1, 2], ol.LAGRANGE, 1)
    op = ol.Operator(sp, eps=1.0, c=0.3, data=1, dirichlet_mask=ALL2, strong_dirichlet=True)
    mask, g = op.dirichlet()
    b = -op.apply(np.zeros(sp.size))
    it, x, hist = op.cg(b, np.where(mask, g, 0.0), 1e-30, 3)
    assert it == -3 and len(hist) == 3
    it2, x2, hist2 = op.cg(b, np.where(mask, g, 0.0), 1e-10, 500)
    assert 0 < it2 < 500
    r = op.apply(x2, linear=True) - b
    assert np.linalg.norm(r) <= 1e-10 * 1.01
    np.testing.assert_allclose(hist2[:3], hist, rtol=1e-14)


def test_mol_inverse_mass_is_projection_scaling():
    """MOLGalerkinOperator (schemes/molgalerkin.hh:100-124): with the mass integrand L[u] = (u, v) the method-of-lines
    operator M^-1 L is the identity on a DG Legendre space -- the check that pins `referenceVolume / volume`
    (operator/1order/localmassmatrix.hh:304-311) in the oracle."""
    n, lo, hi = [3, 4, 2], [-1.0, 0.0, 0.5], [1.0, 3.0, 1.0]
    sp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    u = np.random.default_rng(2).uniform(-1, 1, sp.size)
    op = ol.Operator(sp, eps=0.0, c=1.0)
    op.setInverseMass(True)
    np.testing.assert_allclose(op.apply(u), u, rtol=0, atol=1e-13)
    lsp = ol.Space([2, 2], [0, 0], [1, 1], ol.LAGRANGE, 1)
    with pytest.raises(ValueError):
        ol.Operator(lsp).setInverseMass(True)


def test_bicgstab_oracle_solves_nonsymmetric_system():
    """LinearSolver::bicgstab restatement (solver/linear/bicgstab.hh:64-214) on the advection-diffusion DG operator:
    converges, returns the signed iteration count, and its last `res` is the true residual norm."""
    sp = ol.Space([4, 4, 4], [0, 0, 0], [1, 1, 1], ol.DG_LEGENDRE_HIER, 1)
    kw = dict(eps=0.1, b=(1.0, 0.5, 0.2), c=1.0, beta=20.0, dirichlet_mask=0b111111, data=2)
    op = ol.Operator(sp, skeleton=True, boundary=True, **kw)
    b = -op.apply(np.zeros(sp.size))
    it, x, hist = op.bicgstab(b, np.zeros(sp.size), 1e-10, 500)
    assert 0 < it < 500 and hist[-1] < 1e-10
    assert abs(np.linalg.norm(op.apply(x, linear=True) - b) - hist[-1]) < 1e-12
    it2, _, hist2 = op.bicgstab(b, np.zeros(sp.size), 1e-30, 7)
    assert it2 == -7 and np.allclose(hist2, hist[:7], rtol=1e-12)
    # relative criterion scales the tolerance by sqrt(b.b) (bicgstab.hh:88-92)
    it3, _, hist3 = op.bicgstab(b, np.zeros(sp.size), 1e-6, 500, tolcrit=1)
    assert hist3[-1] < 1e-6 * np.linalg.norm(b) <= hist3[-2]


def test_gmres_oracle_solves_nonsymmetric_system():
    """LinearSolver::gmres restatement (solver/linear/gmres.hh:117-301): the last |g[j+1]| is the true residual norm, a
    full-length Krylov space converges faster than restarted cycles, maxIterations yields a negative count."""
    sp = ol.Space([4, 4, 4], [0, 0, 0], [1, 1, 1], ol.DG_LEGENDRE_HIER, 1)
    kw = dict(eps=0.1, b=(1.0, 0.5, 0.2), c=1.0, beta=20.0, dirichlet_mask=0b111111, data=2)
    op = ol.Operator(sp, skeleton=True, boundary=True, **kw)
    b = -op.apply(np.zeros(sp.size))
    it20, x, hist = op.gmres(b, np.zeros(sp.size), 1e-10, 1000, restart=20)
    assert it20 > 0 and abs(np.linalg.norm(op.apply(x, linear=True) - b) - hist[-1]) < 1e-12
    it5, _, _ = op.gmres(b, np.zeros(sp.size), 1e-10, 1000, restart=5)
    assert it5 > it20
    it, _, h7 = op.gmres(b, np.zeros(sp.size), 1e-30, 7, restart=5)
    assert it == -7 and np.all(np.diff(h7[:5]) <= 0)           # GMRES residuals decrease monotonically inside a cycle


def test_difference_quotient_jacobian_oracle():
    """AutomaticDifferenceLinearOperator restatement: for a LINEAR operator the quotient reproduces A v (up to the rounding
    noise 1/eps amplifies), and Newton-GMRES on the cubic reaction model converges quadratically."""
    sp = ol.Space([3, 3, 3], [0, 0, 0], [1, 1, 1], ol.DG_LEGENDRE_HIER, 1)
    lin = ol.Operator(sp, skeleton=True, boundary=True, eps=0.5, b=(0.3, 0.2, 0.1), c=1.0, beta=20.0, dirichlet_mask=0b111111, data=2)
    rng = np.random.default_rng(1)
    u, v = rng.uniform(-1, 1, sp.size), rng.uniform(-1, 1, sp.size)
    lin.linearize(u)
    w, eps = lin.applyJacobian(v)
    av = lin.apply(v, linear=True)
    assert np.abs(w - av).max() < 1e-6 * np.abs(av).max()
    assert abs(eps - math.sqrt((1 + np.linalg.norm(u)) * np.finfo(float).eps / np.dot(v, v))) < 1e-20
    op = ol.Operator(sp, skeleton=True, boundary=True, eps=0.5, b=(0.3, 0.2, 0.1), c=1.0, gamma=2.0, beta=20.0, dirichlet_mask=0b111111, data=2)
    x, norms = np.zeros(sp.size), []
    for _ in range(6):
        r = op.apply(x)
        norms.append(np.linalg.norm(r))
        if norms[-1] < 1e-7:          # the quotient's rounding noise (~1e-8) is the floor of a Jacobian-free Newton method
            break
        op.linearize(x)
        it, d, _ = op.gmres_jacobian(-r, np.zeros(sp.size), 1e-7, 400, tolcrit=2, restart=30)
        assert it > 0
        x += d
    assert len(norms) == 4 and norms[-1] < 1e-7 and norms[2] < 1e-2 * norms[1] and norms[3] < 1e-3 * norms[2]


def test_jacobi_preconditioned_cg_oracle():
    """preconditioned branch of LinearSolver::cg (solver/linear/cg.hh:52-56, 72-107) with B = diag(A)^-1: same solution as plain
    CG in fewer iterations on the Q2 Lagrange Laplacian; the diagonal probed with unit vectors is positive."""
    sp = ol.Space([4, 4, 4], [0, 0, 0], [1, 1, 1], ol.LAGRANGE, 2)
    kw = dict(eps=1.0, c=0.1, data=1, dirichlet_mask=0b111111, strong_dirichlet=True)
    op = ol.Operator(sp, **kw)
    b = -op.apply(np.zeros(sp.size))
    mask, g = op.dirichlet()
    x0 = np.where(mask, g, 0.0)
    d = op.diagonal()
    assert d.min() > 0 and np.all(d[mask != 0] == 1.0)
    it, x, _ = op.cg(b, x0, 1e-10, 2000)
    itp, xp, hist = op.pcg(d, b, x0, 1e-10, 2000)
    assert 0 < itp < it and np.abs(x - xp).max() < 1e-8 and hist[-1] <= 1e-10


@pytest.mark.parametrize("order,hier", [(1, True), (2, False), (2, True), (3, True)])
def test_kronecker_factorisation_probed_from_dense_loop(order, hier):
    """The claim the GPU Kronecker kernels rest on: for linear constant-coefficient integrands on a uniform box the reference
    loop factorises into 1-D operators acting along one tensor axis.  The matrices are PROBED from the dense restatement on a
    3x3x3 mesh (no derivation), then applied in Kronecker form on another mesh: identical to the dense loop to rounding,
    including boundary elements, both dof orderings, affine part."""
    n, lo, hi = [5, 4, 3], [-1, -1, -1], [1, 1.5, 1]
    kind = ol.DG_LEGENDRE_HIER if hier else ol.DG_LEGENDRE
    kw = dict(eps=0.3, b=(1.0, -0.5, 0.25), c=0.7, beta=20.0 * order ** 2, dirichlet_mask=0b011011, data=1)
    sp = ol.Space(n, lo, hi, kind, order)
    op = ol.Operator(sp, skeleton=True, boundary=True, **kw)
    u = np.random.default_rng(1).uniform(-1, 1, sp.size)
    k = ol.KroneckerCpu(n, lo, hi, kind, order, threads=3, **kw)
    ref = op.apply(u, linear=True)
    assert np.abs(k.apply(u) - ref).max() < 1e-12 * np.abs(ref).max()
    b = -op.apply(np.zeros(sp.size))
    full = op.apply(u)
    assert np.abs(k.apply(u, b) - full).max() < 1e-12 * np.abs(full).max()
