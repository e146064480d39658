"""Pins the oracle's restatements against the reference ITSELF where the reference compiles from its own source files
(oracle/_ref/libdunefem_ref.so, built by `make -C oracle ref` from /root/reference): the Krylov loops of
dune/fem/solver/linear/{cg,bicgstab,gmres}.hh run on the oracle's operators, the Gauss tables, the Legendre polynomials and
the orthonormal cube bases.  CPU only; skipped where neither the built library nor the reference tree exists."""
import numpy as np
import pytest

import oracle_lib as ol
import reference_lib as rl
import solver_cases as sc

pytestmark = pytest.mark.skipif(not rl.available(), reason="oracle/_ref not built and /root/reference absent")


_poisson = sc.poisson


@pytest.mark.parametrize("dim,order,n", [(2, 1, [24, 24]), (2, 2, [10, 9]), (3, 2, [5, 4, 3])])
@pytest.mark.parametrize("tolcrit", [0, 1, 2])
def test_cg_restatement_is_the_reference_loop(dim, order, n, tolcrit):
    sp, op, b = _poisson(dim, order, n)
    A = lambda u: op.apply(u, linear=True)
    x0 = np.zeros(sp.size)
    it_r, x_r, h_r = rl.cg(A, b, x0, 1e-9, 60, tolcrit)
    it_o, x_o, h_o = op.cg(b, x0, 1e-9, 60, tolcrit)
    assert it_r == it_o and len(h_r) == abs(it_o)
    np.testing.assert_array_equal(h_o, h_r)          # same operations in the same order: bit-identical
    np.testing.assert_array_equal(x_o, x_r)
    # not converged: negative iteration count (cg.hh:116)
    it_r, _, _ = rl.cg(A, b, x0, 1e-14, 5, tolcrit)
    it_o, _, _ = op.cg(b, x0, 1e-14, 5, tolcrit)
    assert it_r == it_o == -5


def test_preconditioned_cg_restatement_is_the_reference_loop():
    sp, op, b = _poisson(3, 2, [4, 4, 3])
    A = lambda u: op.apply(u, linear=True)
    d = op.diagonal()
    x0 = np.zeros(sp.size)
    it_r, x_r, h_r = rl.cg(A, b, x0, 1e-10, 80, 0, precon=lambda r: r / d)
    it_o, x_o, h_o = op.pcg(d, b, x0, 1e-10, 80, 0)
    assert it_r == it_o
    np.testing.assert_allclose(h_o, h_r, rtol=1e-12)   # the oracle multiplies by 1/d, the callback divides
    np.testing.assert_allclose(x_o, x_r, rtol=1e-11, atol=1e-14)


@pytest.mark.parametrize("measure", [0, 1])
def test_legacy_conjugate_gradient_solver_gives_the_same_iterates(measure):
    # the reference holds a second CG, ConjugateGradientSolver (solver/cginverseoperator.hh:595-720, behind the legacy CGInverseOperator):
    # same recurrence and stopping rule as LinearSolver::cg -- the restatement (and with it the device loop) agrees with both
    sp, op, b = _poisson(3, 2, [4, 4, 3])
    A = lambda u: op.apply(u, linear=True)
    x0 = np.zeros(sp.size)
    for maxit in (7, 400):
        it_l, x_l = rl.legacy_cg(A, b, x0, 1e-9, maxit, measure)
        it_o, x_o, _ = op.cg(b, x0, 1e-9, maxit, measure)
        assert it_l == abs(it_o) and (it_o > 0) == (maxit == 400)
        np.testing.assert_allclose(x_o, x_l, rtol=0, atol=1e-13 * np.abs(x_l).max())
    d = op.diagonal()
    it_l, x_l = rl.legacy_cg(A, b, x0, 1e-9, 400, measure, precon=lambda r: r / d)
    it_o, x_o, _ = op.pcg(d, b, x0, 1e-9, 400, measure)
    assert it_l == it_o
    np.testing.assert_allclose(x_o, x_l, rtol=0, atol=1e-12 * np.abs(x_l).max())


_advdiff = sc.advdiff


@pytest.mark.parametrize("order", [1, 2])
@pytest.mark.parametrize("tolcrit", [0, 1, 2])
def test_bicgstab_restatement_is_the_reference_loop(order, tolcrit):
    sp, op, b = _advdiff(order, eps=1.0)
    A = lambda u: op.apply(u, linear=True)
    x0 = np.zeros(sp.size)
    # a fixed number of steps: same recurrence, same scalars (BiCGStab amplifies rounding differences -- FMA contraction in
    # the vector updates -- too quickly for a comparison of whole convergence histories)
    it_r, x_r, h_r = rl.bicgstab(A, b, x0, 1e-12, 12, tolcrit)
    it_o, x_o, h_o = op.bicgstab(b, x0, 1e-12, 12, tolcrit)
    assert it_r == it_o == -12
    # the reference reports `res` only for iterations that continue (bicgstab.hh:186-189): its log is one entry shorter
    assert len(h_r) == 11 and len(h_o) == 12
    np.testing.assert_allclose(h_o[:11], h_r, rtol=1e-9)
    np.testing.assert_allclose(h_o[:4], h_r[:4], rtol=1e-13)
    np.testing.assert_allclose(x_o, x_r, rtol=0, atol=1e-9 * np.abs(x_r).max())
    # run to convergence: positive counts, same solution
    it_r, x_r, _ = rl.bicgstab(A, b, x0, 1e-8, 2000, tolcrit)
    it_o, x_o, _ = op.bicgstab(b, x0, 1e-8, 2000, tolcrit)
    assert it_r > 0 and it_o > 0
    np.testing.assert_allclose(x_o, x_r, rtol=0, atol=1e-5 * np.abs(x_r).max())


@pytest.mark.parametrize("order,restart", [(1, 5), (2, 20)])
@pytest.mark.parametrize("tolcrit", [0, 1, 2])
def test_gmres_restatement_is_the_reference_loop(order, restart, tolcrit):
    sp, op, b = _advdiff(order)
    A = lambda u: op.apply(u, linear=True)
    x0 = np.zeros(sp.size)
    it_r, x_r, h_r = rl.gmres(A, b, x0, 1e-7, 600, tolcrit, restart)
    it_o, x_o, h_o = op.gmres(b, x0, 1e-7, 600, tolcrit, restart)
    assert it_r == it_o and it_r > 0 and len(h_r) == len(h_o)
    np.testing.assert_allclose(h_o, h_r, rtol=1e-9)
    np.testing.assert_allclose(h_o[:10], h_r[:10], rtol=1e-13)
    np.testing.assert_allclose(x_o, x_r, rtol=1e-9, atol=1e-12)
    # Known deviation, on purpose: when maxIterations is hit inside a restart cycle the reference keeps restarting with
    # one-step Krylov spaces and finally tests a stale g[last] (gmres.hh:246-289) -- it returns -(maxIterations + 1) after one
    # more apply; the oracle and the device solver stop at -maxIterations with the iterate of that moment.
    it_r, _, _ = rl.gmres(A, b, x0, 1e-14, 7, tolcrit, restart)
    it_o, _, _ = op.gmres(b, x0, 1e-14, 7, tolcrit, restart)
    assert it_o == -7 and it_r in (-7, -8)


@pytest.mark.parametrize("eps", [0.0, 1e-6])
def test_difference_quotient_restatement_is_the_reference_operator(eps):
    # AutomaticDifferenceOperator / AutomaticDifferenceLinearOperator (operator/common/automaticdifferenceoperator.hh:110-166) on the
    # oracle's NON-LINEAR operator (gamma u^3): same epsilon rule, same order of operations -> bit-identical
    sp = ol.Space([5, 4, 3], [0.0] * 3, [1.0] * 3, ol.LAGRANGE, 2)
    op = ol.Operator(sp, eps=1.0, c=0.5, gamma=2.0, data=2, dirichlet_mask=0b111111, strong_dirichlet=True)
    rng = np.random.default_rng(5)
    u = rng.uniform(-1, 1, sp.size)
    args = np.stack([rng.uniform(-1, 1, sp.size), 1e-9 * rng.uniform(-1, 1, sp.size), 1e4 * rng.uniform(-1, 1, sp.size), np.zeros(sp.size)])
    ref = rl.difference_quotient(lambda v: op.apply(v), u, args, eps=eps)
    op.linearize(u, eps)
    for a, r in zip(args, ref):
        np.testing.assert_array_equal(op.applyJacobian(a)[0], r)
    # the reference's default comes from the parameter file; absent -> 0 -> the dynamic choice (automaticdifferenceoperator.hh:99-101)
    if eps == 0.0:
        np.testing.assert_array_equal(rl.difference_quotient(lambda v: op.apply(v), u, args, from_parameter=True), ref)
    # the quotient approximates the derivative of the cubic term
    lin = op.apply(u + 1e-7 * args[0]) - op.apply(u)
    assert np.abs(ref[0] - lin / 1e-7).max() < 1e-5 * np.abs(ref[0]).max()


_newton_keys = sc.newton_keys


_reaction_diffusion = sc.reaction_diffusion


def test_newton_restatement_of_the_gpu_tests_is_the_reference_loop():
    # the configuration and the restatement tests/test_gpu_jit.py::test_newton_inverse_operator checks the device against, here against
    # Dune::Fem::NewtonInverseOperator itself (difference-quotient Jacobian and GMRES of the reference, oracle operator through callbacks)
    from test_gpu_jit import _oracle_newton
    sp = ol.Space([4, 4, 3], [-1.0] * 3, [1.0] * 3, ol.DG_LEGENDRE_HIER, 2)
    op = ol.Operator(sp, skeleton=True, boundary=True, eps=0.5, b=(1.0, 0.0, 0.0), c=1.0, gamma=2.0, beta=80.0, dirichlet_mask=0b000011, data=1)
    it, lit, delta, w = _oracle_newton(op, np.zeros(sp.size), 1e-7, 2 ** 31 - 1, 1e-7, 4000, 30)
    it_r, lit_r, fail_r, delta_r, w_r = rl.newton(lambda v: op.apply(v), np.zeros(sp.size), _newton_keys(1e-7, 2 ** 31 - 1, 1e-7, 4000, 30, False))
    assert (it, lit) == (it_r, lit_r) and fail_r == 0 and 2 <= it <= 12
    assert delta == delta_r
    np.testing.assert_array_equal(w, w_r)
    assert ol.newton(op, np.zeros(sp.size), 1e-7, 2 ** 31 - 1, 1e-7, 4000, 30, 2)[:4] == (it_r, lit_r, 0, delta_r)


@pytest.mark.parametrize("gamma,amp,c,seed,line_search,maxit,failure", [
    (10.0, 2.0, -8.0, 4, False, 40, 0),      # plain Newton
    (10.0, 2.0, -8.0, 4, True, 40, 0),       # the line search halves the step in iterations 5 and 6 and saves 8 iterations
    (5.0, 4.0, -8.0, 4, True, 3, 5),         # NewtonFailure::TooManyIterations
    (10.0, 2.0, -12.0, 3, True, 40, 7),      # NewtonFailure::LinearSolverFailed (budget shared by all steps, :745-757)
])
def test_newton_line_search_and_failure_codes_are_the_reference_ones(gamma, amp, c, seed, line_search, maxit, failure):
    sp, op = _reaction_diffusion(gamma, c)
    w0 = amp * np.random.default_rng(seed).uniform(-1, 1, sp.size)
    it_r, lit_r, fail_r, delta_r, w_r = rl.newton(lambda v: op.apply(v), w0, _newton_keys(1e-7, maxit, 1e-8, 20000, 48, line_search))
    trace = []
    it, lit, fail, delta, w = ol.newton(op, w0, 1e-7, maxit, 1e-8, 20000, 48, 2, line_search, trace=trace)
    assert (it, fail) == (it_r, fail_r) and fail == failure
    # the reference GMRES reports -(maxIterations + 1) where oracle and device report -maxIterations
    assert lit == lit_r or (fail == 7 and lit == lit_r + 1)
    np.testing.assert_allclose(delta, delta_r, rtol=1e-12)
    np.testing.assert_allclose(w, w_r, rtol=0, atol=1e-12 * np.abs(w_r).max())
    if line_search and maxit == 40:
        assert max(trace) >= 1


def test_newton_on_a_linear_operator_and_with_a_right_hand_side():
    # op_->nonlinear() == false: one step, no residual re-evaluation, iterations() stays 0 (:761, 791-792); u != 0: solves L[w] = u
    sp, op = _reaction_diffusion(0.0, 1.0)
    u = np.random.default_rng(8).uniform(-1, 1, sp.size)
    it_r, lit_r, fail_r, delta_r, w_r = rl.newton(lambda v: op.apply(v), np.zeros(sp.size), _newton_keys(1e-8, 10, 1e-7, 5000, 30, False, "relative"), u=u, nonlinear=False)
    it, lit, fail, delta, w = ol.newton(op, np.zeros(sp.size), 1e-8, 10, 1e-7, 5000, 30, 1, u=u, nonlinear=False)
    assert (it, lit, fail) == (it_r, lit_r, fail_r) and it == 0 and lit > 0
    np.testing.assert_allclose(delta, delta_r, rtol=1e-13)     # still the INITIAL residual norm (numpy sums in another order)
    np.testing.assert_allclose(w, w_r, rtol=0, atol=1e-13 * np.abs(w_r).max())
    assert np.abs(op.apply(w) - u).max() < 1e-5
    # the defaults of the reference when no key is given: tolerance 1e-6, linear tolerance 1e-8 absolute, method gmres (first of the list), restart 20
    sp, op = _reaction_diffusion(2.0, 1.0)
    it_r, lit_r, fail_r, delta_r, w_r = rl.newton(lambda v: op.apply(v), np.zeros(sp.size), {})
    it, lit, fail, delta, w = ol.newton(op, np.zeros(sp.size), 1e-6, 2 ** 31 - 1, 1e-8, 2 ** 31 - 1, 20, 0)
    assert (it, lit, fail) == (it_r, lit_r, fail_r) and fail == 0
    np.testing.assert_allclose(w, w_r, rtol=0, atol=1e-13 * np.abs(w_r).max())


def test_gauss_rules_are_the_reference_tables():
    lib = rl.lib()
    assert lib.ref_gauss_maxp() == 10
    for m in range(1, 11):
        x, w, order = rl.gauss_rule(m)
        assert order == 2 * m - 1
        xo, wo = ol.quadrature(1, order)
        assert len(wo) == m
        np.testing.assert_array_equal(xo[:, 0], x)
        np.testing.assert_array_equal(wo, w)
        # the order selection of femquadratures_inline.hh:59-70: smallest rule with order >= requested
        xo2, _ = ol.quadrature(1, max(order - 1, 0))
        assert len(xo2) == m


def test_legendre_polynomials_are_the_reference_evaluation():
    assert rl.lib().ref_legendre_max_order() == 11
    xs = np.random.default_rng(5).uniform(0, 1, 40).tolist() + [0.0, 0.5, 1.0]
    for num in range(11):
        for x in xs:
            for deriv in (0, 1):
                assert ol.lib().fo_legendre(num, x, deriv) == rl.legendre(num, x, deriv), (num, x, deriv)


@pytest.mark.parametrize("dim,max_order", [(2, 4), (3, 4)])
def test_dgonb_shape_functions_are_the_reference_functions(dim, max_order):
    """The oracle's `dgonb` space (products of orthonormal Legendre polynomials of total degree <= k, graded ordering) against
    the reference's expanded polynomials eval_/grad_quadrilateral_2d, eval_/grad_hexahedron_3d (orthonormalbase_{2,3}d.hh).
    Tolerance: the reference evaluates monomial expansions with coefficients up to 1e3 -- rounding only."""
    rng = np.random.default_rng(7)
    for order in range(1, max_order + 1):
        sp = ol.Space([2] * dim, [0.0] * dim, [1.0] * dim, ol.DG_ONB, order)
        assert sp.local_size == (order + 1) * (order + 2) // 2 if dim == 2 else sp.local_size == (order + 1) * (order + 2) * (order + 3) // 6
        for _ in range(25):
            x = np.zeros(3)
            x[:dim] = rng.uniform(0, 1, dim)
            phi, dphi = sp.shape(x[:dim])
            for i in range(sp.local_size):
                v, g = rl.onb_cube(dim, i, x, grad=True)
                assert abs(v - phi[i]) < 1e-12 * max(1.0, abs(v))
                assert np.abs(g - dphi[i, :dim]).max() < 1e-11 * max(1.0, np.abs(g).max())


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_lagrange_basis_and_local_numbering_are_the_reference_ones(dim, order):
    """GenericLagrangePoint / GenericLagrangeBaseFunction of the cube, compiled from the reference: the oracle's local numbering
    (coordinate 0 fastest), node positions and basis values / reference gradients are the reference's"""
    sp = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.LAGRANGE, order)
    x, codim, sub, num = rl.lagrange_cube_points(dim, order)
    mi = sp.multiindex()[:, :dim]
    assert np.abs(x - mi / order).max() < 1e-15                       # local node l sits at multiIndex[l] / order
    rng = np.random.default_rng(dim * 10 + order)
    for xp in rng.uniform(0, 1, (5, dim)):
        phi, dphi = sp.shape(xp)
        for b in range(sp.local_size):
            v, dv = rl.lagrange_cube_evaluate(dim, order, b, xp)
            assert abs(v - phi[b]) < 1e-13 and np.abs(dv - dphi[b, :dim]).max() < 1e-12


@pytest.mark.parametrize("dim,n", [(2, [3, 2]), (3, [2, 2, 2])])
@pytest.mark.parametrize("order", [1, 2, 3])
def test_lagrange_dof_numbering_follows_the_reference_sub_entity_rule(dim, n, order):
    """The global numbering `offset[type] + numDofs(entity) * index(entity) + dofNumber` (space/mapper/indexsetdofmapper.hh:414-427)
    with the reference's own (codim, subEntity, dofNumber) of every local node (GenericLagrangePoint::dofSubEntity): the oracle's
    closed-form YaspGrid numbering must be consistent with it -- nodes of one sub-entity of an element are numbered contiguously,
    in the order of the reference's dofNumber, blocks ordered by entity dimension (vertices first), and two elements sharing a
    sub-entity agree on all its dofs (no twists on Cartesian grids, lagrange/space.hh:68-71)."""
    sp = ol.Space(n, [0.0] * dim, [1.0] * dim, ol.LAGRANGE, order)
    x, codim, sub, num = rl.lagrange_cube_points(dim, order)
    per_entity = [(order - 1) ** (dim - c) if order > 1 or c == dim else 0 for c in range(dim + 1)]     # dofs inside an entity of codim c
    first_of_dim = {}                                                  # entity dimension -> smallest global dof seen
    for e in range(sp.elements):
        g = sp.dofmap(e)
        for c in range(dim + 1):
            for s in np.unique(sub[codim == c]):
                sel = np.where((codim == c) & (sub == s))[0]
                assert len(sel) == per_entity[c]
                dofs = g[sel][np.argsort(num[sel])]                    # in the order of the reference's dof number inside the entity
                assert (np.diff(dofs) == 1).all() and dofs[0] % len(sel) == (first_of_dim.setdefault(dim - c, dofs[0]) % len(sel))
                first_of_dim[dim - c] = min(first_of_dim[dim - c], dofs[0])
    dims = sorted(first_of_dim)
    assert all(first_of_dim[a] < first_of_dim[b] for a, b in zip(dims, dims[1:]))           # vertices | edges | faces | cells


@pytest.mark.parametrize("dim", [2, 3])
def test_first_touch_numbering_visits_sub_entities_in_the_reference_order(dim):
    """AdaptiveLeafIndexSet numbers the sub-entities of an element in reference-element order (gridpart/adaptiveleafindexset.hh:884-906).
    On a one-element mesh the first-touch index of a sub-entity IS its reference number, so the oracle's adaptive-leaf dof of the P2
    node inside sub-entity (codim, sub) must be offset[dimension] + sub with the reference's own (codim, sub) of that node
    (GenericLagrangePoint::dofSubEntity) -- this pins the sub-entity order tables of the oracle (Space::buildSubEntityOrder) and of the
    product (capi.cu: build_adaptive_leaf_map, unstructured.cu: sub_entity_order), which the unstructured numbering rests on too."""
    sp = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.LAGRANGE, 2, numbering=ol.NUMBERING_ADAPTIVE_LEAF)
    x, codim, sub, num = rl.lagrange_cube_points(dim, 2)
    counts = [int((codim == dim - p).sum()) for p in range(dim + 1)]          # entities per dimension p (one P2 node each)
    offset = np.concatenate([[0], np.cumsum(counts)])
    g = sp.dofmap(0)
    for l in range(sp.local_size):
        assert g[l] == offset[dim - codim[l]] + sub[l] and num[l] == 0
    # the same mesh through the unstructured path (host-only numbering of the product)
    import ctypes as C
    from dune_fem_b200 import _capi
    coords, elems = ol.cartesian_as_unstructured([1] * dim, [0.0] * dim, [1.0] * dim)
    coords, elems = np.ascontiguousarray(coords), np.ascontiguousarray(elems, dtype=np.int64)
    size = C.c_int64()
    dofs = np.empty((1, 3 ** dim), dtype=np.int32)
    assert _capi.lib().b200fem_unstructured_numbering(dim, len(coords), _capi.ptr(coords), 1, _capi.ptr(elems, np.int64), 2, C.byref(size), _capi.ptr(dofs, np.int32), None, None) == 0
    assert (dofs[0] == g).all()


@pytest.mark.parametrize("dim", [2, 3])
@pytest.mark.parametrize("order", [1, 2, 3, 5])
@pytest.mark.parametrize("hierarchical", [False, True])
def test_legendre_shape_function_sets_are_the_reference_sets(dim, order, hierarchical):
    """LegendreShapeFunctionSet compiled from the reference (multi-index recursion, hierarchical sort by (order, lexicographic
    multi-index), space/shapefunctionset/legendre.hh:145-250): the oracle's DG spaces hold the same functions in the same ORDER --
    this is the local numbering of the dof vector the device kernels read"""
    sp = ol.Space([1] * dim, [0.0] * dim, [1.0] * dim, ol.DG_LEGENDRE_HIER if hierarchical else ol.DG_LEGENDRE, order)
    for xp in np.random.default_rng(dim + order).uniform(0, 1, (4, dim)):
        phi_ref, dphi_ref = rl.legendre_set(dim, order, hierarchical, xp)
        phi, dphi = sp.shape(xp)
        assert len(phi_ref) == sp.local_size
        assert np.abs(phi - phi_ref).max() < 1e-12 * max(1.0, np.abs(phi_ref).max())
        assert np.abs(dphi[:, :dim] - dphi_ref).max() < 1e-11 * max(1.0, np.abs(dphi_ref).max())


@pytest.mark.parametrize("dim", [1, 2, 3])
def test_cube_quadratures_are_the_reference_rules(dim):
    """CubeQuadrature compiled from the reference (quadrature/femquadratures_inline.hh:33-95): rule selection (smallest Gauss rule with
    order >= requested), tensor construction with x0 fastest, weights = products -- the oracle's cubeQuadrature point for point"""
    for order in range(0, 14):
        x_ref, w_ref, exact = rl.cube_quadrature(dim, order)
        x, w = ol.quadrature(dim, order)
        assert len(w) == len(w_ref) and exact >= order
        assert np.abs(x[:, :dim] - x_ref).max() == 0.0 and np.abs(w - w_ref).max() <= 4e-16      # same table entries; weights up to the product order
