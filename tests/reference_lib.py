"""ctypes binding of oracle/_ref/libdunefem_ref.so: the pieces of the reference (DUNE-FEM) that compile from their own source
files where they lie under /root/reference (oracle/ref_bind.cpp, oracle/Makefile target `ref`).  Test infrastructure only.

The library is built in the development container (where /root/reference exists) and travels to the GPU box as a built file;
`available()` is False where neither the library nor the reference tree is present."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_SO = os.path.join(_ORACLE_DIR, "_ref", "libdunefem_ref.so")
REFERENCE = os.environ.get("B200FEM_REFERENCE", "/root/reference")
_LIB = None

APPLY_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def available():
    return os.path.exists(_SO) or os.path.isdir(os.path.join(REFERENCE, "dune", "fem"))


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    src = os.path.join(_ORACLE_DIR, "ref_bind.cpp")
    if os.path.isdir(os.path.join(REFERENCE, "dune", "fem")) and (not os.path.exists(_SO) or os.path.getmtime(src) > os.path.getmtime(_SO)):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s", "ref", f"REFERENCE={REFERENCE}"])
    L = C.CDLL(_SO)
    L.ref_cg.restype = C.c_int
    L.ref_cg.argtypes = [APPLY_FN, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _lp, C.c_int64, _dp, _dp, C.c_double, C.c_int, C.c_int,
                         _dp, C.c_int, C.POINTER(C.c_int)]
    L.ref_bicgstab.restype = C.c_int
    L.ref_bicgstab.argtypes = [APPLY_FN, C.c_void_p, C.c_int64, _lp, C.c_int64, _dp, _dp, C.c_double, C.c_int, C.c_int, _dp, C.c_int, C.POINTER(C.c_int)]
    L.ref_gmres.restype = C.c_int
    L.ref_gmres.argtypes = [APPLY_FN, C.c_void_p, C.c_int64, _lp, C.c_int64, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_int, _dp, C.c_int, C.POINTER(C.c_int)]
    L.ref_gauss_maxp.restype = C.c_int
    L.ref_gauss_rule.restype = C.c_int
    L.ref_gauss_rule.argtypes = [C.c_int, _dp, _dp]
    L.ref_legendre_max_order.restype = C.c_int
    L.ref_legendre.restype = C.c_double
    L.ref_legendre.argtypes = [C.c_int, C.c_double, C.c_int]
    L.ref_onb_cube.restype = C.c_double
    L.ref_onb_cube.argtypes = [C.c_int, C.c_int, _dp, C.c_void_p]
    _ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
    L.ref_lagrange_cube_points.restype = C.c_int
    L.ref_lagrange_cube_points.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_cube_quadrature.restype = C.c_int
    L.ref_cube_quadrature.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
    L.ref_legendre_set.restype = C.c_int
    L.ref_legendre_set.argtypes = [C.c_int, C.c_int, C.c_int, _dp, C.c_void_p, C.c_void_p]
    L.ref_lagrange_cube_evaluate.restype = C.c_int
    L.ref_lagrange_cube_evaluate.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp, _dp]
    L.ref_difference_quotient.restype = C.c_int
    L.ref_difference_quotient.argtypes = [APPLY_FN, C.c_void_p, C.c_int64, _lp, C.c_int64, _dp, C.c_double, C.c_int, _dp, C.c_int, _dp]
    L.ref_legacy_cg.restype = C.c_int
    L.ref_legacy_cg.argtypes = [APPLY_FN, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, _lp, C.c_int64, _dp, _dp, C.c_double, C.c_int, C.c_int]
    L.ref_newton.restype = C.c_int
    L.ref_newton.argtypes = [APPLY_FN, C.c_void_p, C.c_int, C.c_int64, _lp, C.c_int64, C.c_void_p, _dp, C.c_char_p, np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS"),
                             C.POINTER(C.c_double)]
    _LIB = L
    return L


def _wrap(apply, n):
    def cb(pu, pw, _ctx):
        u = np.ctypeslib.as_array(pu, shape=(n,))
        w = np.ctypeslib.as_array(pw, shape=(n,))
        w[:] = apply(u)
    return APPLY_FN(cb)


def _aux(aux):
    return np.ascontiguousarray([] if aux is None else aux, dtype=np.int64)


def cg(apply, b, x0, eps, maxit, tolcrit=0, precon=None, aux=None):
    """Dune::Fem::LinearSolver::cg on python callables (apply: u -> A u; precon: r -> B r or None)."""
    n = len(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    hist, nh = np.zeros(max(maxit, 1)), C.c_int(0)
    f = _wrap(apply, n)
    p = _wrap(precon, n) if precon is not None else None
    a = _aux(aux)
    it = lib().ref_cg(f, None, C.cast(p, C.c_void_p) if p is not None else None, None, n, a, len(a), x, np.ascontiguousarray(b, dtype=np.float64),
                      eps, maxit, tolcrit, hist, len(hist), C.byref(nh))
    return it, x, hist[:nh.value]


def bicgstab(apply, b, x0, tol, maxit, tolcrit=0, aux=None):
    n = len(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    hist, nh = np.zeros(max(maxit, 1)), C.c_int(0)
    a = _aux(aux)
    it = lib().ref_bicgstab(_wrap(apply, n), None, n, a, len(a), x, np.ascontiguousarray(b, dtype=np.float64), tol, maxit, tolcrit, hist, len(hist), C.byref(nh))
    return it, x, hist[:nh.value]


def gmres(apply, b, x0, tol, maxit, tolcrit=0, restart=20, aux=None):
    n = len(b)
    x = np.array(x0, dtype=np.float64, copy=True)
    hist, nh = np.zeros(max(maxit, 1) + restart), C.c_int(0)
    a = _aux(aux)
    it = lib().ref_gmres(_wrap(apply, n), None, n, a, len(a), x, np.ascontiguousarray(b, dtype=np.float64), restart, tol, maxit, tolcrit, hist, len(hist), C.byref(nh))
    return it, x, hist[:nh.value]


def legacy_cg(apply, b, x0, eps, maxit, error_measure=0, precon=None, aux=None):
    """Dune::Fem::ConjugateGradientSolver::solve (solver/cginverseoperator.hh, the class behind the legacy CGInverseOperator);
    error_measure 0 absolute, 1 relative to |b|.  Returns (iterations, x)."""
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.array(x0, dtype=np.float64, copy=True)
    n = len(b)
    a = _aux(aux)
    fn = _wrap(apply, n)
    pfn = _wrap(precon, n) if precon is not None else None
    it = lib().ref_legacy_cg(fn, None, C.cast(pfn, C.c_void_p) if pfn is not None else None, None, n, a, len(a), x, b, float(eps), int(maxit), int(error_measure))
    return it, x


def difference_quotient(apply, u, args, eps=0.0, from_parameter=False, aux=None):
    """Dune::Fem::AutomaticDifferenceOperator::jacobian(u, jOp), then jOp(arg) for every row of args (python callable apply: u -> L[u]).
    from_parameter: eps comes from the parameter file ("fem.differenceoperator.eps", absent -> 0 = chosen per argument)."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    args = np.ascontiguousarray(np.atleast_2d(args), dtype=np.float64)
    n = len(u)
    out = np.zeros_like(args)
    a = _aux(aux)
    fn = _wrap(apply, n)
    lib().ref_difference_quotient(fn, None, n, a, len(a), u, float(eps), int(from_parameter), args, args.shape[0], out)
    return out


def newton(apply, w0, parameters, u=None, nonlinear=True, aux=None):
    """Dune::Fem::NewtonInverseOperator (Jacobian = the reference's difference quotient, linear solves = the reference's Krylov loops)
    on a python callable, configured through the reference's parameter KEYS (dict, e.g. {"fem.solver.nonlinear.tolerance": 1e-7}).
    Returns (iterations, linearIterations, failure code, |residual|, w)."""
    w = np.array(w0, dtype=np.float64, copy=True)
    n = len(w)
    a = _aux(aux)
    fn = _wrap(apply, n)
    text = "\n".join(f"{k}: {str(v).lower() if isinstance(v, bool) else repr(float(v)) if isinstance(v, float) else v}" for k, v in parameters.items())
    out = np.zeros(3, dtype=np.int32)
    res = C.c_double()
    uu = None if u is None else np.ascontiguousarray(u, dtype=np.float64)
    rc = lib().ref_newton(fn, None, int(nonlinear), n, a, len(a), None if uu is None else uu.ctypes.data_as(C.c_void_p), w, text.encode(), out, C.byref(res))
    if rc != 0:
        raise RuntimeError("the reference's Newton solver raised an exception (see stderr)")
    return int(out[0]), int(out[1]), int(out[2]), res.value, w


def gauss_rule(m):
    x, w = np.empty(m), np.empty(m)
    order = lib().ref_gauss_rule(m, x, w)
    return x, w, order


def legendre(num, x, deriv=0):
    return lib().ref_legendre(num, float(x), deriv)


def onb_cube(dim, i, x, grad=False):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if not grad:
        return lib().ref_onb_cube(dim, i, x, None)
    g = np.zeros(3)
    v = lib().ref_onb_cube(dim, i, x, g.ctypes.data_as(C.c_void_p))
    return v, g[:dim]


def lagrange_cube_points(dim, order):
    """GenericLagrangePoint of the dim-cube (space/lagrange/genericlagrangepoints.hh): per local node its reference coordinates, the
    codimension and number of the sub-entity it lies in and its number inside that sub-entity"""
    n = lib().ref_lagrange_cube_points(dim, order, None, None, None, None)
    assert n == (order + 1) ** dim
    x = np.empty((n, dim))
    codim, sub, num = (np.empty(n, dtype=np.int32) for _ in range(3))
    lib().ref_lagrange_cube_points(dim, order, x.ctypes.data_as(C.c_void_p), codim.ctypes.data_as(C.c_void_p), sub.ctypes.data_as(C.c_void_p), num.ctypes.data_as(C.c_void_p))
    return x, codim, sub, num


def lagrange_cube_evaluate(dim, order, base, x):
    """GenericLagrangeBaseFunction::evaluate (space/lagrange/genericbasefunctions.hh): value and reference gradient"""
    phi, dphi = np.empty(1), np.empty(dim)
    rc = lib().ref_lagrange_cube_evaluate(dim, order, base, np.ascontiguousarray(x, dtype=np.float64), phi, dphi)
    assert rc == 0
    return phi[0], dphi


def legendre_set(dim, order, hierarchical, x):
    """LegendreShapeFunctionSet< FunctionSpace, hierarchical >( order ) (space/shapefunctionset/legendre.hh): values and reference
    gradients of all shape functions at x, in the set's own order"""
    x = np.ascontiguousarray(x, dtype=np.float64)
    n = lib().ref_legendre_set(dim, order, int(hierarchical), x, None, None)
    phi, dphi = np.empty(n), np.empty((n, dim))
    lib().ref_legendre_set(dim, order, int(hierarchical), x, phi.ctypes.data_as(C.c_void_p), dphi.ctypes.data_as(C.c_void_p))
    return phi, dphi


def cube_quadrature(dim, order):
    """CubeQuadrature< double, dim > of the reference: (points [n][dim], weights [n], order of the selected rule)"""
    exact = C.c_int()
    n = lib().ref_cube_quadrature(dim, order, None, None, C.byref(exact))
    x, w = np.empty((n, dim)), np.empty(n)
    lib().ref_cube_quadrature(dim, order, x.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p), C.byref(exact))
    return x, w, exact.value
