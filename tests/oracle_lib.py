"""ctypes binding of the CPU oracle (oracle/libfem_oracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never imported by the product.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB = None

LAGRANGE, DG_LEGENDRE, DG_LEGENDRE_HIER, DG_ONB = 0, 1, 2, 3
NUMBERING_YASP, NUMBERING_ADAPTIVE_LEAF = 0, 1

_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_lp = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_bp = np.ctypeslib.ndpointer(dtype=np.uint8, flags="C_CONTIGUOUS")


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    so = os.path.join(_ORACLE_DIR, "libfem_oracle.so")
    src = os.path.join(_ORACLE_DIR, "fem_oracle.cpp")
    native = os.environ.get("B200FEM_ORACLE_SO")      # bench.py's CPU legs: the same source built -march=native on the timed host
    if native and os.path.exists(native) and os.path.getmtime(native) >= os.path.getmtime(src):
        so = native
    elif not os.path.exists(so) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _ORACLE_DIR, "-s"])
    L = C.CDLL(so)
    L.fo_space_create.restype = C.c_void_p
    L.fo_space_create.argtypes = [C.c_int, _ip, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    L.fo_space_destroy.argtypes = [C.c_void_p]
    L.fo_space_set_periodic.argtypes = [C.c_void_p, C.c_int]
    L.fo_space_size.restype = C.c_int64
    L.fo_space_size.argtypes = [C.c_void_p]
    L.fo_space_local_size.restype = C.c_int
    L.fo_space_local_size.argtypes = [C.c_void_p]
    L.fo_space_elements.restype = C.c_int64
    L.fo_space_elements.argtypes = [C.c_void_p]
    L.fo_space_dofmap.argtypes = [C.c_void_p, C.c_int64, _lp]
    L.fo_space_multiindex.argtypes = [C.c_void_p, _ip]
    L.fo_quadrature.restype = C.c_int
    L.fo_quadrature.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    L.fo_legendre.restype = C.c_double
    L.fo_legendre.argtypes = [C.c_int, C.c_double, C.c_int]
    L.fo_shape_evaluate.argtypes = [C.c_void_p, _dp, _dp, _dp]
    L.fo_operator_create.restype = C.c_void_p
    L.fo_operator_create.argtypes = [C.c_void_p, _dp, _ip]
    L.fo_operator_destroy.argtypes = [C.c_void_p]
    L.fo_operator_create_user.restype = C.c_void_p
    L.fo_operator_create_user.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, _dp, C.c_int]
    L.fo_vector_operator_create.restype = C.c_void_p
    L.fo_vector_operator_create.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, _dp, C.c_int]
    L.fo_vector_operator_destroy.argtypes = [C.c_void_p]
    L.fo_vector_operator_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    L.fo_unstructured_create.restype = C.c_void_p
    L.fo_unstructured_create.argtypes = [C.c_int, C.c_int64, _dp, C.c_int64, _lp, C.c_int, _dp, _ip]
    L.fo_unstructured_destroy.argtypes = [C.c_void_p]
    L.fo_unstructured_set_user.argtypes = [C.c_void_p, C.c_void_p, _dp, C.c_int]
    L.fo_unstructured_size.restype = C.c_int64
    L.fo_unstructured_size.argtypes = [C.c_void_p]
    L.fo_unstructured_local_size.restype = C.c_int
    L.fo_unstructured_local_size.argtypes = [C.c_void_p]
    L.fo_unstructured_dofmap.argtypes = [C.c_void_p, C.c_int64, _lp]
    L.fo_unstructured_nodes.argtypes = [C.c_void_p, _dp, _bp]
    L.fo_unstructured_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    L.fo_operator_set_threads.argtypes = [C.c_void_p, C.c_int]
    L.fo_operator_set_inverse_mass.argtypes = [C.c_void_p, C.c_int]
    L.fo_operator_set_inverse_mass.restype = C.c_int
    L.fo_operator_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int]
    L.fo_operator_apply_box.argtypes = [C.c_void_p, _dp, _dp, C.c_int, _ip, _ip]
    L.fo_dirichlet.argtypes = [C.c_void_p, _bp, _dp]
    L.fo_cg.restype = C.c_int
    L.fo_cg.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.fo_operator_linearize.argtypes = [C.c_void_p, _dp, C.c_double]
    L.fo_operator_apply_jacobian.argtypes = [C.c_void_p, _dp, _dp]
    L.fo_operator_apply_jacobian.restype = C.c_double
    L.fo_gmres_jacobian.restype = C.c_int
    L.fo_gmres_jacobian.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.fo_operator_diagonal.argtypes = [C.c_void_p, _dp]
    L.fo_pcg_diagonal.restype = C.c_int
    L.fo_pcg_diagonal.argtypes = [C.c_void_p, _dp, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.fo_kron_apply.argtypes = [_ip, C.c_int, _ip, _dp, C.c_double, _dp, _dp, C.c_void_p, C.c_int]
    L.fo_kron_apply.restype = None
    L.fo_gmres.restype = C.c_int
    L.fo_gmres.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.fo_bicgstab.restype = C.c_int
    L.fo_bicgstab.argtypes = [C.c_void_p, _dp, _dp, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.fo_dot.restype = C.c_double
    L.fo_dot.argtypes = [_dp, _dp, C.c_int64]
    L.fo_interpolate.argtypes = [C.c_void_p, C.c_int, _dp]
    L.fo_l2error.restype = C.c_double
    L.fo_l2error.argtypes = [C.c_void_p, _dp, C.c_int]
    L.fo_assemble_dense.argtypes = [C.c_void_p, _dp]
    L.fo_time_apply.restype = C.c_double
    L.fo_time_apply.argtypes = [C.c_void_p, _dp, _dp, C.c_int, C.c_int]
    _LIB = L
    return L


class Space:
    def __init__(self, n, lo, hi, kind, order, numbering=NUMBERING_YASP, interior_order=0, surface_order=0):
        self.dim = len(n)
        self.n = np.ascontiguousarray(n, dtype=np.int32)
        self.lo = np.ascontiguousarray(lo, dtype=np.float64)
        self.hi = np.ascontiguousarray(hi, dtype=np.float64)
        self.kind, self.order = kind, order
        self._h = lib().fo_space_create(self.dim, self.n, self.lo, self.hi, kind, order, numbering,
                                        interior_order, surface_order)
        self.size = lib().fo_space_size(self._h)
        self.local_size = lib().fo_space_local_size(self._h)
        self.elements = lib().fo_space_elements(self._h)

    def set_periodic(self, mask):
        """bit d: periodic along axis d (DG spaces)"""
        lib().fo_space_set_periodic(self._h, mask)

    def dofmap(self, e):
        out = np.empty(self.local_size, dtype=np.int64)
        lib().fo_space_dofmap(self._h, e, out)
        return out

    def multiindex(self):
        out = np.empty((self.local_size, 3), dtype=np.int32)
        lib().fo_space_multiindex(self._h, out)
        return out

    def shape(self, x):
        x3 = np.zeros(3)
        x3[:len(x)] = x
        phi = np.empty(self.local_size)
        dphi = np.empty((self.local_size, 3))
        lib().fo_shape_evaluate(self._h, x3, phi, dphi)
        return phi, dphi

    def node_positions(self):
        """positions of the Lagrange nodes, indexed by global dof (fem_oracle.cpp: Space::nodePosition)"""
        assert self.kind == LAGRANGE
        mi, h = self.multiindex()[:, :self.dim], (self.hi - self.lo) / self.n
        x = np.zeros((self.size, self.dim))
        for e in range(self.elements):
            ec, r = [], e
            for d in range(self.dim):
                ec.append(r % self.n[d])
                r //= self.n[d]
            x[self.dofmap(e)] = self.lo + h * (np.array(ec) + mi / self.order)
        return x

    def interpolate(self, data):
        out = np.zeros(self.size)
        lib().fo_interpolate(self._h, data, out)
        return out

    def l2error(self, u, data):
        return lib().fo_l2error(self._h, np.ascontiguousarray(u), data)

    def __del__(self):
        try:
            lib().fo_space_destroy(self._h)
        except Exception:
            pass


class Operator:
    """ADR integrands (see fem_oracle.cpp): eps, b, c, gamma, beta + flags."""

    def __init__(self, space, eps=1.0, b=(0.0, 0.0, 0.0), c=0.0, gamma=0.0, beta=0.0, dirichlet_mask=0, data=0,
                 skeleton=False, boundary=False, strong_dirichlet=False, threads=1):
        self.space = space
        bb = list(b) + [0.0] * (3 - len(b))
        params = np.array([eps, bb[0], bb[1], bb[2], c, gamma, beta], dtype=np.float64)
        iparams = np.array([dirichlet_mask, data, int(skeleton), int(boundary), int(strong_dirichlet)], dtype=np.int32)
        self._h = lib().fo_operator_create(space._h, params, iparams)
        if threads > 1:
            lib().fo_operator_set_threads(self._h, threads)

    def setInverseMass(self, on=True):
        """MOLGalerkinOperator: w = M^-1 L[u] (schemes/molgalerkin.hh)"""
        if lib().fo_operator_set_inverse_mass(self._h, int(on)) != 0:
            raise ValueError("inverse mass: DG spaces only")

    def apply(self, u, linear=False):
        w = np.empty(self.space.size)
        lib().fo_operator_apply(self._h, np.ascontiguousarray(u, dtype=np.float64), w, int(linear))
        return w

    def apply_box(self, u, lo, hi, linear=False):
        w = np.empty(self.space.size)
        lo3 = np.array(list(lo) + [0] * (3 - len(lo)), dtype=np.int32)
        hi3 = np.array(list(hi) + [1] * (3 - len(hi)), dtype=np.int32)
        lib().fo_operator_apply_box(self._h, np.ascontiguousarray(u, dtype=np.float64), w, int(linear), lo3, hi3)
        return w

    def dirichlet(self):
        mask = np.zeros(self.space.size, dtype=np.uint8)
        vals = np.zeros(self.space.size)
        lib().fo_dirichlet(self._h, mask, vals)
        return mask, vals

    def cg(self, b, x0, eps, maxit, tolcrit=0):
        x = np.array(x0, dtype=np.float64, copy=True)
        hist = np.zeros(max(maxit, 1))
        it = lib().fo_cg(self._h, np.ascontiguousarray(b, dtype=np.float64), x, eps, maxit, tolcrit,
                         hist.ctypes.data_as(C.c_void_p))
        return it, x, hist[:abs(it)]

    def diagonal(self):
        d = np.empty(self.space.size)
        lib().fo_operator_diagonal(self._h, d)
        return d

    def pcg(self, diag, b, x0, eps, maxit, tolcrit=0):
        x = np.array(x0, dtype=np.float64, copy=True)
        hist = np.zeros(max(maxit, 1))
        it = lib().fo_pcg_diagonal(self._h, np.ascontiguousarray(diag, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64), x, eps, maxit,
                                   tolcrit, hist.ctypes.data_as(C.c_void_p))
        return it, x, hist[:abs(it)]

    def linearize(self, u, eps=0.0):
        """AutomaticDifferenceLinearOperator::set(u, op, eps) (operator/common/automaticdifferenceoperator.hh:152-166)"""
        lib().fo_operator_linearize(self._h, np.ascontiguousarray(u, dtype=np.float64), float(eps))

    def applyJacobian(self, v):
        w = np.empty(self.space.size)
        eps = lib().fo_operator_apply_jacobian(self._h, np.ascontiguousarray(v, dtype=np.float64), w)
        return w, eps

    def gmres_jacobian(self, b, x0, eps, maxit, tolcrit=0, restart=20):
        x = np.array(x0, dtype=np.float64, copy=True)
        hist = np.zeros(max(maxit, 1))
        it = lib().fo_gmres_jacobian(self._h, np.ascontiguousarray(b, dtype=np.float64), x, restart, eps, maxit, tolcrit,
                                     hist.ctypes.data_as(C.c_void_p))
        return it, x, hist[:abs(it)]

    def gmres(self, b, x0, eps, maxit, tolcrit=0, restart=20):
        x = np.array(x0, dtype=np.float64, copy=True)
        hist = np.zeros(max(maxit, 1))
        it = lib().fo_gmres(self._h, np.ascontiguousarray(b, dtype=np.float64), x, restart, eps, maxit, tolcrit,
                            hist.ctypes.data_as(C.c_void_p))
        return it, x, hist[:abs(it)]

    def bicgstab(self, b, x0, eps, maxit, tolcrit=0):
        x = np.array(x0, dtype=np.float64, copy=True)
        hist = np.zeros(max(maxit, 1))
        it = lib().fo_bicgstab(self._h, np.ascontiguousarray(b, dtype=np.float64), x, eps, maxit, tolcrit,
                         hist.ctypes.data_as(C.c_void_p))
        return it, x, hist[:abs(it)]

    def assemble_dense(self):
        n = self.space.size
        A = np.empty((n, n))
        lib().fo_assemble_dense(self._h, A)
        return A

    def time_apply(self, u, linear=False, reps=3):
        w = np.empty(self.space.size)
        return lib().fo_time_apply(self._h, np.ascontiguousarray(u, dtype=np.float64), w, int(linear), reps)

    def __del__(self):
        try:
            lib().fo_operator_destroy(self._h)
        except Exception:
            pass


class UserOperator(Operator):
    """GalerkinOperator over user-supplied integrands: `source` is the text handed to b200fem_operator_create_jit, compiled here
    for the HOST (g++, __device__ defined away) into three callbacks the oracle integrates."""
    _PRELUDE = ("#include <cmath>\nusing namespace std;\n#define __device__\n#define __forceinline__ inline\n"
                "struct PointValue { double u; double du[3]; };\nstruct PointRange { double s; double F[3]; };\nnamespace user {\n")
    _EPILOGUE = ("\n}\nextern \"C\" {\n"
                 "void u_interior(const double* x, const PointValue* u, PointRange* r, const double* c, int dim) { user::interior(x, *u, *r, c, dim); }\n"
                 "#ifdef HAS_SKELETON\nvoid u_skeleton(const double* x, int axis, double sign, double ihe, const PointValue* in, const PointValue* out, PointRange* rin, PointRange* rout, const double* c, int dim) { user::skeleton(x, axis, sign, ihe, *in, *out, *rin, *rout, c, dim); }\n#endif\n"
                 "#ifdef HAS_BOUNDARY\nvoid u_boundary(const double* x, int axis, int side, double ihbnd, const PointValue* u, PointRange* r, const double* c, int dim) { user::boundary(x, axis, side, ihbnd, *u, *r, c, dim); }\n#endif\n}\n")

    def __init__(self, space, source, constants=(), skeleton=True, boundary=True, threads=1):
        self.space = space
        flags = (["-DHAS_SKELETON"] if skeleton else []) + (["-DHAS_BOUNDARY"] if boundary else [])
        self._user = _compile_user(self._PRELUDE + source + self._EPILOGUE, flags)
        fn = lambda name, on: C.cast(getattr(self._user, name), C.c_void_p) if on else None
        c = np.zeros(32)
        c[:len(constants)] = constants
        self._h = lib().fo_operator_create_user(space._h, fn("u_interior", True), fn("u_skeleton", skeleton), fn("u_boundary", boundary), c, 32)
        if threads > 1:
            lib().fo_operator_set_threads(self._h, threads)



def _compile_user(text, flags):
    import hashlib
    import tempfile
    tag = hashlib.sha1((text + " ".join(flags)).encode()).hexdigest()[:16]
    d = os.path.join(tempfile.gettempdir(), "b200fem_oracle_user")
    os.makedirs(d, exist_ok=True)
    so = os.path.join(d, f"user_{tag}.so")
    if not os.path.exists(so):
        src = os.path.join(d, f"user_{tag}.{os.getpid()}.cpp")
        with open(src, "w") as f:
            f.write(text)
        tmp = f"{so}.{os.getpid()}.tmp"            # (several ranks may compile the same text at once: private name, atomic rename)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", tmp, src] + flags)
        os.replace(tmp, so)
    return C.CDLL(so)


class VectorUserOperator:
    """GalerkinOperator on a range-R space (dofs = blocks of R components over the scalar space `space`) with user-supplied
    integrands over VectorValue / VectorRange (include/b200fem.h); the SAME source text as the device, compiled for the host."""
    _PRELUDE = ("#include <cmath>\nusing namespace std;\n#define __device__\n#define __forceinline__ inline\n"
                "template <int R> struct PointValueV { double u[R]; double du[R][3]; };\ntemplate <int R> struct PointRangeV { double s[R]; double F[R][3]; };\n"
                "namespace user {\nconstexpr int dimRange = DIM_RANGE;\nusing VectorValue = PointValueV<dimRange>;\nusing VectorRange = PointRangeV<dimRange>;\n")
    _EPILOGUE = ("\n}\nusing user::VectorValue; using user::VectorRange;\nextern \"C\" {\n"
                 "void u_interior(const double* x, const VectorValue* u, VectorRange* r, const double* c, int dim) { user::interior(x, *u, *r, c, dim); }\n"
                 "#ifdef HAS_SKELETON\nvoid u_skeleton(const double* x, int axis, double sign, double ihe, const VectorValue* in, const VectorValue* out, VectorRange* rin, VectorRange* rout, const double* c, int dim) { user::skeleton(x, axis, sign, ihe, *in, *out, *rin, *rout, c, dim); }\n#endif\n"
                 "#ifdef HAS_BOUNDARY\nvoid u_boundary(const double* x, int axis, int side, double ihbnd, const VectorValue* u, VectorRange* r, const double* c, int dim) { user::boundary(x, axis, side, ihbnd, *u, *r, c, dim); }\n#endif\n}\n")

    def __init__(self, space, dim_range, source, constants=(), skeleton=True, boundary=True):
        self.space, self.R, self.size = space, dim_range, space.size * dim_range
        flags = [f"-DDIM_RANGE={dim_range}"] + (["-DHAS_SKELETON"] if skeleton else []) + (["-DHAS_BOUNDARY"] if boundary else [])
        self._user = _compile_user(self._PRELUDE + source + self._EPILOGUE, flags)
        fn = lambda name, on: C.cast(getattr(self._user, name), C.c_void_p) if on else None
        c = np.zeros(32)
        c[:len(constants)] = constants
        self._h = lib().fo_vector_operator_create(space._h, dim_range, fn("u_interior", True), fn("u_skeleton", skeleton), fn("u_boundary", boundary), c, 32)

    def apply(self, u, linear=False):
        w = np.empty(self.size)
        lib().fo_vector_operator_apply(self._h, np.ascontiguousarray(u, dtype=np.float64), w, int(linear))
        return w

    def __del__(self):
        try:
            lib().fo_vector_operator_destroy(self._h)
        except Exception:
            pass



def newton(oop, w0, tol, maxit, lin_tol, lin_maxit, restart=20, tolcrit=0, line_search=False, u=None, nonlinear=True, trace=None):
    """NewtonInverseOperator::operator() and ::lineSearch (solver/newtoninverseoperator.hh:690-803, 588-629) restated on the oracle's
    pieces: difference-quotient Jacobian (Operator.linearize), GMRES on it, forcing "none".  Pinned against the reference's own class in
    tests/test_reference_pieces.py.  Returns (iterations, linearIterations, NewtonFailure code, |residual|, w);
    trace (a list) receives the number of halvings the line search took in every step."""
    big = np.finfo(np.float64).max

    def failed(delta, it, lit, completed):                     # :568-584
        if not (delta < big) or np.isnan(delta):
            return 1
        if it >= maxit:
            return 5
        if lit >= lin_maxit:
            return 6
        if lit < 0:
            return 7
        return 0 if completed else 4

    def residual(w):
        r = oop.apply(w)
        return r if u is None else r - u

    w = np.array(w0, dtype=np.float64, copy=True)
    res = residual(w)
    delta = np.sqrt(res @ res)
    it = lit = 0
    completed = True
    while True:
        oop.linearize(w)
        if lin_maxit - lit <= 0:
            break
        li, dw, _ = oop.gmres_jacobian(res, np.zeros_like(w), lin_tol, lin_maxit - lit, tolcrit, restart)
        if li < 0:
            lit = li
            break
        lit += li
        w -= dw
        if not nonlinear:
            break
        res = residual(w)
        delta_old, delta, ls, halvings = delta, np.sqrt(res @ res), 0, 0
        if line_search:
            if failed(delta, it, lit, completed) == 1:
                test = dw @ dw
                if not (test < big and not np.isnan(test)):
                    delta = 2.0 * delta_old
            factor, ls = 1.0, (1 if delta < delta_old else 0)
            while delta >= delta_old:
                delta_prev = delta
                factor *= 0.5
                if abs(delta - delta_old) < 1e-5 * delta:
                    ls = -1
                    break
                w += factor * dw
                res = residual(w)
                delta = np.sqrt(res @ res)
                halvings += 1
                if abs(delta - delta_prev) < 1e-15:
                    ls = -1
                    break
                if failed(delta, it, lit, completed) == 1:
                    delta = 2.0 * delta_old
        if trace is not None:
            trace.append(halvings)
        completed = ls >= 0
        it += 1
        if delta < tol or failed(delta, it, lit, completed) != 0:
            break
    return it, lit, failed(delta, it, lit, completed), delta, w


def cartesian_as_unstructured(n, lo, hi):
    """vertex coordinates and element -> vertex arrays (cube reference order, elements x fastest) of a Cartesian mesh"""
    dim = len(n)
    nv = [k + 1 for k in n]
    axes = [np.linspace(lo[d], hi[d], nv[d]) for d in range(dim)]
    grids = np.meshgrid(*axes, indexing="ij")
    vid = np.arange(int(np.prod(nv))).reshape(nv[::-1]).transpose(range(dim - 1, -1, -1))      # vid[i0, i1(, i2)], x fastest
    coords = np.zeros((int(np.prod(nv)), dim))
    for d in range(dim):
        coords[vid.ravel(), d] = grids[d].ravel()
    elems = []
    rng = [range(k) for k in n]
    import itertools
    for idx in itertools.product(*rng[::-1]):                 # last axis slowest
        ec = idx[::-1]
        elems.append([vid[tuple(ec[d] + ((v >> d) & 1) for d in range(dim))] for v in range(1 << dim)])
    return coords, np.array(elems, dtype=np.int64)


class UnstructuredOperator:
    """Lagrange space + ADR operator on an unstructured cube mesh (fem_oracle.cpp: UnstructuredLagrange)"""

    def __init__(self, coords, elems, order, eps=1.0, b=(0.0, 0.0, 0.0), c=0.0, gamma=0.0, data=0, strong_dirichlet=False,
                 user_source=None, constants=()):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        elems = np.ascontiguousarray(elems, dtype=np.int64)
        self.dim = coords.shape[1]
        bb = list(b) + [0.0] * (3 - len(b))
        params = np.array([eps, bb[0], bb[1], bb[2], c, gamma, 0.0], dtype=np.float64)
        iparams = np.array([0, data, 0, 0, int(strong_dirichlet)], dtype=np.int32)
        self._h = lib().fo_unstructured_create(self.dim, coords.shape[0], coords.ravel(), elems.shape[0], elems.ravel(), order, params, iparams)
        self.size = lib().fo_unstructured_size(self._h)
        self.local_size = lib().fo_unstructured_local_size(self._h)
        self.elements = elems.shape[0]
        if user_source is not None:          # the interior() of a run-time compiled integrands source, compiled for the host
            self._user = _compile_user(UserOperator._PRELUDE + user_source + UserOperator._EPILOGUE, [])
            cc = np.zeros(32)
            cc[:len(constants)] = constants
            lib().fo_unstructured_set_user(self._h, C.cast(self._user.u_interior, C.c_void_p), cc, 32)

    def dofmap(self, e):
        out = np.empty(self.local_size, dtype=np.int64)
        lib().fo_unstructured_dofmap(self._h, e, out)
        return out

    def nodes(self):
        x = np.empty((self.size, 3))
        bnd = np.empty(self.size, dtype=np.uint8)
        lib().fo_unstructured_nodes(self._h, x.reshape(-1), bnd)
        return x[:, :self.dim], bnd

    def apply(self, u, linear=False):
        w = np.empty(self.size)
        lib().fo_unstructured_apply(self._h, np.ascontiguousarray(u, dtype=np.float64), w, int(linear))
        return w

    def __del__(self):
        try:
            lib().fo_unstructured_destroy(self._h)
        except Exception:
            pass


def quadrature(dim, order):
    n = lib().fo_quadrature(dim, order, None, None)
    x = np.empty((n, 3))
    w = np.empty(n)
    lib().fo_quadrature(dim, order, x.ctypes.data_as(C.c_void_p), w.ctypes.data_as(C.c_void_p))
    return x, w


class KroneckerCpu:
    """Kronecker-form CPU apply of a linear, constant-coefficient DG operator on a uniform box (fem_oracle.cpp: fo_kron_apply).
    The 1-D matrices are obtained by PROBING the dense oracle operator on a 3x3x3 mesh of the same cell size: unit vectors whose
    tensor mode varies along one axis only, read back on the same modes of the centre / boundary elements."""

    def __init__(self, n, lo, hi, kind, order, threads=1, **model):
        assert kind in (DG_LEGENDRE, DG_LEGENDRE_HIER) and len(n) == 3
        self.n, self.N, self.threads = list(n), order + 1, threads
        N, nb = self.N, (order + 1) ** 3
        h = [(hi[d] - lo[d]) / n[d] for d in range(3)]
        self.space = Space(n, lo, hi, kind, order)
        probe = Space([3, 3, 3], lo, [lo[d] + 3 * h[d] for d in range(3)], kind, order)
        op = Operator(probe, skeleton=True, boundary=True, **dict(model, data=0))
        mi = np.zeros(3 * nb, dtype=np.int32)
        lib().fo_space_multiindex(probe._h, mi)
        mi = mi.reshape(nb, 3)
        self.tensor_of_stored = np.ascontiguousarray((mi[:, 0] * N + mi[:, 1]) * N + mi[:, 2], dtype=np.int32)
        stored_of = {tuple(m): l for l, m in enumerate(mi.tolist())}

        def mode(d, j):                     # stored local index of the tensor mode that is j along axis d and 0 elsewhere
            m = [0, 0, 0]
            m[d] = j
            return stored_of[tuple(m)]

        def elem(c):
            return c[0] + 3 * (c[1] + 3 * c[2])

        def block(d, src, dst):             # response matrix [i][j] of element dst (modes along d) to unit vectors in element src
            A = np.zeros((N, N))
            for j in range(N):
                e = np.zeros(probe.size)
                e[elem(src) * nb + mode(d, j)] = 1.0
                w = op.apply(e, linear=True)
                for i in range(N):
                    A[i, j] = w[elem(dst) * nb + mode(d, i)]
            return A

        mats = np.zeros((3, 5, N, N))
        for d in range(3):
            c, lo_e, hi_e = [1, 1, 1], [1, 1, 1], [1, 1, 1]
            lo_e[d], hi_e[d] = 0, 2
            S = block(d, c, c)
            mats[d, 0] = S
            mats[d, 1] = block(d, lo_e, c)                     # L_d: u_{K-e_d} -> w_K
            mats[d, 2] = block(d, hi_e, c)                     # R_d
            mats[d, 3] = block(d, lo_e, lo_e) - S              # Dlo_d
            mats[d, 4] = block(d, hi_e, hi_e) - S              # Dhi_d
        # every S above carries the (0,0) entries of the two other axes on its diagonal: S_d(probed) = S_d + (C - c_d) I, C = sum_d c_d,
        # so  sum_d S_d(probed) = sum_d S_d + 2 C I  and  C = S(probed)[0][0] for any axis
        self.c2 = 2.0 * mats[0, 0, 0, 0]
        self.mats = np.ascontiguousarray(mats)

    def apply(self, u, bvec=None, out=None):
        w = np.empty(self.space.size) if out is None else out
        n3 = np.array(self.n, dtype=np.int32)
        bp = None if bvec is None else np.ascontiguousarray(bvec, dtype=np.float64).ctypes.data_as(C.c_void_p)
        lib().fo_kron_apply(n3, self.N, self.tensor_of_stored, self.mats.ravel(), self.c2, np.ascontiguousarray(u, dtype=np.float64), w, bp, self.threads)
        return w
