"""Galerkin operators.  Mirrors dune.fem.operator.galerkin (python/dune/fem/operator/__init__.py:77-207) and the C++
Dune::Fem::GalerkinOperator interface it registers (dune/fempy/py/operator.hh:207-280): __call__(u, w),
setCommunicate, setQuadratureOrders, plus the affine shift b = -L[0] (:347-350)."""
import ctypes as C

import numpy as np

from . import _capi as capi


class GalerkinOperator:
    def __init__(self, space, eps=1.0, b=(0.0, 0.0, 0.0), c=0.0, gamma=0.0, beta=None, dirichlet_mask=0, data=0,
                 skeleton=None, boundary=None, strong_dirichlet=False, kernel=capi.KERNEL_AUTO):
        self.space = space
        is_dg = space.kind != capi.LAGRANGE
        m = capi.Model()
        m.eps = eps
        bb = list(b) + [0.0] * (3 - len(b))
        m.b[0], m.b[1], m.b[2] = bb
        m.c, m.gamma = c, gamma
        m.beta = 20.0 * space.order ** 2 if beta is None else beta     # pydemo/advectiondiffusion.py:41
        m.dirichlet_mask, m.data = dirichlet_mask, data
        m.has_skeleton = int(is_dg if skeleton is None else skeleton)
        m.has_boundary = int(is_dg if boundary is None else boundary)
        m.strong_dirichlet = int(strong_dirichlet)
        self.model = m
        self.handle = C.c_void_p()
        capi.check(capi.lib().b200fem_operator_create(space.handle, C.byref(m), C.byref(self.handle)))
        self._apply_dev = capi.lib().b200fem_operator_apply_dev          # bound once: the call sits in 40 us loops
        if kernel != capi.KERNEL_AUTO:
            self.setKernel(kernel)

    # --- Dune::Fem::Operator interface (operator/common/operator.hh:55) ---
    def __call__(self, u, w):
        capi.check(capi.lib().b200fem_operator_apply(self.handle, capi.ptr(u), capi.ptr(w)))

    def applyLinear(self, u, w):
        capi.check(capi.lib().b200fem_operator_apply_linear(self.handle, capi.ptr(u), capi.ptr(w)))

    def apply_dev(self, u_ptr, w_ptr, linear=False):
        rc = self._apply_dev(self.handle, u_ptr, w_ptr, 1 if linear else 0)
        if rc:
            capi.check(rc)

    def loadVector(self):
        bvec = np.empty(self.space.size)
        capi.check(capi.lib().b200fem_operator_load_vector(self.handle, capi.ptr(bvec)))
        return bvec

    def setCommunicate(self, communicate):
        capi.check(capi.lib().b200fem_operator_set_communicate(self.handle, int(communicate)))

    def setQuadratureOrders(self, interior, surface):
        capi.check(capi.lib().b200fem_operator_set_quadrature_orders(self.handle, interior, surface))

    def setKernel(self, kernel):
        capi.check(capi.lib().b200fem_operator_set_kernel(self.handle, kernel))

    def setHostPipeline(self, chunks):
        """z-slabs of the H2D / compute / D2H pipeline of __call__ on DG spaces (0: one copy each way)"""
        capi.check(capi.lib().b200fem_operator_set_host_pipeline(self.handle, int(chunks)))

    def dirichlet(self):
        mask = np.zeros(self.space.size, dtype=np.uint8)
        vals = np.zeros(self.space.size)
        capi.check(capi.lib().b200fem_operator_dirichlet(self.handle, capi.ptr(mask, np.uint8), capi.ptr(vals)))
        return mask, vals

    def timing(self):
        t = capi.Timing()
        capi.check(capi.lib().b200fem_operator_timing(self.handle, C.byref(t)))
        return {"last_apply_ms": t.last_apply_ms, "last_exchange_ms": t.last_exchange_ms, "applies": t.applies,
                "kernel": t.kernel, "launches_per_apply": t.launches_per_apply}

    def dot_dev(self, x_ptr, y_ptr):
        r = C.c_double()
        capi.check(capi.lib().b200fem_dot_dev(self.handle, C.c_void_p(x_ptr), C.c_void_p(y_ptr), C.byref(r)))
        return r.value

    def diagonal(self):
        """diag(A) of the homogeneous linear part (what DiagonalPreconditioner needs, solver/diagonalpreconditioner.hh)"""
        d = np.empty(self.space.size)
        capi.check(capi.lib().b200fem_operator_diagonal(self.handle, capi.ptr(d)))
        return d

    def setInverseMass(self, on=True):
        capi.check(capi.lib().b200fem_operator_set_inverse_mass(self.handle, int(on)))

    # --- AutomaticDifferenceOperator::jacobian(u, jOp) (operator/common/automaticdifferenceoperator.hh:108-111) ---
    def linearize(self, u=None, eps=0.0):
        """jOp.set(u, op, eps): afterwards applyLinear and the Krylov solvers act on the difference-quotient Jacobian
        J(u) v = (L[u + eps v] - L[u]) / eps (dynamic eps when eps <= 0); linearize(None) drops the linearisation."""
        capi.check(capi.lib().b200fem_operator_linearize(self.handle, capi.ptr(u) if u is not None else None, float(eps)))

    def communicate_dev(self, v_ptr):
        capi.check(capi.lib().b200fem_communicate_dev(self.handle, C.c_void_p(v_ptr)))

    def close(self):
        if self.handle:
            capi.lib().b200fem_operator_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def nonlinear(self):
        return self.model.gamma != 0.0


class JitGalerkinOperator(GalerkinOperator):
    """GalerkinOperator< Integrands > for user-supplied integrands: the reference generates the Integrands class from UFL and
    compiles the operator for it (python/dune/fem/operator/__init__.py:148-181, models/integrands/model.py); here `source` is
    CUDA C++ defining interior / skeleton / boundary (include/b200fem.h: b200fem_operator_create_jit), compiled by NVRTC into the
    generic quadrature kernel.  `constants` play the role of dune.ufl.Constant coefficients (setConstants: no recompilation)."""

    def __init__(self, space, source, constants=(), skeleton=True, boundary=True):
        self.space = space
        self.model = None
        c = np.ascontiguousarray(constants, dtype=np.float64)
        self.handle = C.c_void_p()
        capi.check(capi.lib().b200fem_operator_create_jit(space.handle, source.encode(), capi.ptr(c) if len(c) else None, len(c),
                                                          int(skeleton), int(boundary), C.byref(self.handle)))
        self._apply_dev = capi.lib().b200fem_operator_apply_dev

    @property
    def nonlinear(self):
        return True                      # (a compiled form is not known to be linear: Operator::nonlinear() defaults to true)

    def setConstants(self, constants):
        c = np.ascontiguousarray(constants, dtype=np.float64)
        capi.check(capi.lib().b200fem_operator_set_constants(self.handle, capi.ptr(c) if len(c) else None, len(c)))


def galerkinJit(space, source, constants=(), skeleton=True, boundary=True):
    return JitGalerkinOperator(space, source, constants, skeleton, boundary)


def galerkin(space, **kwargs):
    return GalerkinOperator(space, **kwargs)


def molGalerkin(space, **kwargs):
    """dune.fem.operator.molGalerkin (python/dune/fem/operator/__init__.py:203-207): Dune::Fem::MOLGalerkinOperator
    (schemes/molgalerkin.hh), i.e. w = M^-1 L[u] -- the operator explicit time stepping applies.  DG spaces only."""
    op = GalerkinOperator(space, **kwargs)
    op.setInverseMass(True)
    return op
