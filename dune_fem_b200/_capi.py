"""ctypes binding of include/b200fem.h (dune_fem_b200/lib/libb200fem.so).

There is no Python or CPU fallback: if the shared library is missing, or no CUDA device is usable, the calls fail.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200fem.so")

# every symbol include/b200fem.h declares (checked by tests/test_capi_symbols.py)
SYMBOLS = [
    "b200fem_last_error", "b200fem_version", "b200fem_ctx_create", "b200fem_ctx_destroy", "b200fem_ctx_synchronize",
    "b200fem_malloc", "b200fem_free", "b200fem_memcpy_h2d", "b200fem_memcpy_d2h", "b200fem_mesh_cartesian",
    "b200fem_mesh_cartesian_distributed", "b200fem_mesh_destroy", "b200fem_partition_box", "b200fem_mesh_local_box", "b200fem_march_schedule", "b200fem_space_create", "b200fem_space_destroy",
    "b200fem_space_size", "b200fem_space_local_size", "b200fem_space_elements", "b200fem_space_dofmap",
    "b200fem_operator_create", "b200fem_operator_destroy", "b200fem_operator_apply", "b200fem_operator_apply_linear",
    "b200fem_operator_apply_dev", "b200fem_operator_load_vector", "b200fem_operator_set_communicate",
    "b200fem_operator_set_quadrature_orders", "b200fem_operator_set_kernel", "b200fem_operator_set_host_pipeline", "b200fem_operator_set_inverse_mass", "b200fem_operator_linearize", "b200fem_operator_linearize_dev", "b200fem_operator_dirichlet",
    "b200fem_operator_timing", "b200fem_cg_solve", "b200fem_cg_solve_dev", "b200fem_bicgstab_solve", "b200fem_bicgstab_solve_dev", "b200fem_gmres_solve", "b200fem_gmres_solve_dev", "b200fem_operator_diagonal", "b200fem_pcg_solve", "b200fem_pcg_solve_dev", "b200fem_dot_dev", "b200fem_axpy_dev",
    "b200fem_ctx_set_nccl", "b200fem_nccl_unique_id", "b200fem_nccl_init", "b200fem_ctx_transport", "b200fem_communicate_dev",
    "b200fem_operator_create_jit", "b200fem_operator_set_constants", "b200fem_jit_compile_check", "b200fem_device_count", "b200fem_mesh_set_periodic", "b200fem_newton_solve", "b200fem_newton_solve_dev", "b200fem_space_create_vector", "b200fem_space_dim_range", "b200fem_jit_compile_check_space", "b200fem_mesh_unstructured", "b200fem_jit_compile_check_unstructured", "b200fem_unstructured_numbering",
]

OK, ERR_INVALID, ERR_NOT_IMPLEMENTED, ERR_CUDA, ERR_COMM = 0, -1, -2, -3, -4
LAGRANGE, DG_LEGENDRE, DG_LEGENDRE_HIER, DG_ONB = 0, 1, 2, 3
NUMBERING_YASP, NUMBERING_ADAPTIVE_LEAF = 0, 1
KERNEL_AUTO, KERNEL_QUADRATURE, KERNEL_KRONECKER, KERNEL_KRONECKER_TILE = 0, 1, 2, 3
TOL_ABSOLUTE, TOL_RELATIVE, TOL_RESIDUAL_REDUCTION = 0, 1, 2


class Model(C.Structure):
    _fields_ = [("eps", C.c_double), ("b", C.c_double * 3), ("c", C.c_double), ("gamma", C.c_double),
                ("beta", C.c_double), ("dirichlet_mask", C.c_int32), ("data", C.c_int32),
                ("has_skeleton", C.c_int32), ("has_boundary", C.c_int32), ("strong_dirichlet", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("last_apply_ms", C.c_double), ("last_exchange_ms", C.c_double), ("applies", C.c_int64),
                ("kernel", C.c_int32), ("launches_per_apply", C.c_int32)]


class B200FemError(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"b200fem error {code}: {message}")
        self.code = code


_lib = None


def lib():
    """Load the CUDA library.  Fails loudly if it has not been built (python __graft_entry__.py / build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build the CUDA extension first (__graft_entry__.build()); "
                          "dune_fem_b200 has no CPU fallback")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    L.b200fem_last_error.restype = C.c_char_p
    sig = {
        "b200fem_device_count": [P(C.c_int)], "b200fem_ctx_create": [C.c_int, vp, P(vp)], "b200fem_ctx_destroy": [vp], "b200fem_ctx_synchronize": [vp],
        "b200fem_malloc": [vp, i64, P(vp)], "b200fem_free": [vp, vp], "b200fem_memcpy_h2d": [vp, vp, vp, i64],
        "b200fem_memcpy_d2h": [vp, vp, vp, i64],
        "b200fem_mesh_cartesian": [vp, C.c_int, P(i32), P(dbl), P(dbl), P(vp)],
        "b200fem_mesh_cartesian_distributed": [vp, C.c_int, P(i32), P(dbl), P(dbl), P(i32), C.c_int, P(vp)],
        "b200fem_mesh_destroy": [vp], "b200fem_mesh_set_periodic": [vp, C.c_int],
        "b200fem_partition_box": [C.c_int, P(i32), P(i32), C.c_int, C.c_int, P(i32)],
        "b200fem_mesh_local_box": [vp, C.c_int, P(i32)],
        "b200fem_mesh_unstructured": [vp, C.c_int, i64, vp, i64, vp, P(vp)],
        "b200fem_unstructured_numbering": [C.c_int, i64, vp, i64, vp, C.c_int, P(i64), vp, vp, vp],
        "b200fem_jit_compile_check_unstructured": [C.c_char_p, C.c_int, C.c_int, C.c_char_p, C.c_int],
        "b200fem_march_schedule": [P(i32), C.c_int, C.c_int, P(i32), i32, P(i32), P(i32)],
        "b200fem_space_create": [vp, C.c_int, C.c_int, C.c_int, P(vp)], "b200fem_space_destroy": [vp],
        "b200fem_space_create_vector": [vp, C.c_int, C.c_int, C.c_int, C.c_int, P(vp)], "b200fem_space_dim_range": [vp, P(i32)],
        "b200fem_jit_compile_check_space": [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int],
        "b200fem_space_size": [vp, P(i64)], "b200fem_space_local_size": [vp, P(i32)], "b200fem_space_elements": [vp, P(i64)],
        "b200fem_space_dofmap": [vp, i64, P(i64)],
        "b200fem_operator_create": [vp, P(Model), P(vp)], "b200fem_operator_destroy": [vp],
        "b200fem_operator_apply": [vp, vp, vp], "b200fem_operator_apply_linear": [vp, vp, vp],
        "b200fem_operator_apply_dev": [vp, vp, vp, C.c_int], "b200fem_operator_load_vector": [vp, vp],
        "b200fem_operator_set_communicate": [vp, C.c_int], "b200fem_operator_set_quadrature_orders": [vp, C.c_uint, C.c_uint],
        "b200fem_operator_set_kernel": [vp, C.c_int], "b200fem_operator_set_inverse_mass": [vp, C.c_int],
        "b200fem_operator_set_host_pipeline": [vp, C.c_int], "b200fem_ctx_transport": [vp, P(C.c_int)],
        "b200fem_operator_linearize": [vp, vp, dbl], "b200fem_operator_linearize_dev": [vp, vp, dbl], "b200fem_operator_dirichlet": [vp, vp, vp],
        "b200fem_operator_timing": [vp, P(Timing)],
        "b200fem_cg_solve": [vp, vp, vp, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_cg_solve_dev": [vp, vp, vp, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_bicgstab_solve": [vp, vp, vp, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_bicgstab_solve_dev": [vp, vp, vp, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_operator_diagonal": [vp, vp],
        "b200fem_pcg_solve": [vp, vp, vp, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_pcg_solve_dev": [vp, vp, vp, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_gmres_solve": [vp, vp, vp, C.c_int, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_gmres_solve_dev": [vp, vp, vp, C.c_int, dbl, C.c_int, C.c_int, P(C.c_int), vp],
        "b200fem_dot_dev": [vp, vp, vp, P(dbl)], "b200fem_axpy_dev": [vp, dbl, vp, vp],
        "b200fem_ctx_set_nccl": [vp, vp, C.c_int, C.c_int], "b200fem_nccl_unique_id": [vp],
        "b200fem_nccl_init": [vp, vp, C.c_int, C.c_int], "b200fem_communicate_dev": [vp, vp],
        "b200fem_operator_create_jit": [vp, C.c_char_p, vp, C.c_int, C.c_int, C.c_int, P(vp)],
        "b200fem_operator_set_constants": [vp, vp, C.c_int],
        "b200fem_newton_solve": [vp, vp, vp, dbl, C.c_int, C.c_int, dbl, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_int), P(C.c_int), P(dbl), P(C.c_int)],
        "b200fem_newton_solve_dev": [vp, vp, vp, dbl, C.c_int, C.c_int, dbl, C.c_int, C.c_int, C.c_int, C.c_int, P(C.c_int), P(C.c_int), P(dbl), P(C.c_int)], "b200fem_jit_compile_check": [C.c_char_p, C.c_int, C.c_char_p, C.c_int],
    }
    for name, argtypes in sig.items():
        fn = getattr(L, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = L
    return L


def check(code):
    if code != 0:
        raise B200FemError(code, lib().b200fem_last_error().decode())


def ptr(a, dtype=np.float64):
    """numpy array -> void*.  The C ABI reads raw memory: the array must be C-contiguous and of the expected dtype (dof
    vectors float64, masks uint8) -- anything else would be silently reinterpreted."""
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise TypeError(f"expected a C-contiguous numpy array of dtype {np.dtype(dtype).name}, got "
                        f"{type(a).__name__}{'' if not isinstance(a, np.ndarray) else ' of dtype ' + a.dtype.name}")
    return a.ctypes.data_as(C.c_void_p)
