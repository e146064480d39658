"""Grid views.  Mirrors dune.grid.structuredGrid(lo, hi, n) (YaspGrid) as used by pydemo/advectiondiffusion.py:113."""
import ctypes as C

import numpy as np

from . import _capi as capi


class Context:
    """Device context (one per process and GPU).  stream: optional cudaStream_t as int (e.g. torch stream)."""
    _default = {}

    def __init__(self, device=0, stream=None):
        self.handle = C.c_void_p()
        capi.check(capi.lib().b200fem_ctx_create(device, C.c_void_p(stream) if stream else None, C.byref(self.handle)))
        self.device = device
        self.rank, self.world = 0, 1

    @classmethod
    def default(cls, device=0):
        if device not in cls._default:
            cls._default[device] = Context(device)
        return cls._default[device]

    def synchronize(self):
        capi.check(capi.lib().b200fem_ctx_synchronize(self.handle))

    def close(self):
        """destroys the context (stream, communicator, peer mappings); every grid / space / operator made from it must
        have been closed or dropped before"""
        if self.handle:
            capi.lib().b200fem_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    @property
    def peer_memory(self):
        """True when halo exchange and scalar sums use peer-mapped mailboxes over NVLink, False for the NCCL transport"""
        v = C.c_int()
        capi.check(capi.lib().b200fem_ctx_transport(self.handle, C.byref(v)))
        return bool(v.value)

    def init_nccl(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 128)
        capi.check(capi.lib().b200fem_nccl_init(self.handle, buf, rank, world))
        self.rank, self.world = rank, world

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        capi.check(capi.lib().b200fem_nccl_unique_id(buf))
        return buf.raw


class GridView:
    def __init__(self, lo, hi, n, ctx=None, proc=None, rank=0, periodic=None):
        self.ctx = ctx or Context.default()
        self.dim = len(n)
        self.n, self.lo, self.hi = list(n), list(lo), list(hi)
        n_a = (C.c_int32 * 3)(*(list(n) + [1] * (3 - self.dim)))
        lo_a = (C.c_double * 3)(*(list(lo) + [0.0] * (3 - self.dim)))
        hi_a = (C.c_double * 3)(*(list(hi) + [1.0] * (3 - self.dim)))
        self.handle = C.c_void_p()
        if proc is None:
            capi.check(capi.lib().b200fem_mesh_cartesian(self.ctx.handle, self.dim, n_a, lo_a, hi_a, C.byref(self.handle)))
        else:
            p_a = (C.c_int32 * 3)(*(list(proc) + [1] * (3 - len(proc))))
            capi.check(capi.lib().b200fem_mesh_cartesian_distributed(self.ctx.handle, self.dim, n_a, lo_a, hi_a, p_a, rank,
                                                                     C.byref(self.handle)))
        self.proc, self.rank = proc, rank
        self.periodic = 0
        if periodic is not None and any(periodic):       # dune.grid.structuredGrid(..., periodic=[...]) -> YaspGrid's periodic bitset
            self.periodic = sum(1 << d for d, on in enumerate(periodic) if on)
            capi.check(capi.lib().b200fem_mesh_set_periodic(self.handle, self.periodic))

    def close(self):
        if self.handle:
            capi.lib().b200fem_mesh_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):       # (spaces keep their grid view alive through self.gridView, grid views their context)
        try:
            self.close()
        except Exception:
            pass


def structuredGrid(lo, hi, n, ctx=None, proc=None, rank=0, periodic=None):
    return GridView(lo, hi, n, ctx=ctx, proc=proc, rank=rank, periodic=periodic)


class UnstructuredGridView(GridView):
    """Unstructured conforming cube mesh -- dune.alugrid.aluCubeGrid({"vertices": ..., "cubes": ...}) behind an adaptive leaf grid
    view: `vertices` [nv][dim] coordinates, `cubes` [ne][2^dim] vertex numbers in the cube reference element's order."""

    def __init__(self, vertices, cubes, ctx=None):
        self.ctx = ctx or Context.default()
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cubes, dtype=np.int64)
        assert v.ndim == 2 and c.ndim == 2 and c.shape[1] == 1 << v.shape[1], "vertices [nv][dim], cubes [ne][2^dim]"
        self.dim, self.vertices, self.cubes = v.shape[1], v, c
        self.proc, self.rank, self.periodic = None, 0, 0
        self.handle = C.c_void_p()
        capi.check(capi.lib().b200fem_mesh_unstructured(self.ctx.handle, self.dim, v.shape[0], capi.ptr(v), c.shape[0], capi.ptr(c, np.int64), C.byref(self.handle)))


def unstructuredGrid(vertices, cubes, ctx=None):
    return UnstructuredGridView(vertices, cubes, ctx=ctx)


def partition_box(n_global, proc, rank, overlap):
    """host-only: (origin, extents, own_lo, own_hi) of `rank`'s box, see b200fem_partition_box"""
    dim = len(n_global)
    n_a = (C.c_int32 * 3)(*(list(n_global) + [1] * (3 - dim)))
    p_a = (C.c_int32 * 3)(*(list(proc) + [1] * (3 - dim)))
    out = (C.c_int32 * 12)()
    capi.check(capi.lib().b200fem_partition_box(dim, n_a, p_a, rank, int(overlap), out))
    v = list(out)
    return v[0:3], v[3:6], v[6:9], v[9:12]
