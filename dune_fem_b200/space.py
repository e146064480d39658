"""Discrete function spaces.  Mirrors dune.fem.space.lagrange / dglegendre / dgonb (python/dune/fem/space/_spaces.py:106,183-229)."""
import ctypes as C

import numpy as np

from . import _capi as capi


class DiscreteFunctionSpace:
    def __init__(self, gridView, kind, order, numbering=capi.NUMBERING_YASP, dimRange=1):
        self.gridView, self.kind, self.order, self.dimRange = gridView, kind, order, dimRange
        self.handle = C.c_void_p()
        if dimRange == 1:
            capi.check(capi.lib().b200fem_space_create(gridView.handle, kind, order, numbering, C.byref(self.handle)))
        else:    # create.space(name, grid, dimRange=dimR, order=order): size = blocks * dimRange, dof (block, c) = block * dimRange + c
            capi.check(capi.lib().b200fem_space_create_vector(gridView.handle, kind, order, numbering, dimRange, C.byref(self.handle)))
        v = C.c_int64()
        capi.check(capi.lib().b200fem_space_size(self.handle, C.byref(v)))
        self.size = v.value
        capi.check(capi.lib().b200fem_space_elements(self.handle, C.byref(v)))
        self.elements = v.value
        nb = C.c_int32()
        capi.check(capi.lib().b200fem_space_local_size(self.handle, C.byref(nb)))
        self.localBlockSize = nb.value

    def close(self):
        if self.handle:
            capi.lib().b200fem_space_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):       # (operators keep their space alive through self.space)
        try:
            self.close()
        except Exception:
            pass

    def mapper(self, element):
        """blockMapper().map(entity) -- global dof indices of an element"""
        out = np.empty(self.localBlockSize, dtype=np.int64)
        capi.check(capi.lib().b200fem_space_dofmap(self.handle, element, out.ctypes.data_as(C.POINTER(C.c_int64))))
        return out

    def function(self, name="uh", values=None):
        """a discrete function = host-owned dof vector in the reference layout (numpy storage)"""
        a = np.zeros(self.size) if values is None else np.ascontiguousarray(values, dtype=np.float64)
        assert a.shape == (self.size,)
        return a


def lagrange(gridView, order=1, numbering=capi.NUMBERING_YASP, dimRange=1):
    return DiscreteFunctionSpace(gridView, capi.LAGRANGE, order, numbering, dimRange=dimRange)


def dglegendre(gridView, order=1, hierarchical=True, dimRange=1):
    return DiscreteFunctionSpace(gridView, capi.DG_LEGENDRE_HIER if hierarchical else capi.DG_LEGENDRE, order, dimRange=dimRange)


def dgonb(gridView, order=1, dimRange=1):
    """orthonormal P_k on cubes -- the space pydemo/advectiondiffusion.py:9 imports (shapefunctionset/orthonormal.hh:55-60)"""
    return DiscreteFunctionSpace(gridView, capi.DG_ONB, order, dimRange=dimRange)
