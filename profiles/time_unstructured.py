"""times the unstructured-mesh Lagrange kernel (P1 / P2, 3-D) beside the Cartesian quadrature and lattice kernels on the same mesh"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dune_fem_b200 as fem
from dune_fem_b200 import _capi

dev = torch.device("cuda:0")
ctx = fem.grid.Context.default()
def timed(o, size, reps=20, linear=True):
    uu = [torch.rand(size, dtype=torch.float64, device=dev) for _ in range(3)]; ww = [torch.empty(size, dtype=torch.float64, device=dev) for _ in range(3)]
    for i in range(6): o.apply_dev(uu[i % 3].data_ptr(), ww[i % 3].data_ptr(), linear)
    ctx.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import ctypes
    torch.cuda.synchronize(); 
    import time; t0 = time.perf_counter()
    for i in range(reps): o.apply_dev(uu[i % 3].data_ptr(), ww[i % 3].data_ptr(), linear)
    ctx.synchronize(); return (time.perf_counter() - t0) / reps
for order, cu in ((1, 64), (2, 40)):
    ax = np.linspace(0.0, 1.0, cu + 1)
    X = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1)
    vid = (np.arange(cu + 1)[:, None, None] + (cu + 1) * (np.arange(cu + 1)[None, :, None] + (cu + 1) * np.arange(cu + 1)[None, None, :]))
    coords = np.zeros(((cu + 1) ** 3, 3)); coords[vid.ravel()] = X.reshape(-1, 3)
    e0, e1, e2 = np.meshgrid(np.arange(cu), np.arange(cu), np.arange(cu), indexing="ij")
    oe = np.argsort((e0 + cu * (e1 + cu * e2)).ravel())
    cubes = np.stack([vid[e0 + (v & 1), e1 + ((v >> 1) & 1), e2 + (v >> 2)].ravel() for v in range(8)], axis=1)[oe].astype(np.int64)
    import time; t0 = time.perf_counter()
    sp = fem.space.lagrange(fem.unstructuredGrid(coords, cubes, ctx=ctx), order=order)
    print(f"order {order} cells {cu}^3: setup {time.perf_counter() - t0:.2f} s, dofs {sp.size}")
    kw = dict(eps=1.0, data=2, strong_dirichlet=True)
    o = fem.operator.galerkin(sp, dirichlet_mask=1, **kw)
    t = timed(o, sp.size); print(f"  unstructured: {t * 1e6:.1f} us, {sp.size / t / 1e9:.2f} GDoF/s, launches {o.timing()['launches_per_apply']}")
    ss = fem.space.lagrange(fem.structuredGrid([0.0] * 3, [1.0] * 3, [cu] * 3, ctx=ctx), order=order)
    for name, k in (("quadrature", _capi.KERNEL_QUADRATURE), ("lattice", _capi.KERNEL_KRONECKER)):
        o2 = fem.operator.galerkin(ss, dirichlet_mask=63, kernel=k, **kw)
        t = timed(o2, ss.size); print(f"  cartesian {name}: {t * 1e6:.1f} us, {ss.size / t / 1e9:.2f} GDoF/s")
