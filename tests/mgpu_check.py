"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):
    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py
Each rank owns one box of the grid; results (operator apply incl. halo exchange, CG) are compared against the
single-domain CPU oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dune_fem_b200 as fem          # noqa: E402
from dune_fem_b200 import _capi      # noqa: E402
from dune_fem_b200.grid import Context, partition_box   # noqa: E402
import oracle_lib as ol              # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = Context(device=lr)
    ids = [Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.init_nccl(ids[0], rank, world)
    procs = {2: [[1, 1, 2], [2, 1, 1]], 4: [[1, 2, 2], [2, 2, 1]], 8: [[1, 2, 4], [2, 2, 2]]}[world]
    worst = 0.0
    for proc in procs:
        n, lo, hi = [8, 6, 8], [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
        # ---------------- DG Q2, both kernels ----------------
        origin, ext, olo, ohi = partition_box(n, proc, rank, overlap=1)
        grid = fem.structuredGrid(lo, hi, n, ctx=ctx, proc=proc, rank=rank)
        space = fem.space.dglegendre(grid, order=2, hierarchical=True)
        osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
        kw = dict(eps=0.05, b=(1.0, 0.5, 0.25), beta=80.0, dirichlet_mask=0b000011, data=1)
        oop = ol.Operator(osp, skeleton=True, boundary=True, **kw)
        ug = np.random.default_rng(5).uniform(-1, 1, osp.size)
        wg = oop.apply(ug)
        nb = 27
        # local (ghosted) box <- global vector
        lidx = np.empty(ext[0] * ext[1] * ext[2], dtype=np.int64)
        k = 0
        for z in range(ext[2]):
            for y in range(ext[1]):
                for x in range(ext[0]):
                    lidx[k] = (origin[0] + x) + n[0] * ((origin[1] + y) + n[1] * (origin[2] + z))
                    k += 1
        gather = (lidx[:, None] * nb + np.arange(nb)[None, :]).ravel()
        ul = np.ascontiguousarray(ug[gather])
        assert ul.size == space.size
        for kernel in (_capi.KERNEL_QUADRATURE, _capi.KERNEL_KRONECKER):
            op = fem.operator.galerkin(space, kernel=kernel, **kw)
            wl = np.empty(space.size)
            op(ul, wl)                               # includes the Copy halo exchange of w
            err = np.abs(wl - wg[gather]).max() / np.abs(wg).max()      # owned AND ghost copies must match
            worst = max(worst, err)
            assert err < 1e-12, (proc, kernel, err)
        # ---------------- DG Q3 (slab Kronecker kernel, BASELINE config 5) ----------------
        space3 = fem.space.dglegendre(grid, order=3, hierarchical=True)
        osp3 = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 3)
        kw3 = dict(eps=0.05, b=(1.0, 0.5, 0.25), beta=180.0, dirichlet_mask=0b000011, data=1)
        ug3 = np.random.default_rng(8).uniform(-1, 1, osp3.size)
        wg3 = ol.Operator(osp3, skeleton=True, boundary=True, threads=4, **kw3).apply(ug3)
        gather3 = (lidx[:, None] * 64 + np.arange(64)[None, :]).ravel()
        op3 = fem.operator.galerkin(space3, kernel=_capi.KERNEL_KRONECKER, **kw3)
        wl3 = np.empty(space3.size)
        op3(np.ascontiguousarray(ug3[gather3]), wl3)
        err = np.abs(wl3 - wg3[gather3]).max() / np.abs(wg3).max()
        worst = max(worst, err)
        assert err < 1e-12, ("q3 slab", proc, err)
        # CG on an SPD DG operator, distributed dots
        kw2 = dict(eps=1.0, c=1.0, beta=80.0, dirichlet_mask=0, data=2)
        op = fem.operator.galerkin(space, **kw2)
        oop2 = ol.Operator(osp, skeleton=True, boundary=True, **kw2)
        bg = -oop2.apply(np.zeros(osp.size))
        inv = fem.solver.CgInverseOperator({"tolerance": 1e-30, "maxiterations": 8})
        inv.bind(op)
        xl = np.zeros(space.size)
        it = inv(np.ascontiguousarray(bg[gather]), xl)
        it_ref, x_ref, hist_ref = oop2.cg(bg, np.zeros(osp.size), 1e-30, 8)
        assert it == it_ref
        np.testing.assert_allclose(inv.residuals, hist_ref, rtol=1e-8)
        assert np.abs(xl - x_ref[gather]).max() / np.abs(x_ref).max() < 1e-9
        # BiCGStab and GMRES on the non-symmetric advection-diffusion operator, MOL scaling (distributed dots via ncclAllReduce)
        kw4 = dict(eps=0.1, b=(1.0, 0.5, 0.2), c=1.0, beta=80.0, dirichlet_mask=0b111111, data=2)
        opn = fem.operator.galerkin(space, **kw4)
        oopn = ol.Operator(osp, skeleton=True, boundary=True, **kw4)
        bgn = -oopn.apply(np.zeros(osp.size))
        for name, solver, ref in (("bicgstab", fem.solver.BicgstabInverseOperator({"tolerance": 1e-30, "maxiterations": 6}), oopn.bicgstab(bgn, np.zeros(osp.size), 1e-30, 6)),
                                  ("gmres", fem.solver.GmresInverseOperator({"tolerance": 1e-30, "maxiterations": 9, "gmres.restart": 4}), oopn.gmres(bgn, np.zeros(osp.size), 1e-30, 9, restart=4))):
            solver.bind(opn)
            xl = np.zeros(space.size)
            it = solver(np.ascontiguousarray(bgn[gather]), xl)
            assert it == ref[0], (name, it, ref[0])
            np.testing.assert_allclose(solver.residuals, ref[2], rtol=1e-7)
            assert np.abs(xl - ref[1][gather]).max() / np.abs(ref[1]).max() < 1e-8, name
        opn.setInverseMass(True)
        oopn.setInverseMass(True)
        wl = np.empty(space.size)
        opn(ul, wl)
        wm = oopn.apply(ug)
        err = np.abs(wl - wm[gather]).max() / np.abs(wm).max()
        worst = max(worst, err)
        assert err < 1e-12, ("mol", proc, err)
        # ---------------- Lagrange P2 (Add on shared dofs) ----------------
        origin, ext, olo, ohi = partition_box(n, proc, rank, overlap=0)
        lspace = fem.space.lagrange(grid, order=2)
        losp = ol.Space(n, lo, hi, ol.LAGRANGE, 2)
        l2g = np.full(lspace.size, -1, dtype=np.int64)
        for z in range(ext[2]):
            for y in range(ext[1]):
                for x in range(ext[0]):
                    el = x + ext[0] * (y + ext[1] * z)
                    eg = (origin[0] + x) + n[0] * ((origin[1] + y) + n[1] * (origin[2] + z))
                    l2g[lspace.mapper(el)] = losp.dofmap(eg)
        assert (l2g >= 0).all()
        kwl = dict(eps=1.0, c=0.3, data=1, dirichlet_mask=0b111111, strong_dirichlet=True)
        lop = fem.operator.galerkin(lspace, **kwl)
        loop_ = ol.Operator(losp, **kwl)
        ulg = np.random.default_rng(6).uniform(-1, 1, losp.size)
        wlg = loop_.apply(ulg)
        wl = np.empty(lspace.size)
        lop(np.ascontiguousarray(ulg[l2g]), wl)
        err = np.abs(wl - wlg[l2g]).max() / np.abs(wlg).max()
        worst = max(worst, err)
        assert err < 1e-12, ("lagrange", proc, err)
        bl = lop.loadVector()
        blg = -loop_.apply(np.zeros(losp.size))
        assert np.abs(bl - blg[l2g]).max() / np.abs(blg).max() < 1e-12
        mask, g = loop_.dirichlet()
        x0g = np.where(mask, g, 0.0)
        inv = fem.solver.CgInverseOperator({"tolerance": 1e-30, "maxiterations": 10})
        inv.bind(lop)
        xl = np.ascontiguousarray(x0g[l2g])
        it = inv(np.ascontiguousarray(blg[l2g]), xl)
        it_ref, x_ref, hist_ref = loop_.cg(blg, x0g, 1e-30, 10)
        assert it == it_ref
        np.testing.assert_allclose(inv.residuals, hist_ref, rtol=1e-8)
        assert np.abs(xl - x_ref[l2g]).max() / np.abs(x_ref).max() < 1e-9
        # matrix-free diagonal (partial sums on interface nodes, completed by the Add exchange) and Jacobi-preconditioned CG
        dg_ref = loop_.diagonal()
        inv = fem.solver.JacobiCgInverseOperator({"tolerance": 1e-30, "maxiterations": 7})
        inv.bind(lop)
        xl = np.ascontiguousarray(x0g[l2g])
        it = inv(np.ascontiguousarray(blg[l2g]), xl)
        it_ref, x_ref, hist_ref = loop_.pcg(dg_ref, blg, x0g, 1e-30, 7)
        assert it == it_ref
        np.testing.assert_allclose(inv.residuals, hist_ref, rtol=1e-8)
        assert np.abs(xl - x_ref[l2g]).max() / np.abs(x_ref).max() < 1e-9
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"mgpu_check OK: world={world}, worst relative difference vs single-domain oracle = {t.item():.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
