"""Multi-GPU parity check (run under torchrun on a box with >= 2 GPUs):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py
Each rank owns one box of the grid; results (operator apply incl. halo exchange, Krylov solvers with distributed scalar
products) are compared against the single-domain CPU oracle.  tests/test_gpu_multi.py wraps this script for pytest -m gpu;
bench.py --gpus N calls parity_block() after its timed regions and prints the outcome in its JSON line."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dune_fem_b200 as fem          # noqa: E402
from dune_fem_b200 import _capi      # noqa: E402
from dune_fem_b200.grid import Context, partition_box   # noqa: E402
import oracle_lib as ol              # noqa: E402

TOL = 1e-12


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def dg_gather(n, proc, rank, nb):
    """local (ghosted) box -> indices into the global dof vector"""
    origin, ext, olo, ohi = partition_box(n, proc, rank, overlap=1)
    z, y, x = np.meshgrid(np.arange(ext[2]), np.arange(ext[1]), np.arange(ext[0]), indexing="ij")
    lidx = ((origin[0] + x) + n[0] * ((origin[1] + y) + n[1] * (origin[2] + z))).ravel()
    return (lidx[:, None] * nb + np.arange(nb)[None, :]).ravel()


def lagrange_l2g(lspace, losp, n, proc, rank):
    dim = len(n)
    origin, ext, olo, ohi = partition_box(n, proc[:dim], rank, overlap=0)
    n3, e3, o3 = list(n) + [1] * (3 - dim), list(ext[:dim]) + [1] * (3 - dim), list(origin[:dim]) + [0] * (3 - dim)
    l2g = np.full(lspace.size, -1, dtype=np.int64)
    for z in range(e3[2]):
        for y in range(e3[1]):
            for x in range(e3[0]):
                el = x + e3[0] * (y + e3[1] * z)
                eg = (o3[0] + x) + n3[0] * ((o3[1] + y) + n3[1] * (o3[2] + z))
                l2g[lspace.mapper(el)] = losp.dofmap(eg)
    assert (l2g >= 0).all()
    return l2g


def check_dg_apply(ctx, rank, proc, n, order, kernels, kw, seed):
    lo, hi = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    nb = (order + 1) ** 3
    grid = fem.structuredGrid(lo, hi, n, ctx=ctx, proc=proc, rank=rank)
    space = fem.space.dglegendre(grid, order=order, hierarchical=True)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, order)
    ug = np.random.default_rng(seed).uniform(-1, 1, osp.size)
    wg = ol.Operator(osp, skeleton=True, boundary=True, threads=4, **kw).apply(ug)
    gather = dg_gather(n, proc, rank, nb)
    ul = np.ascontiguousarray(ug[gather])
    assert ul.size == space.size
    worst = 0.0
    for kernel in kernels:
        op = fem.operator.galerkin(space, kernel=kernel, **kw)
        for rep in range(3):                         # repeated: mailbox parity, sequence numbers
            wl = np.full(space.size, np.nan)
            op(ul, wl)                               # includes the Copy halo exchange of w
            err = rel(wl, wg[gather])                # owned AND ghost copies must match
            assert err < TOL, ("dg apply", proc, order, kernel, rep, err)
            worst = max(worst, err)
    return worst, (grid, space, osp, gather)


def check_dg_onb_and_compiled_integrands(ctx, rank, proc, n):
    """round 2: the `dgonb` P_2 space (10 dofs per element: halo blocks of another size) and run-time compiled integrands on a
    distributed mesh, both through the generic quadrature kernel followed by the Copy exchange"""
    lo, hi = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    grid = fem.structuredGrid(lo, hi, n, ctx=ctx, proc=proc, rank=rank)
    worst = 0.0
    space, osp = fem.space.dgonb(grid, order=2), ol.Space(n, lo, hi, ol.DG_ONB, 2)
    kw = dict(eps=0.05, b=(1.0, 0.5, 0.25), beta=80.0, dirichlet_mask=0b000011, data=1)
    ug = np.random.default_rng(21).uniform(-1, 1, osp.size)
    gather = dg_gather(n, proc, rank, 10)
    wl = np.full(space.size, np.nan)
    fem.operator.galerkin(space, **kw)(np.ascontiguousarray(ug[gather]), wl)
    worst = max(worst, rel(wl, ol.Operator(osp, skeleton=True, boundary=True, threads=4, **kw).apply(ug)[gather]))
    src = open(os.path.join(ROOT, "tests", "integrands", "adr_variable.cuh")).read()
    const = [0.05, 1.0, -0.5, 0.25, 80.0, 0.3, 0.7]
    space, osp = fem.space.dglegendre(grid, order=2), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    ug = np.random.default_rng(22).uniform(-1, 1, osp.size)
    gather = dg_gather(n, proc, rank, 27)
    wl = np.full(space.size, np.nan)
    fem.operator.galerkinJit(space, src, const)(np.ascontiguousarray(ug[gather]), wl)
    worst = max(worst, rel(wl, ol.UserOperator(osp, src, const).apply(ug)[gather]))
    assert worst < TOL, ("dgonb / compiled integrands", proc, worst)
    # vector-valued DG space (dimRange 2): the Copy exchange moves element blocks of n_b * dimRange doubles
    R = 2
    srcv = open(os.path.join(ROOT, "tests", "integrands", "system_dg.cuh")).read()
    constv = [0.05, 0.02, 0.7, 80.0]
    space, osp = fem.space.dglegendre(grid, order=2, dimRange=R), ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    ug = np.random.default_rng(23).uniform(-1, 1, osp.size * R)
    gather = dg_gather(n, proc, rank, 27 * R)
    assert gather.size == space.size
    op = fem.operator.galerkinJit(space, srcv, constv)
    wref = ol.VectorUserOperator(osp, R, srcv, constv).apply(ug)[gather]
    for rep in range(2):
        wl = np.full(space.size, np.nan)
        op(np.ascontiguousarray(ug[gather]), wl)
        worst = max(worst, rel(wl, wref))
    assert worst < TOL, ("vector-valued DG space", proc, worst)
    # vector-valued Lagrange space (P2, dimRange 2): the Add exchange sums blocks of dimRange components on shared nodes
    lspace, losp = fem.space.lagrange(grid, order=2, dimRange=R), ol.Space(n, lo, hi, ol.LAGRANGE, 2)
    l2g = lagrange_l2g(fem.space.lagrange(grid, order=2), losp, n, proc, rank)
    l2gv = (l2g[:, None] * R + np.arange(R)[None, :]).ravel()
    assert l2gv.size == lspace.size
    ug = np.random.default_rng(24).uniform(-1, 1, losp.size * R)
    lop = fem.operator.galerkinJit(lspace, srcv, constv, skeleton=False, boundary=True)
    wref = ol.VectorUserOperator(losp, R, srcv, constv, skeleton=False, boundary=True).apply(ug)[l2gv]
    for rep in range(2):
        wl = np.full(lspace.size, np.nan)
        lop(np.ascontiguousarray(ug[l2gv]), wl)
        worst = max(worst, rel(wl, wref))
    assert worst < TOL, ("vector-valued Lagrange space", proc, worst)
    # scalar P2 with WEAK boundary terms through the quadrature kernel (non-linear model): rank interfaces are not domain boundaries
    kwb = dict(eps=0.7, b=(1.0, -0.5, 0.25), c=0.3, gamma=0.5, beta=80.0, dirichlet_mask=0b011011, data=1, boundary=True)
    sspace = fem.space.lagrange(grid, order=2)
    ug = np.random.default_rng(25).uniform(-1, 1, losp.size)
    wl = np.full(sspace.size, np.nan)
    fem.operator.galerkin(sspace, **kwb)(np.ascontiguousarray(ug[l2g]), wl)
    worst = max(worst, rel(wl, ol.Operator(losp, **kwb).apply(ug)[l2g]))
    assert worst < TOL, ("Lagrange space with boundary integrals", proc, worst)
    return worst


def check_dg_solvers(ctx, rank, proc, n):
    lo, hi = [-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]
    grid = fem.structuredGrid(lo, hi, n, ctx=ctx, proc=proc, rank=rank)
    space = fem.space.dglegendre(grid, order=2, hierarchical=True)
    osp = ol.Space(n, lo, hi, ol.DG_LEGENDRE_HIER, 2)
    gather = dg_gather(n, proc, rank, 27)
    # CG on an SPD DG operator, distributed dots (20 iterations: more than one replayed graph chunk)
    kw2 = dict(eps=1.0, c=1.0, beta=80.0, dirichlet_mask=0, data=2)
    op = fem.operator.galerkin(space, **kw2)
    oop2 = ol.Operator(osp, skeleton=True, boundary=True, threads=4, **kw2)
    bg = -oop2.apply(np.zeros(osp.size))
    inv = fem.solver.CgInverseOperator({"tolerance": 1e-30, "maxiterations": 20})
    inv.bind(op)
    xl = np.zeros(space.size)
    it = inv(np.ascontiguousarray(bg[gather]), xl)
    it_ref, x_ref, hist_ref = oop2.cg(bg, np.zeros(osp.size), 1e-30, 20)
    assert it == it_ref
    np.testing.assert_allclose(inv.residuals, hist_ref, rtol=1e-9)
    worst = rel(xl, x_ref[gather])
    assert worst < 1e-10, ("dg cg", worst)
    # BiCGStab and GMRES on the non-symmetric advection-diffusion operator, MOL scaling
    kw4 = dict(eps=0.1, b=(1.0, 0.5, 0.2), c=1.0, beta=80.0, dirichlet_mask=0b111111, data=2)
    opn = fem.operator.galerkin(space, **kw4)
    oopn = ol.Operator(osp, skeleton=True, boundary=True, threads=4, **kw4)
    bgn = -oopn.apply(np.zeros(osp.size))
    for name, solver, ref in (("bicgstab", fem.solver.BicgstabInverseOperator({"tolerance": 1e-30, "maxiterations": 6}), oopn.bicgstab(bgn, np.zeros(osp.size), 1e-30, 6)),
                              ("gmres", fem.solver.GmresInverseOperator({"tolerance": 1e-30, "maxiterations": 9, "gmres.restart": 4}), oopn.gmres(bgn, np.zeros(osp.size), 1e-30, 9, restart=4))):
        solver.bind(opn)
        xl = np.zeros(space.size)
        it = solver(np.ascontiguousarray(bgn[gather]), xl)
        assert it == ref[0], (name, it, ref[0])
        np.testing.assert_allclose(solver.residuals, ref[2], rtol=1e-7)
        assert rel(xl, ref[1][gather]) < 1e-8, name
    ug = np.random.default_rng(5).uniform(-1, 1, osp.size)
    opn.setInverseMass(True)
    oopn.setInverseMass(True)
    wl = np.empty(space.size)
    opn(np.ascontiguousarray(ug[gather]), wl)
    err = rel(wl, oopn.apply(ug)[gather])
    assert err < TOL, ("mol", proc, err)
    return err


def check_lagrange(ctx, rank, proc, n, order, cg_iters=20, jacobi=True):
    dim = len(n)
    lo, hi = [0.0] * dim, [1.0] * dim
    grid = fem.structuredGrid(lo, hi, n, ctx=ctx, proc=proc[:dim], rank=rank)
    lspace = fem.space.lagrange(grid, order=order)
    losp = ol.Space(n, lo, hi, ol.LAGRANGE, order)
    l2g = lagrange_l2g(lspace, losp, n, proc, rank)
    kwl = dict(eps=1.0, c=0.3, data=1, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
    lop = fem.operator.galerkin(lspace, **kwl)
    loop_ = ol.Operator(losp, threads=4, **kwl)
    ulg = np.random.default_rng(6).uniform(-1, 1, losp.size)
    wlg = loop_.apply(ulg)
    worst = 0.0
    for rep in range(3):
        wl = np.full(lspace.size, np.nan)
        lop(np.ascontiguousarray(ulg[l2g]), wl)       # Add exchange on shared nodes, then the Dirichlet wrapper
        err = rel(wl, wlg[l2g])
        assert err < TOL, ("lagrange apply", proc, order, rep, err)
        worst = max(worst, err)
    bl = lop.loadVector()
    blg = -loop_.apply(np.zeros(losp.size))
    assert rel(bl, blg[l2g]) < TOL
    mask, g = loop_.dirichlet()
    x0g = np.where(mask, g, 0.0)
    inv = fem.solver.CgInverseOperator({"tolerance": 1e-30, "maxiterations": cg_iters})
    inv.bind(lop)
    xl = np.ascontiguousarray(x0g[l2g])
    it = inv(np.ascontiguousarray(blg[l2g]), xl)
    it_ref, x_ref, hist_ref = loop_.cg(blg, x0g, 1e-30, cg_iters)
    assert it == it_ref
    np.testing.assert_allclose(inv.residuals, hist_ref, rtol=1e-9)
    err = rel(xl, x_ref[l2g])
    assert err < 1e-10, ("lagrange cg", err)
    if jacobi:
        # matrix-free diagonal (partial sums on interface nodes, completed by the Add exchange) and Jacobi-preconditioned CG
        dg_ref = loop_.diagonal()
        inv = fem.solver.JacobiCgInverseOperator({"tolerance": 1e-30, "maxiterations": 7})
        inv.bind(lop)
        xl = np.ascontiguousarray(x0g[l2g])
        it = inv(np.ascontiguousarray(blg[l2g]), xl)
        it_ref, x_ref, hist_ref = loop_.pcg(dg_ref, blg, x0g, 1e-30, 7)
        assert it == it_ref
        np.testing.assert_allclose(inv.residuals, hist_ref, rtol=1e-8)
        assert rel(xl, x_ref[l2g]) < 1e-9
    return worst, err


def check_large_messages(ctx, rank, world, proc):
    """Copy exchange with face messages far beyond what one resident wave of blocks can hold (the round-1 kernel made a block's
    receive wait for ALL of the peer's send blocks): ghost copies must equal the owner's values, checked in closed form."""
    import torch
    n = [200 * proc[0], 200 * proc[1], 4 * proc[2]]             # Q3: a z-face message is 200 x 200 elements x 64 doubles = 2.56 M doubles = 1250 blocks
    grid = fem.structuredGrid([0.0] * 3, [1.0] * 3, n, ctx=ctx, proc=proc, rank=rank)
    space = fem.space.dglegendre(grid, order=3, hierarchical=True)
    op = fem.operator.galerkin(space, eps=1.0, beta=180.0)
    origin, ext, olo, ohi = partition_box(n, proc, rank, overlap=1)
    z, y, x = np.meshgrid(np.arange(ext[2]), np.arange(ext[1]), np.arange(ext[0]), indexing="ij")
    gid = ((origin[0] + x) + n[0] * ((origin[1] + y) + n[1] * (origin[2] + z))).ravel().astype(np.float64)
    owned = ((x >= olo[0]) & (x < ohi[0]) & (y >= olo[1]) & (y < ohi[1]) & (z >= olo[2]) & (z < ohi[2])).ravel()
    want = (gid[:, None] * 64 + np.arange(64)[None, :]).ravel()
    have = np.where(np.repeat(owned, 64), want, -1.0)
    for rep in range(3):
        v = torch.from_numpy(have.copy()).cuda()
        if rep == 1 and rank == world - 1:
            time.sleep(3.0)                                    # deliberate rank skew: the peers wait (the old kernel gave up after ~2 s)
        op.communicate_dev(v.data_ptr())
        ctx.synchronize()
        assert np.array_equal(v.cpu().numpy(), want), ("large exchange", rep)
    return 0.0


def parity_block(ctx, rank, world):
    """what bench.py --gpus N reports: small Q2 + Q3 + P2 applies and CG iterations on bench.py's own process grids"""
    dg_proc = {2: [1, 1, 2], 4: [1, 2, 2], 8: [1, 2, 4]}[world]
    lag_proc = {2: [1, 1, 2], 4: [1, 2, 2], 8: [2, 2, 2]}[world]
    cases = {}
    kw = dict(eps=0.05, b=(1.0, 0.5, 0.25), beta=80.0, dirichlet_mask=0b000011, data=1)
    cases["dg_q2_apply_march_fused_exchange"], _ = check_dg_apply(ctx, rank, dg_proc, [16, 8, 16], 2, (_capi.KERNEL_KRONECKER,), kw, 5)
    cases["dg_q2_apply_quadrature"], _ = check_dg_apply(ctx, rank, dg_proc, [8, 6, 8], 2, (_capi.KERNEL_QUADRATURE,), kw, 5)
    cases["dg_q3_apply_tensor_core"], _ = check_dg_apply(ctx, rank, dg_proc, [8, 6, 8], 3, (_capi.KERNEL_KRONECKER,), dict(kw, beta=180.0), 8)
    cases["lagrange_p2_apply"], cases["lagrange_p2_cg10_x"] = check_lagrange(ctx, rank, lag_proc, [8, 6, 8], 2, cg_iters=10, jacobi=False)
    return {"max_rel_err": max(cases.values()), "cases": cases, "tolerance": "apply 1e-12 of max|w|; CG x 1e-10, residual history 1e-9, iteration counts equal",
            "transport": "peer memory" if ctx.peer_memory else "nccl"}


def main():
    import torch
    import torch.distributed as dist
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    ctx = Context(device=lr)
    ids = [Context.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.init_nccl(ids[0], rank, world)
    procs = {2: [[1, 1, 2], [2, 1, 1], [1, 2, 1]], 4: [[1, 2, 2], [2, 2, 1]], 8: [[1, 2, 4], [2, 2, 2]]}[world]
    worst = 0.0
    kw = dict(eps=0.05, b=(1.0, 0.5, 0.25), beta=80.0, dirichlet_mask=0b000011, data=1)
    for proc in procs:
        # DG Q2: generic kernel, marching kernel (fused exchange when x is not split), plain tile kernel
        w, _ = check_dg_apply(ctx, rank, proc, [8, 6, 8], 2, (_capi.KERNEL_QUADRATURE, _capi.KERNEL_KRONECKER, _capi.KERNEL_KRONECKER_TILE), kw, 5)
        worst = max(worst, w)
        # partial tiles in x / y, several columns per CTA row, odd plane counts
        w, _ = check_dg_apply(ctx, rank, proc, [34, 20, 14], 2, (_capi.KERNEL_KRONECKER,), kw, 15)
        worst = max(worst, w)
        w, _ = check_dg_apply(ctx, rank, proc, [8, 6, 8], 3, (_capi.KERNEL_KRONECKER,), dict(kw, beta=180.0), 8)   # Q3 tensor-core kernel on boxes with ghost layers (BASELINE config 5)
        worst = max(worst, w)
        worst = max(worst, check_dg_solvers(ctx, rank, proc, [8, 6, 8]))
        if proc is procs[0]:
            worst = max(worst, check_dg_onb_and_compiled_integrands(ctx, rank, proc, [8, 6, 8]))
        w, _ = check_lagrange(ctx, rank, proc, [8, 6, 8], 2)
        worst = max(worst, w)
        w, _ = check_lagrange(ctx, rank, proc, [8, 6, 8], 1)
        worst = max(worst, w)
        if proc is procs[0]:
            w, _ = check_lagrange(ctx, rank, proc, [4, 4, 6], 3, cg_iters=12)    # order 3: quadrature kernels + Add exchange on a 3 n + 1 lattice
            worst = max(worst, w)
        if proc[2] == 1:
            w, _ = check_lagrange(ctx, rank, proc, [12, 10], 1)          # 2-D lattice
            worst = max(worst, w)
    check_large_messages(ctx, rank, world, procs[0])
    pb = parity_block(ctx, rank, world)
    worst = max(worst, pb["max_rel_err"])
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"mgpu_check OK: world={world}, transport={'peer memory' if ctx.peer_memory else 'nccl'}, worst relative difference vs single-domain oracle = {t.item():.3e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
