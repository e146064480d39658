"""N > 1 host logic on CPU (gloo, world_size 2): the block partition exposed by the C ABI, the decomposition
invariance of the operator (rank-local applies with ghost neighbours reproduce the single-domain result) and the
primary-dof dot product.  The device halo exchange itself is covered by tests/mgpu_check.py on real GPUs."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_lib as ol
    from dune_fem_b200.grid import partition_box

    n, proc = [6, 4, 5], [1, 1, 2]
    # ---- DG: Copy semantics, overlap 1 ----
    origin, ext, olo, ohi = partition_box(n, proc, rank, overlap=1)
    glo = [origin[d] + olo[d] for d in range(3)]
    ghi = [origin[d] + ohi[d] for d in range(3)]
    sp = ol.Space(n, [-1, -1, -1], [1, 1, 1], ol.DG_LEGENDRE, 2)
    kw = dict(eps=0.1, b=(1.0, 0.5, 0.25), beta=80.0, dirichlet_mask=0b000011, data=1, skeleton=True, boundary=True)
    op = ol.Operator(sp, **kw)
    u = np.random.default_rng(1).uniform(-1, 1, sp.size)          # same on every rank
    w_local = op.apply_box(u, glo, ghi)                           # owned elements of this rank, ghosts as neighbours
    nb = sp.local_size
    owned = np.zeros(sp.size, dtype=bool)
    for e in range(sp.elements):
        c = (e % n[0], (e // n[0]) % n[1], e // (n[0] * n[1]))
        if all(glo[d] <= c[d] < ghi[d] for d in range(3)):
            owned[e * nb:(e + 1) * nb] = True
    contrib = torch.from_numpy(np.where(owned, w_local, 0.0))
    count = torch.from_numpy(owned.astype(np.float64))
    dist.all_reduce(contrib)
    dist.all_reduce(count)
    ok_cover = bool((count.numpy() == 1).all())                   # the owned boxes tile the grid exactly once
    w_ref = op.apply(u)
    err_dg = float(np.abs(contrib.numpy() - w_ref).max() / np.abs(w_ref).max())
    # primary-dof dot product: sum over ranks of owned-dof dots == global dot (scalarproducts.hh:115-127)
    d = torch.tensor([float(np.dot(u[owned], w_ref[owned]))], dtype=torch.float64)
    dist.all_reduce(d)
    err_dot = abs(d.item() - float(np.dot(u, w_ref))) / abs(float(np.dot(u, w_ref)))
    # ---- Lagrange: Add semantics, no overlap ----
    origin, ext, olo, ohi = partition_box(n, proc, rank, overlap=0)
    sl = ol.Space(n, [0, 0, 0], [1, 1, 1], ol.LAGRANGE, 2)
    opl = ol.Operator(sl, eps=1.0, c=0.5, data=2)
    ul = np.random.default_rng(2).uniform(-1, 1, sl.size)
    wl = torch.from_numpy(opl.apply_box(ul, origin, [origin[d] + ext[d] for d in range(3)]))
    dist.all_reduce(wl)                                            # communicate() with DFCommunicationOperation::Add
    wl_ref = opl.apply(ul)
    err_lag = float(np.abs(wl.numpy() - wl_ref).max() / np.abs(wl_ref).max())
    if rank == 0:
        results.put((ok_cover, err_dg, err_dot, err_lag))
    dist.destroy_process_group()


def test_two_rank_decomposition_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok_cover, err_dg, err_dot, err_lag = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert ok_cover
    assert err_dg < 1e-12 and err_lag < 1e-12 and err_dot < 1e-12


def test_partition_box_properties():
    from dune_fem_b200.grid import partition_box
    for n, proc in [([64, 64, 128], [1, 1, 2]), ([7, 5, 9], [2, 1, 3]), ([133, 266, 532], [1, 2, 4])]:
        world = proc[0] * proc[1] * proc[2]
        cover = np.zeros(n, dtype=int)
        for r in range(world):
            origin, ext, olo, ohi = partition_box(n, proc, r, overlap=1)
            g0 = [origin[d] + olo[d] for d in range(3)]
            g1 = [origin[d] + ohi[d] for d in range(3)]
            cover[g0[0]:g1[0], g0[1]:g1[1], g0[2]:g1[2]] += 1
            for d in range(3):                                    # one ghost layer exactly where a neighbour rank exists
                assert olo[d] == (1 if g0[d] > 0 else 0)
                assert ext[d] - ohi[d] == (1 if g1[d] < n[d] else 0)
        assert (cover == 1).all()
