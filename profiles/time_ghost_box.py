"""One GPU: the marching kernel on the rank-local boxes of a distributed grid (no communicator attached, so no exchange):
isolates what the ghost layers cost the compute kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dune_fem_b200 as fem
from dune_fem_b200.grid import Context

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = Context(device=0, stream=stream.cuda_stream)
MODEL = dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1)
for proc, rank in (([1, 1, 1], 0), ([1, 1, 2], 0), ([1, 1, 2], 1), ([1, 2, 2], 0), ([1, 2, 4], 3), ([1, 2, 4], 2)):
    n = [64 * p for p in proc]
    g = fem.structuredGrid([-1.0] * 3, [-1.0 + 2.0 * p for p in proc], n, ctx=ctx, proc=proc if proc != [1, 1, 1] else None, rank=rank)
    sp = fem.space.dglegendre(g, order=2, hierarchical=True)
    op = fem.operator.galerkin(sp, **MODEL)
    us = [torch.rand(sp.size, dtype=torch.float64, device=dev) for _ in range(6)]
    ws = [torch.empty(sp.size, dtype=torch.float64, device=dev) for _ in range(6)]
    res = []
    for linear in (False, True):
        for i in range(12):
            op.apply_dev(us[i % 6].data_ptr(), ws[i % 6].data_ptr(), linear)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(500):
            op.apply_dev(us[i % 6].data_ptr(), ws[i % 6].data_ptr(), linear)
        e1.record(stream); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) * 2.0)
    print(f"proc {proc} rank {rank}: size {sp.size}  affine {res[0]:.2f} us  linear {res[1]:.2f} us", flush=True)
    del op, sp, g, us, ws
