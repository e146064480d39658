#!/usr/bin/env python
"""Device-resident apply timing of the Lagrange Kronecker kernel (C3: P2 128^3, C1: P1 256^2 and a 4096^2 variant)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dune_fem_b200 as fem
from dune_fem_b200.grid import Context

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = Context(device=0, stream=stream.cuda_stream)
which = sys.argv[1:] or ["c3"]
cfgs = {"c3": (3, 128, 2), "c1": (2, 256, 1), "c1big": (2, 4096, 1), "p1_3d": (3, 256, 1)}
for name in which:
    dim, cells, order = cfgs[name]
    g = fem.structuredGrid([0.0] * dim, [1.0] * dim, [cells] * dim, ctx=ctx)
    sp = fem.space.lagrange(g, order=order)
    op = fem.operator.galerkin(sp, eps=1.0, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
    n = sp.size
    us = [torch.rand(n, dtype=torch.float64, device=dev) for _ in range(3)]
    ws = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
    for lin in (True, False):
        for i in range(3): op.apply_dev(us[i % 3].data_ptr(), ws[i % 3].data_ptr(), lin)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for i in range(reps): op.apply_dev(us[i % 3].data_ptr(), ws[i % 3].data_ptr(), lin)
        e1.record(stream); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        print(json.dumps({name: {"dofs": n, "linear": lin, "us": t * 1e6, "gdofs": n / t / 1e9, "gbs_16B": 16 * n / t / 1e9}}), flush=True)
