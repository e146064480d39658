#!/usr/bin/env python
"""Device-resident apply timing of the generic quadrature kernels (DG Q2 64^3 = BASELINE config 2, Q3, Q5; Lagrange P2)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dune_fem_b200 as fem
from dune_fem_b200 import _capi
from dune_fem_b200.grid import Context

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
ctx = Context(device=0, stream=stream.cuda_stream)
cases = {"q2": (2, 64, dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1)),
         "q2nl": (2, 64, dict(eps=1e-2, b=(1.0, 0.0, 0.0), gamma=0.5, beta=80.0, dirichlet_mask=0b000011, data=1)),
         "q3": (3, 48, dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=180.0, dirichlet_mask=0b000011, data=1)),
         "q5": (5, 24, dict(eps=1.0, beta=500.0, dirichlet_mask=0b111111, data=2))}
for name in (sys.argv[1:] or ["q2"]):
    order, cells, model = cases[name]
    g = fem.structuredGrid([-1.0] * 3, [1.0] * 3, [cells] * 3, ctx=ctx)
    sp = fem.space.dglegendre(g, order=order, hierarchical=True)
    op = fem.operator.galerkin(sp, kernel=_capi.KERNEL_QUADRATURE, **model)
    n = sp.size
    us = [torch.rand(n, dtype=torch.float64, device=dev) for _ in range(3)]
    ws = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(3)]
    for lin in (True, False):
        for i in range(3): op.apply_dev(us[i % 3].data_ptr(), ws[i % 3].data_ptr(), lin)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        e0.record(stream)
        for i in range(reps): op.apply_dev(us[i % 3].data_ptr(), ws[i % 3].data_ptr(), lin)
        e1.record(stream); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        print(json.dumps({name: {"dofs": n, "linear": lin, "us": round(t * 1e6, 1), "gdofs": round(n / t / 1e9, 2)}}), flush=True)
