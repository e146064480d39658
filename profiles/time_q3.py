#!/usr/bin/env python
"""C5-type timing of the DG Q3 Kronecker apply (64^3 and 133^3 cells), device-resident, both Q3 kernels (B200FEM_Q3_SLAB=1 selects
the slab kernel)."""
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
for cells in [int(a) for a in (sys.argv[1:] or ["64", "133"])]:
    g = fem.structuredGrid([-1.0] * 3, [1.0] * 3, [cells] * 3, ctx=ctx)
    sp = fem.space.dglegendre(g, order=3, hierarchical=True)
    op = fem.operator.galerkin(sp, kernel=_capi.KERNEL_KRONECKER, eps=1e-5, b=(1.0, 0.0, 0.0), beta=180.0, dirichlet_mask=0b000011, data=1)
    n = sp.size
    npairs = 3 if cells > 100 else 6
    us = [torch.rand(n, dtype=torch.float64, device=dev) for _ in range(npairs)]
    ws = [torch.empty(n, dtype=torch.float64, device=dev) for _ in range(npairs)]
    for lin in (True, False):
        for i in range(npairs): op.apply_dev(us[i].data_ptr(), ws[i].data_ptr(), lin)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for i in range(reps): op.apply_dev(us[i % npairs].data_ptr(), ws[i % npairs].data_ptr(), lin)
        e1.record(stream); torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3 / reps
        print(json.dumps({"cells": cells, "dofs": n, "linear": lin, "us": round(t * 1e6, 1), "gdofs": round(n / t / 1e9, 2), "frac_hbm_16B": round(16 * n / t / 1e9 / 6543.1, 3),
                          "tflops_kron": round(4608 * cells ** 3 / t / 1e12, 2)}), flush=True)
    del us, ws, op, sp, g
