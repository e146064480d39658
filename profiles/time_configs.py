#!/usr/bin/env python
"""Device-resident apply timings of the other BASELINE configs (C4: DG Q5 48^3 SIPG Laplace, C5: DG Q3 133^3
advection-diffusion, C3/C1 Lagrange applies) -- the numbers bench.py reports under "other_configs".

  python profiles/time_configs.py [c4] [c5] [c5small] [c3] [quad]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import dune_fem_b200 as fem
from dune_fem_b200 import _capi
from dune_fem_b200.grid import Context


def time_apply(op, size, stream, dev, reps, linear, npairs=3):
    us = [torch.rand(size, dtype=torch.float64, device=dev) * 2 - 1 for _ in range(npairs)]
    ws = [torch.empty(size, dtype=torch.float64, device=dev) for _ in range(npairs)]
    for i in range(3):
        op.apply_dev(us[i % npairs].data_ptr(), ws[i % npairs].data_ptr(), linear)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(reps):
        op.apply_dev(us[i % npairs].data_ptr(), ws[i % npairs].data_ptr(), linear)
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e-3 / reps


def run(name, cells, order, model, kernel, reps, ctx, stream, dev, peak):
    grid = fem.structuredGrid([-1.0] * 3, [1.0] * 3, [cells] * 3, ctx=ctx)
    space = fem.space.dglegendre(grid, order=order, hierarchical=os.environ.get("B200FEM_BENCH_LEX") is None)
    op = fem.operator.galerkin(space, kernel=kernel, **model)
    out = {"dofs": space.size, "kernel": {1: "quadrature", 2: "kronecker"}[kernel]}
    for label, linear in (("affine", False), ("linear", True)):
        t = time_apply(op, space.size, stream, dev, reps, linear)
        out[label] = {"ms": t * 1e3, "dofs_per_s": space.size / t, "gbs_16B": 16 * space.size / t / 1e9, "frac_hbm": 16 * space.size / t / 1e9 / peak}
    n = order + 1
    flop = 2 * 9 * n ** 4 * cells ** 3                      # Kronecker form: 9 n^4 FMA per element
    out["kron_tflops_linear"] = flop / (out["linear"]["ms"] * 1e-3) / 1e12
    print(json.dumps({name: out}), flush=True)
    del op, space, grid
    torch.cuda.empty_cache()


def main():
    which = set(sys.argv[1:]) or {"c4", "c5small"}
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = Context(device=0, stream=stream.cuda_stream)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    adv = lambda k: dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=20.0 * k * k, dirichlet_mask=0b000011, data=1)
    sipg = lambda k: dict(eps=1.0, b=(0.0, 0.0, 0.0), beta=20.0 * k * k, dirichlet_mask=0b111111, data=2)
    kq = _capi.KERNEL_QUADRATURE if "quad" in which else _capi.KERNEL_KRONECKER
    if "c4" in which:
        run("C4 DG Q5 48^3 SIPG Laplace", 48, 5, sipg(5), kq, 20, ctx, stream, dev, peak)
    if "c5small" in which:
        run("C5/8 DG Q3 64^3 advection-diffusion", 64, 3, adv(3), kq, 20, ctx, stream, dev, peak)
    if "c5" in which:
        run("C5 DG Q3 133^3 advection-diffusion (150.6 M dofs)", 133, 3, adv(3), kq, 10, ctx, stream, dev, peak)
    if "q4" in which:
        run("DG Q4 48^3 advection-diffusion", 48, 4, adv(4), kq, 20, ctx, stream, dev, peak)


if __name__ == "__main__":
    main()
