#!/usr/bin/env python
"""bench.py -- operator-apply throughput (DoF/s, FP64) of the matrix-free Galerkin operator on B200.

Workload (BASELINE.json configs[1], "C2"): advection-diffusion, DG Q2 (dglegendre, hierarchical), 3-D cube
[-1,1]^3 with 64^3 cells per GPU, SIPG + upwind integrands of pydemo/advectiondiffusion.py, explicit operator
apply w = L[u] = A u - b (complete affine operator; b = -L[0] precomputed once and streamed by the kernel).
A "step" is one operator application over one synthetic dof vector (u ~ U(-1,1), PCG64 seed 20261017).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU, torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference, all host threads

Prints ONE JSON line (rank 0).  See the task contract for the keys.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

SEED = 20261017
ORDER = 2
CELLS = 64
MODEL = dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=20.0 * ORDER ** 2, dirichlet_mask=0b000011, data=1)
ALGORITHMIC_BYTES_PER_DOF = 16.0       # read u once + write w once


def proc_grid(n):
    return {1: [1, 1, 1], 2: [1, 1, 2], 4: [1, 2, 2], 8: [1, 2, 4]}[n]   # x (the contiguous axis) is never split


class ClockSampler(threading.Thread):
    """samples SM clock and throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, period, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80, "sync_boost": 0x10, "applications_clocks": 0x2}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(self.period)

    def result(self):
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def cpu_reference_apply(threads, sample_cells, reps):
    """times the CPU oracle (restatement of the reference's GalerkinOperator::evaluate) -- checker/baseline only"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    n = sample_cells
    h = 2.0 / CELLS
    hi = [-1 + n[0] * h, -1 + n[1] * h, -1 + n[2] * h]
    sp = ol.Space(n, [-1, -1, -1], hi, ol.DG_LEGENDRE_HIER, ORDER)
    op = ol.Operator(sp, skeleton=True, boundary=True, threads=threads, **MODEL)
    u = np.random.default_rng(SEED).uniform(-1, 1, sp.size)
    times = []
    w = None
    for _ in range(reps):
        t0 = time.perf_counter()
        w = op.apply(u)
        times.append(time.perf_counter() - t0)
    return sp.size, times, w


def time_other_configs(fem, _capi, ctx, stream, dev, peak):
    """Device-resident apply of the other BASELINE configs on one GPU (not bench lines of their own: they explain where
    the remaining kernels stand).  C4: DG Q5 48^3 SIPG Laplace, C5: DG Q3 133^3 advection-diffusion (150.6 M dofs, one
    GPU's share of the 8-GPU weak-scaling config), both through the slab Kronecker kernel; C3/C1 Lagrange applies."""
    import torch
    out = {}

    def timed(op, size, reps, linear):
        us = [torch.rand(size, dtype=torch.float64, device=dev) * 2 - 1 for _ in range(3)]
        ws = [torch.empty(size, dtype=torch.float64, device=dev) for _ in range(3)]
        for i in range(3):
            op.apply_dev(us[i % 3].data_ptr(), ws[i % 3].data_ptr(), linear)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for i in range(reps):
            op.apply_dev(us[i % 3].data_ptr(), ws[i % 3].data_ptr(), linear)
        e1.record(stream)
        torch.cuda.synchronize()
        del us, ws
        return e0.elapsed_time(e1) * 1e-3 / reps

    dg = (("C4 DG Q5 48^3 SIPG Laplace apply", 48, 5, dict(eps=1.0, b=(0.0, 0.0, 0.0), beta=500.0, dirichlet_mask=0b111111, data=2), [0.0] * 3, 20),
          ("C5 DG Q3 133^3 advection-diffusion apply (one GPU's share)", 133, 3, dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=180.0, dirichlet_mask=0b000011, data=1), [-1.0] * 3, 10))
    for name, cells, order, model, lo, reps in dg:
        grid = fem.structuredGrid(lo, [1.0] * 3, [cells] * 3, ctx=ctx)
        space = fem.space.dglegendre(grid, order=order, hierarchical=True)
        op = fem.operator.galerkin(space, **model)
        t_aff, t_lin = timed(op, space.size, reps, False), timed(op, space.size, reps, True)
        n = order + 1
        out[name] = {"dofs": space.size, "kernel": "dg_kronecker_slab_kernel<%d>" % n, "affine_ms": t_aff * 1e3, "dofs_per_s": space.size / t_aff,
                     "linear_ms": t_lin * 1e3, "linear_dofs_per_s": space.size / t_lin, "frac_hbm_16B": 16 * space.size / t_lin / 1e9 / peak,
                     "fp64_tflops_kronecker_form": 2 * 9 * n ** 4 * cells ** 3 / t_lin / 1e12, "fp64_peak_tflops_measured": 37.1}
        del op, space, grid
        torch.cuda.empty_cache()
    for name, dim, cells, order in (("C3 Poisson P2 Lagrange 3D 128^3 apply", 3, 128, 2), ("C1 Poisson P1 Lagrange 2D 256^2 apply", 2, 256, 1),
                                    ("C1 scaled up: P1 Lagrange 2D 4096^2 apply", 2, 4096, 1)):
        grid = fem.structuredGrid([0.0] * dim, [1.0] * dim, [cells] * dim, ctx=ctx)
        space = fem.space.lagrange(grid, order=order)
        op = fem.operator.galerkin(space, eps=1.0, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
        t_lin = timed(op, space.size, 20, True)
        out[name] = {"dofs": space.size, "kernel": "lagrange_kronecker_kernel<%d>" % order, "linear_ms": t_lin * 1e3, "linear_dofs_per_s": space.size / t_lin,
                     "frac_hbm_16B": 16 * space.size / t_lin / 1e9 / peak}
        del op, space, grid
        torch.cuda.empty_cache()
    return out


def cpu_kronecker_apply(threads, reps):
    """times the Kronecker-form CPU apply of the oracle (fem_oracle.cpp: fo_kron_apply; 1-D matrices probed from the dense
    loop) on the FULL 64^3 mesh: the "sum-factorised, to be fair to the CPU" comparator of BASELINE.md section 3.1"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as ol
    k = ol.KroneckerCpu([CELLS] * 3, [-1.0] * 3, [1.0] * 3, ol.DG_LEGENDRE_HIER, ORDER, threads=threads, **MODEL)
    u = np.random.default_rng(SEED).uniform(-1, 1, k.space.size)
    w = np.zeros(k.space.size)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        k.apply(u, out=w)
        times.append(time.perf_counter() - t0)
    return k.space.size, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    # bounded sample: a 64 x 64 x nz slab of the 64^3 mesh per step (same elements, same integrands).  nz = 16 (a quarter of
    # the work) unless --steps is so large that the whole run would take more than ~90 s: then the slab is thinned (the
    # metric is DoF/s, the sample size is reported in config.sample_cells)
    nz = 16
    _, t_probe, _ = cpu_reference_apply(threads, [CELLS, CELLS, nz], 2)
    est = min(t_probe) * (args.warmup + args.steps)
    while est > 90.0 and nz > 2:
        nz //= 2
        est /= 2
    cells = [CELLS, CELLS, nz]
    ndof, times, _ = cpu_reference_apply(threads, cells, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = ndof * len(timed) / total
    line = {
        "impl": "reference", "metric": "operator-apply DoF/s (FP64)", "value": value, "unit": "DoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 advection-diffusion DG Q2 3D 64^3 explicit operator apply (CPU restatement of the reference; "
                               "DUNE-FEM itself cannot be built in this image)", "sample_cells": cells, "dofs_per_step": ndof},
        "cpu_baseline": {"value": value, "unit": "DoF/s", "cores": threads, "kind": "port",
                         "sample": f"{len(timed)} applies of a {cells[0]}x{cells[1]}x{cells[2]} slab of the 64^3 mesh, {threads} threads"},
        "e2e": {"value": value, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--kernel", default="auto", choices=["auto", "quadrature", "kronecker"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cg", action="store_true", help="skip the CG s/iteration measurement (BASELINE configs 1 and 3)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the apply timings of BASELINE configs 3, 4, 5")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import dune_fem_b200 as fem
    from dune_fem_b200 import _capi
    from dune_fem_b200.grid import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # a dedicated (non-default) stream shared by torch and the library: CUDA events below are recorded on the very
    # stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = Context(device=local_rank, stream=stream.cuda_stream)
    proc = proc_grid(world)
    if world > 1:
        ids = [Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_nccl(ids[0], rank, world)
    n_global = [CELLS * p for p in proc]
    hi = [-1.0 + 2.0 * p for p in proc]                 # every rank keeps a [-1,1]^3-sized box: h is the same at all N
    grid = fem.structuredGrid([-1.0, -1.0, -1.0], hi, n_global, ctx=ctx, proc=proc if world > 1 else None, rank=rank)
    space = fem.space.dglegendre(grid, order=ORDER, hierarchical=True)
    kernel = {"auto": _capi.KERNEL_AUTO, "quadrature": _capi.KERNEL_QUADRATURE, "kronecker": _capi.KERNEL_KRONECKER}[args.kernel]
    op = fem.operator.galerkin(space, kernel=kernel, **MODEL)
    ndof_local = CELLS ** 3 * (ORDER + 1) ** 3            # owned dofs per rank
    ndof_total = ndof_local * world

    # rotating device buffers: 6 (u, w) pairs of 2 x 56.6 MB each = 680 MB >> 126 MB L2
    npairs = 6
    rng = np.random.default_rng(SEED + rank)
    us = [torch.from_numpy(rng.uniform(-1, 1, space.size)).to(dev) for _ in range(npairs)]
    ws = [torch.empty(space.size, dtype=torch.float64, device=dev) for _ in range(npairs)]
    if world > 1:
        for u in us:
            op.communicate_dev(u.data_ptr())             # consistent ghost copies of the input, as the reference assumes

    uptr = [t.data_ptr() for t in us]
    wptr = [t.data_ptr() for t in ws]

    def step(i, linear=False):
        op.apply_dev(uptr[i % npairs], wptr[i % npairs], linear)

    host_issue_us = [0.0]

    def timed(nsteps, linear=False):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        for i in range(nsteps):
            step(i, linear)
        host_issue_us[0] = 1e6 * (time.perf_counter() - h0) / nsteps
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for i in range(max(args.warmup, 3)):
        step(i)
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms = timed(args.steps)
    host_us = host_issue_us[0]
    sampler.stop_flag = True
    sampler.join()
    ms_linear = timed(args.steps, linear=True)
    tinfo = op.timing()            # (also switches the per-apply timing events on: keep it after the timed loops)
    # multi-GPU diagnostics: the halo exchange alone and the local kernels alone (same stream, same buffers)
    diag = None
    if world > 1:
        def loop(fn, n):
            torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for i in range(n):
                fn(i)
            b.record(stream)
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return 1e3 * float(t.item()) / n
        op.apply_dev(uptr[0], wptr[0])
        tinfo = op.timing()        # now with event times of a distributed apply
        ex_us = loop(lambda i: op.communicate_dev(ws[i % npairs].data_ptr()), 200)
        op.setCommunicate(False)
        comp_us = loop(lambda i: step(i), 200)
        op.setCommunicate(True)
        diag = {"exchange_only_us": ex_us, "local_kernels_only_us": comp_us}

    value = ndof_total * args.steps / (ms * 1e-3)
    launches = tinfo["launches_per_apply"] * args.steps

    # end to end through the host-pointer C ABI call (pinned host dof vectors, H2D + kernel + D2H per step)
    e2e = None
    uh = torch.from_numpy(rng.uniform(-1, 1, space.size)).pin_memory()
    wh = torch.empty(space.size, dtype=torch.float64).pin_memory()
    un, wn = uh.numpy(), wh.numpy()
    op(un, wn)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        op(un, wn)
    t1 = time.perf_counter()
    e2e_t = torch.tensor([t1 - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e = {"value": ndof_total * args.e2e_steps / float(e2e_t.item()), "unit": "DoF/s", "h2d_bytes_per_step": 8 * space.size,
           "d2h_bytes_per_step": 8 * space.size, "steps": args.e2e_steps,
           "api": "b200fem_operator_apply(op, u_host, w_host) with pinned host buffers"}

    # CG seconds/iteration (second half of BASELINE's metric): C3 Poisson P2 Lagrange 3D 128^3 and C1 P1 2D 256^2,
    # 100 fixed iterations of the device-resident CG (tolerance 0 so that no iteration is skipped)
    cg = None
    if world == 1 and not args.no_cg:
        import ctypes as C
        cg = {}
        for name, dim, cells, order in (("C3 Poisson P2 Lagrange 3D 128^3", 3, 128, 2), ("C1 Poisson P1 Lagrange 2D 256^2", 2, 256, 1)):
            g = fem.structuredGrid([0.0] * dim, [1.0] * dim, [cells] * dim, ctx=ctx)
            sp = fem.space.lagrange(g, order=order)
            lop = fem.operator.galerkin(sp, eps=1.0, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
            bt = torch.from_numpy(lop.loadVector()).to(dev)
            mask, gv = lop.dirichlet()
            x0 = torch.from_numpy(np.where(mask, gv, 0.0)).to(dev)
            iters, its = 100, C.c_int()
            for rep in range(2):
                xt = x0.clone()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _capi.check(_capi.lib().b200fem_cg_solve_dev(lop.handle, C.c_void_p(bt.data_ptr()), C.c_void_p(xt.data_ptr()), 0.0, iters,
                                                             _capi.TOL_ABSOLUTE, C.byref(its), None))
                e1.record(stream)
                torch.cuda.synchronize()
                cg_ms = e0.elapsed_time(e1)
            cg[name] = {"dofs": sp.size, "iterations": abs(its.value), "s_per_iteration": cg_ms * 1e-3 / iters,
                        "dofs_per_s": sp.size * iters / (cg_ms * 1e-3), "schedule": ("one cooperative kernel launch per 16 iterations (grid-wide barriers, cg_coop2d.cuh)" if dim == 2 else
                                     "CUDA graph of 16 iterations, %d launches per iteration" % (lop.timing()["launches_per_apply"] + 3))}
            del bt, x0, xt, lop, sp, g

    other = None
    if world == 1 and not args.no_other_configs:
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs") or 6650.0
        except Exception:
            pk = 6650.0
        other = time_other_configs(fem, _capi, ctx, stream, dev, pk)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("hbm_gbs")
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    if not peak:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_launch_s = ms * 1e-3 / args.steps
    achieved = ALGORITHMIC_BYTES_PER_DOF * ndof_local / per_launch_s / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(
            "dg_kronecker" if tinfo["kernel"] == _capi.KERNEL_KRONECKER else "dg_quadrature")
    except Exception:
        pass
    kernel_name = {1: "dg_quadrature_kernel<3>", 2: {"march": "dg_kronecker_march_kernel<3> (z-marching, TMA planes)"}.get(os.environ.get("B200FEM_KRON_VARIANT", "march"), "dg_kronecker_tensor_kernel<3> (TMA tensor tiles)")}.get(tinfo["kernel"], "?")
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": kernel_name, "peak_source": peak_src, "algorithmic_bytes_per_dof": ALGORITHMIC_BYTES_PER_DOF,
                "compulsory_bytes_per_dof": 24.0, "achieved_compulsory_gbs": 24.0 * ndof_local / per_launch_s / 1e9,
                "frac_compulsory": 24.0 * ndof_local / per_launch_s / 1e9 / peak,
                "note": "the affine step also streams the precomputed load vector b (8 B/dof) that the 16 B/dof figure does not count "
                        "(compulsory DRAM traffic of w = A u - b is 24 B/dof: frac_compulsory); the homogeneous apply A u moves "
                        "exactly 16 B/dof, see linear_apply"}
    lin_s = ms_linear * 1e-3 / args.steps
    linear_apply = {"value": ndof_total / lin_s, "unit": "DoF/s", "ms_per_step": ms_linear / args.steps,
                    "achieved_gbs": ALGORITHMIC_BYTES_PER_DOF * ndof_local / lin_s / 1e9,
                    "frac": ALGORITHMIC_BYTES_PER_DOF * ndof_local / lin_s / 1e9 / peak}

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        cells = [CELLS, CELLS, 16]
        nd, times, _ = cpu_reference_apply(threads, cells, 3)
        best = min(times[1:])
        cpu_baseline = {"value": nd / best, "unit": "DoF/s", "cores": threads, "kind": "port",
                        "sample": f"best of 2 applies of a {cells[0]}x{cells[1]}x{cells[2]} slab of the 64^3 mesh, {threads} threads"}
        try:        # the fair comparator: same Kronecker arithmetic as the GPU kernel, homogeneous part A u, full mesh
            ndk, tk = cpu_kronecker_apply(threads, 4)
            cpu_baseline["kronecker_form"] = {"value": ndk / min(tk[1:]), "unit": "DoF/s", "cores": threads, "kind": "port (Kronecker form, matrices probed from the dense loop)",
                                              "sample": f"best of 3 applies A u on the full 64^3 mesh, {threads} threads"}
        except Exception as ex:      # never let the extra comparator take the bench line down
            cpu_baseline["kronecker_form"] = {"error": str(ex)}

    line = {
        "metric": "operator-apply DoF/s (FP64)", "value": value, "unit": "DoF/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 advection-diffusion DG Q2 (dglegendre hierarchical) 3D 64^3 cells per GPU, explicit operator apply "
                               "w = L[u] = A u - b (pydemo/advectiondiffusion.py integrands, eps=1e-5)",
                   "dofs_per_gpu": ndof_local, "process_grid": proc, "halo_exchange": world > 1,
                   "l2": f"{npairs} rotating (u,w) buffer pairs = {npairs * 2 * 8 * space.size / 1e6:.0f} MB > 126 MB L2",
                   "kernel": kernel_name},
        "roofline": roofline, "linear_apply": linear_apply, "cg": cg, "other_configs": other, "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches, "host_issue_us_per_step": host_us, "exchange_ms_last": tinfo.get("last_exchange_ms"), "apply_ms_last": tinfo.get("last_apply_ms"), "multi_gpu_diag": diag,
        "clocks": sampler.result(),
    }
    print(json.dumps(line), flush=True)
    del op
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
