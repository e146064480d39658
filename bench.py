#!/usr/bin/env python
"""bench.py -- operator-apply throughput (DoF/s, FP64) of the matrix-free Galerkin operator on B200.

Headline workload (BASELINE.json configs[1], "C2"): advection-diffusion, DG Q2 (dglegendre, hierarchical), 3-D cube
[-1,1]^3 with 64^3 cells per GPU, SIPG + upwind integrands of pydemo/advectiondiffusion.py, explicit operator
apply w = L[u] = A u - b (complete affine operator; b = -L[0] precomputed once and streamed by the kernel).
A "step" is one operator application over one synthetic dof vector (u ~ U(-1,1), PCG64 seed 20261017); with N > 1
every rank owns a 64^3 box (weak scaling) and a step ends with the Copy exchange of w's ghost layers.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU, torchrun for N>1)
  python bench.py --impl reference --steps K --warmup W    # the CPU restatement of the reference, all host threads

Prints ONE JSON line (rank 0).  Besides the contract keys it carries, at every N: the other BASELINE configs
(`configs`: C1/C3/C4/C5 applies with their own roofline blocks), CG seconds per iteration (`cg`; C3 is run strong-scaled
over the N ranks), the C5 weak-scaling apply (`weak_scaling_c5`, ~150 M dofs per GPU) and, for N > 1, a parity block
(small Q2 / Q3 / P2 applies and 10 CG iterations against the single-domain CPU oracle, checker only).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
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
CG_BYTES_PER_DOF = 72.0                # apply 16 + fused update 56 (dot fused into the apply / update sweeps)
FP64_PEAK_TFLOPS = 37.1                # measured on this pool's B200: profiles/micro/dfma_bench.cu / dmma_bench.cu (DFMA and DMMA agree)
FP64_PEAK_SOURCE = "measured: profiles/micro/dmma_bench_b200.txt (DFMA 37.1, mma.sync.m8n8k4.f64 37.1 TFLOP/s)"


def dg_proc_grid(n):
    return {1: [1, 1, 1], 2: [1, 1, 2], 4: [1, 2, 2], 8: [1, 2, 4]}[n]   # x (the contiguous axis) is never split


def lagrange_proc_grid(n):
    return {1: [1, 1, 1], 2: [1, 1, 2], 4: [1, 2, 2], 8: [2, 2, 2]}[n]


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


# --------------------------------------------------------------------------------------------------------------------
# CPU legs (checker / baseline only: the one place besides tests/ and smoke() that touches oracle/)
def oracle_module():
    """the oracle, rebuilt with -O3 -march=native for THIS host when a compiler is present"""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    native = os.path.join(ROOT, "oracle", "_native", "libfem_oracle.so")
    try:
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "-s", "native"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        if os.path.exists(native):
            os.environ["B200FEM_ORACLE_SO"] = native
    except Exception:
        pass
    import oracle_lib as ol
    return ol, ("-O3 -march=native" if os.environ.get("B200FEM_ORACLE_SO") else "-O3 -march=x86-64-v3")


def cpu_reference_apply(ol, threads, cells, reps):
    """times the CPU oracle (restatement of the reference's GalerkinOperator::evaluate) on a cells[0] x cells[1] x cells[2] slab
    of the 64^3 mesh (same cell size, same integrands)"""
    h = 2.0 / CELLS
    hi = [-1 + cells[0] * h, -1 + cells[1] * h, -1 + cells[2] * h]
    sp = ol.Space(cells, [-1, -1, -1], hi, ol.DG_LEGENDRE_HIER, ORDER)
    op = ol.Operator(sp, skeleton=True, boundary=True, threads=threads, **MODEL)
    u = np.random.default_rng(SEED).uniform(-1, 1, sp.size)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        op.apply(u)
        times.append(time.perf_counter() - t0)
    return sp.size, times


def cpu_kronecker_apply(ol, threads, reps):
    """Kronecker-form CPU apply of the oracle (fo_kron_apply; 1-D matrices probed from the dense loop) on the full 64^3 mesh:
    the "sum-factorised, to be fair to the CPU" comparator of BASELINE.md section 3.1"""
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
    ol, flags = oracle_module()
    threads = os.cpu_count() or 1
    # the full 64^3 mesh of the headline workload per step; only if --steps is so large that the run would exceed ~2 minutes
    # the slab is thinned (the metric is DoF/s; the sample is reported)
    nz = CELLS
    _, t_probe = cpu_reference_apply(ol, threads, [CELLS, CELLS, 8], 2)
    est = min(t_probe) * (nz / 8) * (args.warmup + args.steps)
    while est > 120.0 and nz > 8:
        nz //= 2
        est /= 2
    cells = [CELLS, CELLS, nz]
    ndof, times = cpu_reference_apply(ol, threads, cells, args.warmup + args.steps)
    timed = times[args.warmup:]
    total = sum(timed)
    value = ndof * len(timed) / total
    line = {
        "impl": "reference", "metric": "operator-apply DoF/s (FP64)", "value": value, "unit": "DoF/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 advection-diffusion DG Q2 3D 64^3 explicit operator apply (CPU restatement of the reference; "
                               "DUNE-FEM itself cannot be built in this image)", "sample_cells": cells, "dofs_per_step": ndof,
                   "same_config": nz == CELLS, "compiler_flags": flags},
        "cpu_baseline": {"value": value, "unit": "DoF/s", "cores": threads, "kind": "port",
                         "sample": f"{len(timed)} applies of the {cells[0]}x{cells[1]}x{cells[2]} mesh, {threads} threads, {flags}"},
        "e2e": {"value": value, "unit": "DoF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------------------------------
def load_peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        if p.get("hbm_gbs"):
            return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        return {}


def roofline_block(kernel, ndof, seconds, peak, peak_src, traffic, flops=None, flop_formula=None, bytes_per_dof=ALGORITHMIC_BYTES_PER_DOF):
    """HBM fraction at the contract's algorithmic bytes AND FP64 fraction with the stated flop formula"""
    achieved = bytes_per_dof * ndof / seconds / 1e9
    blk = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic.get(kernel),
           "kernel": kernel, "peak_source": peak_src, "algorithmic_bytes_per_dof": bytes_per_dof}
    if flops is not None:
        tf = flops / seconds / 1e12
        blk["fp64"] = {"achieved_tflops": tf, "peak_tflops": FP64_PEAK_TFLOPS, "frac": tf / FP64_PEAK_TFLOPS, "flops_per_launch": flops,
                       "formula": flop_formula, "peak_source": FP64_PEAK_SOURCE}
    return blk


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--kernel", default="auto", choices=["auto", "quadrature", "kronecker"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cg", action="store_true", help="skip the CG s/iteration measurements (BASELINE configs 1 and 3)")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the apply timings of BASELINE configs 1, 3, 4, 5")
    ap.add_argument("--no-parity", action="store_true", help="skip the multi-GPU parity block")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import dune_fem_b200 as fem
    from dune_fem_b200 import _capi
    from dune_fem_b200.grid import Context, partition_box

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
    if world > 1:
        ids = [Context.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_nccl(ids[0], rank, world)
    peak, peak_src = load_peaks()
    traffic = load_traffic()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed_loop(fn, nsteps):
        """EXACTLY nsteps calls between two CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks (ms)"""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        h0 = time.perf_counter()
        for i in range(nsteps):
            fn(i)
        host_us = 1e6 * (time.perf_counter() - h0) / max(nsteps, 1)
        e1.record(stream)
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)), host_us

    # ------------------------------------------------------------------------------------------ headline: C2, weak scaling
    proc = dg_proc_grid(world)
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

    # warm-up: EVERY buffer pair with BOTH kernel variants (affine and homogeneous), so that tensor-map encoding, module
    # loading and function attributes are all outside the timed regions -- also under the driver's short --steps 20 run
    nwarm = max(args.warmup, 3, 2 * npairs)
    for i in range(nwarm):
        step(i, linear=False)
    for i in range(max(3, npairs)):
        step(i, linear=True)
    for i in range(npairs):
        step(i, linear=False)
    sampler = ClockSampler(local_rank)
    if os.environ.get("B200FEM_BENCH_STEPTIMES"):      # diagnostic: one event pair per step (not a bench value)
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
        barrier()
        evs[0].record(stream)
        for i in range(args.steps):
            step(i, False)
            evs[i + 1].record(stream)
        barrier()
        sys.stderr.write("step times (us): " + " ".join(f"{1e3 * evs[i].elapsed_time(evs[i + 1]):.1f}" for i in range(args.steps)) + "\n")
    if not os.environ.get("B200FEM_BENCH_NOSAMPLER"):
        sampler.start()
    ms, host_us = timed_loop(lambda i: step(i, False), args.steps)
    sampler.stop_flag = True
    if sampler.is_alive():
        sampler.join()
    for i in range(npairs):
        step(i, linear=True)
    ms_linear, _ = timed_loop(lambda i: step(i, True), args.steps)

    # multi-GPU diagnostics: the halo exchange alone and the local kernels alone (same stream, same buffers)
    diag = None
    if world > 1:
        ex_ms, _ = timed_loop(lambda i: op.communicate_dev(wptr[i % npairs]), 200)
        op.setCommunicate(False)
        for i in range(npairs):
            step(i)
        comp_ms, _ = timed_loop(lambda i: step(i), 200)
        op.setCommunicate(True)
        for i in range(npairs):
            step(i)
        diag = {"exchange_only_us": 1e3 * ex_ms / 200, "local_kernels_only_us": 1e3 * comp_ms / 200,
                "transport": "peer memory (fused into the marching kernel)" if ctx.peer_memory else "nccl send/recv"}

    tinfo = op.timing()            # (switches the per-apply timing events on, which serialises launches: only after ALL timed loops)
    launches = tinfo["launches_per_apply"] * args.steps
    value = ndof_total * args.steps / (ms * 1e-3)
    per_launch_s = ms * 1e-3 / args.steps
    kernel_name = {1: "dg_quadrature_kernel<3>", 2: "dg_kronecker_march_kernel<3>"}.get(tinfo["kernel"], "?")
    q2_flops = 2 * 9 * 3 ** 4 * CELLS ** 3
    roofline = roofline_block(kernel_name, ndof_local, per_launch_s, peak, peak_src, traffic, q2_flops, "Kronecker form: 2 * 9 n^4 flop per element (n = 3)")
    roofline.update({"compulsory_bytes_per_dof": 24.0, "frac_compulsory": 24.0 * ndof_local / per_launch_s / 1e9 / peak,
                     "note": "the affine step also streams the precomputed load vector b (8 B/dof) that the 16 B/dof figure does not count; "
                             "the homogeneous apply A u moves exactly 16 B/dof, see linear_apply"})
    lin_s = ms_linear * 1e-3 / args.steps
    linear_apply = {"value": ndof_total / lin_s, "unit": "DoF/s", "ms_per_step": ms_linear / args.steps,
                    "achieved_gbs": ALGORITHMIC_BYTES_PER_DOF * ndof_local / lin_s / 1e9,
                    "frac": ALGORITHMIC_BYTES_PER_DOF * ndof_local / lin_s / 1e9 / peak}

    # end to end through the host-pointer C ABI call (pinned host dof vectors, H2D + kernel + D2H per step)
    uh = torch.from_numpy(rng.uniform(-1, 1, space.size)).pin_memory()
    wh = torch.empty(space.size, dtype=torch.float64).pin_memory()
    un, wn = uh.numpy(), wh.numpy()
    op(un, wn)
    op(un, wn)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        op(un, wn)
    e2e_t = max_over_ranks(time.perf_counter() - t0)
    e2e = {"value": ndof_total * args.e2e_steps / e2e_t, "unit": "DoF/s", "h2d_bytes_per_step": 8 * space.size,
           "d2h_bytes_per_step": 8 * space.size, "steps": args.e2e_steps,
           "api": "b200fem_operator_apply(op, u_host, w_host) with pinned host buffers"}
    del uh, wh, us, ws, op, space, grid
    torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------ helpers for the other configs
    def timed_apply(o, size, reps, linear, nbuf=3):
        uu = [torch.rand(size, dtype=torch.float64, device=dev) * 2 - 1 for _ in range(nbuf)]
        ww = [torch.empty(size, dtype=torch.float64, device=dev) for _ in range(nbuf)]
        if world > 1:
            for t in uu:
                o.communicate_dev(t.data_ptr())
        for i in range(2 * nbuf):
            o.apply_dev(uu[i % nbuf].data_ptr(), ww[i % nbuf].data_ptr(), linear)
        t_ms, _ = timed_loop(lambda i: o.apply_dev(uu[i % nbuf].data_ptr(), ww[i % nbuf].data_ptr(), linear), reps)
        del uu, ww
        return t_ms * 1e-3 / reps

    # ------------------------------------------------------------------------------------------ C5: weak scaling, ~150 M dofs per GPU
    weak_c5 = None
    if not args.no_other_configs:
        c5 = 133
        g = fem.structuredGrid([-1.0] * 3, [-1.0 + 2.0 * p for p in proc], [c5 * p for p in proc], ctx=ctx, proc=proc if world > 1 else None, rank=rank)
        sp = fem.space.dglegendre(g, order=3, hierarchical=True)
        o = fem.operator.galerkin(sp, eps=1e-5, b=(1.0, 0.0, 0.0), beta=180.0, dirichlet_mask=0b000011, data=1)
        nd = c5 ** 3 * 64
        t_aff = timed_apply(o, sp.size, 10, False, nbuf=2)
        t_lin = timed_apply(o, sp.size, 10, True, nbuf=2)
        t_local = None
        if world > 1:            # the same rank-local apply without the halo exchange: what the exchange costs at this size
            o.setCommunicate(False)
            t_local = timed_apply(o, sp.size, 10, False, nbuf=2)
            o.setCommunicate(True)
        blk = roofline_block("dg_kronecker_mma_kernel (FP64 tensor cores, mma.sync.m8n8k4.f64)", nd, t_lin, peak, peak_src, traffic, 2 * 9 * 4 ** 4 * c5 ** 3, "Kronecker form: 2 * 9 n^4 flop per element (n = 4)")
        weak_c5 = {"workload": "C5 DG Q3 advection-diffusion, 133^3 cells = 150.6 M dofs per GPU, apply incl. halo exchange", "n_gpus": world,
                   "process_grid": proc, "dofs_per_gpu": nd, "affine_ms": t_aff * 1e3, "linear_ms": t_lin * 1e3,
                   "value": nd * world / t_aff, "linear_value": nd * world / t_lin, "unit": "DoF/s", "per_gpu_dofs_per_s": nd / t_aff,
                   "roofline_linear": blk, "transport": ("peer memory, send + receive kernels" if ctx.peer_memory else "nccl") if world > 1 else None,
                   "local_kernels_only_ms": None if t_local is None else t_local * 1e3,
                   "note": "BASELINE config 5 (the weak-scaling config): parallel efficiency at N GPUs = per_gpu_dofs_per_s of this line / per_gpu_dofs_per_s of the N = 1 line; "
                           "the headline value stays on config 2 at every N so that the driver's own efficiency is computed on one workload"}
        del o, sp, g
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------ CG seconds per iteration
    cg = None
    if not args.no_cg:
        cg = {}
        cases = [("C3 Poisson P2 Lagrange 3D 128^3", 3, 128, 2)]
        if world == 1:
            cases.append(("C1 Poisson P1 Lagrange 2D 256^2", 2, 256, 1))
        for name, dim, cells, order in cases:
            lp = lagrange_proc_grid(world)[:dim]
            g = fem.structuredGrid([0.0] * dim, [1.0] * dim, [cells] * dim, ctx=ctx, proc=lp if world > 1 else None, rank=rank)
            sp = fem.space.lagrange(g, order=order)
            lop = fem.operator.galerkin(sp, eps=1.0, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
            bt = torch.from_numpy(lop.loadVector()).to(dev)
            mask, gv = lop.dirichlet()
            x0 = torch.from_numpy(np.where(mask, gv, 0.0)).to(dev)
            iters, its = 100, C.c_int()
            cg_ms = None
            for rep in range(3):
                xt = x0.clone()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                _capi.check(_capi.lib().b200fem_cg_solve_dev(lop.handle, C.c_void_p(bt.data_ptr()), C.c_void_p(xt.data_ptr()), 0.0, iters,
                                                             _capi.TOL_ABSOLUTE, C.byref(its), None))
                e1.record(stream)
                barrier()
                cg_ms = max_over_ranks(e0.elapsed_time(e1))
            ndof_global = (order * cells + 1) ** dim
            s_it = cg_ms * 1e-3 / iters
            launches_it = lop.timing()["launches_per_apply"] + 3
            cg[name] = {"dofs": ndof_global, "n_gpus": world, "process_grid": lp if world > 1 else None, "iterations": abs(its.value), "s_per_iteration": s_it,
                        "dofs_per_s": ndof_global / s_it,
                        "roofline": roofline_block("cg_iteration", ndof_global / world, s_it, peak, peak_src, traffic, bytes_per_dof=CG_BYTES_PER_DOF),
                        "schedule": ("one cooperative kernel launch per 16 iterations (grid-wide barriers, cg_coop2d.cuh)" if dim == 2 and world == 1 else
                                     "CUDA graph of 16 iterations, %d launches per iteration%s" % (launches_it, "" if world == 1 else " (halo exchange and global sums inside the kernels, peer memory)" if ctx.peer_memory else " (NCCL: not captured)"))}
            del bt, x0, xt, lop, sp, g
            torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------ other single-GPU configs
    other = None
    if world == 1 and not args.no_other_configs:
        other = {}
        c4 = 48
        g = fem.structuredGrid([0.0] * 3, [1.0] * 3, [c4] * 3, ctx=ctx)
        sp = fem.space.dglegendre(g, order=5, hierarchical=True)
        o = fem.operator.galerkin(sp, eps=1.0, b=(0.0, 0.0, 0.0), beta=500.0, dirichlet_mask=0b111111, data=2)
        t_aff, t_lin = timed_apply(o, sp.size, 20, False), timed_apply(o, sp.size, 20, True)
        other["C4 DG Q5 48^3 SIPG Laplace apply"] = {
            "dofs": sp.size, "affine_ms": t_aff * 1e3, "linear_ms": t_lin * 1e3, "dofs_per_s": sp.size / t_aff, "linear_dofs_per_s": sp.size / t_lin,
            "roofline": roofline_block("dg_kronecker_slab_kernel<6>", sp.size, t_lin, peak, peak_src, traffic, 2 * 9 * 6 ** 4 * c4 ** 3, "Kronecker form: 2 * 9 n^4 flop per element (n = 6)")}
        del o, sp, g
        torch.cuda.empty_cache()
        # the generic quadrature kernel on C2 (what a non-linear / variable-coefficient form runs through)
        g = fem.structuredGrid([-1.0] * 3, [1.0] * 3, [CELLS] * 3, ctx=ctx)
        sp = fem.space.dglegendre(g, order=ORDER, hierarchical=True)
        o = fem.operator.galerkin(sp, kernel=_capi.KERNEL_QUADRATURE, **MODEL)
        t_q = timed_apply(o, sp.size, 10, False)
        other["C2 through the generic quadrature kernel"] = {
            "dofs": sp.size, "affine_ms": t_q * 1e3, "dofs_per_s": sp.size / t_q,
            "roofline": roofline_block("dg_quadrature_kernel<3>", sp.size, t_q, peak, peak_src, traffic, 7.2e3 * CELLS ** 3, "SURVEY.md 8(d) sum-factorised quadrature model: 7.2 kflop per Q2 element")}
        del o, sp, g
        torch.cuda.empty_cache()
        for name, dim, cells, order in (("C3 Poisson P2 Lagrange 3D 128^3 apply", 3, 128, 2), ("C1 Poisson P1 Lagrange 2D 256^2 apply", 2, 256, 1),
                                        ("C1 scaled up: P1 Lagrange 2D 4096^2 apply", 2, 4096, 1)):
            g = fem.structuredGrid([0.0] * dim, [1.0] * dim, [cells] * dim, ctx=ctx)
            sp = fem.space.lagrange(g, order=order)
            o = fem.operator.galerkin(sp, eps=1.0, data=2, dirichlet_mask=(1 << (2 * dim)) - 1, strong_dirichlet=True)
            t_lin = timed_apply(o, sp.size, 20, True)
            W = 2 * order + 1
            fl = sp.size * 2.0 * (5 * W if dim == 3 else 3 * W)        # z: 2W, y: 3W (2W in 2-D... counted as 3W upper bound), x: 2W FMA per node
            other[name] = {"dofs": sp.size, "linear_ms": t_lin * 1e3, "linear_dofs_per_s": sp.size / t_lin,
                           "roofline": roofline_block("lagrange_lattice_kernel<%d>" % order, sp.size, t_lin, peak, peak_src, traffic, fl, "lattice stencil: 2 * (7 W) flop per node in 3-D (W = 2k+1), 2 * 3 W in 2-D")}
            del o, sp, g
            torch.cuda.empty_cache()

        # (f)4: the same P2 space on an UNSTRUCTURED cube mesh (vertex / element arrays, distorted vertices): index arrays and
        # per-element geometry instead of closed forms -- SURVEY.md 8(d) reports this traffic apart from the headline figure
        cu = 64
        ax = np.linspace(0.0, 1.0, cu + 1)
        X = np.stack(np.meshgrid(ax, ax, ax, indexing="ij"), axis=-1)                       # X[i0, i1, i2] = vertex coordinates
        vid = (np.arange(cu + 1)[:, None, None] + (cu + 1) * (np.arange(cu + 1)[None, :, None] + (cu + 1) * np.arange(cu + 1)[None, None, :]))
        coords = np.zeros(((cu + 1) ** 3, 3))
        coords[vid.ravel()] = X.reshape(-1, 3)
        coords += np.random.default_rng(20261017).uniform(-0.15, 0.15, coords.shape) / cu
        e0, e1, e2 = np.meshgrid(np.arange(cu), np.arange(cu), np.arange(cu), indexing="ij")
        order_e = np.argsort((e0 + cu * (e1 + cu * e2)).ravel())                              # elements x fastest
        cubes = np.stack([vid[e0 + (v & 1), e1 + ((v >> 1) & 1), e2 + (v >> 2)].ravel() for v in range(8)], axis=1)[order_e].astype(np.int64)
        g = fem.unstructuredGrid(coords, cubes, ctx=ctx)
        sp = fem.space.lagrange(g, order=2)
        o = fem.operator.galerkin(sp, eps=1.0, data=2, dirichlet_mask=1, strong_dirichlet=True)
        t_lin = timed_apply(o, sp.size, 20, True)
        per_elem = 4 * 27 + 24 * 8                                                          # index array + vertex coordinates, B per element
        bpd = 16.0 + per_elem * cu ** 3 / sp.size
        other["P2 Lagrange 3D 64^3 cells as an unstructured cube mesh (index arrays + per-element geometry) apply"] = {
            "dofs": sp.size, "elements": cu ** 3, "linear_ms": t_lin * 1e3, "linear_dofs_per_s": sp.size / t_lin, "launches_per_apply": o.timing()["launches_per_apply"],
            "roofline": roofline_block("lagrange_unstructured_kernel<3, 27>", sp.size, t_lin, peak, peak_src, traffic, 2.0 * cu ** 3 * 27 * (27 * 8 + 60),
                                       "dense tabulated contraction: 27 points x (27 x 4 evaluate + 27 x 4 axpy + ~60 geometry / integrand) FMA per element", bytes_per_dof=bpd),
            "note": "algorithmic bytes = 16 B/dof + 108 B of indices + 192 B of vertex coordinates per element (SURVEY.md 8d: reported apart from the headline)"}
        del o, sp, g
        torch.cuda.empty_cache()

    # ------------------------------------------------------------------------------------------ multi-GPU parity (checker only)
    parity = None
    if world > 1 and not args.no_parity:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import mgpu_check
        local = float("inf")
        try:
            parity = mgpu_check.parity_block(ctx, rank, world)
            local = parity["max_rel_err"]
        except Exception as ex:          # the bench line must survive a failing checker; the failure is reported in the line
            parity = {"error": repr(ex)}
        worst = max_over_ranks(local)
        parity["max_rel_err_all_ranks"] = worst
        parity["ok"] = bool(worst <= 1e-12)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    cpu_baseline = None
    if not args.no_cpu_baseline and world == 1:
        ol, flags = oracle_module()
        threads = os.cpu_count() or 1
        cells = [CELLS, CELLS, CELLS]
        nd, times = cpu_reference_apply(ol, threads, cells, 3)
        best = min(times[1:])
        cpu_baseline = {"value": nd / best, "unit": "DoF/s", "cores": threads, "kind": "port",
                        "sample": f"best of 2 applies of the full {cells[0]}x{cells[1]}x{cells[2]} mesh, {threads} threads, {flags}"}
        try:        # the fair comparator: same Kronecker arithmetic as the GPU kernel, homogeneous part A u, full mesh
            ndk, tk = cpu_kronecker_apply(ol, threads, 4)
            cpu_baseline["kronecker_form"] = {"value": ndk / min(tk[1:]), "unit": "DoF/s", "cores": threads, "kind": "port (Kronecker form, matrices probed from the dense loop)",
                                              "sample": f"best of 3 applies A u on the full 64^3 mesh, {threads} threads"}
        except Exception as ex:      # never let the extra comparator take the bench line down
            cpu_baseline["kronecker_form"] = {"error": str(ex)}

    line = {
        "metric": "operator-apply DoF/s (FP64)", "value": value, "unit": "DoF/s", "n_gpus": world, "steps": args.steps,
        "warmup": nwarm, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 advection-diffusion DG Q2 (dglegendre hierarchical) 3D 64^3 cells per GPU, explicit operator apply "
                               "w = L[u] = A u - b (pydemo/advectiondiffusion.py integrands, eps=1e-5)" + ("" if world == 1 else ", halo exchange of w included"),
                   "dofs_per_gpu": ndof_local, "process_grid": proc, "halo_exchange": world > 1,
                   "l2": f"{npairs} rotating (u,w) buffer pairs = {npairs * 2 * 8 * ndof_local / 1e6:.0f} MB > 126 MB L2",
                   "kernel": kernel_name},
        "roofline": roofline, "linear_apply": linear_apply, "cg": cg, "weak_scaling_c5": weak_c5, "configs": other, "parity": parity,
        "cpu_baseline": cpu_baseline, "e2e": e2e, "gpu_launches": launches,
        "host_issue_us_per_step": host_us, "multi_gpu_diag": diag, "clocks": sampler.result(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
