"""torchrun probe: what makes the local marching kernel slower inside a multi-process run?  Prints per-rank times of the SAME
single-box apply at successive stages of the distributed set-up."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import dune_fem_b200 as fem
from dune_fem_b200.grid import Context

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
MODEL = dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1)


def timed(op, size, tag, linear=False, n=400, comm_u=False):
    us = [torch.rand(size, dtype=torch.float64, device=dev) for _ in range(6)]
    ws = [torch.empty(size, dtype=torch.float64, device=dev) for _ in range(6)]
    if comm_u:
        for t in us:
            op.communicate_dev(t.data_ptr())
    for i in range(12):
        op.apply_dev(us[i % 6].data_ptr(), ws[i % 6].data_ptr(), linear)
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    h0 = time.perf_counter()
    for i in range(n):
        op.apply_dev(us[i % 6].data_ptr(), ws[i % 6].data_ptr(), linear)
    h1 = time.perf_counter()
    e1.record(stream); torch.cuda.synchronize()
    print(f"[rank {rank}] {tag}: {e0.elapsed_time(e1) * 1e3 / n:.2f} us/apply (host issue {(h1 - h0) * 1e6 / n:.1f} us)", flush=True)
    del us, ws


def single_box_op(ctx):
    g = fem.structuredGrid([-1.0] * 3, [1.0] * 3, [64] * 3, ctx=ctx)
    sp = fem.space.dglegendre(g, order=2, hierarchical=True)
    return fem.operator.galerkin(sp, **MODEL), sp


ctx0 = Context(device=lr, stream=stream.cuda_stream)
op0, sp0 = single_box_op(ctx0)
timed(op0, sp0.size, "1 single box, before any distributed set-up")
dist.init_process_group("nccl", device_id=dev)
t = torch.ones(1, device=dev); dist.all_reduce(t); torch.cuda.synchronize()
timed(op0, sp0.size, "2 single box, after torch NCCL init + all_reduce")
ctx = Context(device=lr, stream=stream.cuda_stream)
ids = [Context.nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
ctx.init_nccl(ids[0], rank, world)
timed(op0, sp0.size, "3 single box (old ctx), after library NCCL init + peer mappings")
op1, sp1 = single_box_op(ctx)
timed(op1, sp1.size, "4 single box on the distributed ctx")
proc = {2: [1, 1, 2], 4: [1, 2, 2], 8: [1, 2, 4]}[world]
g = fem.structuredGrid([-1.0] * 3, [-1.0 + 2.0 * p for p in proc], [64 * p for p in proc], ctx=ctx, proc=proc, rank=rank)
sp = fem.space.dglegendre(g, order=2, hierarchical=True)
op = fem.operator.galerkin(sp, **MODEL)
op.setCommunicate(False)
timed(op, sp.size, "5 rank-local box, communicate off")
op.setCommunicate(True)
timed(op, sp.size, "6 rank-local box, fused exchange", comm_u=True)
timed(op, sp.size, "7 rank-local box, fused exchange, linear", linear=True, comm_u=True)
dist.barrier()
dist.destroy_process_group()
