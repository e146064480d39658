"""torchrun probe: fused exchange of the marching kernel, per-rank step time; B200FEM_MARCH_TS=1 prints the tail timeline"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import dune_fem_b200 as fem
from dune_fem_b200.grid import Context
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
dist.init_process_group("nccl", device_id=dev)
ctx = Context(device=lr, stream=stream.cuda_stream)
ids = [Context.nccl_unique_id() if rank == 0 else None]; dist.broadcast_object_list(ids, src=0); ctx.init_nccl(ids[0], rank, world)
MODEL = dict(eps=1e-5, b=(1.0, 0.0, 0.0), beta=80.0, dirichlet_mask=0b000011, data=1)
proc = {2: [1, 1, 2], 4: [1, 2, 2], 8: [1, 2, 4]}[world]
g = fem.structuredGrid([-1.0] * 3, [-1.0 + 2.0 * p for p in proc], [64 * p for p in proc], ctx=ctx, proc=proc, rank=rank)
sp = fem.space.dglegendre(g, order=2, hierarchical=True)
op = fem.operator.galerkin(sp, **MODEL)
us = [torch.rand(sp.size, dtype=torch.float64, device=dev) for _ in range(6)]
ws = [torch.empty(sp.size, dtype=torch.float64, device=dev) for _ in range(6)]
for t in us: op.communicate_dev(t.data_ptr())
for linear in (False, True):
    for i in range(12): op.apply_dev(us[i % 6].data_ptr(), ws[i % 6].data_ptr(), linear)
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(400): op.apply_dev(us[i % 6].data_ptr(), ws[i % 6].data_ptr(), linear)
    e1.record(stream); torch.cuda.synchronize()
    print(f"[rank {rank}] fused exchange, linear={linear}: {e0.elapsed_time(e1) * 1e3 / 400:.2f} us/apply", flush=True)
dist.barrier()
del op
dist.destroy_process_group()
