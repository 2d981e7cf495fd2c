"""Latency of the gradient all-reduce (5.9 K floats) on N GPUs: NCCL vs the one-shot peer exchange.
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29560 scripts/allreduce_latency.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
import crux_b200 as crux
from crux_b200.device import ptr

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = 5800 + 128
for mode in ("nccl", "peer"):
    ctx = crux.Context(local)
    ctx.init_distributed(rank, world, peer_floats=n if mode == "peer" else 0)
    buf = torch.ones(n, device=ctx.device)
    for _ in range(20):
        ctx.check(ctx.lib.crux_nccl_allreduce_f32(ctx.h, ptr(buf), n)); buf.fill_(1.0)
    torch.cuda.synchronize(); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 500
    a.record()
    for _ in range(reps):
        ctx.check(ctx.lib.crux_nccl_allreduce_f32(ctx.h, ptr(buf), n))
    b.record(); torch.cuda.synchronize()
    if rank == 0:
        print(f"{mode}: {1e3 * a.elapsed_time(b) / reps:.2f} us per all-reduce of {n} floats on {world} GPUs (back to back)", flush=True)
    dist.barrier()
dist.destroy_process_group()
