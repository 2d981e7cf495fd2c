"""Per-k-tile cost of gemm_tc5_kernel with parts switched off (CRUX_G5_DEBUG bits: 1 no refills, 2 no split passes, 4 no MMAs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crux_b200 as crux
from crux_b200.device import ptr
ctx = crux.Context(0)
rng = np.random.default_rng(0)
os.environ["CRUX_GEMM_TC5"] = "1"
B, K, N = 2048, 4096, 256
net = crux.ContinuousNetwork(crux.Chain(crux.Dense(K, N, crux.relu, rng=rng)), ctx=ctx)
x = torch.randn((B, K), device=ctx.device); y = torch.empty((B, N), device=ctx.device)
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    os.environ["CRUX_G5_DEBUG"] = str(dbg)
    for _ in range(3):
        ctx.check(ctx.lib.crux_mlp_forward(net.mlp.h, ptr(x), B, ptr(y)))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(50):
        ctx.check(ctx.lib.crux_mlp_forward(net.mlp.h, ptr(x), B, ptr(y)))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 50 * 1e3
    print(f"dbg={dbg}: {us:7.1f} us  = {us / (K // 32) * 1.9e3:6.0f} cycles per k-tile (at 1.9 GHz)")
