"""Per-launch time of the layer engine's forward GEMM (one Dense layer through crux_mlp_forward) for a few shapes, FFMA vs tcgen05:
   python scripts/gemm_tc5_bench.py        (CRUX_GEMM_TC5 is switched per call inside)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crux_b200 as crux
from crux_b200.device import ptr

ctx = crux.Context(0)
rng = np.random.default_rng(0)
for (B, K, N) in [(2048, 256, 256), (2048, 2048, 256), (2048, 393, 256), (2048, 376, 256), (16384, 256, 256), (2048, 256, 34)]:
    net = crux.ContinuousNetwork(crux.Chain(crux.Dense(K, N, crux.relu, rng=rng)), ctx=ctx)
    x = torch.randn((B, K), device=ctx.device)
    y = torch.empty((B, N), device=ctx.device)
    res = {}
    for mode in ("0", "1"):
        os.environ["CRUX_GEMM_TC5"] = mode
        for _ in range(5):
            ctx.check(ctx.lib.crux_mlp_forward(net.mlp.h, ptr(x), B, ptr(y)))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(200):
            ctx.check(ctx.lib.crux_mlp_forward(net.mlp.h, ptr(x), B, ptr(y)))
        e1.record(); torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 200 * 1e3
    fl = 2.0 * B * K * N
    print(f"B={B:6d} K={K:5d} N={N:4d}: ffma {res['0']:7.1f} us ({fl / res['0'] / 1e6:6.1f} TFLOP/s)   tcgen05 {res['1']:7.1f} us ({fl / res['1'] / 1e6:6.1f} TFLOP/s)")
