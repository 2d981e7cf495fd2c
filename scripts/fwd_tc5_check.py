"""tcgen05 forward kernel (CRUX_FWD_TC5=1) against the oracle and the mma.sync forward kernel: python scripts/fwd_tc5_check.py"""
import os, sys
os.environ.setdefault("CRUX_FWD_TC5", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import crux_b200 as crux
from oracle import crux_oracle as o
from gpu_util import make_mlp, dev, host, p

ctx = crux.Context(0)
rng = np.random.default_rng(3)
for (I, O, B) in ((17, 1, 131072), (17, 6, 40000), (24, 1, 19000 + 77), (5, 3, 148 * 128)):
    ref = o.MLP([I, 64, 64, O], [1, 1, 0], rng)
    h = make_mlp(ctx, ref.dims, ref.acts, ref.flat())
    x = rng.standard_normal((B, I)).astype(np.float32)
    xd = dev(ctx, x)
    y = ctx.empty((B, O)); y.fill_(float("nan"))
    l0 = ctx.launch_count()
    ctx.check(ctx.lib.crux_mlp_forward(h, p(xd), B, p(y)))
    torch.cuda.synchronize()
    want = ref(x).detach().numpy()
    err = np.abs(host(y) - want)
    print(f"I={I} O={O} B={B}: max |err| {err.max():.3e}  (max |ref| {np.abs(want).max():.3f})  nan={int(np.isnan(host(y)).sum())}", flush=True)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): ctx.check(ctx.lib.crux_mlp_forward(h, p(xd), B, p(y)))
    a.record()
    for _ in range(20): ctx.check(ctx.lib.crux_mlp_forward(h, p(xd), B, p(y)))
    b.record(); torch.cuda.synchronize()
    print(f"   {1e3 * a.elapsed_time(b) / 20:.1f} us per forward", flush=True)
