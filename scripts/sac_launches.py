"""Kernel launch list of one SAC update (BASELINE configs[3]: 376-obs / 17-act, 256-256, batch 2048) for
   ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_sac.csv python scripts/sac_launches.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crux_b200 as crux
from crux_b200.device import ptr

F32 = np.float32
ctx = crux.Context(0)
obs, act, hid, Bs = 376, 17, 256, 2048
rng = np.random.default_rng(4)
g = torch.Generator(device=ctx.device).manual_seed(3)
D = crux.Dense
Apol = crux.SquashedGaussianPolicy(crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 2 * act, rng=rng)), ctx=ctx))
Q = lambda: crux.ContinuousNetwork(crux.Chain(D(obs + act, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
pi = crux.ActorCritic(Apol, crux.DoubleNetwork(Q(), Q()))
S4 = crux.SAC(pi, crux.ContinuousSpace(obs), N=10, dN=1, c_opt=dict(batch_size=Bs, epochs=1), buffer_size=Bs, buffer_init=Bs)
s = torch.randn((Bs, obs), device=ctx.device, generator=g)
a = torch.tanh(torch.randn((Bs, act), device=ctx.device, generator=g))
sp = torch.randn((Bs, obs), device=ctx.device, generator=g)
r = torch.randn(Bs, device=ctx.device, generator=g)
dn = (torch.rand(Bs, device=ctx.device, generator=g) < 0.01).to(torch.uint8)
if "--time" in sys.argv:      # live timing instead of a profiler bracket: python scripts/sac_launches.py --time
    for k in range(1, 6):
        ctx.check(ctx.lib.crux_sac_train(S4._sac, ptr(s), ptr(a), ptr(sp), ptr(r), ptr(dn), Bs, F32(0.99), None, None, None, 4, 3 * k, None, None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(6, 56):
        ctx.check(ctx.lib.crux_sac_train(S4._sac, ptr(s), ptr(a), ptr(sp), ptr(r), ptr(dn), Bs, F32(0.99), None, None, None, 4, 3 * k, None, None))
    e1.record(); torch.cuda.synchronize()
    print("sac ms_per_update", e0.elapsed_time(e1) / 50, "CRUX_GEMM_TC5=" + os.environ.get("CRUX_GEMM_TC5", ""))
    sys.exit(0)
for k in range(1, 4):
    if k == 3:
        torch.cuda.synchronize(); torch.cuda.profiler.start()
    ctx.check(ctx.lib.crux_sac_train(S4._sac, ptr(s), ptr(a), ptr(sp), ptr(r), ptr(dn), Bs, F32(0.99), None, None, None, 4, 3 * k, None, None))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
