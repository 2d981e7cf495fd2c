"""One PPO iteration (device env, BASELINE config[1]) bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off.
   python scripts/ncu_step.py            # one full iteration: rollout + values + GAE + whiten + update
   python scripts/ncu_step.py --gae      # the GAE scan alone at [2048, 16384] (738 MB)"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crux_b200 as crux
import bench
from crux_b200.device import ptr

ctx = crux.Context(0)
if "--gae" in sys.argv:
    T, N = 2048, 16384
    g = torch.Generator(device=ctx.device).manual_seed(2)
    r, vs, vsp = (torch.randn((T, N), device=ctx.device, generator=g) for _ in range(3))
    done = (torch.rand((T, N), device=ctx.device, generator=g) < 0.001).to(torch.uint8)
    ee = done.clone(); ee[999::1000] = 1; ee[-1] = 1
    adv, ret = torch.empty_like(r), torch.empty_like(r)
    run = lambda: ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, ptr(r), ptr(done), ptr(ee), ptr(vs), ptr(vsp), T, N, 0.99, 0.95, ptr(adv), ptr(ret)))
    run(); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    run(); run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:
    S = bench.build_solver(crux, ctx)
    env = crux.DeviceLinQuad(bench.N_ENVS, bench.OBS, bench.ACT, seed=1000, max_steps=1000, ctx=ctx)
    for _ in range(2):
        crux.solve(S, env)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    crux.solve(S, env)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
