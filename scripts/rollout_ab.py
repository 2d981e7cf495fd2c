"""Device time of crux_linquad_rollout (4096 streams x T=32) per kernel variant: CRUX_ROLLOUT_RPT=1|2|4 python scripts/rollout_ab.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crux_b200 as crux
import bench
ctx = crux.Context(0)
S = bench.build_solver(crux, ctx)
env = crux.DeviceLinQuad(bench.N_ENVS, bench.OBS, bench.ACT, seed=1000, max_steps=1000, ctx=ctx)
S.N = bench.N_ENVS * bench.HORIZON
crux.solve(S, env)
D, s = S.buffer, S.sampler
dN = bench.N_ENVS * bench.HORIZON
data = {k: D.column(k)[:dN] for k in D.schema}
ts = []
for _ in range(30):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); s._rollout_device(data, bench.HORIZON, True, 0, True, None); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
print(f"rpt={os.environ.get('CRUX_ROLLOUT_RPT', 'default')}: median {np.median(ts[5:]):.1f} us  min {min(ts):.1f} us")
