"""CTA start / end timeline (%globaltimer) of the minibatch kernels of a few PPO iterations: CRUX_MB6_TRACE=1 python scripts/mb6_trace.py"""
import os, sys
os.environ["CRUX_MB6_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import crux_b200 as crux
import bench
ctx = crux.Context(0)
S = bench.build_solver(crux, ctx)
env = crux.DeviceLinQuad(bench.N_ENVS, bench.OBS, bench.ACT, seed=1000, max_steps=1000, ctx=ctx)
S.N = bench.N_ENVS * bench.HORIZON
for _ in range(7):
    crux.solve(S, env)
