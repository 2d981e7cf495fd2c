"""A/B of the host-env PPO iteration (the e2e leg of bench.py) on one box: wall-clock per iteration + the host-loop split that
CRUX_ROLLOUT_PROFILE=1 prints.  Usage: [CRUX_ROLLOUT_DIRECT=1] python scripts/e2e_ab.py [iters]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crux_b200 as crux
import bench

iters = int(sys.argv[1]) if len(sys.argv) > 1 else 8
ctx = crux.Context(0)
S = bench.build_solver(crux, ctx, seed=2)
env = crux.NativeHostLinQuad(bench.N_ENVS, bench.OBS, bench.ACT, seed=5, n_threads=int(os.environ.get("ENV_THREADS", 0)) or bench.cpu_threads())
S.N = bench.N_ENVS * bench.HORIZON
for _ in range(3):
    crux.solve(S, env)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(iters):
    crux.solve(S, env)
    S.training_info()
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / iters
print("threads %d  ms/iteration %.3f  env-steps/s %.3e  direct=%s" % (env.n_threads, dt * 1e3, S.N / dt, bool(os.environ.get("CRUX_ROLLOUT_DIRECT"))))
