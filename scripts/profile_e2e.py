"""cProfile of the host-env PPO iteration (e2e leg of bench.py) on the GPU box."""
import cProfile
import pstats
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import crux_b200 as crux
import bench

ctx = crux.Context(0)
S = bench.build_solver(crux, ctx, seed=2)
env = (crux.HostLinQuad if '--numpy' in sys.argv else crux.NativeHostLinQuad)(bench.N_ENVS, bench.OBS, bench.ACT, seed=5)
crux.solve(S, env)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(3):
    crux.solve(S, env)
torch.cuda.synchronize()
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
