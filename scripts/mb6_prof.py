"""Phase timeline (clock64 deltas of CTA 0) of the tcgen05 minibatch kernel: CRUX_MB6_PROF=1 python scripts/mb6_prof.py"""
import os, sys, ctypes as C
os.environ["CRUX_MB6_PROF"] = "1"
os.environ["CRUX_NO_SIDE_STREAM"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import crux_b200 as crux
import bench
from crux_b200.device import ptr
ctx = crux.Context(0)
S = bench.build_solver(crux, ctx)
env = crux.DeviceLinQuad(bench.N_ENVS, bench.OBS, bench.ACT, seed=1000, max_steps=1000, ctx=ctx)
S.N = bench.N_ENVS * bench.HORIZON
crux.solve(S, env)
