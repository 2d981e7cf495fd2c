import os, sys, time, numpy as np
sys.path.insert(0, "/root/repo")
isa = sys.argv[1]
os.environ["CRUX_HOSTENV_ISA"] = isa
import importlib.util
spec = importlib.util.spec_from_file_location("envs", "/root/repo/crux.jl_b200/envs.py")
import crux_b200 as crux
env = crux.NativeHostLinQuad(4096, 17, 6, seed=1, n_threads=int(sys.argv[2]) if len(sys.argv) > 2 else 0)
a = np.random.default_rng(0).standard_normal((4096, 6)).astype(np.float32)
sp = np.zeros((4096, 17), np.float32); r = np.zeros(4096, np.float32); done = np.zeros(4096, np.uint8)
env.reset() if hasattr(env, "reset") else None
for _ in range(50): env.lib().crux_hostenv_step(env.h, a.ctypes.data, sp.ctypes.data, r.ctypes.data, done.ctypes.data)
t0 = time.perf_counter()
for _ in range(500): env.lib().crux_hostenv_step(env.h, a.ctypes.data, sp.ctypes.data, r.ctypes.data, done.ctypes.data)
dt = (time.perf_counter() - t0) / 500
print(isa, env.n_threads, "threads:", round(dt * 1e6, 1), "us per 4096-env step", float(sp.sum()), float(r.sum()))
