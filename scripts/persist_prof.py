"""One persistent-epoch launch per network (critic + actor, 256 minibatches of 128 rows) for ncu:
   ncu --set full --import-source on -k regex:epoch_kernel -c 2 -o gpurun_out/prof_persist python scripts/persist_prof.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import crux_b200 as crux
import test_gpu_ppo as T
ctx = crux.default_context()
n = 16384
rng, pi, cr, handles, D = T._setup(ctx, crux, n, seed=1)
hp = T._hp(crux, actor_batch=128, critic_batch=128, actor_epochs=2, critic_epochs=2)
for _ in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); T._run(ctx, crux, handles, D, hp, None, None, n); b.record(); torch.cuda.synchronize()
print("256 minibatches per network: %.1f us per minibatch (both networks concurrent)" % (a.elapsed_time(b) * 1e3 / 256))
