"""PPO update (4 epochs x 4 minibatches of 32768 rows, actor + critic) for network widths the fused 64-wide kernels do not cover: the layer engine
(tcgen05 GEMMs where a layer fills a tile, FFMA tiles otherwise).  python scripts/bench_generic_ppo.py > gpurun_out/generic_ppo.json"""
import ctypes as C
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import crux_b200 as crux
from crux_b200.device import ptr

ctx = crux.Context(0)
n, mb, epochs, obs, act = 131072, 32768, 4, 17, 6
g = torch.Generator(device=ctx.device).manual_seed(0)
d = {"s": torch.randn((n, obs), device=ctx.device, generator=g), "a": torch.randn((n, act), device=ctx.device, generator=g),
     "logprob": torch.randn(n, device=ctx.device, generator=g) * 0.1 - 6.0, "advantage": torch.randn(n, device=ctx.device, generator=g),
     "return": torch.randn(n, device=ctx.device, generator=g)}
out = {}
for hid, mode in ((64, ""), (64, "nofused"), (128, ""), (128, "ffma"), (256, ""), (256, "ffma")):
    os.environ.pop("CRUX_NO_FUSED", None); os.environ.pop("CRUX_GEMM_TC5", None)
    if mode == "nofused":
        os.environ["CRUX_NO_FUSED"] = "1"
    if mode == "ffma":
        os.environ["CRUX_GEMM_TC5"] = "0"
    rng = np.random.default_rng(1)
    D = crux.Dense
    mu = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.tanh, rng=rng), D(hid, hid, crux.tanh, rng=rng), D(hid, act, rng=rng)), ctx=ctx)
    cr = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.tanh, rng=rng), D(hid, hid, crux.tanh, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
    pi = crux.ActorCritic(crux.GaussianPolicy(mu, np.full(act, -0.5, np.float32)), cr)
    mu.mlp.set_adam(np.float32(3e-4)); cr.mlp.set_adam(np.float32(3e-4))
    hp = crux._abi.PPOHp(eps_clip=0.2, lambda_p=1.0, lambda_e=0.0, target_kl=math.inf, a2c=0, actor_epochs=epochs, actor_batch=mb, critic_epochs=epochs,
                         critic_batch=mb, actor_max_batches=0, critic_max_batches=0)

    def upd(k):
        ctx.check(ctx.lib.crux_ppo_update_async(pi.A.h, cr.mlp.h, ptr(d["s"]), ptr(d["a"]), ptr(d["logprob"]), ptr(d["advantage"]), ptr(d["return"]), n,
                                                C.byref(hp), None, None, k))
    for k in range(2):
        upd(k)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for k in range(5):
        upd(10 + k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out[f"hidden{hid}{'_' + mode if mode else ''}"] = {"ms_per_update": ms, "env_steps_per_s_update_only": n / ms * 1e3}
    print(f"hidden {hid} {mode or 'default'}: {ms:.3f} ms per update", file=sys.stderr, flush=True)
print(json.dumps(out, indent=1))
