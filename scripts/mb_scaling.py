"""Minibatch-kernel time vs tiles per CTA (prologue vs per-tile cost): python scripts/mb_scaling.py"""
import os, sys, ctypes as C, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import crux_b200 as crux
import bench
from crux_b200.device import ptr

ctx = crux.Context(0)
S = bench.build_solver(crux, ctx)
env = crux.DeviceLinQuad(bench.N_ENVS, bench.OBS, bench.ACT, seed=1000, max_steps=1000, ctx=ctx)
crux.solve(S, env)
D = S.buffer
n = bench.N_ENVS * bench.HORIZON
cols = {k: D.column(k)[:n] for k in D.schema}
pi = S.agent.pi if hasattr(S, "agent") else S.pi
A = crux._abi
for tiles_per_cta in (0.25, 0.5, 1, 2, 3, 4):
    bm = int(296 * 64 * tiles_per_cta)
    hp = A.PPOHp(eps_clip=0.2, lambda_p=1.0, lambda_e=0.0, target_kl=float("inf"), a2c=0, actor_epochs=1, actor_batch=bm,
                 critic_epochs=1, critic_batch=bm, actor_max_batches=1, critic_max_batches=1)
    ia, ic = np.zeros((8, 8), np.float32), np.zeros((8, 8), np.float32)
    os.environ["CRUX_NO_SIDE_STREAM"] = "1"
    def run():
        ctx.check(ctx.lib.crux_ppo_update(pi.A.h, pi.C.mlp.h, ptr(cols["s"]), ptr(cols["a"]), ptr(cols["logprob"]), ptr(cols["advantage"]),
                                          ptr(cols["return"]), bm, C.byref(hp), None, None, 1, ptr(ia), ptr(ic)))
    for _ in range(3): run()
    fam_ms, fam_n = (C.c_float * 8)(), (C.c_int32 * 8)()
    ctx.check(ctx.lib.crux_ctx_timing_begin(ctx.h))
    for _ in range(20): run()
    ctx.check(ctx.lib.crux_ctx_timing_end(ctx.h, fam_ms, fam_n))
    print(f"tiles/CTA {tiles_per_cta:4}: rows {bm:6d}  minibatch kernel avg {1e3*fam_ms[0]/fam_n[0]:7.2f} us over {fam_n[0]} launches (actor+critic)", flush=True)
