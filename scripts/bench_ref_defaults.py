"""the PPO step of BASELINE config[1] (4096 env streams x T=32, 17-64-64-6 / 17-64-64-1) once more with the
REFERENCE-DEFAULT training hyper-parameters instead of the large-batch ones the headline uses: TrainingParams defaults
epochs = 80, batch_size = 128 (training.jl:3-6), PPO defaults λe = 0.1, target_kl = 0.012 (ppo.jl:42-45).  That is up to
80 x 1024 minibatch updates of 128 rows per network and iteration: each network's batch_train! runs as ONE launch of the persistent
cluster kernel (csrc/mb_persist.cuh); bench.py reports the result as `ref_defaults`.
Usage (GPU box): python scripts/bench_ref_defaults.py [--iters 2] > gpurun_out/ref_defaults.json"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--no-early-stop", action="store_true", help="target_kl = Inf: all 80 actor epochs run")
    args = ap.parse_args()
    print(json.dumps(run(args.iters, args.no_early_stop)))


def run(iters=2, no_early_stop=False, ctx=None):
    import torch
    import crux_b200 as crux

    class args:
        pass
    args.iters, args.no_early_stop = iters, no_early_stop
    ctx = ctx or crux.Context(0)
    n_envs, T, obs, act, hid = 4096, 32, 17, 6, 64
    rng = np.random.default_rng(1)
    D = crux.Dense
    mu = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.tanh, rng=rng), D(hid, hid, crux.tanh, rng=rng), D(hid, act, rng=rng)), ctx=ctx)
    cr = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.tanh, rng=rng), D(hid, hid, crux.tanh, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
    pi = crux.ActorCritic(crux.GaussianPolicy(mu, np.full(act, -0.5, np.float32)), cr)
    kw = dict(target_kl=float("inf")) if args.no_early_stop else {}
    S = crux.PPO(pi, crux.ContinuousSpace(obs), N=n_envs * T, dN=n_envs * T, max_steps=1000, lam_gae=0.95, log=None, seed=1, **kw)
    assert S.a_opt.epochs == 80 and S.a_opt.batch_size == 128 and S.c_opt.epochs == 80
    env = crux.DeviceLinQuad(n_envs, obs, act, seed=1000, max_steps=1000, ctx=ctx)
    crux.solve(S, env)                       # warm-up iteration
    torch.cuda.synchronize()
    l0 = ctx.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    infos = []
    for _ in range(args.iters):
        crux.solve(S, env)
        infos.append(S.training_info())
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1) / args.iters
    out = {"workload": "PPO, 4096 env streams x T=32, reference-default TrainingParams (epochs 80, batch 128), λe 0.1, target_kl %s"
                       % ("Inf" if args.no_early_stop else "0.012"),
           "env_steps_per_s": n_envs * T / (ms * 1e-3), "ms_per_iteration": ms, "wall_ms_per_iteration": wall * 1e3 / args.iters,
           "launches_per_iteration": (ctx.launch_count() - l0) / args.iters,
           "actor_batches_trained": [i["actor_batches_trained"] for i in infos],
           "critic_batches_trained": [i.get("critic_batches_trained") for i in infos],
           "us_per_minibatch_update": ms * 1e3 / max(1, max(infos[-1]["actor_batches_trained"], infos[-1].get("critic_batches_trained", 0))),
           "last_info": {k: (round(v, 6) if isinstance(v, float) else v) for k, v in infos[-1].items()}}
    return out


if __name__ == "__main__":
    main()
