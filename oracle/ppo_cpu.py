"""CPU restatement of one full PPO iteration of the reference's solve loop (on_policy.jl:91-106):
steps!(ΔN, explore, reset) -> whiten(advantage) -> batch_train!(actor, ppo_loss) -> batch_train!(critic, mse).
TEST INFRASTRUCTURE ONLY (parity tests, bench.py cpu_baseline / --impl reference).

Two rollout modes:
  * ``vectorised=True``  : all N env streams advance together with batched forwards (the strongest CPU baseline;
                           same numbers as N independent reference samplers up to BLAS summation order);
  * ``vectorised=False`` : reference-faithful -- one env, one transition at a time, batch-1 forwards and two batch-1
                           value calls per transition inside fill_gae! (sampler.jl:71-137,262-273).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import crux_oracle as o

F32 = np.float32


class OraclePPO:
    def __init__(self, n_envs, T, obs_dim=17, act_dim=6, hidden=64, seed=1, epochs=4, batch=None, eta=F32(3e-4), eps_clip=0.2,
                 lp=1.0, le=0.0, target_kl=math.inf, lam=0.95, max_steps=1000, env_seed=0, log_sigma=-0.5):
        rng = np.random.default_rng(seed)
        self.mu = o.MLP([obs_dim, hidden, hidden, act_dim], [o.ACT_TANH, o.ACT_TANH, o.ACT_IDENTITY], rng)
        self.critic = o.MLP([obs_dim, hidden, hidden, 1], [o.ACT_TANH, o.ACT_TANH, o.ACT_IDENTITY], rng)
        self.pi = o.GaussianPolicy(self.mu, np.full(act_dim, log_sigma, F32))
        self.spec = o.LinQuadSpec(obs_dim, act_dim, seed=0)
        self.n, self.T, self.obs_dim, self.act_dim = n_envs, T, obs_dim, act_dim
        self.epochs, self.batch = epochs, (n_envs * T // 4 if batch is None else batch)
        self.P = {"eps": F32(eps_clip), "lp": F32(lp), "le": F32(le)}
        self.target_kl, self.lam, self.gamma, self.max_steps = target_kl, F32(lam), self.spec.gamma, max_steps
        self.opt_a, self.opt_c = o.Adam(eta), o.Adam(eta)
        self.env_rng = np.random.default_rng(env_seed)
        self.rng = np.random.default_rng(seed + 1)
        self.state = self._s0(n_envs)
        self.ep_len = np.zeros(n_envs, np.int64)
        self.last = {}

    def _s0(self, n):
        return ((self.env_rng.random((n, self.obs_dim), dtype=F32) * F32(2) - F32(1)) * F32(0.1)).astype(F32)

    # ---- steps!(Nsteps = T*n, explore=true, reset=true)
    def rollout(self, eps=None):
        n, T = self.n, self.T
        D = {k: np.zeros((T, n) + sh, F32) for k, sh in (("s", (self.obs_dim,)), ("a", (self.act_dim,)), ("sp", (self.obs_dim,)),
                                                          ("r", ()), ("logprob", ()))}
        D["done"], D["episode_end"] = np.zeros((T, n), bool), np.zeros((T, n), bool)
        with torch.no_grad():
            for t in range(T):
                e = self.rng.standard_normal((n, self.act_dim)).astype(F32) if eps is None else eps[t]
                a, lp = self.pi.exploration(self.state, e)
                a, lp = a.numpy(), lp.numpy()[:, 0]
                xi = self.env_rng.standard_normal((n, self.obs_dim), dtype=F32)
                sp, r, done = self.spec.step(self.state, a, xi)
                self.ep_len += 1
                end = done | (self.ep_len >= self.max_steps)
                if t == T - 1:
                    end[:] = True
                D["s"][t], D["a"][t], D["sp"][t], D["r"][t], D["logprob"][t] = self.state, a, sp, r, lp
                D["done"][t], D["episode_end"][t] = done, end
                nxt = sp
                if end.any():
                    idx = np.flatnonzero(end)
                    nxt = sp.copy()
                    nxt[idx] = self._s0(len(idx))
                    self.ep_len[idx] = 0
                self.state = nxt
            vs = self.critic(D["s"].reshape(T * n, -1)).numpy().reshape(T, n)
            vsp = self.critic(D["sp"].reshape(T * n, -1)).numpy().reshape(T, n)
        D["advantage"], D["return"] = o.gae_returns_TN(D["r"], D["done"], D["episode_end"], vs, vsp, self.gamma, self.lam)
        return {k: v.reshape((T * n,) + v.shape[2:]) for k, v in D.items()}

    def update(self, D, orders=None):
        n = len(D["r"])
        D = dict(D)
        D["advantage"] = o.whiten(D["advantage"])  # ppo.jl:61
        cols = {k: D[k] for k in ("s", "a", "logprob", "advantage", "return")}
        order = np.arange(n)
        recs_a, recs_c, stop = [], [], False
        for e in range(self.epochs):
            order = order[self.rng.permutation(n)] if orders is None else np.asarray(orders[0][e])
            for st in range(0, n, self.batch):
                idx = order[st:st + self.batch]
                mb = {k: v[idx] for k, v in cols.items()}
                info = {}
                o.train_step(self.pi.params(), lambda inf, mb=mb: o.ppo_loss(self.pi, self.P, mb, inf), self.opt_a, info, "actor_")
                recs_a.append(info)
                if info["kl"] > self.target_kl:
                    stop = True
                    break
            if stop:
                break
        for e in range(self.epochs):
            order = order[self.rng.permutation(n)] if orders is None else np.asarray(orders[1][e])
            for st in range(0, n, self.batch):
                idx = order[st:st + self.batch]
                mb = {k: v[idx] for k, v in cols.items()}
                info = {}
                o.train_step(self.critic.params(), lambda inf, mb=mb: o.value_mse_loss(self.critic, mb), self.opt_c, info, "critic_")
                recs_c.append(info)
        self.last = {"actor": recs_a, "critic": recs_c}
        return self.last

    def iteration(self, eps=None, orders=None):
        D = self.rollout(eps)
        self.update(D, orders)
        return D


def reference_faithful_steps_per_sec(n_steps=300, seed=1, max_steps=1000):
    """The reference's actual execution pattern: one env, one transition at a time, batch-1 forwards,
    and fill_gae! evaluating the critic twice per transition with batch-1 calls.  Returns env-steps/s of the
    sampling part only (rollout + per-episode GAE) on one thread."""
    import time
    torch.set_num_threads(1)
    p = OraclePPO(1, n_steps, seed=seed, max_steps=max_steps)
    t0 = time.perf_counter()
    with torch.no_grad():
        s = p.state
        rows = []
        for t in range(n_steps):
            e = p.rng.standard_normal((1, p.act_dim)).astype(F32)
            a, lp = p.pi.exploration(s, e)
            xi = p.env_rng.standard_normal((1, p.obs_dim), dtype=F32)
            sp, r, done = p.spec.step(s, a.numpy(), xi)
            rows.append((s, sp, r, done))
            s = sp
        A = F32(0)
        for (s_, sp_, r, done) in reversed(rows):  # fill_gae!: 2 batch-1 value calls per transition
            v, vp = float(p.critic(s_)), float(p.critic(sp_))
            A = F32(p.gamma * p.lam) * A + r[0] + (1 - float(done[0])) * float(p.gamma) * vp - v
    return n_steps / (time.perf_counter() - t0)
