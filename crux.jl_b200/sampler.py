"""Host-side mirror of ``src/sampler.jl``: ``Sampler``, ``steps_`` (``steps!``), ``fill_gae_``/``fill_returns_``.

One ``Sampler`` drives N independent env streams at once (the reference steps one env per ``Sampler`` and loops a vector
of samplers sequentially, sampler.jl:157-173).  Rollout rows are laid out ``[T][N]`` (row ``t*N + e``, the reference's own
interleaving ``j = (step-1)*Nenv + env``); every stream follows the single-env rules of ``step!`` (sampler.jl:71-137):
episode_length, ``done || episode_length >= max_steps`` -> ``terminate_episode!`` (episode_end flag, reset), and the
forced termination of ``steps!(reset=true)`` (:148).  Advantages and returns are filled for all episode ranges at once
by the segmented scan ``crux_fill_gae_returns`` after V(s) and V(sp) have been evaluated in two batched critic passes
(the reference calls ``value`` twice per transition, sampler.jl:262-273).
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _abi
from .buffer import ExperienceBuffer, mdp_data, _TORCH
from .device import ptr
from .policies import (ActorCritic, ContinuousNetwork, DiscreteNetwork, FirstExplorePolicy, GaussianPolicy, MixedPolicy, Policy, PolicyParams,
                       action, actor, critic, exploration, fusable)
from .spaces import ContinuousSpace, DiscreteSpace

F32 = np.float32


class Sampler:
    """sampler.jl:1-22.  ``mdp`` is a vectorised environment (crux.jl_b200/envs.py protocol)."""

    def __init__(self, mdp, agent, S=None, max_steps=100, required_columns=(), lam=float("nan"), gamma=None, seed=0, Vc=None):
        self.mdp = mdp
        self.Vc = Vc                # value network of the cost (sampler.jl:16-18), LagrangePPO only
        self.agent = agent if isinstance(agent, PolicyParams) else PolicyParams(agent)
        self.ctx = self.agent.pi.ctx
        self.n = int(mdp.n_envs)
        self.S = S if S is not None else ContinuousSpace(mdp.obs_dim)
        self.max_steps = int(max_steps)
        self.required_columns = list(required_columns)
        self.gamma = F32(mdp.gamma if gamma is None else gamma)
        self.lam = F32(lam)
        self.seed = int(seed)
        self.noise_ctr = 0          # Philox stream position of the exploration noise
        self.on_device = bool(getattr(mdp, "on_device", False))
        self.sdim = int(np.prod(self.S.dims))
        self.episode_length = np.zeros(self.n, dtype=np.int64)
        self.cur = self.ctx.empty((self.n, self.sdim))   # current observation of every stream (device)
        self._pinned = None
        self._scratch = {}
        self.reset_()

    # ---- reset_sampler! (sampler.jl:31-43) for every stream
    def reset_(self):
        if self.on_device:
            assert self.mdp.max_steps == self.max_steps, "a device env bakes max_steps into its step kernel: construct it with the solver's max_steps"
            self.mdp.reset_into(self.cur)
        else:
            o = self._tovec(self.mdp.reset(None))
            self._pin()
            self._pinned["obs"][...] = o
            self.ctx.h2d(self.cur, self._pinned["obs"])
        self.episode_length[:] = 0

    def _tovec(self, o):
        """tovec(o, S) spaces.jl:24-25 for a ContinuousSpace: (o - μ)/σ."""
        o = np.asarray(o, dtype=F32).reshape(-1, self.sdim)
        mu, sg = getattr(self.S, "mu", 0), getattr(self.S, "sigma", 1)
        if np.any(np.asarray(mu) != 0) or np.any(np.asarray(sg) != 1):
            o = ((o - F32(mu)) / F32(sg)).astype(F32)
        return o

    def _pin(self):
        """Pinned host staging buffers (numpy views over cudaMallocHost memory): the env reads actions from and writes
        transitions into them; copies are raw cudaMemcpyAsync calls on the context stream."""
        if self._pinned is not None:
            return
        n, sd, ctx = self.n, self.sdim, self.ctx
        A = self.agent.space
        adim = A.N if isinstance(A, DiscreteSpace) else int(np.prod(A.dims))
        self._pinned = {"obs": ctx.pinned_array((n, sd)), "sp": ctx.pinned_array((n, sd)), "r": ctx.pinned_array((n,)),
                        "done": ctx.pinned_array((n,), np.uint8), "ee": ctx.pinned_array((n,), np.uint8),
                        "t": ctx.pinned_array((n,), np.int64), "a": ctx.pinned_array((n, adim)), "ai": ctx.pinned_array((n,), np.int32)}

    def _tmp(self, name, shape, dtype=torch.float32):
        t = self._scratch.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.ctx.device)
            self._scratch[name] = t
        return t

    # ---- exploration / action for all streams at once (sampler.jl:73)
    def _act(self, obs, a_out, logp_out, explore, i, noise=None):
        """Writes the stored action rows (continuous vector / one-hot) and logprob; returns what the env consumes."""
        ag, ctx = self.agent, self.ctx
        pi, pe = ag.pi, ag.pi_explore
        discrete = isinstance(ag.space, DiscreteSpace)
        ctr = self.noise_ctr
        self.noise_ctr += 1
        if not explore:
            a = action(pi, obs)
            if discrete:
                a_out.zero_()
                a_out.scatter_(1, a.long().reshape(-1, 1), 1.0)
                if logp_out is not None:
                    logp_out.fill_(float("nan"))
                return a
            a_out.copy_(a)
            if logp_out is not None:
                logp_out.fill_(float("nan"))
            return a_out
        if isinstance(pe, FirstExplorePolicy):   # policies.jl:526-534: dispatch on the policy that acts at this step
            pol, plain = pe.resolve(i)
            if plain:
                a = action(pi if pol is None else pol, obs)
                if logp_out is not None:
                    logp_out.fill_(float("nan"))
                if discrete:
                    a_out.zero_()
                    a_out.scatter_(1, a.long().reshape(-1, 1), 1.0)
                    return a
                a_out.copy_(a.reshape(a_out.shape))
                return a_out
            pe = pol
        if isinstance(pe, Policy) and hasattr(pe, "exploration") and not isinstance(pe, (GaussianPolicy, ActorCritic, DiscreteNetwork, ContinuousNetwork)):
            if isinstance(pe, MixedPolicy):
                idx, oh, lp = pe.exploration(obs, pi, i, u=noise, seed=self.seed, ctr=ctr)
                a_out.copy_(oh)
                if logp_out is not None:
                    logp_out.copy_(lp.reshape(logp_out.shape))
                return idx
            a, lp = pe.exploration(obs, pi, i, seed=self.seed, ctr=ctr, **({"eps": noise} if noise is not None else {}))
            a_out.copy_(a.reshape(a_out.shape))
            if logp_out is not None:
                logp_out.fill_(float("nan") if isinstance(lp, float) else 0.0)
            return a_out
        act_pi = actor(pe)
        if isinstance(act_pi, GaussianPolicy):
            # hot path (i): one fused launch sequence writes a and logprob straight into the rollout rows
            eps = None if noise is None else ctx.to_device(noise, torch.float32)
            ctx.check(ctx.lib.crux_rollout_step(act_pi.h, None, ptr(obs), obs.shape[0], ptr(eps), self.seed, ctr, ptr(a_out),
                                                ptr(logp_out), None))
            return a_out
        if isinstance(act_pi, DiscreteNetwork):
            idx, lp = exploration(act_pi, obs, eps=noise, seed=self.seed, ctr=ctr)
            a_out.zero_()
            a_out.scatter_(1, idx.long().reshape(-1, 1), 1.0)
            if logp_out is not None:
                logp_out.copy_(lp.reshape(logp_out.shape))
            return idx
        raise TypeError(f"Sampler: unsupported exploration policy {type(pe).__name__}")

    # ---- steps! (sampler.jl:139-155) over N streams
    def steps_(self, buffer=None, Nsteps=1, explore=False, i=0, reset=False, cb=None, store=None, noise=None):
        """Collect ``Nsteps`` transitions (= ``Nsteps / n`` vector steps) and push them to ``buffer``.
        ``noise``: optional per-vector-step list of injected exploration noise (parity runs).  Returns the rollout
        columns (device views, ``[Nsteps, ...]``)."""
        n = self.n
        assert Nsteps % n == 0, f"Nsteps={Nsteps} must be a multiple of the {n} env streams"
        T = Nsteps // n
        cols = self.required_columns
        in_place = False
        if buffer is not None:
            start0 = buffer.next_ind - 1
            in_place = start0 + Nsteps <= buffer.capacity and all(k in buffer.schema for k in cols)
        if in_place:  # zero-copy: rollout rows are the buffer's own rows (push! would copy them there anyway)
            data = {k: buffer.column(k)[start0:start0 + Nsteps] for k in buffer.schema}
        else:
            data = self._alloc(Nsteps)
        logp = data.get("logprob")
        discrete = isinstance(self.agent.space, DiscreteSpace)
        if self.on_device:
            assert "cost" not in data, "the device env reports no cost: LagrangePPO rollouts use a host env with last_info['cost']"
            self._rollout_device(data, T, explore, i, reset, noise)
        elif self._can_rollout_native(explore, noise, data, discrete):
            self._rollout_host_native(data, T, reset)
        else:
            self._rollout_host(data, T, explore, i, reset, noise, discrete)
        # terminate_episode! bookkeeping for all closed ranges at once (sampler.jl:56-57)
        if "advantage" in data or "return" in data:
            self.fill_gae_returns_(data, T)
        if "cost_advantage" in data or "cost_return" in data:   # sampler.jl:64-66: the same scans with source = :cost and Vc
            self.fill_gae_returns_(data, T, source="cost", adv_key="cost_advantage", ret_key="cost_return", V=self.Vc)
        if cb is not None:
            cb(data)
        if store is not None:
            store.append({k: v.clone() for k, v in data.items()})
        if buffer is not None:
            if in_place:
                buffer.commit_rows_(Nsteps)
            else:
                buffer.push_(data)
        return data

    def _alloc(self, Nsteps):
        schema = mdp_data(self.S, self.agent.space, Nsteps, self.required_columns)
        out = {}
        for k, (dt, dims, init) in schema.items():
            t = self._tmp("col_" + k, (Nsteps, *dims), _TORCH[dt])
            t.fill_(init)
            out[k] = t
        return out

    def _rollout_device(self, data, T, explore, i, reset, noise):
        n, env = self.n, self.mdp
        pe = self.agent.pi_explore
        ap = actor(pe) if isinstance(pe, (GaussianPolicy, ActorCritic)) else None
        if (explore and noise is None and isinstance(ap, GaussianPolicy) and not ap.squashed and ap.log_sigma is not None
                and fusable(ap.mu.mlp) and ap.mu.mlp.dims[0] == env.obs_dim and ap.adim == env.act_dim and hasattr(env, "rollout_into")
                and not getattr(self, "force_step_kernels", False)):
            # the whole steps! loop in ONE persistent launch (crux_linquad_rollout): bit-identical to the per-step kernels below
            env.rollout_into(ap, T, reset, self.cur, data, self.seed, self.noise_ctr)
            self.noise_ctr += T
            return
        s, a, sp, r = data["s"], data["a"], data["sp"], data["r"]
        done, ee, logp = data["done"], data["episode_end"], data.get("logprob")
        s[:n].copy_(self.cur)
        for t in range(T):
            rows = slice(t * n, (t + 1) * n)
            self._act(s[rows], a[rows], None if logp is None else logp[rows], explore, i + t * n, None if noise is None else noise[t])
            nxt = s[(t + 1) * n:(t + 2) * n] if t + 1 < T else self.cur
            env.step_into(s[rows], a[rows], sp[rows], r[rows], done[rows], ee[rows], nxt, force_end=(reset and t == T - 1))

    def _rollout_host(self, data, T, explore, i, reset, noise, discrete):
        n, env, ctx = self.n, self.mdp, self.ctx
        self._pin()
        P = self._pinned
        s, a, sp, r = data["s"], data["a"], data["sp"], data["r"]
        done, ee, logp, tcol, icol = data["done"], data["episode_end"], data.get("logprob"), data.get("t"), data.get("i")
        native = hasattr(env, "step_into") and not discrete
        for t in range(T):
            rows = slice(t * n, (t + 1) * n)
            ctx.h2d(s[rows], P["obs"])                                        # H2D: svec of every stream
            env_a = self._act(s[rows], a[rows], None if logp is None else logp[rows], explore, i + t * n,
                              None if noise is None else noise[t])
            if discrete:
                ctx.d2h(P["ai"], env_a.reshape(-1).to(torch.int32))
            else:
                ctx.d2h(P["a"], env_a)                                        # D2H: the env needs the action
            ctx.sync()
            if native:                                                        # @gen(:sp,:r), isterminal
                env.step_into(P["a"], P["sp"], P["r"], P["done"])
                spv, dn = P["sp"], P["done"].astype(bool)
                if self._needs_tovec():
                    P["sp"][...] = self._tovec(P["sp"])
            else:
                spn, rn, dn = env.step(P["ai"] if discrete else P["a"])
                spv = self._tovec(spn)
                P["sp"][...] = spv
                P["r"][...] = rn
                dn = np.asarray(dn, dtype=bool)
                P["done"][...] = dn
            if "cost" in data:                                                # sampler.jl:76-78,114: info["cost"] of the env step
                data["cost"][rows].copy_(torch.from_numpy(np.asarray(env.last_info["cost"], dtype=np.float32).reshape(n, 1)))
            if tcol is not None:
                P["t"][...] = self.episode_length + 1
            self.episode_length += 1                                          # sampler.jl:130
            end = dn | (self.episode_length >= self.max_steps)
            if reset and t == T - 1:
                end = np.ones(n, dtype=bool)                                  # steps!(reset=true) :148
            P["ee"][...] = end
            ctx.h2d(sp[rows], P["sp"])                                        # H2D: the transition
            ctx.h2d(r[rows], P["r"])
            ctx.h2d(done[rows], P["done"])
            ctx.h2d(ee[rows], P["ee"])
            if tcol is not None:
                ctx.h2d(tcol[rows], P["t"])
            if icol is not None:
                icol[rows].fill_(i + t * n + 1)
            # next observation: sp, or a fresh initial state where the episode ended (terminate_episode! -> reset_sampler!).
            # The staging buffers are rewritten only after the next ctx.sync() (after the action D2H), when the copies
            # queued above have completed -- except obs, whose previous H2D finished before this step's sync.
            P["obs"][...] = P["sp"]
            if end.any():
                idx = np.flatnonzero(end)
                P["obs"][idx] = self._tovec(env.reset(idx))
                self.episode_length[idx] = 0
        ctx.h2d(self.cur, P["obs"])
        ctx.sync()

    # ---- the same loop in one C call (crux_rollout_host) when the env exposes C callbacks and nothing is injected
    def _can_rollout_native(self, explore, noise, data, discrete):
        pe = self.agent.pi_explore
        return (explore and noise is None and not discrete and hasattr(self.mdp, "c_callbacks") and not self._needs_tovec()
                and not getattr(self, "force_python_loop", False)
                and isinstance(actor(pe), GaussianPolicy) and (pe is self.agent.pi or isinstance(pe, (GaussianPolicy, ActorCritic)))
                and not actor(pe).squashed and actor(pe).log_sigma is not None and "t" not in data and "i" not in data
                and "cost" not in data)

    def _rollout_host_native(self, data, T, reset):
        import ctypes as C
        ctx = self.ctx
        self._pin()
        if getattr(self, "_ep_len32", None) is None:
            self._ep_len32 = np.zeros(self.n, dtype=np.int32)
        self._ep_len32[:] = self.episode_length
        step_fn, reset_fn, user = self.mdp.c_callbacks()
        lp = data.get("logprob")
        cols = _abi.RolloutCols(data["s"].data_ptr(), data["a"].data_ptr(), data["sp"].data_ptr(), data["r"].data_ptr(), data["done"].data_ptr(),
                                data["episode_end"].data_ptr(), lp.data_ptr() if lp is not None else None)
        ctr0 = self.noise_ctr
        self.noise_ctr += T
        ctx.check(ctx.lib.crux_rollout_host(actor(self.agent.pi_explore).h, self.n, T, self.max_steps, 1 if reset else 0, step_fn, reset_fn, user,
                                            C.c_void_p(self._pinned["obs"].ctypes.data), C.c_void_p(self._ep_len32.ctypes.data), C.byref(cols),
                                            self.seed, ctr0))
        self.episode_length[:] = self._ep_len32
        ctx.h2d(self.cur, self._pinned["obs"])
        ctx.sync()

    def _needs_tovec(self):
        mu, sg = getattr(self.S, "mu", 0), getattr(self.S, "sigma", 1)
        return bool(np.any(np.asarray(mu) != 0) or np.any(np.asarray(sg) != 1))

    # ---- fill_gae! / fill_returns! (sampler.jl:255-281) for every episode range of every stream
    def fill_gae_returns_(self, data, T, source="r", adv_key="advantage", ret_key="return", V=None):
        ctx, n = self.ctx, self.n
        adv, ret = data.get(adv_key), data.get(ret_key)
        vs = vsp = None
        if adv is not None:
            V = critic(self.agent.pi) if V is None else V
            assert isinstance(V, ContinuousNetwork) and not math.isnan(float(self.lam)), "GAE needs a critic and λ"
            vs, vsp = self._tmp("v_s", (T * n, 1)), self._tmp("v_sp", (T * n, 1))
            V.mlp.forward(data["s"], out=vs)      # value(V, s_i) for every row
            # value(V, sp_i): the stored next observation (pre-reset at boundaries); wherever sp_i is bitwise the next step's s the
            # kernel reuses V(s) of that row instead of evaluating the network again
            V.mlp.value_next(data["sp"], data["s"], vs, T, n, out=vsp)
        else:
            vs = vsp = data[source]  # unused by the kernel when adv is NULL
        ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, ptr(data[source]), ptr(data["done"]), ptr(data["episode_end"]), ptr(vs), ptr(vsp),
                                                T, n, float(self.gamma), float(0.0 if math.isnan(float(self.lam)) else self.lam),
                                                ptr(adv), ptr(ret)))

    # ---- episodes! (sampler.jl:175-200): whole episodes only, each one contiguous in the returned columns
    def _episode_quota(self, Neps):
        """The reference runs ``Neps`` episodes one after the other, each to completion (sampler.jl:181-193).  With n parallel
        streams the unbiased equivalent assigns episode k (0-based) to stream ``k % used`` as that stream's ``k // used``-th
        episode after the reset, ``used = min(n, Neps)``: WHICH episodes count is fixed before the rollout (start order), never
        by finishing order -- picking the first finishers would favour short episodes."""
        used = min(self.n, int(Neps))
        quota = (int(Neps) - np.arange(used) + used - 1) // used
        return used, quota

    def episodes_(self, buffer=None, Neps=1, explore=False, i=0, cb=None, return_episodes=False):
        """Resets every stream, then collects ``Neps`` complete episodes: stream ``e < min(n, Neps)`` contributes its first
        ``quota[e]`` episodes after the reset, every one run to its end however many vector steps that takes (episodes are
        stitched across rollout chunks).  Returns the columns (episodes concatenated, episode k = stream ``k % used``, slot
        ``k // used``); with ``return_episodes`` also the 1-based inclusive (start, end) pairs like ``episodes(data)``."""
        self.reset_()
        n = self.n
        used, quota = self._episode_quota(Neps)
        T = max(1, min(self.max_steps, max(16, (1 << 22) // n)))     # rows per chunk bounded (~4 M), any chunk length stitches
        slot = np.zeros(used, dtype=np.int64)
        parts, tags = [], []
        while np.any(slot < quota):
            data = self.steps_(None, Nsteps=n * T, explore=explore, i=i, reset=False)
            ee = data["episode_end"].reshape(T, n)[:, :used].cpu().numpy().astype(bool)
            idx, tag = [], []
            for e in np.flatnonzero(slot < quota):
                start = 0
                for end in np.flatnonzero(ee[:, e]):
                    idx.append(np.arange(start, end + 1) * n + e)
                    tag.append(np.full(end + 1 - start, e + slot[e] * used))
                    slot[e] += 1
                    start = end + 1
                    if slot[e] >= quota[e]:
                        break
                else:                                                 # the episode in progress continues in the next chunk
                    if start < T:
                        idx.append(np.arange(start, T) * n + e)
                        tag.append(np.full(T - start, e + slot[e] * used))
            if idx:
                sel = torch.as_tensor(np.concatenate(idx), device=self.ctx.device)
                parts.append({k: v.index_select(0, sel) for k, v in data.items()})   # copies: the rollout scratch is reused
                tags.append(np.concatenate(tag))
            i += n * T
        cols = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
        tag = np.concatenate(tags)
        order = np.argsort(tag, kind="stable")                        # episode-major, time order kept inside an episode
        cols = {k: v.index_select(0, torch.as_tensor(order, device=self.ctx.device)) for k, v in cols.items()}
        lens = np.bincount(tag, minlength=int(Neps))
        stops = np.cumsum(lens)
        pairs = [(int(b - l + 1), int(b)) for l, b in zip(lens, stops)]
        self.reset_()                                                 # unfinished episodes of the other streams are dropped
        if cb is not None:
            cb(cols)
        if buffer is not None:
            buffer.push_(cols)
        return (cols, pairs) if return_episodes else cols

    # ---- metrics (sampler.jl:203-251).  Per-episode sums run on the device: the concatenated episodes are ONE stream of the
    # returns scan (crux_fill_gae_returns with adv = NULL), cut by the episode_end flags; the value at an episode's first row is
    # Σ_t γ^t x_t of that episode (γ = 1: the plain sum).
    def episode_sums(self, data, pairs, key="r", gamma=1.0):
        """``[metric_by_key(data, start, stop; key) for (start, stop) in episodes]`` (:206,212) for γ = 1 and
        ``discounted_return(data, start, stop, γ)`` (:221-227) otherwise -> float32 tensor ``[len(pairs)]``."""
        ctx = self.ctx
        x = data[key].reshape(-1).to(torch.float32).contiguous()
        T = x.shape[0]
        if T == 0 or not pairs:
            return ctx.empty((0,))
        out = ctx.empty((T,))
        ee, dn = data["episode_end"].reshape(-1).contiguous(), data["done"].reshape(-1).contiguous()
        ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, ptr(x), ptr(dn), ptr(ee), ptr(x), ptr(x), T, 1, float(gamma), 0.0, None, ptr(out)))
        first = torch.as_tensor([p[0] - 1 for p in pairs], device=ctx.device)
        return out.index_select(0, first)

    def metrics_by_key(self, keys, Neps=100, **kw):
        """``metrics_by_key(s::Sampler; keys, Neps)`` (:208-219): ``sum(data[key]) / Neps`` over ``Neps`` fresh episodes."""
        data, pairs = self.episodes_(Neps=Neps, return_episodes=True, **kw)
        return [float(self.episode_sums(data, pairs, k).sum().item()) / Neps for k in keys]

    def metric_by_key(self, key, Neps=100, **kw):
        return self.metrics_by_key([key], Neps=Neps, **kw)[0]

    def discounted_return(self, Neps=100, **kw):
        """``discounted_return(s::Sampler; Neps)`` (:229-232): mean over episodes of Σ γ^t r_t."""
        data, pairs = self.episodes_(Neps=Neps, return_episodes=True, **kw)
        return float(self.episode_sums(data, pairs, "r", float(self.gamma)).mean().item())

    def failure(self, threshold=0.0, Neps=100, **kw):
        """``failure(s::Sampler; threshold, Neps)`` (:235-240): fraction of episodes whose undiscounted return is below the threshold."""
        data, pairs = self.episodes_(Neps=Neps, return_episodes=True, **kw)
        return float((self.episode_sums(data, pairs, "r") < threshold).to(torch.float32).mean().item())

    def undiscounted_return(self, Neps=10, **kw):
        """``undiscounted_return(s::Sampler; Neps)`` (:216): on a device env through ``episodes_`` + the device scan; on a host env
        the batched greedy loop below
        (which episodes count is fixed by ``_episode_quota``: start order, not finishing order)."""
        if self.on_device:
            return self.metric_by_key("r", Neps=Neps, **kw)
        self.reset_()
        used, quota = self._episode_quota(Neps)
        total = np.zeros(self.n)
        count = np.zeros(self.n, dtype=np.int64)
        result = np.zeros(int(Neps))
        steps = np.zeros(self.n, dtype=np.int64)
        while np.any(count[:used] < quota):
            obs = self.cur
            a = action(self.agent.pi, obs)
            sp, r, dn = self.mdp.step(a.cpu().numpy())
            total += r
            steps += 1
            end = np.asarray(dn, bool) | (steps >= self.max_steps)
            nxt = self._tovec(sp)
            if end.any():
                idx = np.flatnonzero(end)
                for e in idx[idx < used]:
                    if count[e] < quota[e]:
                        result[e + count[e] * used] = total[e]
                count[idx] += 1
                total[idx] = 0
                steps[idx] = 0
                nxt = nxt.copy()
                nxt[idx] = self._tovec(self.mdp.reset(idx))
            self.cur.copy_(torch.from_numpy(nxt))
        self.reset_()
        return float(np.mean(result))


def steps_(sampler, buffer=None, **kw):
    """``steps!(sampler, buffer; Nsteps, explore, i, reset, cb)``."""
    return sampler.steps_(buffer, **kw)


def fill_gae_(data, T, n, V, lam, gamma, ctx=None):
    """``fill_gae!`` over a whole ``[T][n]`` rollout dict (sampler.jl:255-273)."""
    ctx = ctx or V.ctx
    vs, vsp = V.mlp.forward(data["s"]), V.mlp.forward(data["sp"])
    ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, ptr(data["r"]), ptr(data["done"]), ptr(data["episode_end"]), ptr(vs), ptr(vsp), T, n,
                                            float(gamma), float(lam), ptr(data["advantage"]), None))


def fill_returns_(data, T, n, gamma, ctx):
    """``fill_returns!`` (sampler.jl:275-281)."""
    ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, ptr(data["r"]), ptr(data["done"]), ptr(data["episode_end"]), ptr(data["r"]),
                                            ptr(data["r"]), T, n, float(gamma), 0.0, None, ptr(data["return"])))
