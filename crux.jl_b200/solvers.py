"""Host-side mirror of the solver surface: ``TrainingParams`` (src/training.jl:1-11), ``OnPolicySolver`` +
``PPO``/``A2C`` (src/model_free/on_policy.jl, rl/ppo.jl, rl/a2c.jl), ``OffPolicySolver`` + ``DQN``/``SAC``
(src/model_free/off_policy.jl, rl/dqn.jl, rl/sac.jl) and ``solve``.

The reference passes loss closures to a generic ``train!``; here the losses of the four north-star algorithms are
fused CUDA kernels selected by name (``ppo_loss``, ``a2c_loss``, ``td_loss``, ``double_Q_loss`` + ``sac_actor_loss``
+ ``sac_temp_loss``); anything else raises -- there is no generic autodiff fallback.
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np
import torch

from . import _abi
from .buffer import ExperienceBuffer, buffer_like, rand_
from .device import ptr, view
from .policies import (ActorCritic, ContinuousNetwork, DiscreteNetwork, DoubleNetwork, GaussianNoiseExplorationPolicy,
                       GaussianPolicy, LinearDecaySchedule, PolicyParams, SquashedGaussianPolicy, action_space, actor,
                       critic, deepcopy, eps_greedy_policy, polyak_average_)
from .logger import LoggerParams, log_undiscounted_return  # noqa: F401  (re-exported: logging.jl lives in logger.py)
from .sampler import Sampler
from .spaces import DiscreteSpace, dim

F32 = np.float32


class Adam:
    """``Flux.Adam(η, β, ϵ)`` [3P]: float32 moments, Float64 scalars.  ``clip_value`` > 0 is
    ``Flux.Optimiser(ClipValue(clip_value), Adam(η, β, ϵ))`` (examples/rl/atari.jl:10), implemented for the pixel network."""

    def __init__(self, eta=F32(3e-4), beta=(0.9, 0.999), eps=1e-8, clip_value=0.0):
        self.eta, self.beta, self.eps, self.clip_value = float(eta), (float(beta[0]), float(beta[1])), float(eps), float(clip_value)


class TrainingParams:
    """training.jl:1-11."""

    def __init__(self, loss=None, optimizer=None, batch_size=128, epochs=80, update_every=1, early_stopping=None,
                 name="", max_batches=math.inf, target_kl=None):
        self.loss, self.optimizer = loss, optimizer or Adam(F32(3e-4))
        self.batch_size, self.epochs, self.update_every = int(batch_size), int(epochs), int(update_every)
        self.early_stopping, self.name, self.max_batches = early_stopping, name, max_batches
        self.target_kl = target_kl  # the only early-stopping rule the fused update evaluates (rl/ppo.jl:59, a2c.jl:46)


# Philox domain separation: every device RNG consumer keys Philox on (seed, counter, row); the sampler's exploration noise
# uses the solver seed itself, replay sampling and the update-time noise (SAC re-sampling, DDPG/TD3 target smoothing) use disjoint keys,
# so that e.g. the replay index of batch row i never comes from the same 128 bits as env stream i's exploration noise.
_DOM_REPLAY, _DOM_UPDATE = 0xA5A55A5A00010001, 0xC3C33C3C00020002


def _dom(seed, domain):
    return (int(seed) ^ domain) & 0xFFFFFFFFFFFFFFFF


def _set_adam(mlp, opt):
    if getattr(opt, "clip_value", 0.0):
        mlp.set_adam(opt.eta, opt.beta, opt.eps, clip_value=opt.clip_value)    # only the pixel network takes it: others raise TypeError
    else:
        mlp.set_adam(opt.eta, opt.beta, opt.eps)


# =============================================================================================== on-policy
class OnPolicySolver:
    """on_policy.jl:31-54."""

    def __init__(self, agent, S, N=1000, dN=200, max_steps=100, log=None, i=0, a_opt=None, c_opt=None, P=None,
                 post_sample_callback=None, post_batch_callback=None, lam_gae=0.95, required_columns=(), a2c=False, seed=0,
                 interaction_storage=None, weight_column="advantage", Vc=None, cost_opt=None):
        self.agent = agent if isinstance(agent, PolicyParams) else PolicyParams(agent)
        self.S, self.N, self.dN, self.max_steps, self.log, self.i = S, int(N), int(dN), int(max_steps), log, int(i)
        self.a_opt, self.c_opt, self.P = a_opt, c_opt, dict(P or {})
        self.post_sample_callback, self.post_batch_callback = post_sample_callback, post_batch_callback
        self.lam_gae, self.required_columns, self.a2c, self.seed = F32(lam_gae), list(required_columns), a2c, int(seed)
        self.interaction_storage = interaction_storage
        self.update_count = 0
        pi = self.agent.pi
        self.weight_column = weight_column
        self.Vc, self.cost_opt = Vc, cost_opt     # on_policy.jl:51-53: parameters specific to cost constraints
        self._pid_state = None
        if Vc is not None:
            assert isinstance(Vc, ContinuousNetwork) and cost_opt is not None
            _set_adam(Vc.mlp, cost_opt.optimizer)
            self._pid_state = pi.ctx.zeros((5,))  # I, smooth_Δ, smooth_Jc, Jc_prev, penalty (the 1-element arrays of 𝒫, ppo.jl:190-201)
        A = pi.A if isinstance(pi, ActorCritic) else pi
        assert (isinstance(A, GaussianPolicy) and not A.squashed) or isinstance(A, DiscreteNetwork), \
            "the on-policy update supports GaussianPolicy(μ, logΣ vector) and DiscreteNetwork (categorical) actors (rl/ppo.jl, examples/rl/cartpole.jl)"
        assert c_opt is None or (isinstance(pi, ActorCritic) and isinstance(pi.C, ContinuousNetwork)), \
            "a critic optimiser needs ActorCritic(actor, ContinuousNetwork)"
        assert Vc is None or isinstance(A, GaussianPolicy), "LagrangePPO: Gaussian actors"
        self._actor = A
        _set_adam(A.mu.mlp if isinstance(A, GaussianPolicy) else A.mlp, a_opt.optimizer)
        if c_opt is not None:
            _set_adam(pi.C.mlp, c_opt.optimizer)
        self.buffer = None
        self.sampler = None
        self.last_info = None

    def _hp(self):
        a, c = self.a_opt, self.c_opt
        tk = a.target_kl if a.target_kl is not None else math.inf
        return _abi.PPOHp(eps_clip=float(self.P.get("eps", 0.2)), lambda_p=float(self.P.get("lp", 1.0)), lambda_e=float(self.P.get("le", 0.1)),
                          target_kl=float(tk), a2c=1 if self.a2c else 0, actor_epochs=a.epochs, actor_batch=a.batch_size,
                          critic_epochs=0 if c is None else c.epochs, critic_batch=1 if c is None else c.batch_size,
                          actor_max_batches=0 if math.isinf(a.max_batches) else int(a.max_batches),
                          critic_max_batches=0 if (c is None or math.isinf(c.max_batches)) else int(c.max_batches))

    def policy_gradient_training(self, D, orders=None):
        """on_policy.jl:56-78: batch_train!(actor) then batch_train!(critic) as ONE stream-ordered launch sequence.
        ``orders = (order_actor, order_critic)`` injects the shuffles (int32 [epochs, n], 0-based) for parity runs."""
        pi, ctx = self.agent.pi, self.agent.pi.ctx
        n = len(D)
        hp = self._hp()
        oa = oc = None
        if orders is not None:
            oa = None if orders[0] is None else ctx.to_device(np.ascontiguousarray(orders[0], dtype=np.int32), torch.int32)
            oc = None if orders[1] is None else ctx.to_device(np.ascontiguousarray(orders[1], dtype=np.int32), torch.int32)
        self.update_count += 1
        self._hp_last, self._n_last = hp, n
        self._keep = (oa, oc)
        if self.Vc is not None:
            return self._lagrange_training(D, hp, oa, oc, None if orders is None or len(orders) < 3 else orders[2])
        ctx.check(ctx.lib.crux_ppo_update_async(self._actor.h, pi.C.mlp.h if self.c_opt is not None else None, ptr(D["s"]), ptr(D["a"]),
                                                ptr(D["logprob"]), ptr(D[self.weight_column]), ptr(D["return"]), n, C.byref(hp), ptr(oa),
                                                ptr(oc), self.seed * 1000003 + self.update_count))
        return self.training_info  # lazily evaluated: reading it synchronises

    def _lagrange_training(self, D, hp, oa, oc, ok):
        """policy_gradient_training (on_policy.jl:56-78) for LagrangePPO: actor with lagrange_ppo_loss, critic, cost critic."""
        pi, ctx, P, k = self.agent.pi, self.agent.pi.ctx, self.P, self.cost_opt
        n = len(D)
        lhp = _abi.LagrangeHp(target_cost=float(P["target_cost"]), penalty_max=float(P["penalty_max"]), Ki_max=float(P["Ki_max"]),
                              Ki=float(P["Ki"]), Kp=float(P["Kp"]), Kd=float(P["Kd"]), ema_alpha=float(P["ema_alpha"]), cost_epochs=k.epochs,
                              cost_batch=k.batch_size, cost_max_batches=0 if math.isinf(k.max_batches) else int(k.max_batches))
        okd = None if ok is None else ctx.to_device(np.ascontiguousarray(ok, dtype=np.int32), torch.int32)
        nmb_a, nmb_c, nmb_k = -(-n // hp.actor_batch), -(-n // hp.critic_batch), -(-n // k.batch_size)
        ia, il = np.zeros((max(1, hp.actor_epochs * nmb_a), 8), F32), np.zeros((max(1, hp.actor_epochs * nmb_a), 8), F32)
        ic, ik = np.zeros((max(1, hp.critic_epochs * nmb_c), 8), F32), np.zeros((max(1, k.epochs * nmb_k), 8), F32)
        ctx.check(ctx.lib.crux_lagrange_ppo_update(self._actor.h, pi.C.mlp.h, self.Vc.mlp.h, ptr(D["s"]), ptr(D["a"]), ptr(D["logprob"]),
                                                   ptr(D["advantage"]), ptr(D["return"]), ptr(D["cost"]), ptr(D["cost_advantage"]),
                                                   ptr(D["cost_return"]), ptr(D["episode_end"]), n, C.byref(hp), C.byref(lhp),
                                                   ptr(self._pid_state), ptr(oa), ptr(oc), ptr(okd), self.seed * 1000003 + self.update_count,
                                                   ptr(ia), ptr(ic), ptr(il), ptr(ik)))
        valid = np.flatnonzero(ia[:, _abi.PPO_VALID] > 0)
        info = {"actor_batches_trained": int(len(valid))}
        if len(valid):
            last, ll = ia[valid[-1]], il[valid[-1]]
            info.update({"actor_loss": float(last[_abi.PPO_LOSS]), "actor_grad_norm": float(last[_abi.PPO_GRAD_NORM]),
                         "entropy": float(last[_abi.PPO_ENTROPY]), "kl": float(last[_abi.PPO_KL]), "clip_fraction": float(last[_abi.PPO_CLIP_FRAC]),
                         "avg_advantage": float(last[_abi.PPO_AVG_ADV]), "avg_return": float(last[_abi.PPO_AVG_RET]),
                         "penalty": float(ll[0]), "cur_cost": float(ll[1]), "prop_term": float(ll[2]), "deriv_term": float(ll[3]),
                         "integral term": float(ll[4]), "p_loss": float(ll[5]), "cost_loss": float(ll[6])})
        for name, rec in (("critic_", ic), ("cost_critic_", ik)):
            v = np.flatnonzero(rec[:, _abi.PPO_VALID] > 0)
            if len(v):
                info.update({name + "loss": float(rec[v[-1], _abi.PPO_LOSS]), name + "grad_norm": float(rec[v[-1], _abi.PPO_GRAD_NORM])})
        self.last_info = info
        return lambda: info

    def training_info(self):
        """Aggregated info of the last update.  ``batch_train!`` pushes the SAME dict for every minibatch
        (training.jl:43), so the reference's aggregates equal the last trained minibatch's values."""
        if self.Vc is not None:
            return self.last_info
        pi, ctx = self.agent.pi, self.agent.pi.ctx
        hp, n = self._hp_last, self._n_last
        pa, pc = C.c_void_p(), C.c_void_p()
        ctx.check(ctx.lib.crux_ppo_info_ptrs(self._actor.h, C.byref(pa), C.byref(pc)))
        nmb_a, nmb_c = -(-n // hp.actor_batch), -(-n // hp.critic_batch)
        ia = view(pa.value, (max(1, hp.actor_epochs * nmb_a), 8), _abi.F32, ctx.device).cpu().numpy()
        ctx.check_flags()
        valid = np.flatnonzero(ia[:, _abi.PPO_VALID] > 0)
        info = {"actor_batches_trained": int(len(valid))}
        if len(valid):
            last = ia[valid[-1]]
            info.update({"actor_loss": float(last[_abi.PPO_LOSS]), "actor_grad_norm": float(last[_abi.PPO_GRAD_NORM]),
                         "entropy": float(last[_abi.PPO_ENTROPY]), "kl": float(last[_abi.PPO_KL])})
            if not self.a2c:
                info.update({"clip_fraction": float(last[_abi.PPO_CLIP_FRAC]), "avg_advantage": float(last[_abi.PPO_AVG_ADV]),
                             "avg_return": float(last[_abi.PPO_AVG_RET])})
        if self.c_opt is not None and hp.critic_epochs > 0:
            ic = view(pc.value, (hp.critic_epochs * nmb_c, 8), _abi.F32, ctx.device).cpu().numpy()
            v = np.flatnonzero(ic[:, _abi.PPO_VALID] > 0)
            info["critic_batches_trained"] = int(len(v))
            if len(v):
                info.update({"critic_loss": float(ic[v[-1], _abi.PPO_LOSS]), "critic_grad_norm": float(ic[v[-1], _abi.PPO_GRAD_NORM])})
        self.last_info = info
        return info


def _whiten_advantage(D, **_):
    """𝒟[:advantage] .= whiten(𝒟[:advantage])  (ppo.jl:61 post_batch_callback / a2c.jl:48 post_sample_callback)."""
    adv = D["advantage"]
    ctx = getattr(D, "ctx", None)
    if ctx is None:
        from .device import default_context
        ctx = default_context()
    ctx.check(ctx.lib.crux_whiten(ctx.h, ptr(adv), adv.shape[0]))


def PPO(pi, S, eps=0.2, lp=1.0, le=0.1, target_kl=0.012, a_opt=None, c_opt=None, log=None, required_columns=(), **kw):
    """``PPO(;π, ϵ=0.2f0, λp=1f0, λe=0.1f0, target_kl=0.012f0, a_opt, c_opt, ...)`` rl/ppo.jl:40-65."""
    a = TrainingParams(loss="ppo_loss", name="actor_", target_kl=target_kl, **(a_opt or {}))
    c = TrainingParams(loss="mse", name="critic_", **(c_opt or {}))
    cols = list(dict.fromkeys(list(required_columns) + ["return", "logprob", "advantage"]))
    return OnPolicySolver(PolicyParams(pi), S, P={"eps": F32(eps), "lp": F32(lp), "le": F32(le)}, a_opt=a, c_opt=c,
                          post_batch_callback=_whiten_advantage, required_columns=cols, log=log, **kw)


def A2C(pi, S, lp=1.0, le=0.1, a_opt=None, c_opt=None, log=None, required_columns=(), **kw):
    """``A2C(;π, λp, λe, ...)`` rl/a2c.jl:30-51 (early stop at KL > 0.015; advantages whitened in post_sample_callback)."""
    a = TrainingParams(loss="a2c_loss", name="actor_", target_kl=0.015, **(a_opt or {}))
    c = TrainingParams(loss="mse", name="critic_", **(c_opt or {}))
    cols = list(dict.fromkeys(list(required_columns) + ["return", "logprob", "advantage"]))
    return OnPolicySolver(PolicyParams(pi), S, P={"lp": F32(lp), "le": F32(le)}, a_opt=a, c_opt=c, a2c=True,
                          post_sample_callback=_whiten_advantage, required_columns=cols, log=log, **kw)


def _record_avgr(D, info=None, **_):
    """post_sample_callback of LagrangePPO (ppo.jl:181-183): info[:avg_r] = sum(r) / sum(episode_end)."""
    if info is not None:
        info["avg_r"] = float(D["r"].sum().item()) / max(float(D["episode_end"].sum().item()), 1.0)


def LagrangePPO(pi, Vc, S, eps=0.2, lp=1.0, le=0.1, lam_gae=0.95, target_kl=0.012, target_cost=0.025, penalty_scale=1.0,
                penalty_max=math.inf, Ki_max=10.0, Ki=1e-3, Kp=1, Kd=0, ema_alpha=0.95, a_opt=None, c_opt=None, cost_opt=None, log=None,
                required_columns=(), **kw):
    """``LagrangePPO(;π::ActorCritic, Vc::ContinuousNetwork, ϵ, λp, λe, λ_gae, target_kl, target_cost, penalty_max, Ki_max, Ki, Kp,
    Kd, ema_α, a_opt, c_opt, cost_opt, ...)`` rl/ppo.jl:133-214: PPO whose actor loss adds a PID-weighted clipped cost-advantage
    surrogate (``lagrange_ppo_loss`` :70-131), plus a cost critic trained on ``cost_return``.  The rollout needs a host env that
    reports ``last_info["cost"]`` (sampler.jl:76-78,114); ``penalty_scale`` is carried but unused, as in the reference (:112)."""
    a = TrainingParams(loss="lagrange_ppo_loss", name="actor_", target_kl=target_kl, **(a_opt or {}))
    c = TrainingParams(loss="mse", name="critic_", **(c_opt or {}))
    k = TrainingParams(loss="mse", name="cost_critic_", **(cost_opt or {}))
    cols = list(dict.fromkeys(list(required_columns) + ["return", "advantage", "logprob", "cost_advantage", "cost", "cost_return"]))
    P = {"eps": F32(eps), "lp": F32(lp), "le": F32(le), "target_cost": F32(target_cost), "penalty_scale": F32(penalty_scale),
         "penalty_max": F32(penalty_max), "Ki_max": F32(Ki_max), "Ki": F32(Ki), "Kp": Kp, "Kd": Kd, "ema_alpha": float(ema_alpha)}
    return OnPolicySolver(PolicyParams(pi), S, P=P, a_opt=a, c_opt=c, Vc=Vc, cost_opt=k, lam_gae=lam_gae, post_batch_callback=_whiten_advantage,
                          post_sample_callback=_record_avgr, required_columns=cols, log=log, **kw)


def REINFORCE(pi, S, a_opt=None, log=None, required_columns=(), **kw):
    """``REINFORCE(;π, a_opt, ...)`` rl/reinforce.jl:27-39: ``reinforce_loss = -mean(logpdf(π, s, a) .* return)`` (:4-13), early
    stop at KL > 0.015, no critic.  It is the a2c_loss kernel head with the ``return`` column as the per-row weight, λp = 1 and
    λe = 0, so no kernel is specific to it."""
    a = TrainingParams(loss="reinforce_loss", name="actor_", target_kl=0.015, **(a_opt or {}))
    cols = list(dict.fromkeys(list(required_columns) + ["return", "logprob"]))
    return OnPolicySolver(PolicyParams(pi), S, P={"lp": F32(1.0), "le": F32(0.0)}, a_opt=a, c_opt=None, a2c=True,
                          weight_column="return", required_columns=cols, log=log, **kw)


def _solve_on_policy(S, mdp):
    """``POMDPs.solve(𝒮::OnPolicySolver, mdp)`` on_policy.jl:80-109."""
    pi = S.agent.pi
    if S.buffer is None:
        S.buffer = ExperienceBuffer(S.S, S.agent.space, S.dN, S.required_columns, ctx=pi.ctx)
    if S.sampler is None or S.sampler.mdp is not mdp:
        S.sampler = Sampler(mdp, S.agent, S=S.S, required_columns=S.required_columns, lam=S.lam_gae, max_steps=S.max_steps, seed=S.seed,
                            Vc=S.Vc)
    D, s = S.buffer, S.sampler
    if S.log is not None and S.log.sampler is None:
        S.log.sampler = s
    if S.log is not None:
        S.log.log(S.i, solver=S)
    i0 = S.i
    for S.i in range(i0, i0 + S.N - S.dN + 1, S.dN):
        info = {}
        cb = (lambda data: S.post_sample_callback(data, info=info, solver=S)) if S.post_sample_callback else None
        if D.next_ind != 1:
            pass  # capacity == ΔN: every rollout overwrites the whole buffer (push! wraps back to row 1)
        s.steps_(D, Nsteps=S.dN, explore=True, i=S.i, reset=True, cb=_wrap_cb(cb, pi.ctx), store=S.interaction_storage)
        if S.post_batch_callback:
            S.post_batch_callback(D, info=info, solver=S)
        training_info = S.policy_gradient_training(D)
        if S.log is not None:
            S.log.log((S.i + 1, S.i + S.dN), lambda **_: training_info(), info, solver=S)
    S.i += S.dN
    return pi


class _DataView(dict):
    """A rollout dict that also knows its context (so callbacks written against a buffer work on it)."""
    ctx = None


def _wrap_cb(cb, ctx):
    if cb is None:
        return None

    def f(data):
        d = _DataView(data)
        d.ctx = ctx
        return cb(d)
    return f


# =============================================================================================== off-policy
class OffPolicySolver:
    """off_policy.jl:37-64."""

    def __init__(self, agent, S, N=1000, dN=4, max_steps=100, log=None, i=0, a_opt=None, c_opt=None, P=None, kind="dqn",
                 tau=0.005, buffer_size=1000, required_columns=(), buffer=None, buffer_init=None, prioritized=False,
                 priority_params=None, weighted_loss=False, seed=0, post_sample_callback=None):
        self.agent, self.S = agent, S
        self.N, self.dN, self.max_steps, self.log, self.i = int(N), int(dN), int(max_steps), log, int(i)
        self.a_opt, self.c_opt, self.P, self.kind, self.tau = a_opt, c_opt, dict(P or {}), kind, F32(tau)
        self.required_columns = list(required_columns)
        ctx = agent.pi.ctx
        self.buffer = buffer if buffer is not None else ExperienceBuffer(S, agent.space, buffer_size, self.required_columns,
                                                                         prioritized=prioritized, priority_params=priority_params, ctx=ctx)
        self.buffer_init = int(buffer_init) if buffer_init is not None else max(c_opt.batch_size, 200)
        self.weighted_loss, self.seed = weighted_loss, int(seed)
        self.post_sample_callback = post_sample_callback
        self.sampler = None
        self.train_count = 0
        self.last_info = {}
        self._sac = None
        self._ddpg = None
        if kind in ("dqn", "softq"):
            _set_adam(agent.pi.mlp, c_opt.optimizer)
        elif kind in ("ddpg", "td3"):
            pi, tg = agent.pi, agent.pi_target
            twin = isinstance(pi.C, DoubleNetwork)
            crit, crit_t = ([pi.C.N1, pi.C.N2], [tg.C.N1, tg.C.N2]) if twin else ([pi.C], [tg.C])
            _set_adam(pi.A.mlp, a_opt.optimizer)
            for c in crit:
                _set_adam(c.mlp, c_opt.optimizer)
            st = C.c_void_p()
            ctx.check(ctx.lib.crux_ddpg_create(pi.A.mlp.h, tg.A.mlp.h, crit[0].mlp.h, crit_t[0].mlp.h, crit[1].mlp.h if twin else None,
                                               crit_t[1].mlp.h if twin else None, self.tau, C.byref(st)))
            self._ddpg = st
        else:
            pi = agent.pi
            _set_adam(pi.A.mu.mlp, a_opt.optimizer)
            _set_adam(pi.C.N1.mlp, c_opt.optimizer)
            _set_adam(pi.C.N2.mlp, c_opt.optimizer)
            st = C.c_void_p()
            tg = agent.pi_target
            ctx.check(ctx.lib.crux_sac_create(pi.A.h, pi.C.N1.mlp.h, pi.C.N2.mlp.h, tg.C.N1.mlp.h, tg.C.N2.mlp.h,
                                              F32(self.P["SAC_log_alpha"]), F32(self.P["SAC_H_target"]), float(self.P.get("alpha_eta", F32(3e-4))),
                                              self.tau, C.byref(st)))
            self._sac = st

    def value_training(self, D, gamma, draws=None, noise=None):
        """off_policy.jl:66-111.  ``draws[epoch]`` / ``noise[epoch]`` inject the sampling draws and SAC noise (parity runs)."""
        ctx, lib = self.agent.pi.ctx, self.agent.pi.ctx.lib
        B = D.capacity
        infos = []
        if self.kind in ("sac", "ddpg", "td3"):
            # off_policy.jl:83-89,93 for these solvers is not fused into crux_sac_train / crux_ddpg_train: refuse instead of silently
            # leaving every priority at max_priority / ignoring the importance weights
            if self.buffer.isprioritized() or self.weighted_loss:
                raise NotImplementedError(f"{self.kind}: prioritized replay / weighted_loss are only wired for DQN and SoftQ (off_policy.jl:83-93)")
            if self.kind == "sac" and (self.c_opt.update_every != 1 or self.a_opt.update_every != 1):
                raise NotImplementedError("sac: c_opt/a_opt.update_every != 1 is not supported by the fused SAC update (off_policy.jl:91,96)")
        # Training info of the cycle's LAST epoch is read back when a logger is attached (one synchronisation per value_training call; the
        # reference aggregates every epoch's Dict, off_policy.jl:104-110 -- reading all of them would put a host sync behind every train! step)
        want_info = self.log is not None or getattr(self, "collect_info", False)
        for epoch in range(self.c_opt.epochs):
            info_buf = np.zeros(8, F32) if (want_info and epoch == self.c_opt.epochs - 1) else None
            self.train_count += 1
            rand_(D, self.buffer, i=self.i, draws=None if draws is None else [draws[epoch]], seed=_dom(self.seed, _DOM_REPLAY),
                  ctr=2 * self.train_count)
            s, a, sp, r, dn = D.column("s"), D.column("a"), D.column("sp"), D.column("r"), D.column("done")
            if self.kind in ("dqn", "softq"):
                pi, tgt = self.agent.pi, self.agent.pi_target
                nA = len(pi.outputs)
                y = ctx.empty((B,))
                q_sp = tgt.mlp.forward(sp)
                if self.kind == "dqn":
                    ctx.check(lib.crux_dqn_target(ctx.h, ptr(r), ptr(dn), ptr(q_sp), B, nA, float(gamma), ptr(y)))  # rl/dqn.jl:4-6
                else:
                    ctx.check(lib.crux_softq_target(ctx.h, ptr(r), ptr(dn), ptr(q_sp), B, nA, float(gamma), float(self.P["alpha"]),
                                                    ptr(y)))                                                          # rl/softq.jl:13-17
                if self.buffer.isprioritized():                                                                       # off_policy.jl:83
                    q = pi.mlp.forward(s)
                    qsa, td = ctx.empty((B,)), ctx.empty((B,))
                    ctx.check(lib.crux_discrete_q_sa(ctx.h, ptr(q), ptr(a), B, nA, ptr(qsa)))
                    ctx.check(lib.crux_td_error(ctx.h, ptr(qsa), ptr(y), B, ptr(td)))
                    self.buffer.update_priorities_(D.indices_dev(), td)
                w = D.column("weight") if (self.weighted_loss and "weight" in D.schema) else None
                if epoch % self.c_opt.update_every == 0:
                    pi.mlp.train_dqn(s, a, y, w, B, info_buf)                                                         # off_policy.jl:91-93
                    if info_buf is not None:
                        n_ = self.c_opt.name
                        self.last_info = {n_ + "loss": float(info_buf[0]), n_ + "grad_norm": float(info_buf[1]), "Qavg": float(info_buf[2])}
            elif self.kind in ("ddpg", "td3"):
                sm = self.P.get("pi_smooth")                      # None: plain ddpg_target (rl/ddpg.jl:6-8)
                e = None if noise is None else ctx.to_device(noise[epoch], torch.float32)
                lo, hi = sm.bounds() if sm is not None else (None, None)
                do_c = epoch % self.c_opt.update_every == 0       # off_policy.jl:91
                do_a = epoch % self.a_opt.update_every == 0       # off_policy.jl:96 (TD3's delayed actor = a_opt.update_every 2)
                ctx.check(lib.crux_ddpg_train(self._ddpg, ptr(s), ptr(a), ptr(sp), ptr(r), ptr(dn), B, float(gamma), 0 if sm is None else 1,
                                              0.0 if sm is None else float(F32(sm.sigma(self.i))), -math.inf if sm is None else float(sm.eps_min),
                                              math.inf if sm is None else float(sm.eps_max), ptr(lo), 0 if lo is None else lo.size, ptr(hi),
                                              0 if hi is None else hi.size, ptr(e), _dom(self.seed, _DOM_UPDATE), 3 * self.train_count, 1 if do_c else 0,
                                              1 if do_a else 0, None, ptr(info_buf)))
                if info_buf is not None:
                    cn, an = self.c_opt.name, self.a_opt.name
                    self.last_info = {cn + "loss": float(info_buf[1]), cn + "grad_norm": float(info_buf[2]), "Q1avg": float(info_buf[6])}
                    if do_a:
                        self.last_info.update({an + "loss": float(info_buf[3]), an + "grad_norm": float(info_buf[4])})
                    if self.kind == "td3":
                        self.last_info["Q2avg"] = float(info_buf[7])
            else:
                e = (None, None, None) if noise is None else [ctx.to_device(x, torch.float32) for x in noise[epoch]]
                ctx.check(lib.crux_sac_train(self._sac, ptr(s), ptr(a), ptr(sp), ptr(r), ptr(dn), B, float(gamma), ptr(e[0]), ptr(e[1]), ptr(e[2]),
                                             _dom(self.seed, _DOM_UPDATE), 3 * self.train_count, None, ptr(info_buf)))
                if info_buf is not None:
                    cn, an = self.c_opt.name, self.a_opt.name
                    self.last_info = {"temp_loss": float(info_buf[0]), cn + "loss": float(info_buf[1]), cn + "grad_norm": float(info_buf[2]),
                                      an + "loss": float(info_buf[3]), an + "grad_norm": float(info_buf[4]), "entropy": float(info_buf[5]),
                                      "Q1avg": float(info_buf[6]), "Q2avg": float(info_buf[7])}
        if self.kind in ("dqn", "softq"):  # no separate actor: target update after the epoch loop (off_policy.jl:108)
            polyak_average_(self.agent.pi_target, self.agent.pi, self.tau)
        return [self.last_info] if (want_info and self.last_info) else infos

    def __del__(self):
        try:
            if self._sac is not None and self.agent.pi.ctx.h:
                self.agent.pi.ctx.lib.crux_sac_destroy(self._sac)
                self._sac = None
            if self._ddpg is not None and self.agent.pi.ctx.h:
                self.agent.pi.ctx.lib.crux_ddpg_destroy(self._ddpg)
                self._ddpg = None
        except Exception:
            pass


def DQN(pi, S, N, dN=4, pi_explore=None, c_opt=None, log=None, **kw):
    """``DQN(;π::DiscreteNetwork, N, ΔN=4, π_explore=ϵGreedyPolicy(LinearDecaySchedule(1., 0.1, N÷2), π.outputs), c_opt, ...)``
    rl/dqn.jl:29-46 (c_opt.epochs = ΔN: one gradient step per env step)."""
    assert isinstance(pi, DiscreteNetwork)
    pe = pi_explore or eps_greedy_policy(LinearDecaySchedule(1.0, 0.1, N // 2), pi.outputs)
    c = TrainingParams(**{"loss": "td_loss", "name": "critic_", "epochs": dN, **(c_opt or {})})
    agent = PolicyParams(pi, pi_explore=pe, pi_target=deepcopy(pi))
    return OffPolicySolver(agent, S, N=N, dN=dN, c_opt=c, kind="dqn", log=log, **kw)


def SoftQ(pi, S, N, dN=4, c_opt=None, log=None, alpha=1.0, **kw):
    """``SoftQ(;π::DiscreteNetwork, N, ΔN=4, c_opt, α=1f0, ...)`` rl/softq.jl:36-58: the policy becomes always-stochastic with
    ``softmax(value(π, s) ./ α)`` logits (:47-48), the target is ``r + γ(1-done)·α·logsumexp(Q⁻(sp)/α)`` (:8,13-17), the critic
    trains with td_loss like DQN (default ``c_opt = (;epochs=4)`` :26)."""
    assert isinstance(pi, DiscreteNetwork)
    pi.always_stochastic, pi.temperature = True, float(F32(alpha))
    c = TrainingParams(**{"loss": "td_loss", "name": "critic_", **({"epochs": 4} if c_opt is None else c_opt)})
    agent = PolicyParams(pi, pi_target=deepcopy(pi))
    return OffPolicySolver(agent, S, N=N, dN=dN, c_opt=c, kind="softq", P={"alpha": F32(alpha)}, log=log, **kw)


def SAC(pi, S, N=1000, dN=50, SAC_alpha=1.0, SAC_H_target=None, pi_explore=None, SAC_alpha_opt=None, a_opt=None, c_opt=None,
        log=None, **kw):
    """``SAC(;π::ActorCritic{_, DoubleNetwork}, ΔN=50, SAC_α=1f0, SAC_H_target=-dim(A), π_explore=GaussianNoiseExplorationPolicy(0.1f0), ...)``
    rl/sac.jl:77-106."""
    assert isinstance(pi, ActorCritic) and isinstance(pi.A, SquashedGaussianPolicy) and isinstance(pi.C, DoubleNetwork)
    H = F32(-int(np.prod(dim(action_space(pi))))) if SAC_H_target is None else F32(SAC_H_target)
    a = TrainingParams(**{"loss": "sac_actor_loss", "name": "actor_", **(a_opt or {})})
    c = TrainingParams(**{"loss": "double_Q_loss", "name": "critic_", "epochs": dN, **(c_opt or {})})
    t = TrainingParams(**{"loss": "sac_temp_loss", "name": "temp_", **(SAC_alpha_opt or {})})
    agent = PolicyParams(pi, pi_explore=pi_explore or GaussianNoiseExplorationPolicy(F32(0.1)), pi_target=deepcopy(pi))
    P = {"SAC_log_alpha": F32(math.log(SAC_alpha)), "SAC_H_target": H, "alpha_eta": t.optimizer.eta}
    return OffPolicySolver(agent, S, N=N, dN=dN, a_opt=a, c_opt=c, P=P, kind="sac", log=log, **kw)


def DDPG(pi, S, N=1000, dN=50, pi_explore=None, a_opt=None, c_opt=None, pi_smooth=None, smoothed_target=False, log=None, **kw):
    """``DDPG(;π::ActorCritic, ΔN=50, π_explore=GaussianNoiseExplorationPolicy(0.1f0), a_loss=ddpg_actor_loss, c_loss=td_loss(),
    target_fn=ddpg_target, π_smooth=GaussianNoiseExplorationPolicy(0.1f0, ϵ_min=-0.5f0, ϵ_max=0.5f0), ...)`` rl/ddpg.jl:45-67.
    ``smoothed_target=True`` selects ``smoothed_ddpg_target`` (:14-17) as ``target_fn``."""
    assert isinstance(pi, ActorCritic) and isinstance(pi.A, ContinuousNetwork) and isinstance(pi.C, ContinuousNetwork)
    sm = pi_smooth or GaussianNoiseExplorationPolicy(F32(0.1), eps_min=-0.5, eps_max=0.5)
    a = TrainingParams(**{"loss": "ddpg_actor_loss", "name": "actor_", **(a_opt or {})})
    c = TrainingParams(**{"loss": "td_loss", "name": "critic_", "epochs": dN, **(c_opt or {})})
    agent = PolicyParams(pi, pi_explore=pi_explore or GaussianNoiseExplorationPolicy(F32(0.1)), pi_target=deepcopy(pi))
    return OffPolicySolver(agent, S, N=N, dN=dN, a_opt=a, c_opt=c, P={"pi_smooth": sm if smoothed_target else None}, kind="ddpg", log=log, **kw)


def TD3(pi, S, N=1000, dN=50, pi_smooth=None, pi_explore=None, a_opt=None, c_opt=None, log=None, **kw):
    """``TD3(;π, ΔN=50, π_smooth=GaussianNoiseExplorationPolicy(0.1f0, ϵ_min=-0.5f0, ϵ_max=0.5f0),
    π_explore=GaussianNoiseExplorationPolicy(0.1f0), a_loss=td3_actor_loss, c_loss=double_Q_loss(), target_fn=td3_target, ...)``
    rl/td3.jl:34-57; the delayed actor of the paper is ``a_opt=(;update_every=2)`` (training.jl:7, off_policy.jl:96; default 1)."""
    assert isinstance(pi, ActorCritic) and isinstance(pi.A, ContinuousNetwork) and isinstance(pi.C, DoubleNetwork)
    sm = pi_smooth or GaussianNoiseExplorationPolicy(F32(0.1), eps_min=-0.5, eps_max=0.5)
    a = TrainingParams(**{"loss": "td3_actor_loss", "name": "actor_", **(a_opt or {})})
    c = TrainingParams(**{"loss": "double_Q_loss", "name": "critic_", "epochs": dN, **(c_opt or {})})
    agent = PolicyParams(pi, pi_explore=pi_explore or GaussianNoiseExplorationPolicy(F32(0.1)), pi_target=deepcopy(pi))
    return OffPolicySolver(agent, S, N=N, dN=dN, a_opt=a, c_opt=c, P={"pi_smooth": sm}, kind="td3", log=log, **kw)


def _solve_off_policy(S, mdp):
    """``POMDPs.solve(𝒮::OffPolicySolver, mdp)`` off_policy.jl:113-150."""
    pi = S.agent.pi
    D = buffer_like(S.buffer, capacity=S.c_opt.batch_size)
    gamma = F32(mdp.gamma)
    extra = [k for k in S.buffer.schema if k not in ("s", "a", "sp", "r", "done", "episode_end")]
    if S.sampler is None or S.sampler.mdp is not mdp:
        S.sampler = Sampler(mdp, S.agent, S=S.S, max_steps=S.max_steps, required_columns=extra, seed=S.seed)
    s, n = S.sampler, S.sampler.n
    if S.log is not None and S.log.sampler is None:
        S.log.sampler = s
    up = lambda k: -(-int(k) // n) * n  # vector envs advance n transitions per step
    Nfill = max(0, S.buffer_init - len(S.buffer))
    istart = S.i
    if Nfill > 0:
        S.i += up(Nfill)
        s.steps_(S.buffer, Nsteps=up(Nfill), explore=True, i=S.i)
    if S.log is not None:
        S.log.log(S.i, solver=S)
    dN = up(S.dN)
    for S.i in range(S.i, istart + S.N - dN + 1, dN):
        s.steps_(S.buffer, Nsteps=dN, explore=True, i=S.i)
        S.value_training(D, gamma)
        if S.log is not None:
            S.log.log((S.i + 1, S.i + dN), S.last_info, solver=S)
    S.i += dN
    pi.ctx.check_flags()
    return pi


def solve(S, mdp):
    """``POMDPs.solve(𝒮, mdp)``."""
    if isinstance(S, OnPolicySolver):
        return _solve_on_policy(S, mdp)
    if isinstance(S, OffPolicySolver):
        return _solve_off_policy(S, mdp)
    raise TypeError(type(S))
