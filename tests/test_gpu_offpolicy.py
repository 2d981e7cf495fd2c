"""-m gpu: off-policy updates (src/model_free/off_policy.jl:66-111): crux_dqn_train (td_loss, utils.jl:76-87) and
crux_sac_train (rl/sac.jl:4-52: target -> temperature -> double-Q critics -> actor -> polyak) and crux_ddpg_train (rl/ddpg.jl,
rl/td3.jl: (smoothed) target -> critic(s) -> deterministic actor -> polyak of the whole target policy) against the oracle."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, assert_params_close, dev, host, make_mlp, mlp_grads, mlp_params, p

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("weighted", [False, True])
def test_dqn_train(ctx, weighted):
    rng = np.random.default_rng(5)
    net = o.MLP([2, 8, 4], [2, 0], rng)  # README example network (README.md:72-82)
    ref = o.DiscreteNetwork(net, range(4))
    h = make_mlp(ctx, net.dims, net.acts, net.flat())
    ctx.check(ctx.lib.crux_mlp_set_adam(h, float(F32(3e-4)), 0.9, 0.999, 1e-8))
    opt = o.Adam(F32(3e-4))
    B = 128
    for step in range(3):
        s = rng.standard_normal((B, 2)).astype(F32)
        oh = np.eye(4, dtype=F32)[rng.integers(0, 4, B)]
        y = rng.standard_normal(B).astype(F32)
        w = rng.random(B).astype(F32) if weighted else None
        info = {}
        o.train_step(net.params(), lambda inf: o.td_loss(ref.value(s, oh), y, w, inf), opt, info)
        out = np.zeros(3, F32)
        ctx.check(ctx.lib.crux_dqn_train(h, p(dev(ctx, s)), p(dev(ctx, oh)), p(dev(ctx, y)), p(dev(ctx, w)) if weighted else None, B, p(out)))
        assert_close(out, [info["loss"], info["grad_norm"], info["Qavg"]], rtol=1e-4, atol=1e-6, what=f"info step {step}")
        assert_params_close(mlp_params(ctx, h), net.flat(), 3e-4, step + 1, what=f"params step {step}")
    ctx.lib.crux_mlp_destroy(h)


def _sac_setup(ctx, sdim, A, H, seed):
    rng = np.random.default_rng(seed)
    actor = o.MLP([sdim, H, H, 2 * A], [2, 2, 0], rng)
    q1, q2 = o.MLP([sdim + A, H, H, 1], [2, 2, 0], rng), o.MLP([sdim + A, H, H, 1], [2, 2, 0], rng)
    q1t, q2t = q1.clone(), q2.clone()
    for t in (q1t, q2t):  # targets that differ from the online nets
        with torch.no_grad():
            for w in t.W:
                w.mul_(0.9)
    hs = [make_mlp(ctx, m.dims, m.acts, m.flat()) for m in (actor, q1, q2, q1t, q2t)]
    for h in hs[:3]:
        ctx.check(ctx.lib.crux_mlp_set_adam(h, float(F32(3e-4)), 0.9, 0.999, 1e-8))
    pol = C.c_void_p()
    ctx.check(ctx.lib.crux_gaussian_create(ctx.h, hs[0], A, None, 1, 1.0, C.byref(pol)))
    return rng, (actor, q1, q2, q1t, q2t), hs, pol


@pytest.mark.parametrize("sdim,A,H,B", [(11, 3, 32, 64), (376, 17, 256, 256)])
def test_sac_train(ctx, sdim, A, H, B):
    rng, nets, hs, pol = _sac_setup(ctx, sdim, A, H, seed=sdim)
    actor, q1, q2, q1t, q2t = nets
    pi = o.SquashedGaussianPolicy(lambda s: actor(s)[:, :A], lambda s: actor(s)[:, A:], 1.0, actor.params())
    log_alpha0, h_target, tau, gamma = F32(math.log(0.2)), F32(-A), F32(0.005), F32(0.99)
    st = C.c_void_p()
    ctx.check(ctx.lib.crux_sac_create(pol, hs[1], hs[2], hs[3], hs[4], log_alpha0, h_target, float(F32(3e-4)), tau, C.byref(st)))
    log_alpha = torch.tensor([float(log_alpha0)], requires_grad=True)
    opt_t, opt_c, opt_a = o.Adam(F32(3e-4)), o.Adam(F32(3e-4)), o.Adam(F32(3e-4))
    for step in range(2):
        D = {"s": rng.standard_normal((B, sdim)).astype(F32), "a": np.tanh(rng.standard_normal((B, A))).astype(F32),
             "sp": rng.standard_normal((B, sdim)).astype(F32), "r": rng.standard_normal(B).astype(F32),
             "done": (rng.random(B) < 0.2).astype(np.uint8)}
        e1, e2, e3 = (rng.standard_normal((B, A)).astype(F32) for _ in range(3))
        # ---- oracle: off_policy.jl:71-101 order
        y = o.sac_target(pi, q1t, q2t, D, gamma, float(log_alpha.detach()[0]), e1)
        info = {}
        o.train_step([log_alpha], lambda inf: o.sac_temp_loss(pi, D, log_alpha[0], h_target, e2), opt_t, info, "temp_")
        sa = np.concatenate([D["s"], D["a"]], 1)

        def closs(inf):
            return 0.5 * (o.td_loss(q1(sa), y, None, inf, "Q1avg") + o.td_loss(q2(sa), y, None, inf, "Q2avg"))
        o.train_step(q1.params() + q2.params(), closs, opt_c, info, "critic_")
        o.train_step(actor.params(), lambda inf: o.sac_actor_loss(pi, q1, q2, D, float(log_alpha.detach()[0]), e3, inf), opt_a, info, "actor_")
        o.polyak_average(q1t.params(), q1.params(), tau)
        o.polyak_average(q2t.params(), q2.params(), tau)
        # ---- device
        out = np.zeros(8, F32)
        yd = ctx.empty((B,))
        ctx.check(ctx.lib.crux_sac_train(st, p(dev(ctx, D["s"])), p(dev(ctx, D["a"])), p(dev(ctx, D["sp"])), p(dev(ctx, D["r"])),
                                         p(dev(ctx, D["done"])), B, gamma, p(dev(ctx, e1)), p(dev(ctx, e2)), p(dev(ctx, e3)), 0, 0,
                                         p(yd), p(out)))
        assert_close(host(yd), y.numpy()[:, 0], rtol=1e-4, atol=1e-4, what="sac_target")
        want = [info["temp_loss"], info["critic_loss"], info["critic_grad_norm"], info["actor_loss"], info["actor_grad_norm"],
                info["entropy"], info["Q1avg"], info["Q2avg"]]
        assert_close(out, want, rtol=2e-3, atol=1e-4, what=f"info step {step}")
        la = np.zeros(1, F32); ctx.check(ctx.lib.crux_sac_log_alpha(st, p(la)))
        assert_close(la[0], float(log_alpha.detach()[0]), rtol=1e-5, atol=1e-6, what="log alpha")
        for name, h, m in (("q1", hs[1], q1), ("q2", hs[2], q2), ("actor", hs[0], actor), ("q1 target", hs[3], q1t), ("q2 target", hs[4], q2t)):
            assert_params_close(mlp_params(ctx, h), m.flat(), 3e-4, step + 1, what=f"{name} params step {step}")
    ctx.lib.crux_sac_destroy(st)


def test_sac_device_noise_runs(ctx):
    rng, nets, hs, pol = _sac_setup(ctx, 11, 3, 32, seed=2)
    st = C.c_void_p()
    ctx.check(ctx.lib.crux_sac_create(pol, hs[1], hs[2], hs[3], hs[4], F32(-1.6), F32(-3), float(F32(3e-4)), F32(0.005), C.byref(st)))
    B = 128
    D = [rng.standard_normal((B, 11)).astype(F32), np.tanh(rng.standard_normal((B, 3))).astype(F32), rng.standard_normal((B, 11)).astype(F32),
         rng.standard_normal(B).astype(F32), (rng.random(B) < 0.2).astype(np.uint8)]
    dd = [dev(ctx, x) for x in D]
    before = mlp_params(ctx, hs[0]).copy()
    out = np.zeros(8, F32)
    for k in range(3):
        ctx.check(ctx.lib.crux_sac_train(st, *[p(x) for x in dd], B, F32(0.99), None, None, None, 11, 3 * k, None, p(out)))
        assert np.isfinite(out).all()
    assert not np.array_equal(before, mlp_params(ctx, hs[0]))
    ctx.lib.crux_sac_destroy(st)


# ------------------------------------------------------------------------------------------------ DDPG / TD3
def _ddpg_setup(ctx, sdim, A, H, twin, seed):
    rng = np.random.default_rng(seed)
    actor = o.MLP([sdim, H, H, A], [o.ACT_RELU, o.ACT_RELU, o.ACT_TANH], rng)
    crit = [o.MLP([sdim + A, H, H, 1], [o.ACT_RELU, o.ACT_RELU, o.ACT_IDENTITY], rng) for _ in range(2 if twin else 1)]
    at, ct = actor.clone(), [c.clone() for c in crit]
    for t in [at] + ct:  # targets that differ from the online nets
        with torch.no_grad():
            for w in t.W:
                w.mul_(0.9)
    mk = lambda m: make_mlp(ctx, m.dims, m.acts, m.flat())
    ha, hc = mk(actor), [mk(c) for c in crit]
    for h in [ha] + hc:
        ctx.check(ctx.lib.crux_mlp_set_adam(h, float(F32(3e-4)), 0.9, 0.999, 1e-8))
    return rng, actor, at, crit, ct, ha, mk(at), hc, [mk(c) for c in ct]


@pytest.mark.parametrize("twin,smooth", [(False, False), (False, True), (True, True)])
def test_ddpg_td3_train(ctx, twin, smooth):
    """twin=False, smooth=False: DDPG (rl/ddpg.jl:6-8,25); smooth=True: smoothed_ddpg_target (:14-17); twin=True: TD3
    (rl/td3.jl:4-12) with the delayed actor (a_opt.update_every = 2: the actor and the target update run on even epochs only)."""
    sdim, A, H, B = 11, 3, 32, 96
    rng, actor, at, crit, ct, ha, hat, hc, hct = _ddpg_setup(ctx, sdim, A, H, twin, seed=7 + twin)
    tau, gamma = F32(0.005), F32(0.99)
    st = C.c_void_p()
    ctx.check(ctx.lib.crux_ddpg_create(ha, hat, hc[0], hct[0], hc[1] if twin else None, hct[1] if twin else None, tau, C.byref(st)))
    opt_c, opt_a = o.Adam(F32(3e-4)), o.Adam(F32(3e-4))
    a_min, a_max = np.full(A, -0.9, F32), np.array([0.8], F32)     # per-dimension and broadcast bounds
    n_actor = 0
    for step in range(3):
        D = {"s": rng.standard_normal((B, sdim)).astype(F32), "a": np.tanh(rng.standard_normal((B, A))).astype(F32),
             "sp": rng.standard_normal((B, sdim)).astype(F32), "r": rng.standard_normal(B).astype(F32),
             "done": (rng.random(B) < 0.2).astype(np.uint8)}
        eps = rng.standard_normal((B, A)).astype(F32)
        sm = dict(eps=eps, sigma=F32(0.4), eps_min=-0.5, eps_max=0.5, a_min=a_min, a_max=a_max) if smooth else None
        do_actor = (not twin) or step % 2 == 0
        # ---- oracle: off_policy.jl:71-101 order
        y = o.ddpg_target(at, ct, D, gamma, sm)
        sa = np.concatenate([D["s"], D["a"]], 1)
        info = {}
        if twin:
            closs = lambda inf: 0.5 * (o.td_loss(crit[0](sa), y, None, inf, "Q1avg") + o.td_loss(crit[1](sa), y, None, inf, "Q2avg"))
        else:
            closs = lambda inf: o.td_loss(crit[0](sa), y, None, inf, "Q1avg")
        o.train_step(sum((c.params() for c in crit), []), closs, opt_c, info, "critic_")
        if do_actor:
            n_actor += 1
            o.train_step(actor.params(), lambda inf: o.ddpg_actor_loss(actor, crit[0], D), opt_a, info, "actor_")
            o.polyak_average(at.params(), actor.params(), tau)
            for c, t in zip(crit, ct):
                o.polyak_average(t.params(), c.params(), tau)
        # ---- device
        out = np.zeros(8, F32)
        yd = ctx.empty((B,))
        ctx.check(ctx.lib.crux_ddpg_train(st, p(dev(ctx, D["s"])), p(dev(ctx, D["a"])), p(dev(ctx, D["sp"])), p(dev(ctx, D["r"])),
                                          p(dev(ctx, D["done"])), B, gamma, 1 if smooth else 0, F32(0.4), F32(-0.5), F32(0.5), p(a_min), A,
                                          p(a_max), 1, p(dev(ctx, eps)), 0, 0, 1, 1 if do_actor else 0, p(yd), p(out)))
        assert_close(host(yd), y.numpy()[:, 0], rtol=1e-4, atol=1e-4, what="target")
        assert_close(out[[1, 2, 6]], [info["critic_loss"], info["critic_grad_norm"], info["Q1avg"]], rtol=2e-3, atol=1e-4, what=f"critic info {step}")
        if twin:
            assert_close(out[7], info["Q2avg"], rtol=2e-3, atol=1e-4, what="Q2avg")
        if do_actor:
            assert_close(out[[3, 4]], [info["actor_loss"], info["actor_grad_norm"]], rtol=2e-3, atol=1e-4, what=f"actor info {step}")
        for name, h, m, k in [("actor", ha, actor, n_actor), ("actor target", hat, at, n_actor)] + \
                [(f"q{j + 1}", hc[j], crit[j], step + 1) for j in range(len(crit))] + [(f"q{j + 1} target", hct[j], ct[j], n_actor) for j in range(len(crit))]:
            assert_params_close(mlp_params(ctx, h), m.flat(), 3e-4, max(k, 1), what=f"{name} params step {step}")
    # device noise: runs, respects the bounds of the smoothed action only through y (finite), and c_opt.update_every can skip the critic
    before = mlp_params(ctx, hc[0]).copy()
    dd = [dev(ctx, D[k]) for k in ("s", "a", "sp", "r", "done")]
    ctx.check(ctx.lib.crux_ddpg_train(st, *[p(x) for x in dd], B, gamma, 1, F32(0.2), F32(-0.5), F32(0.5), None, 0, None, 0, None, 5, 9, 0, 0,
                                      p(yd), p(out)))
    assert np.isfinite(host(yd)).all() and np.array_equal(before, mlp_params(ctx, hc[0]))
    assert ctx.lib.crux_ddpg_train(st, *[p(x) for x in dd], B, gamma, 1, F32(0.2), F32(-0.5), F32(0.5), p(a_min), 2, None, 0, None, 5, 9, 0, 0,
                                   None, None) != 0      # a_min with neither 1 nor adim entries
    ctx.lib.crux_ddpg_destroy(st)


def test_noise_explore_matches_oracle_and_statistics(ctx):
    """exploration(::GaussianNoiseExplorationPolicy) policies.jl:510-514."""
    rng = np.random.default_rng(0)
    B, A = 1000, 6
    a = rng.standard_normal((B, A)).astype(F32); eps = rng.standard_normal((B, A)).astype(F32)
    lo, hi = np.array([-1.0], F32), np.linspace(0.5, 1.5, A).astype(F32)
    ad = dev(ctx, a.copy())
    ctx.check(ctx.lib.crux_noise_explore(ctx.h, p(ad), B, A, F32(0.3), F32(-0.25), F32(0.4), p(lo), 1, p(hi), A, p(dev(ctx, eps)), 0, 0))
    want = o.gaussian_noise_exploration(a, eps, 0.3, -0.25, 0.4, lo, hi).numpy()
    assert np.array_equal(host(ad), want)
    z = dev(ctx, np.zeros((200000, 4), F32))
    ctx.check(ctx.lib.crux_noise_explore(ctx.h, p(z), 200000, 4, F32(0.5), -math.inf, math.inf, None, 0, None, 0, None, 3, 1))
    zz = host(z)
    assert abs(zz.mean()) < 5e-3 and abs(zz.std() - 0.5) < 5e-3 and abs(np.corrcoef(zz[:, 0], zz[:, 1])[0, 1]) < 1e-2
