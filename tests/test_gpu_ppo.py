"""-m gpu: hot path (iii) -- crux_ppo_update (policy_gradient_training, on_policy.jl:56-78 = batch_train! of the actor with
ppo_loss / a2c_loss then of the critic with Flux.mse; training.jl:28-55) against the oracle's train_step loop on the
same minibatch orders.  Tolerance 1e-5 rtol on parameters (north_star), looser on reduced scalars where noted."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, assert_params_close, dev, host, make_mlp, mlp_grads, mlp_params, p

pytestmark = pytest.mark.gpu


def _setup(ctx, crux, n, seed, sdim=17, adim=6, hidden=64):
    rng = np.random.default_rng(seed)
    mu = o.MLP([sdim, hidden, hidden, adim], [1, 1, 0], rng)
    cr = o.MLP([sdim, hidden, hidden, 1], [1, 1, 0], rng)
    ls = np.full(adim, -0.5, F32)
    pi = o.GaussianPolicy(mu, ls)
    hm = make_mlp(ctx, mu.dims, mu.acts, mu.flat())
    hc = make_mlp(ctx, cr.dims, cr.acts, cr.flat())
    h = C.c_void_p()
    ctx.check(ctx.lib.crux_gaussian_create(ctx.h, hm, adim, p(ls), 0, 1.0, C.byref(h)))
    s = rng.standard_normal((n, sdim)).astype(F32)
    # actions/logprobs from a slightly different (older) policy so that ratios differ from 1
    eps = rng.standard_normal((n, adim)).astype(F32)
    old = o.GaussianPolicy(o.MLP(mu.dims, mu.acts, Ws=[w.detach().numpy() * F32(0.97) for w in mu.W],
                                 bs=[b.detach().numpy() for b in mu.b]), ls - F32(0.05))
    a, lp = old.exploration(s, eps)
    D = {"s": s, "a": a.detach().numpy(), "logprob": lp.detach().numpy()[:, 0],
         "advantage": o.whiten(rng.standard_normal(n).astype(F32)), "return": rng.standard_normal(n).astype(F32)}
    return rng, pi, cr, (hm, hc, h), D


def _orders(rng, n, epochs, start=None):
    order = np.arange(n) if start is None else start
    out = []
    for _ in range(epochs):
        order = order[rng.permutation(n)]  # shuffle! permutes the already-shuffled buffer (experience_buffer.jl:118-124)
        out.append(order.copy())
    return np.stack(out).astype(np.int32) if epochs else np.zeros((0, n), np.int32)


def _oracle_train(params, loss_fn, opt, D, orders, batch, stop=None, max_batches=math.inf):
    recs, total, stopped = [], 0, False
    n = orders.shape[1] if len(orders) else 0
    for order in orders:
        for st in range(0, n, batch):
            idx = order[st:st + batch]
            mb = {k: v[idx] for k, v in D.items()}
            info = {}
            o.train_step(params, lambda inf, mb=mb: loss_fn(mb, inf), opt, info)
            recs.append(dict(info))
            total += 1
            if total >= max_batches or (stop is not None and stop(info)):
                stopped = True
                break
        if stopped:
            break
    return recs


def _hp(crux, **kw):
    d = dict(eps_clip=0.2, lambda_p=1.0, lambda_e=0.1, target_kl=math.inf, a2c=0, actor_epochs=2, actor_batch=128,
             critic_epochs=2, critic_batch=128, actor_max_batches=0, critic_max_batches=0)
    d.update(kw)
    return crux._abi.PPOHp(**d)


def _run(ctx, crux, handles, D, hp, oa, oc, n):
    hm, hc, h = handles
    nmb_a = -(-n // hp.actor_batch); nmb_c = -(-n // hp.critic_batch)
    ia = np.zeros((max(1, hp.actor_epochs * nmb_a), 8), F32)
    ic = np.zeros((max(1, hp.critic_epochs * nmb_c), 8), F32)
    d = {k: dev(ctx, v) for k, v in D.items()}
    rc = ctx.lib.crux_ppo_update(h, hc, p(d["s"]), p(d["a"]), p(d["logprob"]), p(d["advantage"]), p(d["return"]), n, C.byref(hp),
                                 p(dev(ctx, oa)) if oa is not None else None, p(dev(ctx, oc)) if oc is not None else None, 42,
                                 p(ia), p(ic))
    ctx.check(rc)
    return ia, ic


@pytest.mark.parametrize("n,ab,cb", [(512, 128, 128), (1000, 128, 256), (300, 300, 64), (4096, 1024, 4096)])
def test_ppo_update_matches_oracle(ctx, crux, n, ab, cb):
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=n)
    hp = _hp(crux, actor_batch=ab, critic_batch=cb)
    oa = _orders(rng, n, hp.actor_epochs)
    oc = _orders(rng, n, hp.critic_epochs, start=oa[-1])
    P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(0.1)}
    ra = _oracle_train(pi.params(), lambda mb, inf: o.ppo_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab)
    rc_ = _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), o.Adam(F32(3e-4)), D, oc, cb)
    ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
    A = crux._abi
    assert len(ra) == ia.shape[0] and len(rc_) == ic.shape[0]
    for k, rec in enumerate(ra):
        assert ia[k, A.PPO_VALID] == 1.0
        assert_close(ia[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, atol=1e-5, what=f"actor loss mb {k}")
        assert_close(ia[k, A.PPO_KL], rec["kl"], rtol=1e-3, atol=2e-6, what=f"kl mb {k}")
        assert_close(ia[k, A.PPO_ENTROPY], rec["entropy"], rtol=1e-5, what="entropy")
        assert_close(ia[k, A.PPO_CLIP_FRAC], rec["clip_fraction"], rtol=0, atol=2.0 / ab, what="clip_fraction")
        assert_close(ia[k, A.PPO_AVG_ADV], rec["avg_advantage"], rtol=1e-3, atol=1e-5, what="avg_advantage")
        assert_close(ia[k, A.PPO_AVG_RET], rec["avg_return"], rtol=1e-3, atol=1e-5, what="avg_return")
        assert_close(ia[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what="grad_norm")
    for k, rec in enumerate(rc_):
        assert_close(ic[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, what=f"critic loss mb {k}")
        assert_close(ic[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what="critic grad_norm")
    hm, hc, h = handles
    assert_params_close(mlp_params(ctx, hm), pi.mu.flat(), 3e-4, len(ra), what="actor params")
    assert_params_close(mlp_params(ctx, hc), cr.flat(), 3e-4, len(rc_), what="critic params")
    lsp = C.c_void_p(); ctx.check(ctx.lib.crux_gaussian_log_sigma_ptr(h, C.byref(lsp)))
    ls = np.empty(6, F32); ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(ls), lsp, 24)); ctx.sync()
    assert_close(ls, pi.log_sigma.detach().numpy(), rtol=1e-5, atol=2e-6, what="logΣ")


def test_kl_early_stop(ctx, crux):
    """rl/ppo.jl:59 + training.jl:46,49: the minibatch whose KL exceeds target_kl is still applied, then training stops."""
    n, ab = 1024, 256
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=3)
    D["advantage"] = (D["advantage"] * F32(30)).astype(F32)  # large steps -> KL grows quickly
    hp = _hp(crux, actor_batch=ab, actor_epochs=6, critic_epochs=0, target_kl=2e-3)
    oa = _orders(rng, n, hp.actor_epochs)
    P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(0.1)}
    ra = _oracle_train(pi.params(), lambda mb, inf: o.ppo_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab,
                       stop=lambda info: info["kl"] > 2e-3)
    ia, _ = _run(ctx, crux, handles, D, hp, oa, None, n)
    valid = ia[:, crux._abi.PPO_VALID]
    assert 0 < len(ra) < ia.shape[0], "the test must actually stop early"
    assert valid[:len(ra)].all() and not valid[len(ra):].any()
    assert ia[len(ra) - 1, crux._abi.PPO_KL] > 2e-3
    assert_params_close(mlp_params(ctx, handles[0]), pi.mu.flat(), 3e-4, len(ra), what="actor params after early stop")


def test_a2c_loss_and_max_batches(ctx, crux):
    n, ab = 640, 128
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=11)
    hp = _hp(crux, a2c=1, actor_batch=ab, actor_epochs=3, critic_epochs=1, critic_batch=n, actor_max_batches=7, lambda_e=0.0)
    oa = _orders(rng, n, 3); oc = _orders(rng, n, 1, start=oa[-1])
    P = {"lp": F32(1.0), "le": F32(0.0)}
    ra = _oracle_train(pi.params(), lambda mb, inf: o.a2c_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab, max_batches=7)
    rc_ = _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), o.Adam(F32(3e-4)), D, oc, n)
    ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
    assert len(ra) == 7 and ia[:7, crux._abi.PPO_VALID].all() and not ia[7:, crux._abi.PPO_VALID].any()
    for k, rec in enumerate(ra):
        assert_close(ia[k, crux._abi.PPO_LOSS], rec["loss"], rtol=1e-4, atol=1e-5, what=f"a2c loss {k}")
    assert_params_close(mlp_params(ctx, handles[0]), pi.mu.flat(), 3e-4, 7)
    assert_params_close(mlp_params(ctx, handles[1]), cr.flat(), 3e-4, 1)


def test_reinforce_loss_is_the_a2c_head_weighted_by_returns(ctx, crux):
    """rl/reinforce.jl:4-13 ``-mean(logpdf .* return)`` with the early stop at KL > 0.015 (:36): the a2c kernel head with the
    ``return`` column as the per-row weight, λp = 1, λe = 0 and no critic -- what ``crux_b200.REINFORCE`` launches."""
    n, ab = 768, 256
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=17)
    D["return"] = (D["return"] * F32(4)).astype(F32)
    hp = _hp(crux, a2c=1, actor_batch=ab, actor_epochs=4, critic_epochs=0, lambda_p=1.0, lambda_e=0.0, target_kl=0.015)
    oa = _orders(rng, n, 4)
    ra = _oracle_train(pi.params(), lambda mb, inf: o.reinforce_loss(pi, {}, mb, inf), o.Adam(F32(3e-4)), D, oa, ab,
                       stop=lambda info: info["kl"] > 0.015)
    d = dict(D); d["advantage"] = D["return"]
    hm, hc, h = handles
    ia, _ = _run(ctx, crux, (hm, None, h), d, hp, oa, None, n)
    A = crux._abi
    valid = ia[:, A.PPO_VALID]
    assert valid[:len(ra)].all() and not valid[len(ra):].any()
    for k, rec in enumerate(ra):
        assert_close(ia[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, atol=1e-5, what=f"reinforce loss {k}")
        assert_close(ia[k, A.PPO_KL], rec["kl"], rtol=1e-3, atol=2e-6, what=f"kl {k}")
        assert_close(ia[k, A.PPO_ENTROPY], rec["entropy"], rtol=1e-5, what="entropy")
        assert_close(ia[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what="grad_norm")
    assert_params_close(mlp_params(ctx, hm), pi.mu.flat(), 3e-4, len(ra), what="actor params")


def test_lagrange_ppo_update_matches_oracle(ctx, crux):
    """LagrangePPO (rl/ppo.jl:70-214): the PID penalty evaluated once per minibatch on sum(cost)/sum(episode_end), the clipped
    cost-advantage surrogate, the (1 + penalty) normalisation, then the critic and the cost critic on cost_return."""
    n, ab, cb = 1024, 256, 512
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=23)
    hm, hc, h = handles
    vc = o.MLP([17, 64, 64, 1], [o.ACT_TANH, o.ACT_TANH, o.ACT_IDENTITY], rng)
    hk = make_mlp(ctx, vc.dims, vc.acts, vc.flat())
    ctx.check(ctx.lib.crux_mlp_set_adam(hk, float(F32(3e-4)), 0.9, 0.999, 1e-8))
    D["cost"] = (rng.random(n) < 0.3).astype(F32) * rng.random(n).astype(F32)
    D["cost_advantage"] = rng.standard_normal(n).astype(F32)
    D["cost_return"] = rng.standard_normal(n).astype(F32)
    D["episode_end"] = (rng.random(n) < 0.1)
    hp = _hp(crux, actor_batch=ab, actor_epochs=3, critic_epochs=1, critic_batch=cb)
    lhp = crux._abi.LagrangeHp(target_cost=0.025, penalty_max=math.inf, Ki_max=10.0, Ki=0.05, Kp=1.0, Kd=0.5, ema_alpha=0.95,
                               cost_epochs=2, cost_batch=cb, cost_max_batches=0)
    oa = _orders(rng, n, 3); oc = _orders(rng, n, 1, start=oa[-1]); ok = _orders(rng, n, 2, start=oc[-1])
    P = o.lagrange_params(eps=0.2, lp=1.0, le=0.1, target_cost=0.025, Ki=0.05, Kp=1, Kd=0.5)
    ra = _oracle_train(pi.params(), lambda mb, inf: o.lagrange_ppo_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab)
    rc_ = _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), o.Adam(F32(3e-4)), D, oc, cb)
    rk = _oracle_train(vc.params(), lambda mb, inf: o.value_mse_loss(vc, dict(mb, **{"return": mb["cost_return"]})), o.Adam(F32(3e-4)), D, ok, cb)
    d = {k: dev(ctx, v.astype(np.uint8) if v.dtype == bool else v) for k, v in D.items()}
    state = dev(ctx, np.zeros(5, F32))
    ia = np.zeros((len(ra), 8), F32); il = np.zeros((len(ra), 8), F32); ic = np.zeros((len(rc_), 8), F32); ik = np.zeros((len(rk), 8), F32)
    ctx.check(ctx.lib.crux_lagrange_ppo_update(h, hc, hk, p(d["s"]), p(d["a"]), p(d["logprob"]), p(d["advantage"]), p(d["return"]), p(d["cost"]),
                                               p(d["cost_advantage"]), p(d["cost_return"]), p(d["episode_end"]), n, C.byref(hp), C.byref(lhp),
                                               p(state), p(dev(ctx, oa)), p(dev(ctx, oc)), p(dev(ctx, ok)), 42, p(ia), p(ic), p(il), p(ik)))
    A = crux._abi
    assert any(rec["penalty"] > 0 for rec in ra), "the test must exercise a non-zero penalty"
    for k, rec in enumerate(ra):
        assert ia[k, A.PPO_VALID] == 1.0 and il[k, 7] == 1.0
        want = [rec["penalty"], rec["cur_cost"], rec["prop_term"], rec["deriv_term"], rec["integral term"], rec["p_loss"], rec["cost_loss"]]
        assert_close(il[k, :7], want, rtol=2e-4, atol=2e-6, what=f"lagrange info mb {k}")
        assert_close(ia[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, atol=1e-5, what=f"actor loss mb {k}")
        assert_close(ia[k, A.PPO_KL], rec["kl"], rtol=1e-3, atol=2e-6, what=f"kl mb {k}")
        assert_close(ia[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what=f"grad_norm mb {k}")
    for k, rec in enumerate(rc_):
        assert_close(ic[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, what=f"critic loss mb {k}")
    for k, rec in enumerate(rk):
        assert_close(ik[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, what=f"cost critic loss mb {k}")
    assert_params_close(mlp_params(ctx, hm), pi.mu.flat(), 3e-4, len(ra), what="actor params")
    assert_params_close(mlp_params(ctx, hc), cr.flat(), 3e-4, len(rc_), what="critic params")
    assert_params_close(mlp_params(ctx, hk), vc.flat(), 3e-4, len(rk), what="cost critic params")
    st = host(state)
    assert_close(st, [P["I"], P["smooth_D"], P["smooth_Jc"], P["Jc_prev"], ra[-1]["penalty"]], rtol=2e-4, atol=2e-6, what="PID state")
    lsp = C.c_void_p(); ctx.check(ctx.lib.crux_gaussian_log_sigma_ptr(h, C.byref(lsp)))
    ls = np.empty(6, F32); ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(ls), lsp, 24)); ctx.sync()
    assert_close(ls, pi.log_sigma.detach().numpy(), rtol=1e-5, atol=2e-6, what="logΣ")


def _setup_categorical(ctx, crux, n, seed, sdim=4, nA=3, hidden=64, act=o.ACT_RELU):
    rng = np.random.default_rng(seed)
    net = o.MLP([sdim, hidden, hidden, nA], [act, act, o.ACT_IDENTITY], rng)
    cr = o.MLP([sdim, hidden, hidden, 1], [act, act, o.ACT_IDENTITY], rng)
    pi = o.DiscreteNetwork(net, list(range(nA)))
    hm = make_mlp(ctx, net.dims, net.acts, net.flat())
    hc = make_mlp(ctx, cr.dims, cr.acts, cr.flat())
    h = C.c_void_p()
    ctx.check(ctx.lib.crux_categorical_create(ctx.h, hm, nA, C.byref(h)))
    s = rng.standard_normal((n, sdim)).astype(F32)
    # actions / log-probabilities from a slightly different (older) policy so that the ratios differ from 1
    old = o.DiscreteNetwork(o.MLP(net.dims, net.acts, Ws=[w.detach().numpy() * F32(0.9) for w in net.W], bs=[b.detach().numpy() for b in net.b]),
                            list(range(nA)))
    ai, lp = old.exploration(s, rng.random(n))
    a_oh = np.eye(nA, dtype=F32)[ai.numpy()]
    D = {"s": s, "a": a_oh, "logprob": lp.detach().numpy()[:, 0].astype(F32), "advantage": o.whiten(rng.standard_normal(n).astype(F32)),
         "return": rng.standard_normal(n).astype(F32)}
    return rng, pi, cr, (hm, hc, h), D


@pytest.mark.parametrize("loss,n,ab,nA", [("ppo", 512, 128, 2), ("ppo", 1000, 256, 5), ("a2c", 600, 200, 3), ("reinforce", 384, 128, 4)])
def test_categorical_actor_update_matches_oracle(ctx, crux, loss, n, ab, nA):
    """ppo_loss / a2c_loss / reinforce_loss with a DiscreteNetwork actor (examples/rl/cartpole.jl:8-9: Chain(Dense(4,64,relu), Dense(64,64,relu),
    Dense(64,nA))): logpdf = categorical_logpdf (policies.jl:135), entropy per sample (:152-155) averaged by e_loss -- crux_categorical_create +
    crux_ppo_update against the oracle's autograd train_step loop on the same minibatch orders."""
    rng, pi, cr, handles, D = _setup_categorical(ctx, crux, n, seed=n + nA, nA=nA)
    hm, hc, h = handles
    a2c = loss != "ppo"
    le = 0.0 if loss == "reinforce" else 0.1
    hp = _hp(crux, actor_batch=ab, critic_batch=ab, a2c=1 if a2c else 0, lambda_e=le)
    oa = _orders(rng, n, hp.actor_epochs); oc = _orders(rng, n, hp.critic_epochs, start=oa[-1])
    P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(le)}
    Dd = dict(D)
    if loss == "reinforce":      # reinforce_loss = -mean(logp .* return): the a2c head with the return as the weight and no entropy term
        Dd["advantage"] = D["return"]
        fn = lambda mb, inf: o.reinforce_loss(pi, P, mb, inf)
    elif loss == "a2c":
        fn = lambda mb, inf: o.a2c_loss(pi, P, mb, inf)
    else:
        fn = lambda mb, inf: o.ppo_loss(pi, P, mb, inf)
    # raw gradient of the FIRST minibatch against autograd (the update below starts from the same parameters)
    mb0 = {k: v[oa[0][:ab]] for k, v in Dd.items()}
    for q_ in pi.params():
        q_.grad = None
    fn(mb0, {}).backward()
    g_want = o.flat_grads(pi.params())
    hp1 = _hp(crux, actor_batch=ab, critic_batch=ab, a2c=1 if a2c else 0, lambda_e=le, actor_epochs=1, critic_epochs=0, actor_max_batches=1)
    saved = mlp_params(ctx, hm).copy()
    _run(ctx, crux, handles, Dd, hp1, oa[:1], None, n)
    g_got = mlp_grads(ctx, hm)
    assert_close(g_got, g_want, rtol=2e-4, atol=2e-7 + 1e-5 * float(np.abs(g_want).max()), what="raw gradient of the first minibatch")
    ctx.check(ctx.lib.crux_mlp_set_params(hm, p(saved)))
    ctx.check(ctx.lib.crux_mlp_set_adam(hm, float(F32(3e-4)), 0.9, 0.999, 1e-8))
    ra = _oracle_train(pi.params(), fn, o.Adam(F32(3e-4)), Dd, oa, ab)
    rc_ = _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), o.Adam(F32(3e-4)), D, oc, ab)
    ia, ic = _run(ctx, crux, handles, Dd, hp, oa, oc, n)
    A = crux._abi
    assert len(ra) == ia.shape[0]
    for k, rec in enumerate(ra):
        assert ia[k, A.PPO_VALID] == 1.0
        assert_close(ia[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, atol=1e-5, what=f"actor loss mb {k}")
        assert_close(ia[k, A.PPO_KL], rec["kl"], rtol=1e-3, atol=2e-6, what=f"kl mb {k}")
        assert_close(ia[k, A.PPO_ENTROPY], rec["entropy"], rtol=1e-5, atol=1e-6, what="entropy")
        assert_close(ia[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what="grad_norm")
        if loss == "ppo":
            assert_close(ia[k, A.PPO_CLIP_FRAC], rec["clip_fraction"], rtol=0, atol=2.0 / ab, what="clip_fraction")
    for k, rec in enumerate(rc_):
        assert_close(ic[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, what=f"critic loss mb {k}")
    assert_params_close(mlp_params(ctx, hm), pi.net.flat(), 3e-4, len(ra), what="actor params")
    assert_params_close(mlp_params(ctx, hc), cr.flat(), 3e-4, len(rc_), what="critic params")
    # the Gaussian entry points refuse the handle instead of misreading it
    out = dev(ctx, np.zeros((4, nA), F32))
    assert ctx.lib.crux_gaussian_action(h, p(dev(ctx, D["s"][:4])), 4, p(out)) != 0
    ctx.lib.crux_gaussian_destroy(h)


def test_device_permutation_is_a_permutation_and_trains(ctx, crux):
    """order == NULL: device-generated shuffles.  Each epoch must visit every row exactly once (checked through a
    critic whose target equals a per-row id-free constant is not observable; instead check determinism + change)."""
    n = 2048
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=5)
    hp = _hp(crux, actor_batch=512, critic_batch=512, actor_epochs=2, critic_epochs=2)
    before = mlp_params(ctx, handles[0]).copy()
    ia, ic = _run(ctx, crux, handles, D, hp, None, None, n)
    assert ia[:, crux._abi.PPO_VALID].all() and np.isfinite(ia).all() and np.isfinite(ic).all()
    after = mlp_params(ctx, handles[0])
    assert not np.array_equal(before, after)
    # full-batch epoch means are permutation invariant: avg_adv summed over an epoch's minibatches == mean(adv)
    tot = ia[:4, crux._abi.PPO_AVG_ADV].mean()
    assert_close(tot, D["advantage"].mean(), rtol=1e-3, atol=1e-5)


def test_argument_errors(ctx, crux):
    n = 64
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=1)
    hm, hc, h = handles
    hp = _hp(crux, actor_batch=0)
    d = {k: dev(ctx, v) for k, v in D.items()}
    args = (p(d["s"]), p(d["a"]), p(d["logprob"]), p(d["advantage"]), p(d["return"]))
    assert ctx.lib.crux_ppo_update(h, hc, *args, n, C.byref(hp), None, None, 0, None, None) == 1
    hp = _hp(crux)
    assert ctx.lib.crux_ppo_update(h, hc, *args, 0, C.byref(hp), None, None, 0, None, None) == 1
    assert ctx.lib.crux_ppo_update(h, hc, args[0], args[1], args[2], args[3], None, n, C.byref(hp), None, None, 0, None, None) == 1


def test_fused_path_equals_generic_engine(ctx, crux):
    """The fused minibatch / forward kernels (ppo_fused.cu) against the generic layer-by-layer engine on the same inputs
    (CRUX_NO_FUSED=1 selects the generic path at call time)."""
    import os
    n = 3000  # not a multiple of the 64-row tile
    results = []
    for no_fused in ("", "1"):
        if no_fused:
            os.environ["CRUX_NO_FUSED"] = "1"
        else:
            os.environ.pop("CRUX_NO_FUSED", None)
        try:
            rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=77)
            hp = _hp(crux, actor_batch=1000, critic_batch=1500, actor_epochs=2, critic_epochs=1)
            oa = _orders(rng, n, 2); oc = _orders(rng, n, 1, start=oa[-1])
            ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
            y = ctx.empty((n, 6))
            ctx.check(ctx.lib.crux_mlp_forward(handles[0], p(dev(ctx, D["s"])), n, p(y)))
            results.append((ia.copy(), ic.copy(), mlp_params(ctx, handles[0]).copy(), mlp_params(ctx, handles[1]).copy(), host(y).copy()))
        finally:
            os.environ.pop("CRUX_NO_FUSED", None)
    f, g = results
    assert_close(f[0][:, :7], g[0][:, :7], rtol=2e-4, atol=2e-6, what="actor info fused vs generic")
    assert_close(f[1][:, :2], g[1][:, :2], rtol=2e-4, atol=2e-6, what="critic info fused vs generic")
    assert_params_close(f[2], g[2], 3e-4, 6, what="actor params fused vs generic")
    assert_params_close(f[3], g[3], 3e-4, 2, what="critic params fused vs generic")
    assert_close(f[4], g[4], rtol=1e-5, atol=1e-6, what="forward fused vs generic")


def test_tensor_core_path_equals_ffma_path(ctx, crux):
    """The 3xTF32 tensor-core GEMMs of the fused minibatch kernel (default) against its all-FFMA variant (CRUX_NO_MMA=1):
    fp32-level agreement on the loss records, the gradient norm and the updated parameters, on a ragged minibatch."""
    import os
    n = 5000  # not a multiple of the 64-row tile
    results = []
    for no_mma in ("", "1"):
        if no_mma:
            os.environ["CRUX_NO_MMA"] = "1"
        else:
            os.environ.pop("CRUX_NO_MMA", None)
        try:
            rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=99)
            hp = _hp(crux, actor_batch=2048, critic_batch=2500, actor_epochs=2, critic_epochs=2)
            oa = _orders(rng, n, 2); oc = _orders(rng, n, 2, start=oa[-1])
            ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
            results.append((ia.copy(), ic.copy(), mlp_params(ctx, handles[0]).copy(), mlp_params(ctx, handles[1]).copy()))
        finally:
            os.environ.pop("CRUX_NO_MMA", None)
    tc, ff = results
    A = crux._abi
    assert_close(tc[0][:, A.PPO_LOSS], ff[0][:, A.PPO_LOSS], rtol=2e-5, atol=1e-6, what="actor loss tc vs ffma")
    assert_close(tc[0][:, A.PPO_GRAD_NORM], ff[0][:, A.PPO_GRAD_NORM], rtol=2e-5, what="actor grad norm tc vs ffma")
    assert_close(tc[1][:, A.PPO_LOSS], ff[1][:, A.PPO_LOSS], rtol=2e-5, what="critic loss tc vs ffma")
    assert_close(tc[1][:, A.PPO_GRAD_NORM], ff[1][:, A.PPO_GRAD_NORM], rtol=2e-5, what="critic grad norm tc vs ffma")
    assert_params_close(tc[2], ff[2], 3e-4, 6, what="actor params tc vs ffma")
    assert_params_close(tc[3], ff[3], 3e-4, 4, what="critic params tc vs ffma")


def test_full_size_update_properties(ctx, crux):
    """BASELINE config[1] update shape (131 072 rows, minibatches of 32 768, actor then critic) through crux_ppo_update:
    (1) bit-reproducible: two runs from the same parameters give identical parameters and info records (fixed-order partial
        reductions, no atomics); (2) the tensor-core kernel and the all-FFMA kernel agree to fp32 accuracy at this size;
    (3) the first minibatch's loss is the oracle's ppo_loss on those rows."""
    import os
    n, mb = 131072, 32768
    out = {}
    for tag, env in (("tc1", ""), ("tc2", ""), ("ffma", "1")):
        if env:
            os.environ["CRUX_NO_MMA"] = env
        try:
            rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=123)
            hp = _hp(crux, actor_batch=mb, critic_batch=mb, actor_epochs=1, critic_epochs=1, lambda_e=0.0)
            oa = _orders(rng, n, 1); oc = _orders(rng, n, 1, start=oa[-1])
            ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
            out[tag] = (ia.copy(), ic.copy(), mlp_params(ctx, handles[0]).copy(), mlp_params(ctx, handles[1]).copy())
            if tag == "tc1":
                idx = oa[0][:mb]
                mbD = {k: v[idx] for k, v in D.items()}
                info = {}
                want = float(o.ppo_loss(pi, {"eps": F32(0.2), "lp": F32(1), "le": F32(0.0)}, mbD, info))
                assert_close(ia[0, crux._abi.PPO_LOSS], want, rtol=2e-5, atol=1e-6, what="first minibatch ppo_loss at full size")
        finally:
            os.environ.pop("CRUX_NO_MMA", None)
    for k in range(4):
        assert np.array_equal(out["tc1"][k], out["tc2"][k]), "the update is not bit-reproducible"
    A = crux._abi
    assert_close(out["tc1"][0][:, A.PPO_LOSS], out["ffma"][0][:, A.PPO_LOSS], rtol=2e-5, atol=1e-6, what="actor loss tc vs ffma (full size)")
    # first minibatch: same parameters on both paths; later ones see parameters after Adam steps that amplify rounding differences
    # in near-zero gradient coordinates (assert_params_close documents the effect)
    assert_close(out["tc1"][0][:1, A.PPO_GRAD_NORM], out["ffma"][0][:1, A.PPO_GRAD_NORM], rtol=2e-5, what="actor grad norm tc vs ffma (first minibatch)")
    assert_close(out["tc1"][0][:, A.PPO_GRAD_NORM], out["ffma"][0][:, A.PPO_GRAD_NORM], rtol=1e-3, what="actor grad norm tc vs ffma (full size)")
    assert_close(out["tc1"][1][:, A.PPO_LOSS], out["ffma"][1][:, A.PPO_LOSS], rtol=2e-5, what="critic loss tc vs ffma (full size)")
    assert_params_close(out["tc1"][2], out["ffma"][2], 3e-4, 4, what="actor params tc vs ffma (full size)")
    assert_params_close(out["tc1"][3], out["ffma"][3], 3e-4, 4, what="critic params tc vs ffma (full size)")
