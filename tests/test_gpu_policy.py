"""-m gpu: GaussianPolicy / SquashedGaussianPolicy / DiscreteNetwork heads and the batched rollout step
(hot path (i), sampler.jl:71-137) against the oracle; shape contracts of test/policy_tests.jl."""
import ctypes as C
import math

import numpy as np
import pytest
import torch

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, dev, host, make_mlp, p

pytestmark = pytest.mark.gpu


def _gauss(ctx, mu_ref, ls, squashed=False, ascale=1.0):
    hm = make_mlp(ctx, mu_ref.dims, mu_ref.acts, mu_ref.flat())
    h = C.c_void_p()
    ls_arr = None if ls is None else np.ascontiguousarray(ls, dtype=F32)
    adim = mu_ref.dims[-1] if ls is not None else mu_ref.dims[-1] // 2
    ctx.check(ctx.lib.crux_gaussian_create(ctx.h, hm, adim, p(ls_arr), 1 if squashed else 0, ascale, C.byref(h)))
    return hm, h


@pytest.mark.parametrize("B", [1, 100, 4096])
def test_gaussian_policy(ctx, B):
    rng = np.random.default_rng(B)
    mu = o.MLP([17, 64, 64, 6], [1, 1, 0], rng)
    ls = np.linspace(-1, 0.5, 6).astype(F32)
    ref = o.GaussianPolicy(mu, ls)
    hm, h = _gauss(ctx, mu, ls)
    s = rng.standard_normal((B, 17)).astype(F32)
    eps = rng.standard_normal((B, 6)).astype(F32)
    a, lp = ctx.empty((B, 6)), ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_explore(h, p(dev(ctx, s)), B, p(dev(ctx, eps)), 0, 0, p(a), p(lp)))
    a0, lp0 = ref.exploration(s, eps)
    assert_close(host(a), a0.detach().numpy(), rtol=1e-5, atol=1e-5, what="a")
    assert_close(host(lp), lp0.detach().numpy()[:, 0], rtol=1e-5, atol=2e-5, what="logprob")
    # logpdf(π, s, a) ≈ logprob (test/policy_tests.jl:101-110,261-269)
    out = ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_logpdf(h, p(dev(ctx, s)), p(a), B, p(out)))
    assert_close(host(out), host(lp), rtol=1e-5, atol=1e-5)
    assert_close(host(out), ref.logpdf(s, host(a)).detach().numpy()[:, 0], rtol=1e-5, atol=2e-5)
    # action = μ(s); entropy: scalar, constant not scaled by the action dim (policies.jl:348)
    ctx.check(ctx.lib.crux_gaussian_action(h, p(dev(ctx, s)), B, p(a)))
    assert_close(host(a), ref.action(s).detach().numpy(), rtol=1e-5, atol=1e-5)
    ent = ctx.empty((1,))
    ctx.check(ctx.lib.crux_gaussian_entropy(h, p(dev(ctx, s)), B, p(ent)))
    assert_close(host(ent)[0], float(ref.entropy()), rtol=1e-6)
    assert_close(host(ent)[0], 1.4189385332046727 + float(ls.sum()), rtol=1e-6)
    ctx.lib.crux_gaussian_destroy(h); ctx.lib.crux_mlp_destroy(hm)


def test_gaussian_device_rng_statistics(ctx):
    # test/policy_tests.jl:277-281: empirical std of 1e5 samples == exp(logΣ); logprob consistent with the sample
    mu = o.MLP([3, 8, 2], [1, 0], np.random.default_rng(0))
    ls = np.array([-0.5, 0.25], F32)
    hm, h = _gauss(ctx, mu, ls)
    B = 100000
    s = dev(ctx, np.zeros((B, 3), F32))
    a, lp = ctx.empty((B, 2)), ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_explore(h, p(s), B, None, 1234, 0, p(a), p(lp)))
    ah = host(a)
    assert np.allclose(ah.std(0), np.exp(ls), rtol=2e-2)
    assert abs(float(((ah - ah.mean(0)) / ah.std(0)).mean())) < 0.02
    out = ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_logpdf(h, p(s), p(a), B, p(out)))
    assert_close(host(out), host(lp), rtol=1e-5, atol=1e-5)
    a2 = ctx.empty((B, 2))
    ctx.check(ctx.lib.crux_gaussian_explore(h, p(s), B, None, 1234, 1, p(a2), p(lp)))  # another counter: fresh noise
    assert not np.array_equal(host(a2), ah)
    ctx.check(ctx.lib.crux_gaussian_explore(h, p(s), B, None, 1234, 0, p(a2), p(lp)))  # same counter: same noise
    assert np.array_equal(host(a2), ah)
    # tails exist (Box-Muller): |z| > 3 at about 0.27 %
    z = (ah - ah.mean(0)) / ah.std(0)
    assert 0.0015 < (np.abs(z) > 3).mean() < 0.004
    ctx.lib.crux_gaussian_destroy(h); ctx.lib.crux_mlp_destroy(hm)


@pytest.mark.parametrize("B", [1, 257])
def test_squashed_gaussian_heads(ctx, B):
    """SquashedGaussianPolicy with [μ | logΣ] heads on a shared trunk (examples/rl/half_cheetah_mujoco.jl:37-43)."""
    rng = np.random.default_rng(B)
    net = o.MLP([11, 32, 32, 8], [2, 2, 0], rng)
    A = 4
    ref = o.SquashedGaussianPolicy(lambda s: net(s)[:, :A], lambda s: net(s)[:, A:], ascale=2.0)
    hm, h = _gauss(ctx, net, None, squashed=True, ascale=2.0)
    s = (rng.standard_normal((B, 11)) * 3).astype(F32)
    eps = (0.5 * rng.standard_normal((B, A))).astype(F32)
    a, lp = ctx.empty((B, A)), ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_explore(h, p(dev(ctx, s)), B, p(dev(ctx, eps)), 0, 0, p(a), p(lp)))
    a0, lp0 = ref.exploration(s, eps)
    assert_close(host(a), a0.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert_close(host(lp), lp0.detach().numpy()[:, 0], rtol=2e-5, atol=5e-5)
    assert np.all(np.abs(host(a)) <= 2.0)  # test/policy_tests.jl:312-313
    out = ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_logpdf(h, p(dev(ctx, s)), p(a), B, p(out)))
    assert_close(host(out), ref.logpdf(s, host(a)).detach().numpy()[:, 0], rtol=1e-4, atol=1e-3)
    ent = ctx.empty((B,))
    ctx.check(ctx.lib.crux_gaussian_entropy(h, p(dev(ctx, s)), B, p(ent)))
    assert_close(host(ent), ref.entropy(s).detach().numpy()[:, 0], rtol=1e-5, atol=1e-5)  # [B] (policies.jl:398)
    ctx.check(ctx.lib.crux_gaussian_action(h, p(dev(ctx, s)), B, p(a)))
    assert_close(host(a), ref.action(s).detach().numpy(), rtol=1e-5, atol=1e-5)
    ctx.lib.crux_gaussian_destroy(h); ctx.lib.crux_mlp_destroy(hm)


def test_squashed_clamp_uses_unclamped_log_sigma_term(ctx):
    """policies.jl:374-386 quirk: σ from clamped logΣ ∈ [-5, 2], the `- logΣ` term unclamped."""
    # a 1-layer identity net whose bias IS the output: [μ | logΣ] = [0, 0, 5, -9]
    dims, acts = [1, 4], [0]
    flat = np.array([0, 0, 0, 0, 0.0, 0.0, 5.0, -9.0], F32)
    hm = make_mlp(ctx, dims, acts, flat)
    h = C.c_void_p()
    ctx.check(ctx.lib.crux_gaussian_create(ctx.h, hm, 2, None, 1, 1.0, C.byref(h)))
    eps = np.array([[0.1, -0.2]], F32)
    a, lp = ctx.empty((1, 2)), ctx.empty((1,))
    ctx.check(ctx.lib.crux_gaussian_explore(h, p(dev(ctx, np.zeros((1, 1), F32))), 1, p(dev(ctx, eps)), 0, 0, p(a), p(lp)))
    sig = np.exp(np.array([2.0, -5.0]))
    ap = eps[0] * sig
    want = np.sum(-(ap ** 2) / (2 * sig ** 2) - 0.9189385332046727 - np.array([5.0, -9.0]) - 2 * (math.log(2) - ap - np.log1p(np.exp(-2 * ap))))
    assert_close(host(a)[0], np.tanh(ap), rtol=1e-5)
    assert_close(host(lp)[0], want, rtol=1e-5)
    ctx.lib.crux_gaussian_destroy(h); ctx.lib.crux_mlp_destroy(hm)


def test_rollout_step(ctx):
    """crux_rollout_step == exploration(actor) + value(critic) for N env streams (sampler.jl:73 + fill_gae's V(s))."""
    rng = np.random.default_rng(7)
    mu = o.MLP([17, 64, 64, 6], [1, 1, 0], rng)
    cr = o.MLP([17, 64, 64, 1], [1, 1, 0], rng)
    ls = np.full(6, -0.5, F32)
    ref = o.GaussianPolicy(mu, ls)
    hm, h = _gauss(ctx, mu, ls)
    hc = make_mlp(ctx, cr.dims, cr.acts, cr.flat())
    for N in (1, 33, 4096):
        s = rng.standard_normal((N, 17)).astype(F32)
        eps = rng.standard_normal((N, 6)).astype(F32)
        a, lp, v = ctx.empty((N, 6)), ctx.empty((N,)), ctx.empty((N,))
        ctx.check(ctx.lib.crux_rollout_step(h, hc, p(dev(ctx, s)), N, p(dev(ctx, eps)), 0, 0, p(a), p(lp), p(v)))
        a0, lp0 = ref.exploration(s, eps)
        assert_close(host(a), a0.detach().numpy(), rtol=1e-5, atol=1e-5, what=f"a N={N}")
        assert_close(host(lp), lp0.detach().numpy()[:, 0], rtol=1e-5, atol=2e-5, what=f"logp N={N}")
        assert_close(host(v), cr(s).detach().numpy()[:, 0], rtol=1e-5, atol=1e-5, what=f"V(s) N={N}")
        # without a critic / without logprob
        ctx.check(ctx.lib.crux_rollout_step(h, None, p(dev(ctx, s)), N, p(dev(ctx, eps)), 0, 0, p(a), None, None))
        assert_close(host(a), a0.detach().numpy(), rtol=1e-5, atol=1e-5)
    assert ctx.lib.crux_rollout_step(h, hc, None, 0, None, 0, 0, None, None, None) == 0
    ctx.lib.crux_gaussian_destroy(h); ctx.lib.crux_mlp_destroy(hm); ctx.lib.crux_mlp_destroy(hc)


def test_discrete_network(ctx):
    rng = np.random.default_rng(0)
    net = o.MLP([2, 8, 4], [2, 0], rng)
    ref = o.DiscreteNetwork(net, [0, 1, 2, 3])
    B, nA = 500, 4
    s = rng.standard_normal((B, 2)).astype(F32)
    q = dev(ctx, net(s).detach().numpy())
    idx = torch.empty(B, dtype=torch.int32, device=ctx.device)
    ctx.check(ctx.lib.crux_discrete_argmax(ctx.h, p(q), B, nA, p(idx)))
    assert np.array_equal(host(idx), ref.action_index(s).numpy())
    u = rng.random(B)
    lp = ctx.empty((B,))
    ctx.check(ctx.lib.crux_discrete_explore(ctx.h, p(q), B, nA, p(dev(ctx, u)), 0, 0, p(idx), p(lp)))
    ai, lp0 = ref.exploration(s, u)
    assert np.array_equal(host(idx), ai.numpy())
    assert_close(host(lp), lp0.detach().numpy()[:, 0], rtol=1e-5, atol=1e-6)
    oh = np.eye(nA, dtype=F32)[host(idx)]
    out = ctx.empty((B,))
    ctx.check(ctx.lib.crux_discrete_logpdf(ctx.h, p(q), p(dev(ctx, oh)), B, nA, p(out)))
    assert_close(host(out), ref.logpdf(s, oh).detach().numpy()[:, 0], rtol=1e-5, atol=1e-6)
    ctx.check(ctx.lib.crux_discrete_entropy(ctx.h, p(q), B, nA, p(out)))
    assert_close(host(out), ref.entropy(s).detach().numpy()[:, 0], rtol=1e-5, atol=1e-6)
    # device RNG: frequencies follow softmax(Q)
    q1 = dev(ctx, np.tile(np.array([[0.0, 1.0, 2.0, -1.0]], F32), (200000, 1)))
    idx = torch.empty(200000, dtype=torch.int32, device=ctx.device)
    ctx.check(ctx.lib.crux_discrete_explore(ctx.h, p(q1), 200000, nA, None, 7, 0, p(idx), None))
    freq = np.bincount(host(idx), minlength=4) / 200000
    pr = np.exp([0, 1, 2, -1.0]); pr /= pr.sum()
    assert np.allclose(freq, pr, atol=5e-3)
    assert ctx.lib.crux_discrete_argmax(ctx.h, p(q), B, 65, p(idx)) == 1


@pytest.mark.parametrize("alpha", [1.0, 0.3, 2.5])
def test_softq_logits_target_and_sampling(ctx, alpha):
    """rl/softq.jl:8,13-17,47-48: soft_value = α·logsumexp(Q/α); logits = softmax(Q/α) drive exploration, logpdf and entropy."""
    rng = np.random.default_rng(5)
    B, nA = 600, 5
    q = (3 * rng.standard_normal((B, nA))).astype(F32)
    r = rng.standard_normal(B).astype(F32); done = (rng.random(B) < 0.3).astype(np.uint8)
    qd = dev(ctx, q)
    y = ctx.empty((B,))
    ctx.check(ctx.lib.crux_softq_target(ctx.h, p(dev(ctx, r)), p(dev(ctx, done)), p(qd), B, nA, 0.97, alpha, p(y)))
    assert_close(host(y), o.softq_target(q, r, done, 0.97, alpha).numpy()[:, 0], rtol=1e-5, atol=1e-6, what="softq target")
    ps = o.softq_logits(q, alpha).numpy()
    u = rng.random(B)
    idx = torch.empty(B, dtype=torch.int32, device=ctx.device); lp = ctx.empty((B,))
    ctx.check(ctx.lib.crux_discrete_explore_t(ctx.h, p(qd), B, nA, alpha, p(dev(ctx, u)), 0, 0, p(idx), p(lp)))
    cdf = np.cumsum(ps.astype(np.float64), axis=1)
    want = np.minimum((cdf < u[:, None]).sum(1), nA - 1)
    got = host(idx)
    edge = np.abs(np.take_along_axis(cdf, np.minimum(got, want)[:, None], 1)[:, 0] - u) < 1e-6   # draws on a CDF step: either side
    assert np.array_equal(got[~edge], want[~edge]) and edge.sum() <= 2
    assert_close(host(lp), np.log(ps[np.arange(B), got]), rtol=1e-5, atol=1e-6, what="logprob")
    oh = np.eye(nA, dtype=F32)[rng.integers(0, nA, B)]
    out = ctx.empty((B,))
    ctx.check(ctx.lib.crux_discrete_logpdf_t(ctx.h, p(qd), p(dev(ctx, oh)), B, nA, alpha, p(out)))
    assert_close(host(out), np.log((ps * oh).sum(1)), rtol=1e-5, atol=1e-6, what="logpdf")
    ctx.check(ctx.lib.crux_discrete_entropy_t(ctx.h, p(qd), B, nA, alpha, p(out)))
    assert_close(host(out), -(ps * np.log(ps + np.finfo(F32).eps)).sum(1), rtol=1e-5, atol=1e-6, what="entropy")
    if alpha == 1.0:   # α = 1 is the default conversion, bit for bit
        out1 = ctx.empty((B,))
        ctx.check(ctx.lib.crux_discrete_entropy(ctx.h, p(qd), B, nA, p(out1)))
        assert np.array_equal(host(out), host(out1))
    assert ctx.lib.crux_softq_target(ctx.h, p(dev(ctx, r)), p(dev(ctx, done)), p(qd), B, nA, 0.97, 0.0, p(y)) != 0


def test_eps_greedy(ctx):
    # exploration(::MixedPolicy) policies.jl:474-494
    rng = np.random.default_rng(1)
    B, nA = 1000, 4
    q = rng.standard_normal((B, nA)).astype(F32)
    u = rng.random((B, 2))
    idx = torch.empty(B, dtype=torch.int32, device=ctx.device)
    oh, lp = ctx.empty((B, nA)), ctx.empty((B,))
    eps = 0.3
    ctx.check(ctx.lib.crux_discrete_eps_greedy(ctx.h, p(dev(ctx, q)), B, nA, eps, p(dev(ctx, u)), 0, 0, p(idx), p(oh), p(lp)))
    want = np.where(u[:, 0] < eps, np.minimum((u[:, 1] * nA).astype(int), nA - 1), q.argmax(1))
    assert np.array_equal(host(idx), want)
    assert np.array_equal(host(oh), np.eye(nA, dtype=F32)[want])
    assert_close(host(lp), np.full(B, o.eps_greedy_logprob(eps, nA)), rtol=1e-6)
    # eps = 0 -> greedy; eps = 1 -> uniform
    ctx.check(ctx.lib.crux_discrete_eps_greedy(ctx.h, p(dev(ctx, q)), B, nA, 0.0, None, 3, 0, p(idx), None, None))
    assert np.array_equal(host(idx), q.argmax(1))
    ctx.check(ctx.lib.crux_discrete_eps_greedy(ctx.h, p(dev(ctx, q)), B, nA, 1.0, None, 3, 0, p(idx), None, None))
    assert np.allclose(np.bincount(host(idx), minlength=nA) / B, 0.25, atol=0.06)
