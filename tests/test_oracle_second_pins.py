"""CPU (not gpu): INDEPENDENT second pins for the pieces of the oracle the reference's own tests leave unpinned (GAE values,
ppo_loss / td targets, Adam steps, whitening -- test/gym/sampler_tests.jl:75-81 asserts nothing).  Each oracle restatement is compared with
an implementation that shares no code with it:

    oracle (line-by-line restatement of)                     independent pin
    ------------------------------------------------------   ---------------------------------------------------------------------
    o.Adam            Flux 0.14 Optimise.Adam [3P]            torch.optim.Adam(eps=1e-8) -- same update rule, third-party code
      (call site src/training.jl:21)
    o.fill_gae / o.gae_returns_TN / oracle/gae_oracle.c       discounted cumulative sum of the TD residuals with scipy.signal.lfilter
      src/sampler.jl:262-273                                   (the closed form A_t = sum_k (γλ)^k δ_{t+k} inside an episode range)
    o.fill_returns    src/sampler.jl:275-281                  lfilter on the rewards
    o.whiten          src/utils.jl:41-42                      float64 numpy with ddof = 1
    o.ppo_loss        src/model_free/rl/ppo.jl:4-21           float64 numpy restatement of the published PPO-clip objective
    o.dqn_target      src/model_free/rl/dqn.jl:4-6            float64 numpy
    MLP + autograd    Flux Dense / Zygote                     finite differences of the float64 forward pass
"""
import math

import numpy as np
import pytest
import torch
from scipy.signal import lfilter

from oracle import c_oracle
from oracle import crux_oracle as o

F32 = np.float32


def test_adam_matches_torch_optim_adam_over_50_steps():
    rng = np.random.default_rng(0)
    shapes = [(17, 64), (64,), (64, 6), (6,)]
    ours = [torch.tensor(rng.standard_normal(s).astype(F32), requires_grad=True) for s in shapes]
    theirs = [torch.tensor(p.detach().numpy().copy(), requires_grad=True) for p in ours]
    opt_o = o.Adam(F32(3e-4))
    opt_t = torch.optim.Adam(theirs, lr=float(F32(3e-4)), betas=(0.9, 0.999), eps=1e-8)
    for step in range(50):
        scale = 10.0 ** rng.uniform(-4, 1)   # gradient scales from 1e-4 to 10: Adam must be invariant up to eps
        for a, b in zip(ours, theirs):
            g = (rng.standard_normal(a.shape) * scale).astype(F32)
            a.grad = torch.tensor(g); b.grad = torch.tensor(g.copy())
        opt_o.step(ours)
        opt_t.step()
        for a, b in zip(ours, theirs):
            np.testing.assert_allclose(a.detach().numpy(), b.detach().numpy(), rtol=2e-6, atol=2e-7, err_msg=f"step {step}")


def _gae_lfilter(r, done, ee, vs, vsp, gamma, lam):
    """Per stream, per episode range (cut only at episode_end): A = discounted cumsum of δ with factor γλ, R = discounted cumsum of r."""
    T, N = r.shape
    adv, ret = np.zeros((T, N)), np.zeros((T, N))
    delta = r + gamma * (1.0 - done) * vsp - vs          # done only masks the bootstrap
    for e in range(N):
        start = 0
        for t in range(T):
            if ee[t, e] or t == T - 1:
                sl = slice(start, t + 1)
                adv[sl, e] = lfilter([1.0], [1.0, -gamma * lam], delta[sl, e][::-1])[::-1]
                ret[sl, e] = lfilter([1.0], [1.0, -gamma], r[sl, e][::-1])[::-1]
                start = t + 1
    return adv, ret


@pytest.mark.parametrize("T,N,seed", [(37, 5, 0), (128, 16, 1), (9, 33, 2)])
def test_gae_and_returns_match_a_discounted_cumsum(T, N, seed):
    rng = np.random.default_rng(seed)
    r, vs, vsp = (rng.standard_normal((T, N)).astype(F32) for _ in range(3))
    done = rng.random((T, N)) < 0.08
    ee = done | (rng.random((T, N)) < 0.05)
    ee[-1] = True
    gamma, lam = F32(0.99), F32(0.95)
    want_a, want_r = _gae_lfilter(r.astype(np.float64), done.astype(np.float64), ee, vs.astype(np.float64), vsp.astype(np.float64), float(gamma), float(lam))
    a_py, r_py = o.gae_returns_TN(r, done, ee, vs, vsp, gamma, lam)
    a_c, r_c = c_oracle.gae_returns(r, done, ee, vs, vsp, gamma, lam)
    for got, want, what in ((a_py, want_a, "advantage (python oracle)"), (r_py, want_r, "return (python oracle)"),
                            (a_c, want_a, "advantage (C oracle)"), (r_c, want_r, "return (C oracle)")):
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-5, err_msg=what)
    # the single-range form the reference calls from terminate_episode! (sampler.jl:56-57)
    ends = np.flatnonzero(ee[:, 0])
    a0 = np.zeros(T, F32)
    o.fill_gae(r[:, 0], done[:, 0], vs[:, 0], vsp[:, 0], lam, gamma, rng=range(0, ends[0] + 1), out=a0)
    np.testing.assert_allclose(a0[:ends[0] + 1], want_a[:ends[0] + 1, 0], rtol=2e-5, atol=2e-5)


def test_whiten_and_dqn_target_against_float64_numpy():
    rng = np.random.default_rng(3)
    v = (rng.standard_normal(5000) * 3 + 1).astype(F32)
    np.testing.assert_allclose(o.whiten(v), (v.astype(np.float64) - v.astype(np.float64).mean()) / v.astype(np.float64).std(ddof=1), rtol=1e-5, atol=1e-6)
    q, r = rng.standard_normal((300, 4)).astype(F32), rng.standard_normal(300).astype(F32)
    done = rng.random(300) < 0.2
    want = r.astype(np.float64) + 0.99 * (1.0 - done) * q.astype(np.float64).max(1)
    np.testing.assert_allclose(o.dqn_target(q, r, done, F32(0.99)).numpy()[:, 0], want, rtol=1e-5, atol=1e-6)


def _mlp64(mlp, x):
    h = x.astype(np.float64)
    for l, act in enumerate(mlp.acts):
        h = h @ mlp.W[l].detach().numpy().astype(np.float64).T + mlp.b[l].detach().numpy().astype(np.float64)
        h = np.tanh(h) if act == o.ACT_TANH else (np.maximum(h, 0) if act == o.ACT_RELU else h)
    return h


def test_ppo_loss_value_and_gradient_against_float64_numpy():
    rng = np.random.default_rng(4)
    n, eps = 400, 0.2
    mu = o.MLP([17, 64, 64, 6], [o.ACT_TANH, o.ACT_TANH, o.ACT_IDENTITY], rng)
    ls = np.full(6, -0.5, F32)
    pi = o.GaussianPolicy(mu, ls)
    D = {"s": rng.standard_normal((n, 17)).astype(F32), "a": rng.standard_normal((n, 6)).astype(F32) * F32(0.7),
         "advantage": rng.standard_normal(n).astype(F32), "return": rng.standard_normal(n).astype(F32)}
    with torch.no_grad():
        D["logprob"] = (pi.logpdf(D["s"], D["a"]).numpy()[:, 0] + rng.standard_normal(n).astype(F32) * F32(0.1)).astype(F32)

    def loss64(w3_00=None):
        m = _mlp64(mu, D["s"])
        if w3_00 is not None:   # perturb W3[0, 0] for the finite-difference check
            m[:, 0] += (w3_00 - float(mu.W[2][0, 0])) * _mlp64(o.MLP(mu.dims[:3], mu.acts[:2], Ws=[w.detach().numpy() for w in mu.W[:2]],
                                                                   bs=[b.detach().numpy() for b in mu.b[:2]]), D["s"])[:, 0]
        var = np.exp(ls.astype(np.float64)) ** 2
        logp = (-(D["a"] - m) ** 2 / (2 * var) - 0.9189385332046727 - ls.astype(np.float64)).sum(1)
        ratio = np.exp(logp - D["logprob"])
        A = D["advantage"].astype(np.float64)
        p_loss = -np.mean(np.minimum(ratio * A, np.clip(ratio, 1 - eps, 1 + eps) * A))
        e_loss = -(1.4189385332046727 + ls.astype(np.float64).sum())
        return 1.0 * p_loss + 0.1 * e_loss, np.mean(D["logprob"] - logp)

    info = {}
    l = o.ppo_loss(pi, {"eps": F32(eps), "lp": F32(1.0), "le": F32(0.1)}, D, info)
    want, want_kl = loss64()
    assert abs(float(l) - want) < 2e-6 * max(1.0, abs(want))
    assert abs(info["kl"] - want_kl) < 2e-6
    for q in pi.params():
        q.grad = None
    l.backward()
    h = 1e-4
    w0 = float(mu.W[2][0, 0])
    fd = (loss64(w0 + h)[0] - loss64(w0 - h)[0]) / (2 * h)
    assert abs(float(mu.W[2].grad[0, 0]) - fd) < 1e-5 * max(1.0, abs(fd)) + 1e-7
