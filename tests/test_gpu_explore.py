"""-m gpu: FirstExplorePolicy (policies.jl:518-534) through the Sampler -- the exploration policy of the first N steps, then the
on-policy action or another exploration policy -- for continuous and discrete agents (round-1 verdict: untested; advisor: it called
value() instead of action() and dropped MixedPolicy's one-hot / logprob), and the forward-kernel variants of value(π, s)."""
import os

import numpy as np
import pytest

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, dev, host, make_mlp, p
from test_gpu_host_api import _oracle_mlp

pytestmark = pytest.mark.gpu


def _det_actor(crux, ctx, seed, obs=17, act=6, hid=32):
    rng = np.random.default_rng(seed)
    D = crux.Dense
    return crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, act, crux.tanh, rng=rng)), ctx=ctx)


@pytest.mark.parametrize("after", [None, "noise"])
def test_first_explore_policy_continuous(crux, ctx, after):
    n, T, N_first = 8, 5, 16   # vector steps 0 and 1 (i = 0, 8 < 16) use the initial policy, steps 2.. the on-policy actor
    A, B = _det_actor(crux, ctx, 1), _det_actor(crux, ctx, 2)
    after_pol = None if after is None else crux.GaussianNoiseExplorationPolicy(F32(0.0), a_min=-0.25, a_max=0.25)
    pe = crux.FirstExplorePolicy(N_first, B, after_pol)
    agent = crux.PolicyParams(A, space=crux.ContinuousSpace(6), pi_explore=pe)
    s = crux.Sampler(crux.HostLinQuad(n, seed=4), agent, max_steps=100, required_columns=["logprob"])
    d = s.steps_(None, Nsteps=n * T, explore=True, i=0)
    S_, a = host(d["s"]).reshape(T, n, 17), host(d["a"]).reshape(T, n, 6)
    mu_a, mu_b = _oracle_mlp(A.mlp), _oracle_mlp(B.mlp)
    for t in range(T):
        want = (mu_b if t * n < N_first else mu_a)(S_[t]).detach().numpy()
        if t * n >= N_first and after is not None:
            want = np.clip(want, -0.25, 0.25)      # clamp(π(s) + 0·ε, a_min, a_max)  policies.jl:510-514
        assert_close(a[t], want, rtol=1e-5, atol=1e-5, what=f"vector step {t}")
    assert np.isnan(host(d["logprob"])).all()      # every branch returns NaN as the log-probability


def test_first_explore_policy_discrete_with_eps_greedy_after(crux, ctx):
    n, T = 16, 4
    rng = np.random.default_rng(0)
    D = crux.Dense
    q_on = crux.DiscreteNetwork(crux.Chain(D(2, 8, crux.relu, rng=rng), D(8, 4, rng=rng)), [0, 1, 2, 3], ctx=ctx)
    q_init = crux.DiscreteNetwork(crux.Chain(D(2, 8, crux.relu, rng=rng), D(8, 4, rng=rng)), [0, 1, 2, 3], ctx=ctx)
    pe = crux.FirstExplorePolicy(n, q_init, crux.eps_greedy_policy(0.0, [0, 1, 2, 3]))   # ε = 0: greedy on the on-policy network
    agent = crux.PolicyParams(q_on, pi_explore=pe)
    s = crux.Sampler(crux.SimpleGridWorld(n, seed=1), agent, S=crux.ContinuousSpace(2), max_steps=100, required_columns=["logprob"])
    d = s.steps_(None, Nsteps=n * T, explore=True, i=0)
    S_, a, lp = host(d["s"]).reshape(T, n, 2), host(d["a"]).reshape(T, n, 4), host(d["logprob"]).reshape(T, n)
    assert np.all(a.sum(-1) == 1) and set(np.unique(a)) <= {0.0, 1.0}
    for t in range(T):
        q = _oracle_mlp((q_init if t * n < n else q_on).mlp)(S_[t]).detach().numpy()
        assert np.array_equal(a[t].argmax(-1), q.argmax(-1)), f"vector step {t}"
    assert np.isnan(lp[0]).all()                                     # action(π.initial_policy, s), NaN
    assert_close(lp[1:], np.zeros((T - 1, n)), rtol=0, atol=1e-6)      # ϵ-greedy with ϵ = 0: log(1) for the greedy action (policies.jl:487-494)


@pytest.mark.parametrize("variant", ["1", "1s", "0"])
def test_value_forward_kernel_variants(ctx, crux, variant):
    """value(π, s) over a whole column (>= 148 tiles of 128 rows): the tcgen05 kernel with the activations in tensor memory (default "1"),
    the tcgen05 kernel with shared-memory activations ("1s") and the mma.sync kernel ("0") against the oracle."""
    rng = np.random.default_rng(5)
    ref = o.MLP([17, 64, 64, 1], [1, 1, 0], rng)
    h = make_mlp(ctx, ref.dims, ref.acts, ref.flat())
    B = 20000 + 37   # ragged last tile
    x = rng.standard_normal((B, 17)).astype(F32)
    y = ctx.empty((B, 1))
    os.environ["CRUX_FWD_TC5"] = variant
    try:
        ctx.check(ctx.lib.crux_mlp_forward(h, p(dev(ctx, x)), B, p(y)))
        got = host(y)
    finally:
        os.environ.pop("CRUX_FWD_TC5", None)
    assert_close(got, ref(x).detach().numpy(), rtol=1e-5, atol=2e-6, what=f"CRUX_FWD_TC5={variant}")
