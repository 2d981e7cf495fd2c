"""-m gpu: the host-side mirror of the reference surface (ExperienceBuffer, Sampler.steps!, PPO/A2C/DQN/SAC + solve)
driving libcrux_cuda.so, against the reference's own test expectations and the CPU oracle."""
import math

import numpy as np
import pytest
import torch

from oracle import crux_oracle as o
from oracle.ppo_cpu import OraclePPO
from gpu_util import F32, assert_close, assert_params_close, host

pytestmark = pytest.mark.gpu


# ------------------------------------------------------------------------------------------------ ExperienceBuffer
def test_experience_buffer_reference_kats(crux, ctx):
    # test/experience_buffer_tests.jl:32-51,66-147,162-174
    b = crux.ExperienceBuffer(crux.ContinuousSpace(2), crux.ContinuousSpace(1), 100, ctx=ctx)
    d = {"s": 2 * np.ones((50, 2)), "a": np.ones((50, 1)), "sp": np.ones((50, 2)), "r": np.ones((50, 1)), "done": np.zeros((50, 1)),
         "weight": np.zeros((50, 1))}  # extra source keys are ignored (keys(b) drives the loop)
    I = b.push_(d)
    assert I.tolist() == list(range(1, 51)) and len(b) == 50 and b.next_ind == 51
    assert b.get_last_N_indices(10).tolist() == list(range(41, 51)) and b.get_last_N_indices(1000).tolist() == list(range(1, 51))
    b.push_(d); b.push_(d)
    assert b.get_last_N_indices(51).tolist() == [100] + list(range(1, 51))
    assert len(b) == 100 and b.total_count == 150 and b.capacity == 100
    b = crux.ExperienceBuffer(crux.ContinuousSpace(2), crux.DiscreteSpace(4), 100, ctx=ctx)
    b.push_({"s": 2 * np.ones((1, 2)), "a": np.ones((1, 4)), "sp": np.ones((1, 2)), "r": np.ones((1, 1)), "done": np.zeros((1, 1))})
    assert len(b) == 1 and tuple(b["s"].shape) == (1, 2) and tuple(b["a"].shape) == (1, 4) and tuple(b["r"].shape) == (1, 1)
    assert torch.all(b["s"] == 2) and torch.all(b["a"] == 1) and torch.all(b["done"] == 0)
    d3 = {"s": 3 * np.ones((3, 2)), "a": (np.random.default_rng(0).random((3, 4)) < 0.5), "sp": 5 * np.ones((3, 2)), "r": 6 * np.ones((3, 1)),
          "done": np.ones((3, 1))}
    b.push_(d3)
    assert len(b) == 4 and np.array_equal(host(b["a"])[1:], d3["a"].astype(F32)) and torch.all(b["sp"][1:] == 5)
    b.push_(b)  # push!(b, b)
    assert len(b) == 8
    for k in b.keys():
        assert torch.equal(b[k][:4], b[k][4:8])
    mb = b.minibatch([1, 2, 4])
    for k in mb:
        assert torch.equal(mb[k], b[k][[0, 1, 3]])
    # b[:key] is a view callers write through (ppo.jl:61)
    b["r"][:] = 7.0
    assert torch.all(b.column("r")[:8] == 7.0)
    t = crux.buffer_like(b, capacity=3)
    crux.uniform_sample_(t, b, ids=[8, 1, 3])
    assert t.indices.tolist() == [8, 1, 3] and torch.equal(t["s"], b["s"][[7, 0, 2]])
    b.clear_()
    assert len(b) == 0 and b.next_ind == 1
    with pytest.raises(AssertionError):
        b.push_({"s": np.ones((2, 3))})  # row shape mismatch (experience_buffer.jl:251 @assert)


def test_prioritized_buffer_and_multi_source(crux, ctx):
    # test/experience_buffer_tests.jl:193-205,225-262
    bp = crux.ExperienceBuffer(crux.ContinuousSpace(2), crux.DiscreteSpace(4), 50, prioritized=True, ctx=ctx)
    assert "weight" in bp and bp.isprioritized()
    bp.update_priorities_([1, 2, 3], [1.0, 2.0, 3.0])
    assert bp.max_priority == 3.0
    assert_close(host(bp.priorities())[:3], np.array([1, 2, 3], F32) ** F32(0.6), rtol=1e-6)
    d = {"s": 3 * np.ones((3, 2)), "a": np.ones((3, 4)), "sp": 5 * np.ones((3, 2)), "r": 6 * np.ones((3, 1)), "done": np.ones((3, 1))}
    bp.push_(d); bp.push_(d)
    assert bp.max_priority == 3.0
    assert_close(host(bp.priorities())[:6], np.full(6, F32(3) ** F32(0.6)), rtol=1e-6)
    bp["s"][:] = torch.rand((6, 2), device=ctx.device)
    bp.update_priorities_(list(range(1, 7)), np.arange(1, 7, dtype=F32))
    t = crux.ExperienceBuffer(crux.ContinuousSpace(2), crux.DiscreteSpace(4), 1000, ["weight"], ctx=ctx)
    crux.rand_(t, bp)
    pr = host(bp.priorities())[:6]
    ids = t.indices
    freqs = np.array([(ids == i).sum() for i in range(1, 7)]) / 1000
    assert np.all(np.abs(freqs - pr / pr.sum()) / (pr / pr.sum()) < 0.01)
    assert torch.equal(t["s"], bp["s"][torch.as_tensor(ids - 1, device=ctx.device)])
    assert torch.all(t["weight"] <= 1.0 + 1e-6)
    srcs = []
    for v in (1.0, 2.0, 3.0):
        s = crux.ExperienceBuffer(crux.ContinuousSpace(2), crux.DiscreteSpace(4), 10, ctx=ctx)
        s.push_({"s": v * np.ones((1, 2)), "a": np.ones((1, 4)), "sp": np.ones((1, 2)), "r": np.ones((1, 1)), "done": np.zeros((1, 1))})
        srcs.append(s)
    t = crux.ExperienceBuffer(crux.ContinuousSpace(2), crux.DiscreteSpace(4), 10, ctx=ctx)
    assert crux.rand_(t, *srcs) == [4, 3, 3]
    assert torch.all(t["s"][:4] == 1) and torch.all(t["s"][4:7] == 2) and torch.all(t["s"][7:] == 3)
    assert t.indices.tolist() == [1, 1, 1]  # only the LAST source's ids survive


# ------------------------------------------------------------------------------------------------ Sampler
def _actor_critic(crux, ctx, seed=0, obs=17, act=6, hid=64):
    rng = np.random.default_rng(seed)
    D = crux.Dense
    mu = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.tanh, rng=rng), D(hid, hid, crux.tanh, rng=rng), D(hid, act, rng=rng)), ctx=ctx)
    cr = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.tanh, rng=rng), D(hid, hid, crux.tanh, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
    return crux.ActorCritic(crux.GaussianPolicy(mu, np.full(act, -0.5, F32)), cr)


def _oracle_mlp(mlp):
    flat, Ws, bs, off = mlp.get_flat(), [], [], 0
    for l in range(len(mlp.acts)):
        i, oo = mlp.dims[l], mlp.dims[l + 1]
        Ws.append(flat[off:off + i * oo].reshape(i, oo).T.copy()); off += i * oo
        bs.append(flat[off:off + oo].copy()); off += oo
    return o.MLP(mlp.dims, mlp.acts, Ws=Ws, bs=bs)


@pytest.mark.parametrize("device_env", [False, True])
def test_sampler_steps_invariants_and_gae(crux, ctx, device_env):
    """steps!/episode boundaries (test/gym/sampler_tests.jl:32-56) + GAE/returns of every stream against the oracle."""
    n, T, max_steps = 16, 40, 7
    pi = _actor_critic(crux, ctx)
    env = crux.DeviceLinQuad(n, seed=5, max_steps=max_steps, ctx=ctx) if device_env else crux.HostLinQuad(n, seed=5)
    cols = ["return", "logprob", "advantage", "t"] if not device_env else ["return", "logprob", "advantage"]
    s = crux.Sampler(env, pi, max_steps=max_steps, required_columns=cols, lam=0.95)
    buf = crux.ExperienceBuffer(crux.ContinuousSpace(17), crux.ContinuousSpace(6), n * T, cols, ctx=ctx)
    data = s.steps_(buf, Nsteps=n * T, explore=True, i=0, reset=True)
    assert len(buf) == n * T and buf.next_ind == 1
    D = {k: host(v).reshape((T, n) + tuple(v.shape[1:])) for k, v in data.items()}
    ee, done = D["episode_end"][..., 0].astype(bool), D["done"][..., 0].astype(bool)
    assert ee[-1].all()                                   # steps!(reset=true) closes the open episode (sampler.jl:148)
    for e in range(n):
        ends = np.flatnonzero(ee[:, e])
        lens = np.diff(np.concatenate([[-1], ends]))
        assert lens.max() <= max_steps
        for k, end in enumerate(ends[:-1]):
            assert done[end, e] or lens[k] == max_steps  # test/gym/sampler_tests.jl:55
        # inside an episode consecutive rows chain: s[t+1] == sp[t]; after an end the next row is a fresh initial state
        for t in range(T - 1):
            if ee[t, e]:
                assert np.all(np.abs(D["s"][t + 1, e]) <= 0.1)
            else:
                assert np.array_equal(D["s"][t + 1, e], D["sp"][t, e])
        if not device_env:
            tcol = D["t"][:, e, 0]
            assert tcol[0] == 1 and all(tcol[t + 1] == (1 if ee[t, e] else tcol[t] + 1) for t in range(T - 1))
    # logprob of the stored action under the policy (test/policy_tests.jl:101-110)
    lp = crux.logpdf(pi, data["s"], data["a"])
    assert_close(host(lp)[:, 0], host(data["logprob"])[:, 0], rtol=1e-5, atol=2e-5)
    # GAE / returns: the reference's per-episode recurrences on every stream with the oracle critic
    V = _oracle_mlp(pi.C.mlp)
    vs = V(D["s"].reshape(T * n, 17)).detach().numpy().reshape(T, n)
    vsp = V(D["sp"].reshape(T * n, 17)).detach().numpy().reshape(T, n)
    a0, r0 = o.gae_returns_TN(D["r"][..., 0], done, ee, vs, vsp, F32(0.99), F32(0.95))
    assert_close(D["advantage"][..., 0], a0, rtol=1e-4, atol=2e-4, what="advantage")
    assert_close(D["return"][..., 0], r0, rtol=1e-5, atol=2e-5, what="return")
    assert buf.episodes()[-1][1] == n * T


def test_sampler_noise_injection_and_second_call_continues(crux, ctx):
    n, T = 8, 5
    pi = _actor_critic(crux, ctx, seed=3)
    env = crux.HostLinQuad(n, seed=1)
    s = crux.Sampler(env, pi, max_steps=100, required_columns=["logprob"])
    eps = [np.random.default_rng(t).standard_normal((n, 6)).astype(F32) for t in range(T)]
    d = s.steps_(None, Nsteps=n * T, explore=True, noise=eps)
    mu = _oracle_mlp(pi.A.mu.mlp)
    ref = o.GaussianPolicy(mu, np.full(6, -0.5, F32))
    S_ = host(d["s"]).reshape(T, n, 17)
    for t in range(T):
        a0, lp0 = ref.exploration(S_[t], eps[t])
        assert_close(host(d["a"]).reshape(T, n, 6)[t], a0.detach().numpy(), rtol=1e-5, atol=1e-5)
    last_sp = host(d["sp"]).reshape(T, n, 17)[-1]
    d2 = s.steps_(None, Nsteps=n, explore=False)  # no reset in between: the streams continue where they stopped
    assert np.array_equal(host(d2["s"]), last_sp)
    assert_close(host(d2["a"]), mu(last_sp).detach().numpy(), rtol=1e-5, atol=1e-5)  # action(π, s) = μ(s)


# ------------------------------------------------------------------------------------------------ PPO end to end
def test_ppo_two_iterations_match_oracle(crux, ctx):
    """The whole on-policy path (steps! -> GAE -> whiten -> batch_train! actor -> batch_train! critic) for two solve
    iterations on identical inputs (same env stream, injected exploration noise and shuffles): final parameters against
    the CPU restatement (cf. test/gym/solver_tests.jl:20-52, which asserts 1e-2 between the reference's CPU and GPU)."""
    n, T, epochs, batch = 32, 16, 2, 128
    ref = OraclePPO(n, T, seed=1, epochs=epochs, batch=batch, le=0.1, env_seed=7, max_steps=6)
    rng = np.random.default_rng(1)
    D_ = crux.Dense

    def net(m):
        return crux.ContinuousNetwork(crux.Chain(*[D_(m.dims[l], m.dims[l + 1], m.acts[l], m.W[l].detach().numpy(), m.b[l].detach().numpy())
                                                   for l in range(3)]), ctx=ctx)
    pi = crux.ActorCritic(crux.GaussianPolicy(net(ref.mu), np.full(6, -0.5, F32)), net(ref.critic))
    opt = dict(epochs=epochs, batch_size=batch, optimizer=crux.Adam(F32(3e-4)))
    S = crux.PPO(pi, crux.ContinuousSpace(17), eps=0.2, lp=1.0, le=0.1, target_kl=math.inf, a_opt=dict(opt), c_opt=dict(opt),
                 N=n * T, dN=n * T, max_steps=6, lam_gae=0.95)
    env = crux.HostLinQuad(n, seed=7)
    s = crux.Sampler(env, S.agent, max_steps=6, required_columns=S.required_columns, lam=0.95)
    buf = crux.ExperienceBuffer(crux.ContinuousSpace(17), crux.ContinuousSpace(6), n * T, S.required_columns, ctx=ctx)
    noise_rng = np.random.default_rng(99)
    for it in range(2):
        eps = [noise_rng.standard_normal((n, 6)).astype(F32) for _ in range(T)]
        oa = np.stack([noise_rng.permutation(n * T) for _ in range(epochs)])
        oc = np.stack([noise_rng.permutation(n * T) for _ in range(epochs)])
        Dref = ref.rollout(eps)
        ref.update(Dref, (oa, oc))
        data = s.steps_(buf, Nsteps=n * T, explore=True, i=it * n * T, reset=True, noise=eps)
        assert_close(host(data["s"]), Dref["s"], rtol=1e-4, atol=1e-5, what=f"s it{it}")
        assert_close(host(data["r"])[:, 0], Dref["r"], rtol=1e-4, atol=1e-5, what=f"r it{it}")
        assert np.array_equal(host(data["episode_end"])[:, 0].astype(bool), Dref["episode_end"])
        assert_close(host(data["advantage"])[:, 0], Dref["advantage"], rtol=1e-3, atol=1e-3, what=f"advantage it{it}")
        S.post_batch_callback(buf)
        assert_close(host(buf["advantage"])[:, 0], o.whiten(Dref["advantage"]), rtol=1e-3, atol=1e-3, what="whitened advantage")
        S.policy_gradient_training(buf, (oa, oc))
        info = S.training_info()
        assert info["actor_batches_trained"] == epochs * (n * T // batch)
        assert_close(info["actor_loss"], ref.last["actor"][-1]["actor_loss"], rtol=2e-3, atol=2e-4, what="actor_loss")
        assert_close(info["kl"], ref.last["actor"][-1]["kl"], rtol=2e-2, atol=2e-5, what="kl")
        assert_close(info["critic_loss"], ref.last["critic"][-1]["critic_loss"], rtol=2e-3, what="critic_loss")
    steps = 2 * epochs * (n * T // batch)
    assert_params_close(pi.A.mu.mlp.get_flat(), ref.mu.flat(), 3e-4, steps, what="actor", rtol=1e-4, atol=1e-5, frac=5e-3)
    assert_params_close(pi.C.mlp.get_flat(), ref.critic.flat(), 3e-4, steps, what="critic", rtol=1e-4, atol=1e-5, frac=5e-3)
    assert_close(host(pi.A.log_sigma), ref.pi.log_sigma.detach().numpy(), rtol=1e-4, atol=1e-5, what="logΣ")


@pytest.mark.parametrize("device_env", [False, True])
def test_solve_ppo_and_a2c_run(crux, ctx, device_env, tmp_path):
    n, T = 64, 8
    for ctor in (crux.PPO, crux.A2C):
        pi = _actor_critic(crux, ctx, seed=4)
        before = pi.A.mu.mlp.get_flat().copy()
        opt = dict(epochs=2, batch_size=128)
        # the reference's default logger: TBLogger(dir, tb_increment) + fns = [log_undiscounted_return(10), log_episode_averages([:r], period)]
        log = crux.LoggerParams(dir=str(tmp_path / "log"), period=n * T, verbose=False)
        S = ctor(pi, crux.ContinuousSpace(17), a_opt=dict(opt), c_opt=dict(opt), N=3 * n * T, dN=n * T, max_steps=50, log=log)
        env = crux.DeviceLinQuad(n, seed=1, max_steps=50, ctx=ctx) if device_env else crux.HostLinQuad(n, seed=1)
        out = crux.solve(S, env)
        assert out is pi and S.i == 3 * n * T
        info = S.training_info()
        assert np.isfinite(info["actor_loss"]) and np.isfinite(info["critic_loss"]) and info["actor_batches_trained"] >= 1
        assert not np.array_equal(before, pi.A.mu.mlp.get_flat())
        assert len(log.history) >= 3 and "actor_loss" in log.history[-1]
        # logging.jl:21,48-54: every logged value went through log_value into a TensorBoard event file that reads back CRC-clean
        rec = crux.read_scalars(log.logger.path)
        tags = {t for _, t, _ in rec}
        assert {"undiscounted_return", "avg_r", "actor_loss", "critic_loss", "kl"} <= tags, tags
        last = log.history[-1]
        back = {t: v for st, t, v in rec if st == last["step"]}
        for k in ("undiscounted_return", "avg_r", "actor_loss"):
            assert back[k] == pytest.approx(np.float32(last[k]), rel=1e-6, nan_ok=True), k
        assert sorted({st for st, _, _ in rec}) == [h["step"] for h in log.history]
        crux.solve(S, env)  # calling solve again continues (on_policy.jl:38,107)
        assert S.i == 6 * n * T


def test_solve_reinforce_runs_without_a_critic(crux, ctx):
    """rl/reinforce.jl:27-39: a bare GaussianPolicy, columns return + logprob only (no value passes, no advantage)."""
    n, T = 64, 8
    pi = _actor_critic(crux, ctx, seed=6).A
    before = pi.mu.mlp.get_flat().copy()
    S = crux.REINFORCE(pi, crux.ContinuousSpace(17), a_opt=dict(epochs=2, batch_size=128), N=2 * n * T, dN=n * T, max_steps=50)
    out = crux.solve(S, crux.DeviceLinQuad(n, seed=2, max_steps=50, ctx=ctx))
    assert out is pi and S.i == 2 * n * T and "advantage" not in S.buffer.schema
    info = S.training_info()
    assert np.isfinite(info["actor_loss"]) and info["actor_batches_trained"] >= 1 and "critic_loss" not in info
    assert not np.array_equal(before, pi.mu.mlp.get_flat())


def test_solve_lagrange_ppo_runs(crux, ctx):
    """rl/ppo.jl:133-214 through solve(): :cost from the env's info, cost GAE / returns with Vc (sampler.jl:64-66), the PID penalty
    and the cost critic."""
    n, T = 32, 8
    pi = _actor_critic(crux, ctx, seed=8)
    rng = np.random.default_rng(9)
    D = crux.Dense
    Vc = crux.ContinuousNetwork(crux.Chain(D(17, 64, crux.tanh, rng=rng), D(64, 64, crux.tanh, rng=rng), D(64, 1, rng=rng)), ctx=ctx)
    before_vc = Vc.mlp.get_flat().copy()
    opt = dict(epochs=2, batch_size=128)
    S = crux.LagrangePPO(pi, Vc, crux.ContinuousSpace(17), a_opt=dict(opt), c_opt=dict(opt), cost_opt=dict(opt), N=3 * n * T, dN=n * T,
                         max_steps=50, target_cost=0.0, Ki=0.5)
    env = crux.HostLinQuad(n, seed=1, cost_threshold=0.05)
    assert crux.solve(S, env) is pi and S.i == 3 * n * T
    info = S.training_info()
    for k in ("actor_loss", "critic_loss", "cost_critic_loss", "penalty", "cur_cost", "integral term", "p_loss", "cost_loss"):
        assert np.isfinite(info[k]), k
    assert info["cur_cost"] > 0 and info["penalty"] > 0          # costs occur, the integral term has wound up
    cost = host(S.buffer["cost"])[:, 0]
    assert set(np.unique(cost)) <= {0.0, 1.0} and cost.sum() > 0
    # cost_return is the discounted cost-to-go inside each episode range: check the recurrence on the stored rollout
    cr, ee = host(S.buffer["cost_return"])[:, 0].reshape(T, n), host(S.buffer["episode_end"])[:, 0].reshape(T, n).astype(bool)
    c2 = cost.reshape(T, n)
    want = np.zeros((T, n), F32); acc = np.zeros(n, F32)
    for t in range(T - 1, -1, -1):
        acc = np.where(ee[t], F32(0), acc)
        acc = (c2[t] + F32(0.99) * acc).astype(F32)
        want[t] = acc
    assert_close(cr, want, rtol=1e-5, atol=1e-6, what="cost_return")
    assert np.abs(host(S.buffer["cost_advantage"])).sum() > 0
    assert not np.array_equal(before_vc, Vc.mlp.get_flat())


# ------------------------------------------------------------------------------------------------ DQN / SAC
def test_dqn_value_training_matches_oracle(crux, ctx):
    """value_training (off_policy.jl:66-111) for DQN with injected sample ids: dqn_target -> td_loss train! x epochs -> polyak."""
    rng = np.random.default_rng(0)
    chain = crux.Chain(crux.Dense(2, 8, crux.relu, rng=rng), crux.Dense(8, 4, rng=rng))
    pi = crux.DiscreteNetwork(chain, [0, 1, 2, 3], ctx=ctx)
    S = crux.DQN(pi, crux.ContinuousSpace(2), N=1000, dN=4, c_opt=dict(batch_size=32), buffer_size=200)
    nb = 150
    d = {"s": rng.standard_normal((nb, 2)).astype(F32), "a": np.eye(4, dtype=F32)[rng.integers(0, 4, nb)], "sp": rng.standard_normal((nb, 2)).astype(F32),
         "r": rng.standard_normal((nb, 1)).astype(F32), "done": (rng.random((nb, 1)) < 0.2)}
    S.buffer.push_(d)
    q = _oracle_mlp(pi.mlp); qt = _oracle_mlp(S.agent.pi_target.mlp)
    refnet = o.DiscreteNetwork(q, range(4))
    opt = o.Adam(F32(3e-4))
    draws = [rng.integers(1, nb + 1, 32) for _ in range(4)]
    for ids in draws:
        mb = {k: v[ids - 1] for k, v in d.items()}
        y = o.dqn_target(qt(mb["sp"]).detach(), mb["r"], mb["done"], F32(0.95))
        o.train_step(q.params(), lambda inf: o.td_loss(refnet.value(mb["s"], mb["a"]), y, None, inf), opt, {})
    o.polyak_average(qt.params(), q.params(), 0.005)
    Dmb = crux.buffer_like(S.buffer, capacity=32)
    S.value_training(Dmb, F32(0.95), draws=draws)
    assert_params_close(pi.mlp.get_flat(), q.flat(), 3e-4, 4, what="online Q")
    assert_params_close(S.agent.pi_target.mlp.get_flat(), qt.flat(), 3e-4, 4, what="target Q")


def test_softq_value_training_matches_oracle(crux, ctx):
    """rl/softq.jl:36-58 through value_training (off_policy.jl:66-111): soft target -> td_loss train! x epochs -> polyak; the
    policy itself turns always-stochastic with softmax(Q/α) logits."""
    rng = np.random.default_rng(1)
    chain = crux.Chain(crux.Dense(2, 8, crux.relu, rng=rng), crux.Dense(8, 4, rng=rng))
    pi = crux.DiscreteNetwork(chain, [0, 1, 2, 3], ctx=ctx)
    alpha = F32(0.5)
    S = crux.SoftQ(pi, crux.ContinuousSpace(2), N=1000, dN=4, c_opt=dict(batch_size=32, epochs=3), alpha=alpha, buffer_size=200)
    assert pi.always_stochastic and pi.temperature == 0.5 and S.agent.pi_target.temperature == 0.5
    nb = 150
    d = {"s": rng.standard_normal((nb, 2)).astype(F32), "a": np.eye(4, dtype=F32)[rng.integers(0, 4, nb)], "sp": rng.standard_normal((nb, 2)).astype(F32),
         "r": rng.standard_normal((nb, 1)).astype(F32), "done": (rng.random((nb, 1)) < 0.2)}
    S.buffer.push_(d)
    q = _oracle_mlp(pi.mlp); qt = _oracle_mlp(S.agent.pi_target.mlp)
    refnet = o.DiscreteNetwork(q, range(4))
    opt = o.Adam(F32(3e-4))
    draws = [rng.integers(1, nb + 1, 32) for _ in range(3)]
    for ids in draws:
        mb = {k: v[ids - 1] for k, v in d.items()}
        y = o.softq_target(qt(mb["sp"]).detach(), mb["r"], mb["done"], F32(0.95), alpha)
        o.train_step(q.params(), lambda inf: o.td_loss(refnet.value(mb["s"], mb["a"]), y, None, inf), opt, {})
    o.polyak_average(qt.params(), q.params(), 0.005)
    Dmb = crux.buffer_like(S.buffer, capacity=32)
    S.value_training(Dmb, F32(0.95), draws=draws)
    assert_params_close(pi.mlp.get_flat(), q.flat(), 3e-4, 3, what="online Q")
    assert_params_close(S.agent.pi_target.mlp.get_flat(), qt.flat(), 3e-4, 3, what="target Q")
    # action(π, s) samples from softmax(Q/α) once always_stochastic is set (policies.jl:124)
    s1 = np.tile(np.array([[0.3, -0.2]], F32), (40000, 1))
    ps = o.softq_logits(q(s1[:1]).detach(), alpha).numpy()[0]
    freq = np.bincount(host(crux.action(pi, s1)), minlength=4) / 40000
    assert np.allclose(freq, ps, atol=1.5e-2)
    S2 = crux.SoftQ(pi, crux.ContinuousSpace(2), N=400, alpha=alpha, buffer_size=500, buffer_init=200)
    assert crux.solve(S2, crux.SimpleGridWorld(4, seed=0)) is pi and S2.i >= 400 and np.isfinite(pi.mlp.get_flat()).all()
    assert np.isfinite(host(S2.buffer["logprob"])[:len(S2.buffer)]).all() if "logprob" in S2.buffer.schema else True


@pytest.mark.parametrize("algo", ["ppo", "a2c", "reinforce"])
def test_solve_on_policy_discrete_actor_gridworld(crux, ctx, algo):
    """examples/rl/cartpole.jl:8-9,17-25: PPO / A2C / REINFORCE with a DiscreteNetwork ACTOR (categorical exploration with its log-probability,
    one-hot action rows, categorical_logpdf / entropy inside the loss) and a ContinuousNetwork critic, through solve() on the grid world."""
    rng = np.random.default_rng(3)
    D = crux.Dense
    acts = [0, 1, 2, 3]
    A = crux.DiscreteNetwork(crux.Chain(D(2, 64, crux.relu, rng=rng), D(64, 64, crux.relu, rng=rng), D(64, 4, rng=rng)), acts, ctx=ctx)
    V = crux.ContinuousNetwork(crux.Chain(D(2, 64, crux.relu, rng=rng), D(64, 64, crux.relu, rng=rng), D(64, 1, rng=rng)), ctx=ctx)
    before_a, before_v = A.mlp.get_flat().copy(), V.mlp.get_flat().copy()
    n, T = 16, 16
    opt = dict(epochs=3, batch_size=64)
    S_ = crux.ContinuousSpace(2)
    if algo == "ppo":
        S = crux.PPO(crux.ActorCritic(A, V), S_, a_opt=dict(opt), c_opt=dict(opt), N=3 * n * T, dN=n * T, max_steps=30, target_kl=1e9)
    elif algo == "a2c":
        S = crux.A2C(crux.ActorCritic(A, V), S_, a_opt=dict(opt), c_opt=dict(opt), N=3 * n * T, dN=n * T, max_steps=30)
    else:
        S = crux.REINFORCE(A, S_, a_opt=dict(opt), N=3 * n * T, dN=n * T, max_steps=30)
    env = crux.SimpleGridWorld(n, seed=0)
    crux.solve(S, env)
    assert S.i == 3 * n * T
    info = S.training_info()
    assert np.isfinite(info["actor_loss"]) and np.isfinite(info["kl"]) and 0.0 < info["entropy"] <= math.log(4) + 1e-5
    a = host(S.buffer["a"])
    assert a.shape[1] == 4 and np.all(a.sum(1) == 1) and set(np.unique(a)) <= {0.0, 1.0}
    # the stored log-probabilities are those of the categorical draw: log softmax(net(s))[a] under the parameters that sampled them
    lp = host(S.buffer["logprob"])[:, 0]
    assert np.all(lp <= 1e-6) and np.all(lp > -20)
    assert not np.array_equal(before_a, A.mlp.get_flat()) and np.isfinite(A.mlp.get_flat()).all()
    if algo != "reinforce":
        assert not np.array_equal(before_v, V.mlp.get_flat()) and np.isfinite(info["critic_loss"])


def test_solve_dqn_gridworld_readme_example(crux, ctx):
    # README.md:72-82 / test/readme.jl (N reduced): DQN(π=DiscreteNetwork(Chain(Dense(2,8,relu), Dense(8,4)), actions), S, N)
    rng = np.random.default_rng(0)
    env = crux.SimpleGridWorld(4, seed=0)
    pi = crux.DiscreteNetwork(crux.Chain(crux.Dense(2, 8, crux.relu, rng=rng), crux.Dense(8, 4, rng=rng)), [0, 1, 2, 3], ctx=ctx)
    before = pi.mlp.get_flat().copy()
    S = crux.DQN(pi, crux.ContinuousSpace(2), N=600, buffer_size=1000, buffer_init=200)
    crux.solve(S, env)
    assert S.i >= 600 and len(S.buffer) >= 600
    assert not np.array_equal(before, pi.mlp.get_flat()) and np.isfinite(pi.mlp.get_flat()).all()
    a = host(S.buffer["a"])
    assert np.all(a.sum(1) == 1) and set(np.unique(a)) <= {0.0, 1.0}
    # prioritized variant (off_policy.jl:83)
    pi2 = crux.DiscreteNetwork(crux.Chain(crux.Dense(2, 8, crux.relu, rng=rng), crux.Dense(8, 4, rng=rng)), [0, 1, 2, 3], ctx=ctx)
    S2 = crux.DQN(pi2, crux.ContinuousSpace(2), N=400, buffer_size=500, buffer_init=200, prioritized=True)
    crux.solve(S2, env)
    pr = host(S2.buffer.priorities())[:len(S2.buffer)]
    assert np.all(pr > 0) and pr.std() > 0


@pytest.mark.parametrize("kind", ["ddpg", "ddpg_smooth", "td3"])
def test_solve_ddpg_td3_run(crux, ctx, kind):
    """rl/ddpg.jl:45-67, rl/td3.jl:34-57 through solve(): Gaussian-noise exploration on a deterministic actor, value_training with
    the (smoothed) target, the delayed TD3 actor and the polyak of the whole target policy."""
    rng = np.random.default_rng(0)
    D = crux.Dense
    obs, act, hid = 17, 6, 32
    A = crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, act, crux.tanh, rng=rng)), ctx=ctx)
    Q = lambda: crux.ContinuousNetwork(crux.Chain(D(obs + act, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
    pi = crux.ActorCritic(A, crux.DoubleNetwork(Q(), Q()) if kind == "td3" else Q())
    before_a, before_t = A.mlp.get_flat().copy(), None
    n = 16
    opt = dict(batch_size=64)
    expl = crux.GaussianNoiseExplorationPolicy(F32(0.2), a_min=-1.0, a_max=1.0)
    if kind == "td3":
        S = crux.TD3(pi, crux.ContinuousSpace(obs), N=40 * n, dN=2 * n, a_opt=dict(opt, update_every=2), c_opt=dict(opt, epochs=4),
                     pi_explore=expl, buffer_size=2000, buffer_init=128)
    else:
        S = crux.DDPG(pi, crux.ContinuousSpace(obs), N=40 * n, dN=2 * n, a_opt=dict(opt), c_opt=dict(opt, epochs=4), pi_explore=expl,
                      smoothed_target=(kind == "ddpg_smooth"), buffer_size=2000, buffer_init=128)
    before_t = S.agent.pi_target.A.mlp.get_flat().copy()
    out = crux.solve(S, crux.HostLinQuad(n, seed=3))
    assert out is pi and S.i >= 40 * n
    a = host(S.buffer["a"])[:len(S.buffer)]
    assert np.all(np.abs(a) <= 1.0) and a.std() > 0.05                       # clamp(π(s) + noise, a_min, a_max)
    after_a, after_t = A.mlp.get_flat(), S.agent.pi_target.A.mlp.get_flat()
    assert np.isfinite(after_a).all() and not np.array_equal(before_a, after_a)
    assert not np.array_equal(before_t, after_t)                             # the actor copy inside π⁻ follows by polyak
    assert np.abs(after_t - before_t).max() < np.abs(after_a - before_a).max()


def test_solve_sac_runs(crux, ctx):
    rng = np.random.default_rng(0)
    D = crux.Dense
    obs, act, hid = 17, 6, 32
    A = crux.SquashedGaussianPolicy(crux.ContinuousNetwork(crux.Chain(D(obs, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 2 * act, rng=rng)), ctx=ctx))
    Q = lambda: crux.ContinuousNetwork(crux.Chain(D(obs + act, hid, crux.relu, rng=rng), D(hid, hid, crux.relu, rng=rng), D(hid, 1, rng=rng)), ctx=ctx)
    pi = crux.ActorCritic(A, crux.DoubleNetwork(Q(), Q()))
    before = A.mu.mlp.get_flat().copy()
    log = crux.LoggerParams(period=64, verbose=False, logger=None, fns=[])
    S = crux.SAC(pi, crux.ContinuousSpace(obs), N=512, dN=64, c_opt=dict(batch_size=64, epochs=4), buffer_size=2000, buffer_init=256, log=log)
    crux.solve(S, crux.HostLinQuad(16, seed=0))
    after = A.mu.mlp.get_flat()
    assert np.isfinite(after).all() and not np.array_equal(before, after)
    assert len(S.buffer) == 512 and S.i == 512  # the initial fill counts toward N (off_policy.jl:122-133)
    # the off-policy training log is filled from the update's info record (off_policy.jl:104-110, logging.jl:48-54)
    last = log.history[-1]
    assert {"temp_loss", "critic_loss", "critic_grad_norm", "actor_loss", "actor_grad_norm", "entropy", "Q1avg", "Q2avg", "noise_std"} <= set(last), last
    assert all(np.isfinite(last[k]) for k in ("critic_loss", "actor_loss", "critic_grad_norm"))


def test_native_rollout_loop_equals_python_loop(crux, ctx):
    """crux_rollout_host (the steps! loop in one C call, env reached through C callbacks) produces exactly the rollout of
    the Python-driven loop: same env streams, same device noise streams, same bookkeeping."""
    n, T, max_steps = 640, 12, 9   # >= 512 streams: the native loop leapfrogs two halves of the vector step
    outs = []
    for force_python in (False, True):
        pi = _actor_critic(crux, ctx, seed=11)
        env = crux.NativeHostLinQuad(n, seed=4, n_threads=2)
        s = crux.Sampler(env, pi, max_steps=max_steps, required_columns=["return", "logprob", "advantage"], lam=0.95, seed=3)
        s.force_python_loop = force_python
        d1 = {k: v.clone() for k, v in s.steps_(None, Nsteps=n * T, explore=True, reset=True).items()}
        d2 = {k: v.clone() for k, v in s.steps_(None, Nsteps=n * T, explore=True, reset=False).items()}  # continues the streams
        outs.append((d1, d2, s.episode_length.copy(), host(s.cur).copy()))
    for a, b in zip(outs[0][:2], outs[1][:2]):
        for k in a:
            assert torch.equal(a[k], b[k]), f"column {k} differs between the native and the Python rollout loop"
    assert np.array_equal(outs[0][2], outs[1][2]) and np.array_equal(outs[0][3], outs[1][3])
    ee = host(outs[0][0]["episode_end"]).reshape(T, n)
    assert ee[-1].all() and 0 < ee[:-1].mean() < 0.5


@pytest.mark.parametrize("n", [1000, 4100, 37, 4736])
def test_persistent_device_rollout_equals_step_kernels(crux, ctx, n):
    """crux_linquad_rollout (T vector steps in one persistent launch) is bit-identical to T x (crux_rollout_step +
    crux_linquad_step): same noise counters, same FMA order, same bookkeeping.  The stream counts exercise 4-, 14- and 16-stream CTA
    tiles (a warp owns 8 streams: full, partial and empty second warps), a single-stream tile and a ragged last CTA."""
    T, max_steps = 20, 7
    outs = []
    for force_steps in (False, True):
        pi = _actor_critic(crux, ctx, seed=21)
        env = crux.DeviceLinQuad(n, seed=17, max_steps=max_steps, ctx=ctx)
        s = crux.Sampler(env, pi, max_steps=max_steps, required_columns=["return", "logprob", "advantage"], lam=0.95, seed=5)
        s.force_step_kernels = force_steps
        d1 = {k: v.clone() for k, v in s.steps_(None, Nsteps=n * T, explore=True, reset=True).items()}
        d2 = {k: v.clone() for k, v in s.steps_(None, Nsteps=n * 3, explore=True, reset=False).items()}
        outs.append((d1, d2, host(s.cur).copy()))
    for a, b in zip(outs[0][:2], outs[1][:2]):
        for k in a:
            if k in ("logprob", "advantage", "return"):
                assert_close(host(a[k]), host(b[k]), rtol=1e-6, atol=1e-6, what=k)
            else:
                assert torch.equal(a[k], b[k]), f"column {k} differs between the persistent and the per-step rollout"
    assert np.array_equal(outs[0][2], outs[1][2])
    ee = host(outs[0][0]["episode_end"]).reshape(T, n)
    assert ee[-1].all() and ee[max_steps - 1].all() and not ee[0].any()


def test_episodes_and_shuffle(crux, ctx):
    # episodes! (sampler.jl:175-200; test/gym/sampler_tests.jl:32-56,94-100) and shuffle! (experience_buffer.jl:118-124)
    n, max_steps = 4, 6
    pi = _actor_critic(crux, ctx, seed=2)
    env = crux.HostLinQuad(n, seed=3)
    s = crux.Sampler(env, pi, max_steps=max_steps, required_columns=["return", "t"])
    buf = crux.ExperienceBuffer(crux.ContinuousSpace(17), crux.ContinuousSpace(6), 500, ["return", "t"], ctx=ctx)
    data, eps = s.episodes_(buf, Neps=7, explore=True, return_episodes=True)
    assert len(eps) == 7 and eps[0][0] == 1 and eps[-1][1] == data["s"].shape[0] == len(buf)
    ee = host(data["episode_end"])[:, 0].astype(bool)
    t = host(data["t"])[:, 0]
    for a, b in eps:                       # every episode is contiguous, starts at t == 1, ends at its only episode_end flag
        assert t[a - 1] == 1 and ee[b - 1] and not ee[a - 1:b - 1].any() and b - a + 1 <= max_steps
        assert np.array_equal(t[a - 1:b], np.arange(1, b - a + 2))
        sp, s_ = host(data["sp"])[a - 1:b - 1], host(data["s"])[a:b]
        assert np.array_equal(sp, s_)      # rows chain inside an episode
    assert buf.episodes() == eps           # episodes(buffer) == episodes from the sampler (test/gym/sampler_tests.jl:94-100)
    before = {k: host(buf[k]).copy() for k in buf.keys()}
    perm = np.random.default_rng(0).permutation(len(buf)) + 1
    buf.shuffle_(perm)
    for k in buf.keys():
        assert np.array_equal(host(buf[k]), before[k][perm - 1])


@pytest.mark.parametrize("device_env", [False, True])
def test_episode_metrics_on_device(crux, ctx, device_env):
    """sampler.jl:203-240: undiscounted / discounted return, metric_by_key and failure over episodes!, with the per-episode sums
    computed by the returns scan on the device (one stream cut at episode_end)."""
    n, max_steps = 8, 12
    pi = _actor_critic(crux, ctx, seed=5)
    env = crux.DeviceLinQuad(n, seed=4, max_steps=max_steps, ctx=ctx) if device_env else crux.HostLinQuad(n, seed=4)
    s = crux.Sampler(env, pi, max_steps=max_steps)
    data, eps = s.episodes_(Neps=19, explore=True, return_episodes=True)
    r = host(data["r"])[:, 0].astype(np.float64)
    g = float(s.gamma)
    und = np.array([r[a - 1:b].sum() for a, b in eps])
    dis = np.array([o.discounted_return(r[a - 1:b].astype(F32), F32(g)) for a, b in eps])
    assert_close(host(s.episode_sums(data, eps, "r")), und, rtol=1e-5, atol=1e-5, what="undiscounted sums")
    assert_close(host(s.episode_sums(data, eps, "r", g)), dis, rtol=1e-5, atol=1e-5, what="discounted sums")
    assert s.episode_sums(data, [], "r").shape[0] == 0
    # the sampler-level metrics draw fresh (greedy) episodes: deterministic given the env seed, finite, consistent with each other
    u = s.undiscounted_return(Neps=10)
    d = s.discounted_return(Neps=10)
    assert np.isfinite(u) and np.isfinite(d)
    assert s.failure(threshold=1e9, Neps=5) == 1.0 and s.failure(threshold=-1e9, Neps=5) == 0.0
    m = s.metrics_by_key(["r", "done"], Neps=6)
    assert np.isfinite(m[0]) and 0.0 <= m[1] <= 1.0      # at most one terminal transition per episode


class _BimodalEnv:
    """Host env whose episode lengths are bimodal and known in advance: episode c (0-based, counted per stream since construction or
    the last full reset) of stream e lasts 2 steps when (e + c) is even and 20 steps otherwise; r = 1 per step.  A selection of the
    first episodes to FINISH would report ~2, the reference's sequential evaluation (sampler.jl:181-193) the true mix."""
    on_device = False

    def __init__(self, n, crux):
        self.n_envs, self.obs_dim, self.act_dim, self.gamma = n, 17, 6, F32(0.99)
        self.action_space = crux.ContinuousSpace(6)
        self.c = np.zeros(n, dtype=np.int64)
        self.t = np.zeros(n, dtype=np.int64)

    def length(self, e, c):
        return np.where((e + c) % 2 == 0, 2, 20)

    def reset(self, idx=None):
        if idx is None:
            self.c[:] = 0
            self.t[:] = 0
            idx = np.arange(self.n_envs)
        else:
            self.c[idx] += 1
            self.t[idx] = 0
        return np.zeros((len(idx), self.obs_dim), dtype=F32)

    def step(self, a):
        self.t += 1
        done = self.t >= self.length(np.arange(self.n_envs), self.c)
        return np.zeros((self.n_envs, self.obs_dim), dtype=F32), np.ones(self.n_envs, dtype=F32), done


@pytest.mark.parametrize("n,Neps", [(16, 10), (4, 10), (3, 7)])
def test_evaluation_is_unbiased_for_bimodal_episode_lengths(crux, ctx, n, Neps):
    """episodes are chosen by start order (stream k % used, slot k // used), never by finishing order, and every
    chosen episode runs to completion across rollout chunks."""
    pi = _actor_critic(crux, ctx, seed=2)
    env = _BimodalEnv(n, crux)
    s = crux.Sampler(env, pi, max_steps=50)
    used = min(n, Neps)
    expect = [int(env.length(k % used, k // used)) for k in range(Neps)]
    assert 2 in expect and 20 in expect
    assert s.undiscounted_return(Neps=Neps) == pytest.approx(np.mean(expect))
    data, eps = s.episodes_(Neps=Neps, return_episodes=True)
    assert [b - a + 1 for a, b in eps] == expect
    ee = host(data["episode_end"])[:, 0].astype(bool)
    assert ee.sum() == Neps and all(ee[b - 1] for _, b in eps)
    assert s.metric_by_key("r", Neps=Neps) == pytest.approx(np.mean(expect))
    assert s.failure(threshold=10.0, Neps=Neps) == pytest.approx(np.mean(np.array(expect) < 10))
