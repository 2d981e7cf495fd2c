"""CPU: the oracle against every known-answer value the reference's own tests hold for the hot path
(tests/golden/ref_kats.json cites each reference test file:line)."""
import json
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import crux_oracle as o

F32 = np.float32


@pytest.fixture(scope="module")
def kats(golden_dir):
    with open(os.path.join(golden_dir, "ref_kats.json")) as f:
        return json.load(f)


def test_circ_inds(kats):
    # test/experience_buffer_tests.jl:23-28
    exp = {(4, 60, 100): list(range(4, 64)), (1, 100, 100): list(range(1, 101)), (1, 101, 100): list(range(1, 101)) + [1],
           (1, 120, 100): list(range(1, 101)) + list(range(1, 21)), (90, 20, 100): list(range(90, 101)) + list(range(1, 10))}
    for start, n, cap in kats["circ_inds"]["cases"]:
        assert o.circ_inds(start, n, cap).tolist() == exp[(start, n, cap)]
        # C oracle is 0-based
        assert (c_oracle.ring_indices(start - 1, n, cap) + 1).tolist() == exp[(start, n, cap)]


def test_split_batches(kats):
    # test/experience_buffer_tests.jl:177-180
    for N, fracs, want in kats["split_batches"]["cases"]:
        assert o.split_batches(N, fracs).tolist() == want
    with pytest.raises(AssertionError):
        o.split_batches(100, 0.4)


def _buf(cap=100, sdim=2, adim=1, **kw):
    return o.ExperienceBuffer.create(sdim, adim, cap, **kw)


def _d(n, sdim=2, adim=1, s=2.0):
    return {"s": s * np.ones((n, sdim), F32), "a": np.ones((n, adim), F32), "sp": np.ones((n, sdim), F32),
            "r": np.ones((n, 1), F32), "done": np.zeros((n, 1), bool)}


def test_get_last_N_indices(kats):
    # test/experience_buffer_tests.jl:32-51
    b = _buf()
    b.push(_d(50))
    for N, (lo, hi) in kats["last_n_partial"]["cases"]:
        assert b.get_last_N_indices(N).tolist() == list(range(lo, hi + 1))
    b.push(_d(50))
    b.push(_d(50))
    assert b.get_last_N_indices(10).tolist() == list(range(41, 51))
    assert b.get_last_N_indices(1).tolist() == [50]
    assert b.get_last_N_indices(50).tolist() == list(range(1, 51))
    assert b.get_last_N_indices(51).tolist() == [100] + list(range(1, 51))
    assert b.get_last_N_indices(100).tolist() == list(range(51, 101)) + list(range(1, 51))
    assert b.get_last_N_indices(1000).tolist() == list(range(51, 101)) + list(range(1, 51))


def test_push_semantics():
    # test/experience_buffer_tests.jl:121-147
    b = _buf(100, 2, 4)
    b.push(_d(1, 2, 4))
    assert len(b) == 1 and np.all(b["s"] == 2) and np.all(b["sp"] == 1) and np.all(b["r"] == 1) and not b["done"].any()
    d = {"s": 3 * np.ones((3, 2), F32), "a": (np.random.default_rng(0).random((3, 4)) < 0.5).astype(F32),
         "sp": 5 * np.ones((3, 2), F32), "r": 6 * np.ones((3, 1), F32), "done": np.ones((3, 1), bool)}
    b.push(d)
    assert len(b) == 4
    assert np.all(b["s"][1:] == 3) and np.array_equal(b["a"][1:], d["a"]) and np.all(b["sp"][1:] == 5)
    assert np.all(b["r"][1:] == 6) and b["done"][1:].all()
    b.push(b)
    assert len(b) == 8
    for k in b.keys():
        assert np.array_equal(b[k][:4], b[k][4:8])
    b.clear()
    assert len(b) == 0 and b.next_ind == 1


def test_push_more_than_capacity_later_rows_win():
    # experience_buffer.jl:236,249-252
    b = _buf(5)
    d = _d(12)
    d["r"] = np.arange(12, dtype=F32).reshape(12, 1)
    I = b.push(d)
    assert I.tolist() == [1, 2, 3, 4, 5, 1, 2, 3, 4, 5, 1, 2]
    assert b["r"][:, 0].tolist() == [10, 11, 7, 8, 9]
    assert len(b) == 5 and b.next_ind == 3 and b.total_count == 12


def test_priorities(kats):
    # test/experience_buffer_tests.jl:193-205
    k = kats["priorities"]
    b = _buf(50, 2, 4, prioritized=True)
    b.update_priorities(np.array(k["I"]), np.array(k["v"], dtype=F32))
    assert b.pp.max_priority == F32(3.0) + o.EPS32  # Julia `== 3.0` holds after Float32 rounding: 3 + eps rounds to ...
    for i, v in zip(k["I"], k["v"]):
        assert np.isclose(b.pp.priorities[i - 1], F32(v) ** F32(k["alpha"]))
    # SURVEY 8c literal values including the +eps
    assert np.allclose(b.pp.priorities[:3], [1.0000001, 1.5157167, 1.9331821], rtol=1e-7)
    d = _d(3, 2, 4)
    b.push(d)
    b.push(d)
    for i in range(6):
        assert np.isclose(b.pp.priorities[i], F32(3.0) ** F32(0.6))


def test_prioritized_sampling_frequencies():
    # test/experience_buffer_tests.jl:247-262
    rng = np.random.default_rng(0)
    b = _buf(50, 2, 4, prioritized=True)
    b.push(_d(6, 2, 4))
    b.data["s"][:6] = rng.random((6, 2), dtype=F32)
    b.update_priorities(np.arange(1, 7), np.arange(1, 7, dtype=F32))
    t = _buf(1000, 2, 4, extras=("weight",))
    o.prioritized_sample(t, b, rng.random(1000))
    pr = b.pp.priorities[:6]
    freqs = np.array([(t.indices == i).sum() for i in range(1, 7)]) / 1000
    probs = pr / pr.sum()
    assert np.all(np.abs(freqs - probs) / probs < 0.01)
    assert np.array_equal(t["s"], b["s"][t.indices - 1])
    assert np.all(t["weight"] <= 1.0)


def test_multi_source_fractions():
    # test/experience_buffer_tests.jl:225-242
    srcs = []
    for v in (1.0, 2.0, 3.0):
        s = _buf(10, 2, 4)
        s.push(_d(1, 2, 4, s=v))
        srcs.append(s)
    t = _buf(10, 2, 4)
    batches = o.rand_b(t, srcs, [np.ones(4, int), np.ones(3, int), np.ones(3, int)])
    assert batches.tolist() == [4, 3, 3]
    assert np.all(t["s"][:4] == 1) and np.all(t["s"][4:7] == 2) and np.all(t["s"][7:] == 3)


def test_schedule():
    # test/util_tests.jl:55-86 (l = LinearDecaySchedule(1.0, 0.1, 10); m(31) == 0.1; m(0) == 1)
    l = o.LinearDecaySchedule(1.0, 0.1, 10)
    assert l(0) == 1.0 and l(31) == 0.1 and l(10) == pytest.approx(0.1) and l(5) == pytest.approx(0.55)


def test_spaces(kats):
    # test/spaces_tests.jl:15,24,34-40
    k = kats["whiten_space"]
    assert o.whiten(np.array([k["x"]], F32), k["mu"], k["sigma"])[0] == F32(k["y"])
    assert o.onehot(3, [1, 2, 3, 4]).tolist() == [False, False, True, False]


def test_discounted_return_recurrence():
    # test/gym/sampler_tests.jl:62-66
    # a single terminal reward ur after n steps: discounted_return == ur * γ^(n-1) (the reference's assertion, ≈)
    for ur, n in [(10.0, 7), (-5.0, 3), (3.0, 1)]:
        r = [0.0] * (n - 1) + [ur]
        assert np.isclose(o.discounted_return(r, 0.95), ur * 0.95 ** (n - 1), rtol=1e-6)
    r = [1.0, 2.0, 3.0]
    assert np.allclose(o.fill_returns(np.array(r, F32), 0.9), [1 + 0.9 * (2 + 0.9 * 3), 2 + 0.9 * 3, 3], rtol=1e-6)


def test_gae_kat(kats):
    # inputs of test/gym/sampler_tests.jl:75-81; values from SURVEY 8c ("parity unpinned" by the reference)
    k = kats["gae_kat"]
    r = np.full(5, 6, F32)
    z = np.zeros(5, F32)
    adv = o.fill_gae(r, np.ones(5), z, z, 0.9, 0.7)
    ret = o.fill_returns(r, 0.7)
    assert adv.tolist() == [F32(x) for x in k["advantage"]]
    assert ret.tolist() == [F32(x) for x in k["return"]]
    ee = np.zeros((5, 1)); ee[-1] = 1
    a2, r2 = c_oracle.gae_returns(r.reshape(5, 1), np.ones((5, 1)), ee, z.reshape(5, 1), z.reshape(5, 1), 0.7, 0.9)
    assert a2[:, 0].tolist() == adv.tolist() and r2[:, 0].tolist() == ret.tolist()


def test_gae_c_equals_numpy_random():
    rng = np.random.default_rng(2)
    T, N = 67, 129
    r, vs, vsp = (rng.standard_normal((T, N)).astype(F32) for _ in range(3))
    done = rng.random((T, N)) < 0.03
    ee = done | (rng.random((T, N)) < 0.05)
    ee[-1] = True
    a1, r1 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.99, 0.95)
    a2, r2 = o.gae_returns_TN(r, done, ee, vs, vsp, 0.99, 0.95)
    assert np.array_equal(a1, a2) and np.array_equal(r1, r2)
    # per-episode single-env recurrence (the reference's own form) on one stream
    e = 5
    ends = np.flatnonzero(ee[:, e])
    start = 0
    for end in ends:
        rg = range(start, end + 1)
        a = o.fill_gae(r[:, e], done[:, e], vs[:, e], vsp[:, e], 0.95, 0.99, rng=rg)
        assert np.array_equal(a[start:end + 1], a1[start:end + 1, e])
        start = end + 1


def test_half_cheetah_fixture_episodes(golden_dir):
    # examples/il/expert_data/half_cheetah_mujoco.bson (first 2000 rows): episodes from t == 1 (experience_buffer.jl:198-200)
    d = np.load(os.path.join(golden_dir, "half_cheetah_2k.npz"))
    assert d["s"].shape == (2000, 17) and d["a"].shape == (2000, 6)
    b = o.ExperienceBuffer({"s": d["s"], "t": d["t"].reshape(-1, 1)})
    assert b.episodes() == [(1, 1000), (1001, 2000)]
    # consecutive rows inside an episode chain: sp[t] == s[t+1]
    assert np.array_equal(d["sp"][:999], d["s"][1:1000])


def test_gaussian_policy_self_consistency():
    # test/policy_tests.jl:101-110,257-281: logpdf(exploration) == logprob, entropy 0-dim, std = exp(logΣ)
    rng = np.random.default_rng(0)
    mu = o.MLP([3, 8, 2], [o.ACT_TANH, o.ACT_IDENTITY], rng)
    p = o.GaussianPolicy(mu, np.array([-0.5, 0.25], F32))
    s = rng.standard_normal((100, 3)).astype(F32)
    a, lp = p.exploration(s, rng.standard_normal((100, 2)).astype(F32))
    assert np.allclose(p.logpdf(s, a.detach().numpy()).detach().numpy(), lp.detach().numpy(), atol=1e-6)
    assert p.entropy(s).ndim == 0
    eps = rng.standard_normal((100000, 2)).astype(F32)
    a, _ = p.exploration(np.zeros((100000, 3), F32), eps)
    assert np.allclose(a.detach().numpy().std(0), np.exp([-0.5, 0.25]), rtol=2e-2)


def test_squashed_policy_self_consistency():
    # test/policy_tests.jl:299-321
    rng = np.random.default_rng(1)
    mu = o.MLP([3, 8, 2], [o.ACT_TANH, o.ACT_IDENTITY], rng)
    ls = o.MLP([3, 8, 2], [o.ACT_TANH, o.ACT_IDENTITY], rng)
    p = o.SquashedGaussianPolicy(mu, ls, ascale=2.0)
    s = rng.standard_normal((64, 3)).astype(F32)
    a, lp = p.exploration(s, rng.standard_normal((64, 2)).astype(F32))
    assert float(a.abs().max()) <= 2.0
    # atanh(tanh(x)) loses precision in float32 as |a| -> ascale: compare away from saturation (the reference's ≈)
    ok = (a.abs().max(dim=1).values < 0.99 * 2.0).numpy()
    assert ok.sum() > 32
    assert np.allclose(p.logpdf(s, a.detach().numpy()).detach().numpy()[ok], lp.detach().numpy()[ok], atol=5e-3)
    assert tuple(p.entropy(s).shape) == (64, 1)


def test_adam_first_step_is_eta_sign():
    # Flux Adam [3P]: first step = eta * g/(|g| + eps*sqrt(1-b2)...) ~ eta*sign(g)
    import torch
    p = torch.tensor([1.0, -2.0, 3.0], requires_grad=True)
    p.grad = torch.tensor([0.5, -0.25, 0.0])
    o.Adam(eta=1e-3).step([p])
    assert np.allclose(p.detach().numpy(), [1 - 1e-3, -2 + 1e-3, 3.0], atol=1e-8)


def test_soft_value_and_reinforce_restatements():
    # rl/softq.jl:8: α·logsumexp(Q/α) -> max(Q) as α -> 0, >= max(Q) always, = log Σ exp Q at α = 1 (float64 cross-check)
    import torch
    rng = np.random.default_rng(3)
    q = (2 * rng.standard_normal((50, 6))).astype(F32)
    v1 = o.soft_value(q, 1.0).numpy()[:, 0]
    assert np.allclose(v1, np.log(np.exp(q.astype(np.float64)).sum(1)), rtol=1e-6)
    v0 = o.soft_value(q, 1e-2).numpy()[:, 0]
    assert np.allclose(v0, q.max(1), atol=0.05) and np.all(v0 >= q.max(1) - 1e-6)
    assert np.allclose(o.softq_logits(q, 0.5).sum(1).numpy(), 1.0, atol=1e-6)
    y = o.softq_target(q, np.ones(50, F32), np.ones(50, bool), 0.9, 0.5).numpy()[:, 0]
    assert np.array_equal(y, np.ones(50, F32))            # done rows never bootstrap
    # rl/reinforce.jl:4-13 == a2c_loss (rl/a2c.jl:4-16) with the return column as the weight, λp = 1, λe = 0
    mu = o.MLP([3, 8, 2], [o.ACT_TANH, o.ACT_IDENTITY], rng)
    pi = o.GaussianPolicy(mu, np.full(2, -0.5, F32))
    D = {"s": rng.standard_normal((32, 3)).astype(F32), "a": rng.standard_normal((32, 2)).astype(F32),
         "logprob": rng.standard_normal((32, 1)).astype(F32), "return": rng.standard_normal((32, 1)).astype(F32)}
    i1, i2 = {}, {}
    l1 = o.reinforce_loss(pi, {}, D, i1)
    l2 = o.a2c_loss(pi, {"lp": F32(1), "le": F32(0)}, dict(D, advantage=D["return"]), i2)
    assert torch.equal(l1, l2) and i1["kl"] == i2["kl"] and i1["entropy"] == i2["entropy"]


def test_lagrange_ppo_loss_restatement():
    # rl/ppo.jl:70-131: with a zero penalty the loss IS ppo_loss; the PID terms follow the hand-computed recurrences
    import torch
    rng = np.random.default_rng(4)
    mu = o.MLP([3, 8, 2], [o.ACT_TANH, o.ACT_IDENTITY], rng)
    pi = o.GaussianPolicy(mu, np.full(2, -0.5, F32))
    n = 64
    D = {"s": rng.standard_normal((n, 3)).astype(F32), "a": rng.standard_normal((n, 2)).astype(F32),
         "logprob": (-3 + 0.1 * rng.standard_normal(n)).astype(F32), "advantage": rng.standard_normal(n).astype(F32),
         "return": rng.standard_normal(n).astype(F32), "cost_advantage": rng.standard_normal(n).astype(F32),
         "cost": (rng.random(n) < 0.5).astype(F32), "episode_end": rng.random(n) < 0.2}
    P0 = o.lagrange_params(target_cost=1e6)                       # Δ << 0: I clamps at 0, Kp·smooth_Δ < 0 -> penalty 0
    i0, i1 = {}, {}
    l0 = o.lagrange_ppo_loss(pi, P0, D, i0)
    l1 = o.ppo_loss(pi, P0, D, i1)
    assert i0["penalty"] == 0.0 and torch.allclose(l0, l1, rtol=0, atol=1e-7) and i0["kl"] == i1["kl"]
    P = o.lagrange_params(target_cost=0.025, Ki=0.1, Kp=2, Kd=1)
    Jc = D["cost"].sum() / D["episode_end"].sum()
    info = {}
    o.lagrange_ppo_loss(pi, P, D, info)
    d = Jc - 0.025
    want_I, want_sd, want_sj = 0.1 * d, 0.05 * d, 0.05 * Jc
    want_pen = 2 * want_sd + want_I + 1 * max(0.0, want_sj - 0.0)
    assert np.isclose(info["cur_cost"], Jc, rtol=1e-6) and np.isclose(info["integral term"], want_I, rtol=1e-5)
    assert np.isclose(info["prop_term"], 2 * want_sd, rtol=1e-5) and np.isclose(info["deriv_term"], want_sj, rtol=1e-5)
    assert np.isclose(info["penalty"], want_pen, rtol=1e-5) and np.isclose(float(P["Jc_prev"]), want_sj, rtol=1e-5)
    info2 = {}
    l2 = o.lagrange_ppo_loss(pi, P, D, info2)                     # second evaluation: the state carries over
    assert np.isclose(info2["integral term"], 2 * want_I, rtol=1e-5) and info2["penalty"] > info["penalty"]
    assert np.isclose(float(l2), (info2["p_loss"] + 0.1 * -info2["entropy"] + info2["cost_loss"]) / (1 + info2["penalty"]), rtol=1e-5)
