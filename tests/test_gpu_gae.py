"""-m gpu: hot path (ii) -- crux_fill_gae_returns / crux_whiten / TD targets against the oracle.
Tolerance: rtol 1e-5 fp32 (north_star); nvcc contracts c*A + x into an FMA, the reference does not."""
import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import crux_oracle as o
from gpu_util import F32, assert_close, dev, host, p

pytestmark = pytest.mark.gpu


def _inputs(T, N, seed, p_done=0.01, p_end=0.02, horizon=None):
    rng = np.random.default_rng(seed)
    r, vs, vsp = (rng.standard_normal((T, N)).astype(F32) for _ in range(3))
    done = rng.random((T, N)) < p_done
    ee = done | (rng.random((T, N)) < p_end)
    if horizon:
        ee |= (np.arange(1, T + 1) % horizon == 0)[:, None]
    if T:
        ee[-1] = True
    return r, done.astype(np.uint8), ee.astype(np.uint8), vs, vsp


def _run(ctx, r, done, ee, vs, vsp, gamma, lam, want_adv=True, want_ret=True):
    T, N = r.shape
    d = [dev(ctx, x) for x in (r, done, ee, vs, vsp)]
    adv = torch.full((T, N), float("nan"), device=ctx.device) if want_adv else None
    ret = torch.full((T, N), float("nan"), device=ctx.device) if want_ret else None
    ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, p(d[0]), p(d[1]), p(d[2]), p(d[3]), p(d[4]), T, N, gamma, lam, p(adv), p(ret)))
    ctx.check_flags()
    return (host(adv) if want_adv else None), (host(ret) if want_ret else None)


def test_kat(ctx):
    # inputs of test/gym/sampler_tests.jl:75-81, values of SURVEY 8c
    r = np.full((5, 1), 6, F32); z = np.zeros((5, 1), F32)
    done = np.ones((5, 1), np.uint8); ee = np.zeros((5, 1), np.uint8); ee[-1] = 1
    adv, ret = _run(ctx, r, done, ee, z, z, 0.7, 0.9)
    assert_close(adv[:, 0], [14.60685920715332, 13.66168212890625, 12.161399841308594, 9.779999732971191, 6.0], rtol=1e-6)
    assert_close(ret[:, 0], [16.638599395751953, 15.197999000549316, 13.139999389648438, 10.199999809265137, 6.0], rtol=1e-6)


@pytest.mark.parametrize("T,N", [(1, 1), (1, 7), (3, 33), (32, 4096), (33, 100), (128, 64), (129, 31), (256, 96), (257, 65),
                                 (1000, 40), (1024, 513), (2049, 17)])
def test_against_oracle(ctx, T, N):
    r, done, ee, vs, vsp = _inputs(T, N, seed=T * 131 + N)
    adv, ret = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.99, 0.95)
    assert_close(adv, a0, rtol=1e-5, atol=2e-5, what=f"advantage [{T},{N}]")
    assert_close(ret, r0, rtol=1e-5, atol=2e-5, what=f"return [{T},{N}]")


@pytest.mark.parametrize("pattern", ["last_only", "every_row", "none_but_done"])
def test_episode_end_patterns(ctx, pattern):
    T, N = 300, 70
    r, done, ee, vs, vsp = _inputs(T, N, seed=5)
    if pattern == "last_only":
        ee[:] = 0; ee[-1] = 1; done[:] = 0
    elif pattern == "every_row":
        ee[:] = 1
    else:
        ee[:] = done; ee[-1] = 1
    adv, ret = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.99, 0.95)
    assert_close(adv, a0, rtol=1e-5, atol=5e-5, what=pattern)
    assert_close(ret, r0, rtol=1e-5, atol=5e-5, what=pattern)
    if pattern == "every_row":  # one-step episodes: A = r + (1-done) γ V(sp) - V(s), R = r
        assert_close(ret, r, rtol=0, atol=0)


def test_single_output_and_empty(ctx):
    r, done, ee, vs, vsp = _inputs(40, 50, seed=9)
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.9, 0.8)
    adv, _ = _run(ctx, r, done, ee, vs, vsp, 0.9, 0.8, want_ret=False)
    _, ret = _run(ctx, r, done, ee, vs, vsp, 0.9, 0.8, want_adv=False)
    assert_close(adv, a0, atol=2e-5); assert_close(ret, r0, atol=2e-5)
    # empty rollouts are no-ops; both outputs NULL is an argument error
    assert ctx.lib.crux_fill_gae_returns(ctx.h, None, None, None, None, None, 0, 5, 0.9, 0.9, None, None) == 0
    assert ctx.lib.crux_fill_gae_returns(ctx.h, None, None, None, None, None, 5, 0, 0.9, 0.9, None, None) == 0
    d = [dev(ctx, x) for x in (r, done, ee, vs, vsp)]
    assert ctx.lib.crux_fill_gae_returns(ctx.h, p(d[0]), p(d[1]), p(d[2]), p(d[3]), p(d[4]), 40, 50, 0.9, 0.8, None, None) == 1


def test_half_cheetah_real_data(ctx, golden_dir):
    # reference fixture rows as ONE env stream (N=1, T=2000); episodes from t == 1 (experience_buffer.jl:198-200)
    import os
    d = np.load(os.path.join(golden_dir, "half_cheetah_2k.npz"))
    T = 2000
    r = d["r"].reshape(T, 1)
    done = d["done"].reshape(T, 1)
    ee = np.zeros((T, 1), np.uint8); ee[999] = 1; ee[1999] = 1
    rng = np.random.default_rng(0)
    vs, vsp = rng.standard_normal((T, 1)).astype(F32), rng.standard_normal((T, 1)).astype(F32)
    adv, ret = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    a_ref = np.zeros(T, F32); r_ref = np.zeros(T, F32)
    for ep in (range(0, 1000), range(1000, 2000)):  # the reference's per-episode calls (sampler.jl:56-57)
        o.fill_gae(r[:, 0], done[:, 0], vs[:, 0], vsp[:, 0], 0.95, 0.99, rng=ep, out=a_ref)
        o.fill_returns(r[:, 0], 0.99, rng=ep, out=r_ref)
    assert_close(adv[:, 0], a_ref, rtol=2e-5, atol=1e-4)
    assert_close(ret[:, 0], r_ref, rtol=2e-5, atol=1e-4)


def test_full_size_properties(ctx):
    """BASELINE GAE microbench shape [2048, 16384] (738 MB): direct oracle compare + linearity in (r, V)."""
    T, N = 2048, 16384
    r, done, ee, vs, vsp = _inputs(T, N, seed=2, p_done=0.001, p_end=0.0, horizon=1000)
    adv, ret = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    c_oracle.set_threads(8)
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.99, 0.95)
    c_oracle.set_threads(1)
    assert_close(adv, a0, rtol=1e-5, atol=1e-4, what="full-size advantage")
    assert_close(ret, r0, rtol=1e-5, atol=1e-4, what="full-size return")
    adv2, ret2 = _run(ctx, 2 * r, done, ee, 2 * vs, 2 * vsp, 0.99, 0.95)  # exact: scaling by 2 commutes with rounding
    assert np.array_equal(adv2, 2 * adv) and np.array_equal(ret2, 2 * ret)


# ---- the TMA-fed streaming scan (gae_tma.cu): forced with CRUX_GAE=tma, segment hand-off exercised through CRUX_GAE_CFG
@pytest.fixture
def gae_env(monkeypatch):
    def set_(mode, cfg=None):
        monkeypatch.setenv("CRUX_GAE", mode)
        if cfg:
            monkeypatch.setenv("CRUX_GAE_CFG", cfg)
        else:
            monkeypatch.delenv("CRUX_GAE_CFG", raising=False)
    return set_


@pytest.mark.parametrize("T,N,cfg", [(1, 16, None), (5, 48, None), (64, 128, None), (100, 144, "16,2,1,2"), (257, 272, "16,3,2,2"),
                                     (300, 4112, "32,2,1,1"), (1000, 1040, "32,4,3,2"), (1024, 4096, None), (2049, 160, "16,4,5,2"),
                                     (513, 16400, None),
                                     # 64-stream tiles (fifth field = chain warps) and explicit 128-stream tiles on a narrow rollout
                                     (300, 4112, "32,3,1,1,1,2"), (257, 272, "16,3,2,2,1,2"), (1024, 4096, "16,6,0,1,2,2"), (700, 3088, "0,0,0,0,1,2"),
                                     # 32-stream tiles (sixth field = streams per chain lane; the default up to 32 x SM-count streams)
                                     (300, 4112, "32,3,2,1,1,1"), (1000, 1040, "64,3,3,2,1,1"), (2049, 160, "32,4,5,2,1,1"), (1024, 4096, "0,0,0,0,0,1"),
                                     (129, 4736, None)])
def test_tma_path_against_oracle(ctx, gae_env, T, N, cfg):
    r, done, ee, vs, vsp = _inputs(T, N, seed=T * 7 + N)
    gae_env("tma", cfg)
    l0 = ctx.launch_count()
    adv, ret = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    assert ctx.launch_count() - l0 == 1
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.99, 0.95)
    assert_close(adv, a0, rtol=1e-5, atol=2e-5, what=f"tma advantage [{T},{N}] cfg={cfg}")
    assert_close(ret, r0, rtol=1e-5, atol=2e-5, what=f"tma return [{T},{N}] cfg={cfg}")
    # the streaming scan IS the sequential recurrence: identical bits for every tiling / segmentation
    gae_env("tma", "32,2,1,1,2,2")
    adv2, ret2 = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    assert np.array_equal(adv, adv2) and np.array_equal(ret, ret2)
    # and within tolerance of the register-resident scan kernel
    gae_env("scan")
    adv3, ret3 = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    assert_close(adv, adv3, rtol=1e-5, atol=2e-5); assert_close(ret, ret3, rtol=1e-5, atol=2e-5)


def test_tma_path_single_output_nan_and_fallback(ctx, crux, gae_env):
    r, done, ee, vs, vsp = _inputs(200, 256, seed=3)
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.9, 0.8)
    gae_env("tma", "16,3,2,2,0,2")
    adv, _ = _run(ctx, r, done, ee, vs, vsp, 0.9, 0.8, want_ret=False)
    _, ret = _run(ctx, r, done, ee, vs, vsp, 0.9, 0.8, want_adv=False)
    assert_close(adv, a0, atol=2e-5); assert_close(ret, r0, atol=2e-5)
    r[17, 33] = np.nan
    d = [dev(ctx, x) for x in (r, done, ee, vs, vsp)]
    out = ctx.empty((200, 256))
    ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, p(d[0]), p(d[1]), p(d[2]), p(d[3]), p(d[4]), 200, 256, 0.9, 0.9, p(out), None))
    with pytest.raises(crux.NaNError):
        ctx.check_flags()
    # N not a multiple of 16: not TMA-encodable (u8 row pitch), the scan kernel takes over silently
    r, done, ee, vs, vsp = _inputs(130, 100, seed=4)
    adv, ret = _run(ctx, r, done, ee, vs, vsp, 0.99, 0.95)
    a0, r0 = c_oracle.gae_returns(r, done, ee, vs, vsp, 0.99, 0.95)
    assert_close(adv, a0, rtol=1e-5, atol=2e-5); assert_close(ret, r0, rtol=1e-5, atol=2e-5)


def test_nan_advantage_flag(ctx, crux):
    r, done, ee, vs, vsp = _inputs(16, 8, seed=1)
    r[3, 2] = np.nan
    d = [dev(ctx, x) for x in (r, done, ee, vs, vsp)]
    adv = ctx.empty((16, 8))
    ctx.check(ctx.lib.crux_fill_gae_returns(ctx.h, p(d[0]), p(d[1]), p(d[2]), p(d[3]), p(d[4]), 16, 8, 0.9, 0.9, p(adv), None))
    with pytest.raises(crux.NaNError):  # sampler.jl:270 @assert !isnan(A)
        ctx.check_flags()
    ctx.check_flags()  # the flag is cleared by the read


@pytest.mark.parametrize("n", [2, 1000, 131072, 1 << 20])
def test_whiten(ctx, n):
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3 + 1.5).astype(F32)
    t = dev(ctx, x)
    ctx.check(ctx.lib.crux_whiten(ctx.h, p(t), n))
    x64 = x.astype(np.float64)
    want = ((x64 - x64.mean()) / x64.std(ddof=1)).astype(F32)  # utils.jl:41-42, Bessel
    assert_close(host(t), want, rtol=1e-5, atol=2e-6)
    assert_close(host(t), o.whiten(x), rtol=1e-4, atol=1e-5)


def test_td_targets(ctx):
    rng = np.random.default_rng(0)
    B, nA = 777, 4
    r = rng.standard_normal(B).astype(F32); done = (rng.random(B) < 0.3).astype(np.uint8)
    q = rng.standard_normal((B, nA)).astype(F32)
    y = ctx.empty((B,))
    ctx.check(ctx.lib.crux_dqn_target(ctx.h, p(dev(ctx, r)), p(dev(ctx, done)), p(dev(ctx, q)), B, nA, 0.95, p(y)))
    want = o.dqn_target(q, r, done, 0.95).numpy()[:, 0]
    assert_close(host(y), want, rtol=1e-6, atol=1e-6)
    # sac target
    q1, q2, lp = (rng.standard_normal(B).astype(F32) for _ in range(3))
    la = dev(ctx, np.array([-0.7], F32))
    ctx.check(ctx.lib.crux_sac_target(ctx.h, p(dev(ctx, r)), p(dev(ctx, done)), p(dev(ctx, q1)), p(dev(ctx, q2)), p(dev(ctx, lp)), B,
                                      0.99, p(la), p(y)))
    want = r + F32(0.99) * (1 - done.astype(F32)) * (np.minimum(q1, q2) - np.exp(F32(-0.7)) * lp)
    assert_close(host(y), want, rtol=1e-5, atol=1e-6)
    # td_error and discrete Q(s,a)
    oh = np.eye(nA, dtype=F32)[rng.integers(0, nA, B)]
    qsa = ctx.empty((B,))
    ctx.check(ctx.lib.crux_discrete_q_sa(ctx.h, p(dev(ctx, q)), p(dev(ctx, oh)), B, nA, p(qsa)))
    assert_close(host(qsa), (q * oh).sum(1), rtol=1e-6)
    e = ctx.empty((B,))
    ctx.check(ctx.lib.crux_td_error(ctx.h, p(qsa), p(y), B, p(e)))
    assert_close(host(e), np.abs(host(qsa) - host(y)), rtol=0, atol=0)


def test_normalize_obs(ctx):
    x = np.random.default_rng(0).standard_normal(1000).astype(F32)
    t = dev(ctx, x)
    ctx.check(ctx.lib.crux_normalize_obs(ctx.h, p(t), 1000, 1.0, 2.0, p(t)))
    assert_close(host(t), (x - F32(1)) / F32(2), rtol=1e-7)  # spaces.jl:24-25; test/spaces_tests.jl:34-40
