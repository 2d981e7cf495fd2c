"""-m gpu: the device-side synthetic LinQuad MDP (bench `value` leg) against its numpy specification."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, dev, host, p

pytestmark = pytest.mark.gpu


def test_linquad_step(ctx):
    spec = o.LinQuadSpec(17, 6, seed=0)
    N = 4096
    env = C.c_void_p()
    ctx.check(ctx.lib.crux_linquad_create(ctx.h, 17, 6, p(spec.A), p(spec.B), N, 5, 99, C.byref(env)))
    obs = ctx.empty((N, 17))
    ctx.check(ctx.lib.crux_linquad_reset(env, p(obs)))
    s0 = host(obs)
    assert np.all(np.abs(s0) <= 0.1) and s0.std() > 0.04  # s0 ~ U(-0.1, 0.1)^17
    rng = np.random.default_rng(0)
    sp, r, nxt = ctx.empty((N, 17)), ctx.empty((N,)), ctx.empty((N, 17))
    done = torch.empty(N, dtype=torch.uint8, device=ctx.device); ee = torch.empty(N, dtype=torch.uint8, device=ctx.device)
    cur = obs
    for t in range(7):
        a = rng.standard_normal((N, 6)).astype(F32)
        if t == 3:
            cur = dev(ctx, host(cur) + F32(6.0) * (np.arange(N) % 2 == 0)[:, None].astype(F32))  # push half the envs over |s_1| > 5
        ctx.check(ctx.lib.crux_linquad_step(env, p(cur), p(dev(ctx, a)), p(sp), p(r), p(done), p(ee), p(nxt), 0))
        s = host(cur)
        mean_sp = s @ spec.A.T + np.tanh(a) @ spec.B.T
        sph = host(sp)
        xi = (sph - mean_sp) / 0.01
        inside = np.abs(mean_sp) < 9.9
        assert abs(xi[inside].mean()) < 0.02 and abs(xi[inside].std() - 1) < 0.03, (xi[inside].mean(), xi[inside].std())
        assert np.all(np.abs(sph) <= 10)
        want_r = 1 - (sph * sph).sum(1) / 17 - 0.1 * (a * a).sum(1) / 6
        assert_close(host(r), want_r, rtol=1e-5, atol=1e-5)
        dn = host(done).astype(bool)
        assert np.array_equal(dn, np.abs(sph[:, 0]) > 5)
        want_end = dn | ((t + 1) % 5 == 0 if t < 3 else False)
        e = host(ee).astype(bool)
        if t < 3:
            assert np.array_equal(e, dn)
        # next obs: sp if the episode continues, a fresh s0 otherwise (reset_sampler! sampler.jl:31-43)
        nx = host(nxt)
        assert np.array_equal(nx[~e], sph[~e])
        if e.any():
            assert np.all(np.abs(nx[e]) <= 0.1)
        cur = nxt.clone()
    # max_steps = 5 terminates every stream that never hit the terminal set (sampler.jl:131)
    ctx.check(ctx.lib.crux_linquad_reset(env, p(obs)))
    cur = obs
    for t in range(5):
        a = np.zeros((N, 6), F32)
        ctx.check(ctx.lib.crux_linquad_step(env, p(cur), p(dev(ctx, a)), p(sp), p(r), p(done), p(ee), p(nxt), 0))
        assert host(ee).all() == (t == 4)
        cur = nxt.clone()
    # force_end marks every row (steps!(reset=true) -> terminate_episode! sampler.jl:148)
    ctx.check(ctx.lib.crux_linquad_step(env, p(cur), p(dev(ctx, a)), p(sp), p(r), p(done), p(ee), p(nxt), 1))
    assert host(ee).all() and not host(done).any()
    ctx.lib.crux_linquad_destroy(env)


def test_host_and_device_env_share_noise_streams(ctx, crux):
    """NativeHostLinQuad (C++) and DeviceLinQuad (CUDA) draw the same Philox streams: same seed + same actions ->
    the same trajectories up to libm-vs-CUDA math rounding."""
    n = 512
    dev_env = crux.DeviceLinQuad(n, seed=21, max_steps=1000, ctx=ctx)
    host_env = crux.NativeHostLinQuad(n, seed=21)
    obs = ctx.empty((n, 17))
    dev_env.reset_into(obs)
    s0 = host_env.reset()
    assert np.array_equal(host(obs), s0)
    rng = np.random.default_rng(0)
    sp, r, nxt = ctx.empty((n, 17)), ctx.empty((n,)), ctx.empty((n, 17))
    done = torch.empty(n, dtype=torch.uint8, device=ctx.device); ee = torch.empty(n, dtype=torch.uint8, device=ctx.device)
    cur = obs
    for t in range(4):
        a = rng.standard_normal((n, 6)).astype(F32)
        dev_env.step_into(cur, dev(ctx, a), sp, r, done, ee, nxt)
        sph, rh, dh = host_env.step(a)
        assert_close(host(sp), sph, rtol=1e-5, atol=2e-6, what=f"sp step {t}")
        assert_close(host(r), rh, rtol=1e-5, atol=1e-5, what=f"r step {t}")
        assert np.array_equal(host(done).astype(bool), dh)
        cur = nxt.clone()
