"""The C client of the ABI (tests/abi_smoke.c) against its ctypes twin: one PPO iteration on the device env -- rollout, values, GAE,
whiten, update with caller-chosen row orders, replay-buffer bookkeeping -- driven once from plain C through include/crux_cuda.h and once
from Python through crux.jl_b200/_abi.py must give the same numbers: the header, the ctypes tables and the struct layouts agree."""
import ctypes as C
import math
import subprocess

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
F32 = np.float32


def _fill(n, phase, scale):
    k = np.arange(n, dtype=np.float64)
    return (np.sin(0.37 * k + phase) * scale).astype(F32)


def _cs(x):
    x = np.asarray(x, dtype=np.float64).reshape(-1)
    return float((x * ((np.arange(x.size) % 7) + 1)).sum())


def _python_twin(crux, ctx, N, T, seed):
    from crux_b200 import _abi
    from crux_b200.device import ptr
    lib, n = ctx.lib, N * T
    out = {}

    def mlp(dims, phase):
        h = C.c_void_p()
        ctx.check(lib.crux_mlp_create(ctx.h, 3, (C.c_int32 * 4)(*dims), (C.c_int32 * 3)(1, 1, 0), C.byref(h)))
        np_ = C.c_int64()
        ctx.check(lib.crux_mlp_num_params(h, C.byref(np_)))
        flat = _fill(np_.value, phase, 0.2)
        ctx.check(lib.crux_mlp_set_params(h, flat.ctypes.data_as(C.c_void_p)))
        ctx.check(lib.crux_mlp_set_adam(h, 3e-4, 0.9, 0.999, 1e-8))
        return h, np_.value
    mu, pa = mlp([17, 64, 64, 6], 0.1)
    V, pc = mlp([17, 64, 64, 1], 1.3)
    ls = np.full(6, -0.5, F32)
    pi = C.c_void_p()
    ctx.check(lib.crux_gaussian_create(ctx.h, mu, 6, ls.ctypes.data_as(C.c_void_p), 0, 1.0, C.byref(pi)))
    A = _fill(17 * 17, 0.5, 0.02).reshape(17, 17)
    A[np.arange(17), np.arange(17)] += F32(0.95)
    B = _fill(17 * 6, 2.1, 0.1)
    A = np.ascontiguousarray(A, dtype=F32)
    env = C.c_void_p()
    ctx.check(lib.crux_linquad_create(ctx.h, 17, 6, A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p), N, 1000, seed + 7, C.byref(env)))
    dev = lambda shape, dt=torch.float32: torch.zeros(shape, dtype=dt, device=ctx.device)
    s, a, sp, r = dev((n, 17)), dev((n, 6)), dev((n, 17)), dev((n,))
    done, ee, logp, adv, ret = dev((n,), torch.uint8), dev((n,), torch.uint8), dev((n,)), dev((n,)), dev((n,))
    obs, v_s, v_sp = dev((N, 17)), dev((n,)), dev((n,))
    cols = _abi.RolloutCols(s.data_ptr(), a.data_ptr(), sp.data_ptr(), r.data_ptr(), done.data_ptr(), ee.data_ptr(), logp.data_ptr())
    ctx.check(lib.crux_linquad_reset(env, ptr(obs)))
    ctx.check(lib.crux_linquad_rollout(env, pi, T, 1, ptr(obs), C.byref(cols), seed, 0))
    ctx.check(lib.crux_mlp_forward(V, ptr(s), n, ptr(v_s)))
    ctx.check(lib.crux_value_next(V, ptr(sp), ptr(s), ptr(v_s), T, N, ptr(v_sp)))
    ctx.check(lib.crux_fill_gae_returns(ctx.h, ptr(r), ptr(done), ptr(ee), ptr(v_s), ptr(v_sp), T, N, 0.99, 0.95, ptr(adv), ptr(ret)))
    ctx.sync()
    for k, x in (("sum_s", s), ("sum_a", a), ("sum_r", r), ("sum_logp", logp), ("sum_adv", adv), ("sum_ret", ret)):
        out[k] = _cs(x.cpu().numpy())
    e = ee.cpu().numpy().astype(np.float64)
    out["sum_ee"] = float((e * ((np.arange(e.size) % 5) + 1)).sum())
    ctx.check(lib.crux_whiten(ctx.h, ptr(adv), n))
    hp = _abi.PPOHp(eps_clip=0.2, lambda_p=1.0, lambda_e=0.1, target_kl=math.inf, a2c=0, actor_epochs=2, actor_batch=n // 2, critic_epochs=2,
                    critic_batch=n // 2, actor_max_batches=0, critic_max_batches=0)
    order = np.stack([(np.arange(n, dtype=np.int64) * 7 + e_) % n for e_ in range(2)]).astype(np.int32)
    od = ctx.to_device(order, torch.int32)
    ia, ic = np.zeros((4, 8), F32), np.zeros((4, 8), F32)
    ctx.check(lib.crux_ppo_update(pi, V, ptr(s), ptr(a), ptr(logp), ptr(adv), ptr(ret), n, C.byref(hp), ptr(od), ptr(od), seed,
                                  ia.ctypes.data_as(C.c_void_p), ic.ctypes.data_as(C.c_void_p)))
    out["ia"], out["ic"] = ia, ic
    for name, h, npar in (("sum_actor_params", mu, pa), ("sum_critic_params", V, pc)):
        flat = np.zeros(npar, F32)
        ctx.check(lib.crux_mlp_get_params(h, flat.ctypes.data_as(C.c_void_p)))
        out[name] = _cs(flat)
    lib.crux_linquad_destroy(env); lib.crux_gaussian_destroy(pi); lib.crux_mlp_destroy(mu); lib.crux_mlp_destroy(V)
    return out


@pytest.mark.parametrize("N,T", [(256, 8), (1000, 5)])
def test_c_client_matches_ctypes_twin(crux, ctx, N, T):
    import __graft_entry__ as g
    exe = g.build_abi_smoke()
    seed = 11
    res = subprocess.run([exe, "run", str(N), str(T), str(seed)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got, ia, ic = {}, np.zeros((4, 8)), np.zeros((4, 8))
    for line in res.stdout.strip().splitlines():
        tok = line.split()
        if tok[0].startswith("actor_mb"):
            m = int(tok[0][8:])
            ia[m, [0, 1, 2, 3, 4, 7]] = [float(tok[i]) for i in (2, 4, 6, 8, 10, 12)]
        elif tok[0].startswith("critic_mb"):
            m = int(tok[0][9:])
            ic[m, [0, 1, 7]] = [float(tok[i]) for i in (2, 4, 6)]
        elif tok[0] == "buffer":
            got["buffer"] = [int(tok[i]) for i in (2, 4, 6, 8)]
        else:
            got[tok[0]] = float(tok[1])
    n = N * T
    assert got["buffer"] == [n, 0, n, n]        # elements, next_ind (0-based, wrapped: capacity == ΔN), total_count, capacity
    assert got["launches"] > 10
    ref = _python_twin(crux, ctx, N, T, seed)
    for k in ("sum_s", "sum_a", "sum_r", "sum_logp", "sum_ee", "sum_adv", "sum_ret", "sum_actor_params", "sum_critic_params"):
        assert got[k] == pytest.approx(ref[k], rel=1e-6, abs=1e-6), k
    assert (ia[:, 7] == 1).all() and (ic[:, 7] == 1).all()
    np.testing.assert_allclose(ia[:, [0, 1, 2, 3, 4]], ref["ia"][:, [0, 1, 2, 3, 4]], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(ic[:, [0, 1]], ref["ic"][:, [0, 1]], rtol=1e-5, atol=1e-7)
