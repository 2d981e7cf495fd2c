"""The persistent small-minibatch epoch kernel (csrc/mb_persist.cuh): batch_train! (training.jl:28-55) of one network as ONE cluster launch
for the reference-default batch_size = 128 (training.jl:5).  Checked against the oracle's step-by-step train! sequence with injected
shuffles -- records of every minibatch, parameters, logΣ, early stop, max_batches, ragged last minibatch, a2c head -- and against the
step-by-step device path (CRUX_NO_PERSIST=1), whose Adam state must continue seamlessly."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, assert_params_close, mlp_params, p
from test_gpu_ppo import _hp, _oracle_train, _orders, _run, _setup

pytestmark = pytest.mark.gpu


def _log_sigma(ctx, h):
    lsp = C.c_void_p(); ctx.check(ctx.lib.crux_gaussian_log_sigma_ptr(h, C.byref(lsp)))
    ls = np.empty(6, F32); ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(ls), lsp, 24)); ctx.sync()
    return ls


def _check_records(crux, ia, ic, ra, rc_, ab):
    A = crux._abi
    for k, rec in enumerate(ra):
        assert ia[k, A.PPO_VALID] == 1.0
        assert_close(ia[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, atol=1e-5, what=f"actor loss mb {k}")
        assert_close(ia[k, A.PPO_KL], rec["kl"], rtol=1e-3, atol=2e-6, what=f"kl mb {k}")
        assert_close(ia[k, A.PPO_ENTROPY], rec["entropy"], rtol=1e-5, what=f"entropy mb {k}")
        if "clip_fraction" in rec:
            assert_close(ia[k, A.PPO_CLIP_FRAC], rec["clip_fraction"], rtol=0, atol=2.0 / ab, what="clip_fraction")
            assert_close(ia[k, A.PPO_AVG_ADV], rec["avg_advantage"], rtol=1e-3, atol=1e-5, what="avg_advantage")
            assert_close(ia[k, A.PPO_AVG_RET], rec["avg_return"], rtol=1e-3, atol=1e-5, what="avg_return")
        assert_close(ia[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what=f"grad_norm mb {k}")
    assert not ia[len(ra):, A.PPO_VALID].any()
    for k, rec in enumerate(rc_):
        assert ic[k, A.PPO_VALID] == 1.0
        assert_close(ic[k, A.PPO_LOSS], rec["loss"], rtol=1e-4, what=f"critic loss mb {k}")
        assert_close(ic[k, A.PPO_GRAD_NORM], rec["grad_norm"], rtol=1e-3, what=f"critic grad_norm mb {k}")


@pytest.mark.parametrize("n,ab,cb,epochs", [(1024, 128, 128, 3), (1000, 128, 64, 2), (520, 96, 128, 4), (4096, 128, 128, 1)])
def test_persistent_epochs_match_oracle(ctx, crux, n, ab, cb, epochs):
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=n + ab)
    hp = _hp(crux, actor_batch=ab, critic_batch=cb, actor_epochs=epochs, critic_epochs=epochs)
    oa = _orders(rng, n, epochs)
    oc = _orders(rng, n, epochs, start=oa[-1])
    P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(0.1)}
    ra = _oracle_train(pi.params(), lambda mb, inf: o.ppo_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab)
    rc_ = _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), o.Adam(F32(3e-4)), D, oc, cb)
    l0 = ctx.launch_count()
    ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
    assert ctx.launch_count() - l0 <= 8, "small minibatches must run as one cluster launch per network, not step by step"
    assert len(ra) == ia.shape[0] and len(rc_) == ic.shape[0]
    _check_records(crux, ia, ic, ra, rc_, ab)
    hm, hc, h = handles
    assert_params_close(mlp_params(ctx, hm), pi.mu.flat(), 3e-4, len(ra), what="actor params")
    assert_params_close(mlp_params(ctx, hc), cr.flat(), 3e-4, len(rc_), what="critic params")
    assert_close(_log_sigma(ctx, h), pi.log_sigma.detach().numpy(), rtol=1e-5, atol=2e-6, what="logΣ")


def test_persistent_early_stop_max_batches_and_a2c(ctx, crux):
    n, ab = 1024, 128
    # KL early stop (rl/ppo.jl:59): the minibatch whose KL exceeds the threshold is applied, later ones are not
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=3)
    D["advantage"] = (D["advantage"] * F32(30)).astype(F32)
    hp = _hp(crux, actor_batch=ab, actor_epochs=6, critic_epochs=0, target_kl=2e-3)
    oa = _orders(rng, n, hp.actor_epochs)
    P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(0.1)}
    ra = _oracle_train(pi.params(), lambda mb, inf: o.ppo_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab, stop=lambda info: info["kl"] > 2e-3)
    ia, _ = _run(ctx, crux, handles, D, hp, oa, None, n)
    valid = ia[:, crux._abi.PPO_VALID]
    assert 0 < len(ra) < ia.shape[0], "the test must actually stop early"
    assert valid[:len(ra)].all() and not valid[len(ra):].any()
    assert_params_close(mlp_params(ctx, handles[0]), pi.mu.flat(), 3e-4, len(ra), what="actor params after early stop")
    # a2c head + max_batches (training.jl:10,44)
    rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=4)
    hp = _hp(crux, actor_batch=ab, critic_batch=ab, actor_epochs=3, critic_epochs=3, a2c=1, actor_max_batches=11, critic_max_batches=9)
    oa = _orders(rng, n, 3); oc = _orders(rng, n, 3, start=oa[-1])
    P = {"lp": F32(1.0), "le": F32(0.1)}
    ra = _oracle_train(pi.params(), lambda mb, inf: o.a2c_loss(pi, P, mb, inf), o.Adam(F32(3e-4)), D, oa, ab, max_batches=11)
    rc_ = _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), o.Adam(F32(3e-4)), D, oc, ab, max_batches=9)
    ia, ic = _run(ctx, crux, handles, D, hp, oa, oc, n)
    assert len(ra) == 11 and len(rc_) == 9
    _check_records(crux, ia, ic, ra, rc_, ab)
    assert not ic[9:, crux._abi.PPO_VALID].any()
    assert_params_close(mlp_params(ctx, handles[0]), pi.mu.flat(), 3e-4, 11, what="actor params (a2c, max_batches)")
    assert_params_close(mlp_params(ctx, handles[1]), cr.flat(), 3e-4, 9, what="critic params (max_batches)")


def test_persistent_path_continues_the_step_by_step_path(ctx, crux):
    """Two updates in a row: the first through the cluster kernel, the second step by step (CRUX_NO_PERSIST=1) -- and the other way round.
    Adam moments, step counters, β-power cache and logΣ must carry over: both orders equal the oracle's uninterrupted sequence."""
    n, ab = 768, 128
    for first_persistent in (True, False):
        rng, pi, cr, handles, D = _setup(ctx, crux, n, seed=11)
        hp = _hp(crux, actor_batch=ab, critic_batch=ab, actor_epochs=2, critic_epochs=2)
        opt_a, opt_c = o.Adam(F32(3e-4)), o.Adam(F32(3e-4))
        P = {"eps": F32(0.2), "lp": F32(1.0), "le": F32(0.1)}
        order = None
        for phase in range(2):
            oa = _orders(rng, n, 2, start=order); oc = _orders(rng, n, 2, start=oa[-1]); order = oc[-1]
            _oracle_train(pi.params(), lambda mb, inf: o.ppo_loss(pi, P, mb, inf), opt_a, D, oa, ab)
            _oracle_train(cr.params(), lambda mb, inf: o.value_mse_loss(cr, mb), opt_c, D, oc, ab)
            persistent = first_persistent == (phase == 0)
            if not persistent:
                os.environ["CRUX_NO_PERSIST"] = "1"
            try:
                l0 = ctx.launch_count()
                _run(ctx, crux, handles, D, hp, oa, oc, n)
                assert (ctx.launch_count() - l0 <= 8) == persistent
            finally:
                os.environ.pop("CRUX_NO_PERSIST", None)
        assert_params_close(mlp_params(ctx, handles[0]), pi.mu.flat(), 3e-4, 24, what=f"actor params (persistent first: {first_persistent})")
        assert_params_close(mlp_params(ctx, handles[1]), cr.flat(), 3e-4, 24, what=f"critic params (persistent first: {first_persistent})")
        assert_close(_log_sigma(ctx, handles[2]), pi.log_sigma.detach().numpy(), rtol=1e-5, atol=2e-6, what="logΣ")
