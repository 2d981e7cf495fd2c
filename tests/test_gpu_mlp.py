"""-m gpu: Dense/Chain engine (value(π, s), value(π, s, a), train! with Flux.mse, Adam, polyak) against the oracle."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import crux_oracle as o
from gpu_util import F32, assert_close, assert_params_close, dev, host, make_mlp, mlp_grads, mlp_params, p

pytestmark = pytest.mark.gpu

NETS = [([17, 64, 64, 6], [1, 1, 0]), ([17, 64, 64, 1], [1, 1, 0]), ([2, 8, 4], [2, 0]), ([393, 256, 256, 1], [2, 2, 0]),
        ([5, 3], [0]), ([376, 256, 256, 34], [2, 2, 0])]


@pytest.mark.parametrize("dims,acts", NETS)
@pytest.mark.parametrize("B", [1, 7, 128, 4096, 20000])
def test_forward(ctx, dims, acts, B):
    rng = np.random.default_rng(B + dims[0])
    ref = o.MLP(dims, acts, rng)
    h = make_mlp(ctx, dims, acts, ref.flat())
    x = rng.standard_normal((B, dims[0])).astype(F32)
    y = ctx.empty((B, dims[-1]))
    ctx.check(ctx.lib.crux_mlp_forward(h, p(dev(ctx, x)), B, p(y)))
    assert_close(host(y), ref(x).detach().numpy(), rtol=1e-5, atol=1e-5, what=f"forward {dims} B={B}")
    ctx.lib.crux_mlp_destroy(h)


def test_param_layout_roundtrip(ctx):
    ref = o.MLP([3, 4, 2], [1, 0], np.random.default_rng(0))
    flat = ref.flat()
    h = make_mlp(ctx, [3, 4, 2], [1, 0], flat)
    assert np.array_equal(mlp_params(ctx, h), flat)
    n = C.c_int64(); ctx.check(ctx.lib.crux_mlp_num_params(h, C.byref(n)))
    assert n.value == 3 * 4 + 4 + 4 * 2 + 2
    # W[out,in] column-major == [in][out] row-major: first `out` floats are W[:, 0]
    assert np.array_equal(flat[:4], ref.W[0].detach().numpy()[:, 0])
    ctx.lib.crux_mlp_destroy(h)


def test_forward_sa(ctx):
    rng = np.random.default_rng(3)
    ref = o.MLP([23, 32, 1], [2, 0], rng)
    h = make_mlp(ctx, [23, 32, 1], [2, 0], ref.flat())
    s, a = rng.standard_normal((100, 17)).astype(F32), rng.standard_normal((100, 6)).astype(F32)
    y = ctx.empty((100, 1))
    ctx.check(ctx.lib.crux_mlp_forward_sa(h, p(dev(ctx, s)), 17, p(dev(ctx, a)), 6, 100, p(y)))
    assert_close(host(y), ref(np.concatenate([s, a], 1)).detach().numpy(), rtol=1e-5, atol=1e-5)  # policies.jl:96 vcat(s, a)
    assert ctx.lib.crux_mlp_forward_sa(h, p(dev(ctx, s)), 17, p(dev(ctx, a)), 5, 100, p(y)) == 1
    ctx.lib.crux_mlp_destroy(h)


@pytest.mark.parametrize("dims,acts,B", [([17, 64, 64, 1], [1, 1, 0], 512), ([2, 8, 4], [2, 0], 33), ([40, 256, 256, 3], [2, 1, 1], 300),
                                         ([17, 64, 64, 1], [1, 1, 0], 32768)])
def test_train_mse_steps(ctx, dims, acts, B):
    """train! (training.jl:15-25) with Flux.mse + Flux Adam: gradients, grad-norm, loss and 3 updates."""
    rng = np.random.default_rng(B)
    ref = o.MLP(dims, acts, rng)
    h = make_mlp(ctx, dims, acts, ref.flat())
    ctx.check(ctx.lib.crux_mlp_set_adam(h, float(F32(3e-4)), 0.9, 0.999, 1e-8))
    opt = o.Adam(F32(3e-4))
    x = rng.standard_normal((B, dims[0])).astype(F32)
    y = rng.standard_normal((B, dims[-1])).astype(F32)
    dx, dy = dev(ctx, x), dev(ctx, y)
    for step in range(3):
        info = {}
        o.train_step(ref.params(), lambda inf: torch.mean((ref(x) - torch.as_tensor(y)) ** 2), opt, info)
        out = np.zeros(2, F32)
        ctx.check(ctx.lib.crux_mlp_train_mse(h, p(dx), p(dy), B, p(out)))
        if step == 0:
            assert_close(mlp_grads(ctx, h), o.flat_grads(ref.params()), rtol=2e-4, atol=1e-7, what="gradient")
        assert_close(out[0], info["loss"], rtol=1e-5, what="loss")
        assert_close(out[1], info["grad_norm"], rtol=1e-4, what="grad_norm")
        assert_params_close(mlp_params(ctx, h), ref.flat(), 3e-4, step + 1, what=f"params after step {step + 1}")
    ctx.lib.crux_mlp_destroy(h)


@pytest.mark.parametrize("mode", ["1", "0"])
@pytest.mark.parametrize("dims,acts,B", [([40, 256, 256, 3], [2, 1, 1], 300), ([393, 256, 256, 34], [1, 1, 0], 2048), ([130, 64, 17], [1, 0], 129),
                                         ([376, 256, 256, 1], [1, 1, 0], 2048), ([100, 132, 36], [1, 0], 515)])
def test_train_mse_steps_tcgen05_gemm(ctx, dims, acts, B, mode, monkeypatch):
    """The same train! steps with the generic engine's GEMMs on tcgen05 (csrc/gemm_tc5.cu, the default for shapes that fill a 128 x 64 tile:
    forward, data-gradient and split-K weight-gradient GEMMs with TMEM accumulators and the A operand in tensor memory, 3xTF32) and,
    CRUX_GEMM_TC5=0, on the FFMA tile kernel: gradients against autograd, loss, grad-norm, Adam updates.  The shapes cover aligned / unaligned
    rows (393, 130), ragged tiles, the 1-output head (skinny forward kernel) and widths that are not multiples of the tile.  (tanh in the deep
    case: the tensor-core accumulation carries ~5e-6 relative error against 2e-7 for FFMA -- the accumulator truncates on every one of the
    3 K / 8 adds -- which is enough to flip a relu unit sitting at zero and move a gradient column by 1e-3; measured, profiles/r2_notes.md.)"""
    monkeypatch.setenv("CRUX_GEMM_TC5", mode)
    l0 = ctx.launch_count()
    test_train_mse_steps(ctx, dims, acts, B)
    assert ctx.launch_count() > l0


def test_polyak_and_copy(ctx):
    rng = np.random.default_rng(0)
    a, b = o.MLP([4, 8, 2], [1, 0], rng), o.MLP([4, 8, 2], [1, 0], rng)
    ha, hb = make_mlp(ctx, [4, 8, 2], [1, 0], a.flat()), make_mlp(ctx, [4, 8, 2], [1, 0], b.flat())
    ctx.check(ctx.lib.crux_mlp_polyak(ha, hb, F32(0.005)))
    want = F32(0.005) * b.flat() + (F32(1) - F32(0.005)) * a.flat()  # policies.jl:54
    assert_close(mlp_params(ctx, ha), want, rtol=1e-7, atol=1e-9)
    ctx.check(ctx.lib.crux_mlp_polyak(ha, hb, F32(1.0)))  # test/policy_tests.jl:35-37: τ=1 copies
    assert np.array_equal(mlp_params(ctx, ha), b.flat())
    ctx.check(ctx.lib.crux_mlp_set_params(ha, p(a.flat())))
    ctx.check(ctx.lib.crux_mlp_copy(ha, hb))
    assert np.array_equal(mlp_params(ctx, ha), b.flat())
    hc = make_mlp(ctx, [4, 7, 2], [1, 0])
    assert ctx.lib.crux_mlp_polyak(ha, hc, F32(0.5)) == 1
    for h in (ha, hb, hc):
        ctx.lib.crux_mlp_destroy(h)


def test_nan_gradient_is_an_error(ctx, crux):
    ref = o.MLP([3, 4, 1], [1, 0], np.random.default_rng(0))
    h = make_mlp(ctx, [3, 4, 1], [1, 0], ref.flat())
    x = np.ones((8, 3), F32); y = np.ones((8, 1), F32); y[2] = np.nan
    out = np.zeros(2, F32)
    rc = ctx.lib.crux_mlp_train_mse(h, p(dev(ctx, x)), p(dev(ctx, y)), 8, p(out))
    assert rc == crux._abi.ERR_NAN  # training.jl:20 error("NaN detected!")
    assert np.array_equal(mlp_params(ctx, h), ref.flat())  # no update was applied
    ctx.lib.crux_mlp_destroy(h)


def test_bad_arguments(ctx):
    h = C.c_void_p()
    lib = ctx.lib
    assert lib.crux_mlp_create(ctx.h, 0, (C.c_int32 * 1)(3), (C.c_int32 * 1)(0), C.byref(h)) == 1
    assert lib.crux_mlp_create(ctx.h, 1, (C.c_int32 * 2)(3, 0), (C.c_int32 * 1)(0), C.byref(h)) == 1
    assert lib.crux_mlp_create(ctx.h, 1, (C.c_int32 * 2)(3, 2), (C.c_int32 * 1)(9), C.byref(h)) == 1
    assert b"activation" in lib.crux_last_error(ctx.h)


def test_whole_column_forward_and_value_next(ctx):
    """Tensor-core forward kernel (whole rollout columns) against the oracle, and crux_value_next: V(sp) over a [T][N] rollout reuses
    V(s)[t+1] wherever sp[t] == s[t+1] bit for bit and evaluates the network elsewhere (resets, last step, perturbed rows)."""
    import os
    rng = np.random.default_rng(5)
    T, N, I = 40, 512, 17   # 20480 rows >= 2 x 148 tiles of 64: the tensor-core path
    ref = o.MLP([I, 64, 64, 1], [1, 1, 0], rng)
    h = make_mlp(ctx, ref.dims, ref.acts, ref.flat())
    s = rng.standard_normal((T, N, I)).astype(F32)
    sp = np.empty_like(s)
    sp[:-1] = s[1:]
    sp[-1] = rng.standard_normal((N, I)).astype(F32)
    reset = rng.random((T, N)) < 0.01             # transitions followed by a reset: sp differs from the next s
    sp[reset] = rng.standard_normal((int(reset.sum()), I)).astype(F32)
    sp[7, 100, 3] = np.nextafter(sp[7, 100, 3], F32(10))   # a one-ulp difference must be noticed
    B = T * N
    sd, spd = dev(ctx, s.reshape(B, I)), dev(ctx, sp.reshape(B, I))
    vs, vsp, vsp_plain = ctx.empty((B, 1)), ctx.empty((B, 1)), ctx.empty((B, 1))
    ctx.check(ctx.lib.crux_mlp_forward(h, p(sd), B, p(vs)))
    ctx.check(ctx.lib.crux_value_next(h, p(spd), p(sd), p(vs), T, N, p(vsp)))
    ctx.check(ctx.lib.crux_mlp_forward(h, p(spd), B, p(vsp_plain)))
    want_s = ref(s.reshape(B, I)).detach().numpy()
    want_sp = ref(sp.reshape(B, I)).detach().numpy()
    assert_close(host(vs), want_s, rtol=1e-5, atol=2e-6, what="V(s), tensor-core forward")
    assert_close(host(vsp_plain), want_sp, rtol=1e-5, atol=2e-6, what="V(sp), plain forward")
    assert_close(host(vsp), want_sp, rtol=1e-5, atol=2e-6, what="V(sp), value_next")
    # tiles without any reset are copies of V(s)[t+1]: bitwise
    v, vn = host(vs).reshape(T, N), host(vsp).reshape(T, N)
    G = 128   # reuse granularity: one 128-row tile of the tcgen05 kernel (a multiple of the 64-row tile of the mma.sync kernel)
    same_tile = ~(reset | (np.arange(T)[:, None] == T - 1)).reshape(-1, G).any(axis=1)
    same_tile[(7 * N + 100) // G] = False
    rows = np.repeat(same_tile, G).reshape(T, N)
    assert rows.sum() > 0.15 * T * N
    shifted = np.roll(v, -1, axis=0)
    assert np.array_equal(vn[rows], shifted[rows])
    # and the FFMA fallback agrees
    os.environ["CRUX_NO_MMA"] = "1"
    try:
        v2 = ctx.empty((B, 1))
        ctx.check(ctx.lib.crux_value_next(h, p(spd), p(sd), p(vs), T, N, p(v2)))
        assert_close(host(v2), want_sp, rtol=1e-5, atol=2e-6, what="V(sp), FFMA path")
    finally:
        os.environ.pop("CRUX_NO_MMA", None)
    ctx.check(ctx.lib.crux_mlp_destroy(h))
