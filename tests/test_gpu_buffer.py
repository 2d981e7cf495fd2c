"""-m gpu: device-resident ExperienceBuffer (src/experience_buffer.jl) through the C ABI against the oracle buffer.
Integer results (ring indices, sample indices given the same prefix array, episode bookkeeping) are bit-exact."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import c_oracle
from oracle import crux_oracle as o
from gpu_util import F32, assert_close, dev, host, p

pytestmark = pytest.mark.gpu

S, A_, SP, R, DONE, EE, W = 0, 1, 2, 3, 4, 5, 9


def _mk(ctx, crux, cap, sdim=2, adim=4, prioritized=False, sdtype=None, alpha=0.6, weight=False):
    A = crux._abi
    sd = A.F32 if sdtype is None else sdtype
    cols = [(S, sd, sdim, 0.0), (A_, A.F32, adim, 0.0), (SP, sd, sdim, 0.0), (R, A.F32, 1, 0.0), (DONE, A.U8, 1, 0.0),
            (EE, A.U8, 1, 0.0)]
    if weight or prioritized:
        cols.append((W, A.F32, 1, 1.0))
    arr = (A.ColDesc * len(cols))(*[A.ColDesc(*c) for c in cols])
    h = C.c_void_p()
    ctx.check(ctx.lib.crux_buffer_create(ctx.h, cap, len(cols), arr, 1 if prioritized else 0, alpha, C.byref(h)))
    return h


def _state(ctx, h):
    v = [C.c_int64() for _ in range(4)]
    ctx.check(ctx.lib.crux_buffer_state(h, *[C.byref(x) for x in v]))
    return tuple(x.value for x in v)  # elements, next_ind(0-based), total_count, capacity


def _col(ctx, crux, h, cid, rows=None):
    ptr, rl, dt = C.c_void_p(), C.c_int64(), C.c_int32()
    ctx.check(ctx.lib.crux_buffer_col(h, cid, C.byref(ptr), C.byref(rl), C.byref(dt)))
    el, _, _, cap = _state(ctx, h)
    rows = el if rows is None else rows
    npdt = {0: np.uint8, 1: np.float32, 2: np.int32, 3: np.int64}[dt.value]
    out = np.empty((rows, rl.value), dtype=npdt)
    if out.size:
        ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(out), ptr, out.nbytes))
        ctx.sync()
    return out


def _push(ctx, h, d, on_host=True, ids=None):
    keys = list(d)
    n = len(next(iter(d.values())))
    arrs = [np.ascontiguousarray(d[k]) for k in keys]
    keep = arrs if on_host else [dev(ctx, a) for a in arrs]
    ptrs = (C.c_void_p * len(keys))(*[a.ctypes.data if on_host else a.data_ptr() for a in keep])
    idp = None
    if ids is not None:
        ids = np.ascontiguousarray(ids, dtype=np.int32)
        idk = ids if on_host else dev(ctx, ids)
        idp = p(idk)
        n = len(ids)
    first = C.c_int64(-1)
    ctx.check(ctx.lib.crux_buffer_push(h, n, len(keys), (C.c_int32 * len(keys))(*keys), ptrs, 1 if on_host else 0, idp, C.byref(first)))
    ctx.sync()
    return first.value


def _rows(n, sdim=2, adim=4, base=0.0, rng=None):
    rng = rng or np.random.default_rng(int(base) + n)
    return {S: rng.standard_normal((n, sdim)).astype(F32) + F32(base), A_: (rng.random((n, adim)) < 0.5).astype(F32),
            SP: rng.standard_normal((n, sdim)).astype(F32), R: rng.standard_normal((n, 1)).astype(F32),
            DONE: (rng.random((n, 1)) < 0.5).astype(np.uint8), EE: np.zeros((n, 1), np.uint8)}


def _names(d):
    m = {S: "s", A_: "a", SP: "sp", R: "r", DONE: "done", EE: "episode_end", W: "weight"}
    return {m[k]: (v.astype(bool) if v.dtype == np.uint8 and k in (DONE, EE) else v) for k, v in d.items()}


@pytest.mark.parametrize("on_host", [True, False])
def test_push_ring_semantics(ctx, crux, on_host):
    """push! (experience_buffer.jl:232-259): ring indices, wrap, N > capacity (later rows win), bookkeeping."""
    cap = 10
    h = _mk(ctx, crux, cap)
    ref = o.ExperienceBuffer.create(2, 4, cap)
    for n in (1, 3, 5, 4, 10, 23, 7):
        d = _rows(n, base=n)
        first = _push(ctx, h, d, on_host)
        I = ref.push(_names(d))
        assert first == I[0] - 1
        el, nxt, tot, c = _state(ctx, h)
        assert (el, nxt + 1, tot, c) == (ref.elements, ref.next_ind, ref.total_count, cap)
        for cid, name in ((S, "s"), (A_, "a"), (SP, "sp"), (R, "r"), (DONE, "done")):
            got = _col(ctx, crux, h, cid)
            assert np.array_equal(got, ref[name].astype(got.dtype)), f"column {name} after pushing {n}"
    # ids gather (push!(b, data, ids=...)), 0-based across the ABI
    d = _rows(6, base=100)
    ids = np.array([5, 0, 0, 3], np.int32)
    _push(ctx, h, d, on_host, ids=ids)
    ref.push(_names(d), ids=ids + 1)
    assert np.array_equal(_col(ctx, crux, h, S), ref["s"])
    ctx.check(ctx.lib.crux_buffer_clear(h))
    assert _state(ctx, h)[:3] == (0, 0, 0)  # clear! :97-104
    ctx.lib.crux_buffer_destroy(h)


def test_reference_push_kats(ctx, crux):
    # test/experience_buffer_tests.jl:121-147 (push 1, push 3, push a buffer onto itself)
    h = _mk(ctx, crux, 100)
    d1 = {S: 2 * np.ones((1, 2), F32), A_: np.ones((1, 4), F32), SP: np.ones((1, 2), F32), R: np.ones((1, 1), F32), DONE: np.zeros((1, 1), np.uint8)}
    _push(ctx, h, d1)
    assert _state(ctx, h)[0] == 1 and np.all(_col(ctx, crux, h, S) == 2) and np.all(_col(ctx, crux, h, R) == 1)
    d3 = {S: 3 * np.ones((3, 2), F32), A_: (np.random.default_rng(0).random((3, 4)) < 0.5).astype(F32), SP: 5 * np.ones((3, 2), F32),
          R: 6 * np.ones((3, 1), F32), DONE: np.ones((3, 1), np.uint8)}
    _push(ctx, h, d3)
    assert _state(ctx, h)[0] == 4
    assert np.all(_col(ctx, crux, h, S)[1:] == 3) and np.array_equal(_col(ctx, crux, h, A_)[1:], d3[A_]) and np.all(_col(ctx, crux, h, DONE)[1:] == 1)
    ctx.check(ctx.lib.crux_buffer_push_from(h, h, 4, None))  # push!(b, b)
    assert _state(ctx, h)[0] == 8
    for cid in (S, A_, SP, R, DONE):
        c = _col(ctx, crux, h, cid)
        assert np.array_equal(c[:4], c[4:8])
    ctx.lib.crux_buffer_destroy(h)


def test_last_n_indices(ctx, crux):
    # test/experience_buffer_tests.jl:32-51
    h = _mk(ctx, crux, 100, 2, 1)
    d = _rows(50, 2, 1)
    _push(ctx, h, d)

    def last(N):
        out = np.empty(200, np.int64); n = C.c_int64()
        ctx.check(ctx.lib.crux_buffer_last_n_indices(h, N, p(out), C.byref(n)))
        return (out[:n.value] + 1).tolist()
    assert last(10) == list(range(41, 51)) and last(1) == [50] and last(50) == list(range(1, 51))
    assert last(51) == list(range(1, 51)) and last(1000) == list(range(1, 51))
    _push(ctx, h, d); _push(ctx, h, d)
    assert last(10) == list(range(41, 51)) and last(1) == [50]
    assert last(51) == [100] + list(range(1, 51))
    assert last(100) == list(range(51, 101)) + list(range(1, 51)) == last(1000)
    ctx.lib.crux_buffer_destroy(h)


def test_u8_nd_states(ctx, crux):
    # test/experience_buffer_tests.jl:271-278 ContinuousSpace((2,2), UInt8); C3-shaped rows (84*84*4 u8)
    for rowlen in (4, 84 * 84 * 4):
        h = _mk(ctx, crux, 16, sdim=rowlen, adim=4, sdtype=crux._abi.U8)
        rng = np.random.default_rng(rowlen)
        d = {S: rng.integers(0, 256, (5, rowlen), dtype=np.uint8), SP: rng.integers(0, 256, (5, rowlen), dtype=np.uint8),
             A_: np.ones((5, 4), F32), R: np.ones((5, 1), F32), DONE: np.zeros((5, 1), np.uint8)}
        _push(ctx, h, d, on_host=False)
        assert np.array_equal(_col(ctx, crux, h, S), d[S]) and np.array_equal(_col(ctx, crux, h, SP), d[SP])
        _push(ctx, h, d, on_host=True, ids=np.array([4, 4, 1], np.int32))
        assert np.array_equal(_col(ctx, crux, h, S)[5:], d[S][[4, 4, 1]])
        ctx.lib.crux_buffer_destroy(h)


def _prs(ctx, h, n, want_cumsum=False):
    pp, cp, mx, mn = C.c_void_p(), C.c_void_p(), C.c_float(), C.c_float()
    ctx.check(ctx.lib.crux_buffer_priorities(h, C.byref(pp), C.byref(cp) if want_cumsum else None, C.byref(mx), C.byref(mn)))
    out = np.empty(n, F32)
    ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(out), pp, out.nbytes)); ctx.sync()
    cs = None
    if want_cumsum:
        cs = np.empty(n, F32)
        ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(cs), cp, cs.nbytes)); ctx.sync()
    return out, cs, mx.value, mn.value


def test_update_priorities_kat(ctx, crux):
    # test/experience_buffer_tests.jl:193-205
    h = _mk(ctx, crux, 50, prioritized=True)
    ref = o.ExperienceBuffer.create(2, 4, 50, prioritized=True)
    idx = dev(ctx, np.array([0, 1, 2], np.int32)); v = dev(ctx, np.array([1, 2, 3], F32))
    ctx.check(ctx.lib.crux_buffer_update_priorities(h, p(idx), p(v), 3))
    ref.update_priorities(np.array([1, 2, 3]), np.array([1, 2, 3], F32))
    prs, _, mx, mn = _prs(ctx, h, 50)
    assert np.array_equal(prs, ref.pp.priorities)  # bit-exact: Float32^Float32 via Float64 pow
    assert_close(prs[:3], [1.0000001, 1.5157167, 1.9331821], rtol=1e-7)
    assert mx == ref.pp.max_priority == F32(3.0) + o.EPS32 and mn == ref.pp.min_priority
    d = _rows(3)
    _push(ctx, h, d); _push(ctx, h, d)
    ref.push(_names(d)); ref.push(_names(d))
    prs, _, mx, mn = _prs(ctx, h, 50)
    assert np.array_equal(prs, ref.pp.priorities) and mx == ref.pp.max_priority and mn == ref.pp.min_priority
    for i in range(6):
        assert np.isclose(prs[i], F32(3.0) ** F32(0.6))
    # duplicate indices: the last write wins (sequential loop semantics, experience_buffer.jl:292-299)
    idx = dev(ctx, np.array([4, 4, 4, 7], np.int32)); v = dev(ctx, np.array([9, 8, 0.5, 2], F32))
    ctx.check(ctx.lib.crux_buffer_update_priorities(h, p(idx), p(v), 4))
    ref.update_priorities(np.array([5, 5, 5, 8]), np.array([9, 8, 0.5, 2], F32))
    prs, _, mx, mn = _prs(ctx, h, 50)
    assert np.array_equal(prs, ref.pp.priorities) and mx == ref.pp.max_priority and mn == ref.pp.min_priority
    ctx.lib.crux_buffer_destroy(h)


@pytest.mark.parametrize("N,B", [(6, 1000), (1000, 32), (5000, 512), (100000, 512)])
def test_prioritized_sample(ctx, crux, N, B):
    """prioritized_sample! (experience_buffer.jl:324-349): indices bit-exact GIVEN the device prefix array, IS weights,
    gathered rows, frequency ∝ priority (test/experience_buffer_tests.jl:247-262)."""
    rng = np.random.default_rng(N)
    src = _mk(ctx, crux, N + 10, prioritized=True)
    ref = o.ExperienceBuffer.create(2, 4, N + 10, prioritized=True)
    d = _rows(N, rng=rng)
    _push(ctx, src, d, on_host=False); ref.push(_names(d))
    td = (np.abs(rng.standard_normal(N)) + 1e-3).astype(F32) if N > 6 else np.arange(1, 7, dtype=F32)
    ctx.check(ctx.lib.crux_buffer_update_priorities(src, p(dev(ctx, np.arange(N, dtype=np.int32))), p(dev(ctx, td)), N))
    ref.update_priorities(np.arange(1, N + 1), td)
    tgt = _mk(ctx, crux, B, weight=True)
    reft = o.ExperienceBuffer.create(2, 4, B, extras=("weight",))
    u = rng.random(B)
    ctx.check(ctx.lib.crux_buffer_sample_prioritized(tgt, src, B, 0.5, W, p(u), 0, 0)); ctx.sync()
    prs, cs, mx, mn = _prs(ctx, src, N, want_cumsum=True)
    assert np.array_equal(prs, ref.pp.priorities[:N])
    # the device scan vs a float64 reference prefix: float32 summation-order differences only
    assert_close(cs, np.cumsum(prs.astype(np.float64)), rtol=2e-5, what="prefix sums")
    assert np.all(np.diff(cs.astype(np.float64)) >= 0)
    ip, n = C.c_void_p(), C.c_int64()
    ctx.check(ctx.lib.crux_buffer_indices(tgt, C.byref(ip), C.byref(n)))
    ids = np.empty(B, np.int32); ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(ids), ip, ids.nbytes)); ctx.sync()
    assert n.value == B
    want = np.minimum(c_oracle.per_indices(cs, B, u), N - 1)  # same prefix array -> identical indices
    assert np.array_equal(ids, want)
    o.prioritized_sample(reft, ref, u, B=B, cumsum=cs)
    assert np.array_equal(np.minimum(reft.indices - 1, N - 1), ids)
    assert np.array_equal(_col(ctx, crux, tgt, S), d[S][ids])
    w_src = _col(ctx, crux, src, W)[:, 0]
    assert_close(w_src[ids], ref["weight"][ids, 0], rtol=1e-5)
    assert np.all(_col(ctx, crux, tgt, W) <= 1.0 + 1e-6)
    if N == 6:
        freqs = np.bincount(ids, minlength=6) / B
        probs = prs / prs.sum()
        assert np.all(np.abs(freqs - probs) / probs < 0.01)
    # device RNG path: strata are respected (one sample per stratum of width ptot/B)
    ctx.check(ctx.lib.crux_buffer_sample_prioritized(tgt, src, B, 0.5, W, None, 5, 1)); ctx.sync()
    ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(ids), ip, ids.nbytes)); ctx.sync()
    assert np.all(np.diff(ids) >= 0) and ids.min() >= 0 and ids.max() < N
    ctx.lib.crux_buffer_destroy(tgt); ctx.lib.crux_buffer_destroy(src)


def test_uniform_sample(ctx, crux):
    # uniform_sample! (experience_buffer.jl:317-321); reference test :214-222 with the draw supplied
    src = _mk(ctx, crux, 100); d = _rows(37); _push(ctx, src, d)
    tgt = _mk(ctx, crux, 3)
    ids = np.array([36, 0, 11], np.int32)
    ctx.check(ctx.lib.crux_buffer_sample_uniform(tgt, src, 3, p(ids), 0, 0)); ctx.sync()
    for cid in (S, A_, SP, R, DONE):
        assert np.array_equal(_col(ctx, crux, tgt, cid), d[cid][ids])
    bad = np.array([37, 0, 1], np.int32)
    assert ctx.lib.crux_buffer_sample_uniform(tgt, src, 3, p(bad), 0, 0) == 1
    big = _mk(ctx, crux, 20000)
    ctx.check(ctx.lib.crux_buffer_sample_uniform(big, src, 20000, None, 9, 0)); ctx.sync()
    ip, n = C.c_void_p(), C.c_int64()
    ctx.check(ctx.lib.crux_buffer_indices(big, C.byref(ip), C.byref(n)))
    got = np.empty(20000, np.int32); ctx.check(ctx.lib.crux_memcpy_d2h(ctx.h, p(got), ip, got.nbytes)); ctx.sync()
    assert got.min() == 0 and got.max() == 36
    assert np.allclose(np.bincount(got, minlength=37) / 20000, 1 / 37, atol=6e-3)
    assert np.array_equal(_col(ctx, crux, big, R), d[R][got])
    for h in (src, tgt, big):
        ctx.lib.crux_buffer_destroy(h)


def test_gather_rows_and_errors(ctx, crux):
    rng = np.random.default_rng(0)
    for rowbytes in (4, 68, 16, 3, 28224):
        src = rng.integers(0, 256, (50, rowbytes), dtype=np.uint8)
        idx = rng.integers(0, 50, 200).astype(np.int32)
        dst = torch.empty((200, rowbytes), dtype=torch.uint8, device=ctx.device)
        ctx.check(ctx.lib.crux_gather_rows(ctx.h, p(dst), p(dev(ctx, src)), p(dev(ctx, idx)), 200, rowbytes))
        assert np.array_equal(host(dst), src[idx])
    h = _mk(ctx, crux, 4)
    assert ctx.lib.crux_buffer_col(h, 33, None, None, None) == 1  # KeyError
    assert ctx.lib.crux_buffer_update_priorities(h, None, None, 1) == 1  # not prioritized
    ctx.lib.crux_buffer_destroy(h)
