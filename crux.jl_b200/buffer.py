"""Host-side mirror of ``src/experience_buffer.jl``: a device-resident structure-of-arrays ``ExperienceBuffer``.

Every column lives in HBM (``crux_buffer_*`` in include/crux_cuda.h); ``b["s"]`` is a zero-copy torch view of the
first ``len(b)`` rows that callers may write through (the reference returns views too, ppo.jl:61).  Arrays are
batch-major ``[rows, features]`` = the memory order of the reference's ``[features, rows]``.  Indices exposed
here are **1-based** like the reference (``next_ind``, ``indices``, ``get_last_N_indices``); the C ABI is 0-based.
Julia's ``f!`` is spelled ``f_``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _abi
from .device import default_context, ptr, view
from .spaces import ContinuousSpace, DiscreteSpace

# column name -> id of the C ABI (crux_col_desc.id)
COLUMN_IDS = {k: i for i, k in enumerate(
    ["s", "a", "sp", "r", "done", "episode_end", "return", "logprob", "advantage", "weight", "t", "i", "value", "cost",
     "cost_advantage", "cost_return", "xlogprob", "importance_weight", "fail", "expert", "id", "s0", "x",
     "fwd_importance_weight", "rev_importance_weight", "cum_importance_weight", "traj_importance_weight", "var_prob",
     "cvar_prob", "f", "grasp_success", "value_s", "value_sp"])}
_F32_ZERO = {"return", "logprob", "xlogprob", "advantage", "cost", "cost_advantage", "cost_return", "value", "var_prob",
             "cvar_prob", "f", "value_s", "value_sp"}
_F32_ONE = {"weight", "importance_weight", "fwd_importance_weight", "rev_importance_weight", "cum_importance_weight",
            "traj_importance_weight"}
_BOOL = {"fail", "grasp_success", "expert"}
_INT = {"t", "i", "id"}
_NP = {_abi.U8: np.uint8, _abi.F32: np.float32, _abi.I32: np.int32, _abi.I64: np.int64}
_TORCH = {_abi.U8: torch.uint8, _abi.F32: torch.float32, _abi.I32: torch.int32, _abi.I64: torch.int64}


def _space_desc(S):
    """(dtype code, dims) of a column holding elements of space S.  One-hot actions are stored as float32 0/1
    (the reference stores Bool; the update kernels consume float one-hot rows)."""
    if isinstance(S, DiscreteSpace):
        return _abi.F32, (S.N,)
    t = np.dtype(S.type)
    code = _abi.U8 if t == np.uint8 else _abi.F32
    return code, tuple(S.dims)


def mdp_data(S, A, capacity, extras=()):
    """Column schema of ``mdp_data`` (experience_buffer.jl:4-35): ordered ``{name: (dtype, dims, init)}``."""
    sd, sdims = _space_desc(S)
    ad, adims = _space_desc(A)
    cols = {"s": (sd, sdims, 0.0), "a": (ad, adims, 0.0), "sp": (sd, sdims, 0.0), "r": (_abi.F32, (1,), 0.0),
            "done": (_abi.U8, (1,), 0.0), "episode_end": (_abi.U8, (1,), 0.0)}
    for k in extras:
        if k in cols:
            continue
        if k in _F32_ZERO:
            cols[k] = (_abi.F32, (1,), 0.0)
        elif k in _F32_ONE:
            cols[k] = (_abi.F32, (1,), 1.0)
        elif k in _BOOL:
            cols[k] = (_abi.U8, (1,), 0.0)
        elif k in _INT:
            cols[k] = (_abi.I64, (1,), 0.0)
        elif k == "s0":
            cols[k] = (sd, sdims, 0.0)
        elif k == "x":
            cols[k] = (ad, adims, 0.0)
        else:
            raise KeyError(f"Unrecognized key: {k}")
    return cols


def split_batches(N, fracs):
    """experience_buffer.jl:126-131 (host integer arithmetic in the library)."""
    fracs = np.atleast_1d(np.asarray(fracs, dtype=np.float64))
    out = (C.c_int64 * len(fracs))()
    rc = _abi.load().crux_split_batches(int(N), (C.c_double * len(fracs))(*fracs), len(fracs), out)
    if rc != 0:
        raise AssertionError("sum(fracs) ≈ 1")
    return list(out)


class PriorityParams:
    """experience_buffer.jl:38-50: α, β(i); priorities / max / min live on the device."""

    def __init__(self, alpha=0.6, beta=None):
        self.alpha = np.float32(alpha)
        self.beta = beta if beta is not None else (lambda i: np.float32(0.5))


class ExperienceBuffer:
    """experience_buffer.jl:53-80."""

    def __init__(self, S, A=None, capacity=None, extras=(), prioritized=False, priority_params=None, ctx=None, _schema=None):
        self.ctx = ctx or default_context()
        if _schema is None:
            extras = list(extras)
            if prioritized and "weight" not in extras:
                extras.append("weight")  # :65
            _schema = mdp_data(S, A, capacity, extras)
        self.schema = dict(_schema)
        self._capacity = int(capacity)
        self.priority_params = None
        if prioritized:
            pp = priority_params or {}
            self.priority_params = pp if isinstance(pp, PriorityParams) else PriorityParams(**pp)
        descs = []
        for name, (dt, dims, init) in self.schema.items():
            descs.append(_abi.ColDesc(COLUMN_IDS[name], dt, int(np.prod(dims)) if len(dims) else 1, float(init)))
        arr = (_abi.ColDesc * len(descs))(*descs)
        h = C.c_void_p()
        alpha = float(self.priority_params.alpha) if prioritized else 0.6
        self.ctx.check(self.ctx.lib.crux_buffer_create(self.ctx.h, self._capacity, len(descs), arr, 1 if prioritized else 0,
                                                       alpha, C.byref(h)))
        self.h = h
        self._cols = {}
        for name, (dt, dims, _) in self.schema.items():
            p = C.c_void_p()
            self.ctx.check(self.ctx.lib.crux_buffer_col(self.h, COLUMN_IDS[name], C.byref(p), None, None))
            self._cols[name] = view(p.value, (self._capacity, *dims), dt, self.ctx.device)

    # ---- bookkeeping ---------------------------------------------------------------------------------------
    def _state(self):
        v = [C.c_int64() for _ in range(4)]
        self.ctx.check(self.ctx.lib.crux_buffer_state(self.h, *[C.byref(x) for x in v]))
        return [x.value for x in v]

    def __len__(self):
        return self._state()[0]

    @property
    def elements(self):
        return self._state()[0]

    @property
    def next_ind(self):
        return self._state()[1] + 1

    @property
    def total_count(self):
        return self._state()[2]

    @property
    def capacity(self):
        return self._capacity

    @property
    def device(self):
        return self.ctx.device

    def keys(self):
        return self.schema.keys()

    def __contains__(self, k):
        return k in self.schema

    def __getitem__(self, k):
        """``b[:key]`` = view of the first ``length(b)`` rows (:173)."""
        return self._cols[k][: len(self)]

    def column(self, k):
        """``b.data[:key]``: the whole column (all ``capacity`` rows)."""
        return self._cols[k]

    @property
    def indices(self):
        """1-based ids of the last sample (``target.indices``; only the LAST source's ids, :313,319,338)."""
        p, n = C.c_void_p(), C.c_int64()
        self.ctx.check(self.ctx.lib.crux_buffer_indices(self.h, C.byref(p), C.byref(n)))
        if n.value == 0:
            return np.zeros(0, dtype=np.int64)
        return view(p.value, (n.value,), _abi.I32, self.ctx.device).cpu().numpy().astype(np.int64) + 1

    def indices_dev(self):
        p, n = C.c_void_p(), C.c_int64()
        self.ctx.check(self.ctx.lib.crux_buffer_indices(self.h, C.byref(p), C.byref(n)))
        return view(p.value, (n.value,), _abi.I32, self.ctx.device)

    def clear_(self):
        """``clear!`` :97-104."""
        self.ctx.check(self.ctx.lib.crux_buffer_clear(self.h))
        return self

    def get_last_N_indices(self, N):
        """:223-229 (1-based)."""
        out = np.empty(max(1, min(int(N), self._capacity)), dtype=np.int64)
        n = C.c_int64()
        self.ctx.check(self.ctx.lib.crux_buffer_last_n_indices(self.h, int(min(N, self._capacity)), ptr(out), C.byref(n)))
        return out[: n.value] + 1

    # ---- push! ---------------------------------------------------------------------------------------------
    def push_(self, data, ids=None):
        """``push!(b, data; ids)`` :232-259.  ``data``: dict of arrays (numpy = host, torch cuda = device; all columns of
        one call on the same side) or another ``ExperienceBuffer``.  ``ids`` 1-based.  Returns the 1-based ring indices."""
        lib, ctx = self.ctx.lib, self.ctx
        start = self._state()[1]
        if isinstance(data, ExperienceBuffer):
            n = len(data) if ids is None else len(ids)
            idt = None if ids is None else ctx.to_device(np.asarray(ids, dtype=np.int64) - 1, torch.int32)
            ctx.check(lib.crux_buffer_push_from(self.h, data.h, n, ptr(idt)))
        else:
            keys = [k for k in data if k in self.schema]
            assert keys, "push!: no matching columns"
            first = data[keys[0]]
            on_host = not (isinstance(first, torch.Tensor) and first.is_cuda)
            arrs, n_src = [], None
            for k in keys:
                dt, dims, _ = self.schema[k]
                v = data[k]
                if on_host:
                    v = np.ascontiguousarray(v.cpu().numpy() if isinstance(v, torch.Tensor) else v, dtype=_NP[dt])
                else:
                    v = v.to(_TORCH[dt]).contiguous()
                rows = v.shape[0]
                assert int(np.prod(v.shape[1:])) == int(np.prod(dims)), f"push!: column {k} has row shape {tuple(v.shape[1:])}, expected {dims}"
                n_src = rows if n_src is None else n_src
                assert rows == n_src, "push!: columns disagree on the number of rows"
                arrs.append(v)
            n = n_src if ids is None else len(ids)
            idp = None
            if ids is not None:
                idn = np.ascontiguousarray(np.asarray(ids, dtype=np.int64) - 1, dtype=np.int32)
                assert idn.min() >= 0 and idn.max() < n_src, "push!: ids out of range (BoundsError)"
                idk = idn if on_host else ctx.to_device(idn, torch.int32)
                idp = ptr(idk)
            ptrs = (C.c_void_p * len(keys))(*[a.ctypes.data if on_host else a.data_ptr() for a in arrs])
            cids = (C.c_int32 * len(keys))(*[COLUMN_IDS[k] for k in keys])
            ctx.check(lib.crux_buffer_push(self.h, n, len(keys), cids, ptrs, 1 if on_host else 0, idp, None))
            if on_host:
                ctx.sync()  # pageable host sources must outlive the copy
        return (start + np.arange(n)) % self._capacity + 1

    def commit_rows_(self, n):
        """Rows ``next_ind .. next_ind+n-1`` were written in place through the column views (zero-copy rollouts):
        advance the ring bookkeeping exactly like ``push!`` would (priorities of new rows included)."""
        dummy_ids, dummy_ptrs = (C.c_int32 * 1)(0), (C.c_void_p * 1)(None)
        self.ctx.check(self.ctx.lib.crux_buffer_push(self.h, int(n), 0, dummy_ids, dummy_ptrs, 0, None, None))

    # ---- minibatches / sampling ------------------------------------------------------------------------------
    def minibatch(self, indices):
        """``minibatch(b, indices)`` :170 (1-based) -> dict of gathered device tensors."""
        idx = self.ctx.to_device(np.asarray(indices, dtype=np.int64) - 1, torch.int32)
        out = {}
        for k, (dt, dims, _) in self.schema.items():
            col = self._cols[k]
            dst = torch.empty((len(idx), *dims), dtype=col.dtype, device=self.ctx.device)
            rb = int(np.prod(dims)) * col.element_size()
            self.ctx.check(self.ctx.lib.crux_gather_rows(self.ctx.h, ptr(dst), ptr(col), ptr(idx), len(idx), rb))
            out[k] = dst
        return out

    def shuffle_(self, perm=None):
        """``shuffle!(b)`` :118-124: permutes every column in place (``perm`` 1-based injects ``shuffle(1:length(b))``).  The
        update kernels never need this (an epoch is an index order consumed by their gather); it exists for API parity."""
        n = len(self)
        perm = np.random.permutation(n) + 1 if perm is None else np.asarray(perm, dtype=np.int64)
        assert sorted(perm.tolist()) == list(range(1, n + 1)), "shuffle!: not a permutation"
        mb = self.minibatch(perm)
        for k, v in mb.items():
            self._cols[k][:n].copy_(v)
        return self

    def isprioritized(self):
        return self.priority_params is not None

    def update_priorities_(self, I, v):
        """``update_priorities!(b, I, v)`` :290-301.  ``I`` 1-based (host) or a 0-based int32 device tensor; ``v`` |td|."""
        if isinstance(I, torch.Tensor) and I.is_cuda:
            idx = I.to(torch.int32).contiguous()
        else:
            idx = self.ctx.to_device(np.asarray(I, dtype=np.int64) - 1, torch.int32)
        vv = v if isinstance(v, torch.Tensor) and v.is_cuda else self.ctx.to_device(np.asarray(v, dtype=np.float32))
        vv = vv.to(torch.float32).contiguous().reshape(-1)
        assert len(idx) == len(vv)
        self.ctx.check(self.ctx.lib.crux_buffer_update_priorities(self.h, ptr(idx), ptr(vv), len(idx)))

    def priorities(self):
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_buffer_priorities(self.h, C.byref(p), None, None, None))
        return view(p.value, (self._capacity,), _abi.F32, self.ctx.device)

    @property
    def max_priority(self):
        m = C.c_float()
        self.ctx.check(self.ctx.lib.crux_buffer_priorities(self.h, None, None, C.byref(m), None))
        return np.float32(m.value)

    @property
    def min_priority(self):
        m = C.c_float()
        self.ctx.check(self.ctx.lib.crux_buffer_priorities(self.h, None, None, None, C.byref(m)))
        return np.float32(m.value)

    def episodes(self):
        """``episodes(b)`` :194-221 -> 1-based inclusive (start, end) pairs (host bookkeeping)."""
        n = len(self)
        if "episode_end" in self.schema:
            ends = list(np.flatnonzero(self["episode_end"].reshape(-1).cpu().numpy()) + 1)
            starts = [1] + [e + 1 for e in ends[:-1]]
        elif "t" in self.schema:
            starts = list(np.flatnonzero(self["t"].reshape(-1).cpu().numpy() == 1) + 1)
            ends = [s - 1 for s in starts[1:]] + [n]
        else:
            raise ValueError("Need :episode_end flag or :t column to determine episodes")
        if n > 0 and (not ends or ends[-1] != n):
            starts.append((ends[-1] + 1) if ends else 1)
            ends.append(n)
        return list(zip(starts, ends))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self._cols = {}
                self.ctx.lib.crux_buffer_destroy(self.h)
                self.h = None
        except Exception:
            pass


def buffer_like(b, capacity=None, device=None):
    """``buffer_like`` :82-85: same columns, new capacity, empty; keeps α/β/max_priority semantics for PER."""
    cap = b.capacity if capacity is None else capacity
    pp = PriorityParams(b.priority_params.alpha, b.priority_params.beta) if b.isprioritized() else None
    return ExperienceBuffer(None, None, cap, prioritized=b.isprioritized(), priority_params=pp, ctx=b.ctx, _schema=b.schema)


def uniform_sample_(target, source, B=None, ids=None, seed=0, ctr=0):
    """``uniform_sample!`` :317-321.  ``ids`` (1-based) injects the draw ``rand(1:length(source), B)`` for parity runs."""
    B = target.capacity if B is None else int(B)
    idn = None if ids is None else np.ascontiguousarray(np.asarray(ids, dtype=np.int64) - 1, dtype=np.int32)
    target.ctx.check(target.ctx.lib.crux_buffer_sample_uniform(target.h, source.h, B, ptr(idn), seed, ctr))


def prioritized_sample_(target, source, i=1, B=None, rands=None, seed=0, ctr=0):
    """``prioritized_sample!`` :324-349.  ``rands`` injects ``rand(B)`` (Float64) for parity runs."""
    assert "weight" in source.schema  # :325
    B = target.capacity if B is None else int(B)
    u = None if rands is None else np.ascontiguousarray(rands, dtype=np.float64)
    beta = float(np.float32(source.priority_params.beta(i)))
    target.ctx.check(target.ctx.lib.crux_buffer_sample_prioritized(target.h, source.h, B, beta, COLUMN_IDS["weight"], ptr(u), seed, ctr))


def rand_(target, *sources, i=1, fracs=None, draws=None, seed=0, ctr=0):
    """``rand!(target, source...; i, fracs)`` :303-315.  ``draws[k]``: injected ids (uniform) / U(0,1) (prioritized)."""
    fracs = np.ones(len(sources)) / len(sources) if fracs is None else np.array(fracs, dtype=np.float64)
    lens = np.array([len(s) for s in sources])
    if np.any(lens == 0):
        fracs[lens == 0] = 0
        fracs = fracs / fracs.sum()
    batches = split_batches(target.capacity, fracs)
    for k, (b, B) in enumerate(zip(sources, batches)):
        if B == 0:
            continue
        d = None if draws is None else draws[k]
        if b.isprioritized():
            prioritized_sample_(target, b, i=i, B=B, rands=d, seed=seed, ctr=ctr + 2 * k)
        else:
            uniform_sample_(target, b, B=B, ids=d, seed=seed, ctr=ctr + 2 * k)
    return batches
