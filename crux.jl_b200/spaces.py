"""Mirror of ``src/spaces.jl``: ``DiscreteSpace``, ``ContinuousSpace``, ``tovec``, ``state_space``."""
from __future__ import annotations

import numpy as np


class AbstractSpace:
    pass


class DiscreteSpace(AbstractSpace):
    """spaces.jl:2-8."""

    def __init__(self, N, vals=None):
        if not isinstance(N, (int, np.integer)):
            vals = list(N)
            N = len(vals)
        self.N = int(N)
        self.vals = list(range(1, self.N + 1)) if vals is None else list(vals)

    @property
    def type(self):
        return np.bool_

    @property
    def dims(self):
        return (self.N,)


class ContinuousSpace(AbstractSpace):
    """spaces.jl:10-16."""

    def __init__(self, dims, type=np.float32, mu=np.float32(0), sigma=np.float32(1)):
        self.dims = tuple(int(d) for d in np.atleast_1d(dims))
        self.type, self.mu, self.sigma = type, mu, sigma


def dim(S):
    return S.dims


def whiten(v, mu=None, sigma=None):
    """utils.jl:41-42 (host arrays; the device version is ``crux_whiten``)."""
    v = np.asarray(v)
    if mu is None:
        mu, sigma = v.mean(), v.std(ddof=1)
    return (v - mu) / sigma


def tovec(v, S):
    """spaces.jl:24-25: one-hot for a DiscreteSpace, ``whiten(v, μ, σ)`` for a ContinuousSpace."""
    if isinstance(S, DiscreteSpace):
        out = np.zeros(S.N, dtype=bool)
        out[S.vals.index(v)] = True
        return out
    return whiten(v, S.mu, S.sigma)


def state_space(o, mu=np.float32(0), sigma=np.float32(1)):
    """spaces.jl:27-31 for an observation array."""
    o = np.asarray(o)
    dims = o.shape
    if len(dims) == 4 and dims[-1] == 1:
        dims = dims[:-1]
    return ContinuousSpace(dims, o.dtype.type, mu, sigma)
