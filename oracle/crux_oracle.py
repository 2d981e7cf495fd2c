"""CPU oracle for the Crux.jl actor-learner hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, on the CPU (NumPy float32 + torch-CPU float32 autograd), the
algorithms of the reference hot path (sisl/Crux.jl @ d1b6ab5).  It is the checker
for the CUDA path: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``crux.jl_b200``) never imports it and has no CPU fallback.

Parity pinning status
---------------------
The reference is pure Julia and Julia is not installed in this image, so the
reference itself cannot be executed here.  The oracle is pinned against every
known-answer value the reference's own tests hold for this path (ring indices,
``split_batches``, priorities, schedules, one-hot/whiten, return recurrence; see
``tests/test_oracle_golden.py``).  GAE values, PPO/TD loss values, Adam steps and
TD targets are NOT pinned by any reference test (``test/gym/sampler_tests.jl:75-81``
asserts nothing): for those functions this oracle is "parity unpinned" and is a
line-by-line restatement only.  Flux/NNlib/Zygote semantics (Dense, Adam, mse,
softmax, softplus) are third-party to the reference and are restated from their
published definitions (Flux 0.14 compat range in ``Project.toml:34-52``).

Conventions
-----------
* Arrays are stored batch-first (``[batch, features]``), which is the *memory*
  order of the reference's column-major ``[features, batch]`` arrays
  (``src/devices.jl:23-34``).
* Dense weights are passed in Flux layout ``W[out, in]``.
* Indices inside ``ExperienceBuffer`` are 1-based like the reference; the
  C ABI is 0-based and tests convert.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

F32 = np.float32
EPS32 = np.float32(np.finfo(np.float32).eps)  # Julia eps(Float32)


# --------------------------------------------------------------------------- #
# index arithmetic  (src/experience_buffer.jl:223-259, test :23-28)
# --------------------------------------------------------------------------- #
def mod1(x, c):
    """Julia ``mod1``: result in 1..c."""
    return (np.asarray(x) - 1) % c + 1


def circ_inds(start, n, cap):
    """``mod1.(start:start+n-1, cap)`` (experience_buffer.jl:236)."""
    return mod1(np.arange(start, start + n), cap)


def split_batches(N, fracs):
    """experience_buffer.jl:126-131."""
    fracs = np.atleast_1d(np.asarray(fracs, dtype=np.float64))
    if np.ndim(fracs) == 0 or not np.isclose(fracs.sum(), 1.0):
        raise AssertionError("sum(fracs) must be ~1")
    batches = np.floor(N * fracs).astype(np.int64)
    batches[0] += N - batches.sum()
    return batches


class LinearDecaySchedule:
    """utils.jl:116-126 (Float64 arithmetic like the reference's `1., 0.1`)."""

    def __init__(self, start, stop, steps):
        self.start, self.stop, self.steps = start, stop, int(steps)

    def __call__(self, i):
        rate = (self.start - self.stop) / self.steps
        return max(self.stop, self.start - i * rate)


# --------------------------------------------------------------------------- #
# spaces / whitening  (src/spaces.jl:10-25, src/utils.jl:41-42)
# --------------------------------------------------------------------------- #
def whiten(v, mu=None, sigma=None):
    """utils.jl:41-42.  ``std`` is Bessel-corrected (n-1), no epsilon."""
    v = np.asarray(v, dtype=F32)
    if mu is None:
        mu = F32(v.mean(dtype=F32))
        sigma = F32(v.std(ddof=1, dtype=F32))
    return ((v - F32(mu)) / F32(sigma)).astype(F32)


def onehot(v, vals):
    """spaces.jl:24 ``Flux.onehot(v, S.vals)``."""
    vals = list(vals)
    out = np.zeros(len(vals), dtype=bool)
    out[vals.index(v)] = True
    return out


# --------------------------------------------------------------------------- #
# ExperienceBuffer  (src/experience_buffer.jl)
# --------------------------------------------------------------------------- #
_F32_ZERO = {"return", "logprob", "xlogprob", "advantage", "cost", "cost_advantage",
             "cost_return", "value", "var_prob", "cvar_prob", "f"}
_F32_ONE = {"weight", "importance_weight", "fwd_importance_weight", "rev_importance_weight",
            "cum_importance_weight", "traj_importance_weight"}
_BOOL = {"fail", "grasp_success", "expert"}
_INT = {"t", "i", "id"}


def mdp_data(sdims, stype, adims, atype, capacity, extras=()):
    """experience_buffer.jl:4-35.  Arrays are [capacity, *dims]."""
    sdims, adims = tuple(np.atleast_1d(sdims)), tuple(np.atleast_1d(adims))
    d = {
        "s": np.zeros((capacity, *sdims), dtype=stype),
        "a": np.zeros((capacity, *adims), dtype=atype),
        "sp": np.zeros((capacity, *sdims), dtype=stype),
        "r": np.zeros((capacity, 1), dtype=F32),
        "done": np.zeros((capacity, 1), dtype=bool),
        "episode_end": np.zeros((capacity, 1), dtype=bool),
    }
    for k in extras:
        if k in _F32_ZERO:
            d[k] = np.zeros((capacity, 1), dtype=F32)
        elif k in _F32_ONE:
            d[k] = np.ones((capacity, 1), dtype=F32)
        elif k in _BOOL:
            d[k] = np.zeros((capacity, 1), dtype=bool)
        elif k in _INT:
            d[k] = np.zeros((capacity, 1), dtype=np.int64)
        elif k == "s0":
            d[k] = np.zeros((capacity, *sdims), dtype=stype)
        elif k == "x":
            d[k] = np.zeros((capacity, *adims), dtype=atype)
        else:
            raise KeyError(f"Unrecognized key: {k}")
    return d


@dataclass
class PriorityParams:
    """experience_buffer.jl:38-50."""
    priorities: np.ndarray
    cumsum: np.ndarray = field(default_factory=lambda: np.zeros(0, dtype=F32))
    cumsum_valid: bool = False
    alpha: np.float32 = F32(0.6)
    beta: object = staticmethod(lambda i: F32(0.5))
    max_priority: np.float32 = F32(1.0)
    min_priority: np.float32 = F32(np.inf)


def pow_f32(x, p):
    """Julia ``Float32 ^ Float32``: evaluated in Float64 and rounded once."""
    return F32(np.float64(F32(x)) ** np.float64(F32(p)))


def cumsum_f32(x):
    """Prefix sum used for PER thresholds.

    Julia's ``cumsum`` on ``Vector{Float32}`` accumulates pairwise in blocks
    (Base, third party to the reference); only *index equality given the same
    prefix array* is required, so the oracle fixes a plain
    float64-accumulated, float32-rounded prefix as its own order."""
    return np.cumsum(np.asarray(x, dtype=np.float64)).astype(F32)


class ExperienceBuffer:
    """Ring buffer of SoA columns (experience_buffer.jl:53-80)."""

    def __init__(self, data, elements=None, next_ind=1, prioritized=False, alpha=0.6, beta=None):
        self.data = data
        cap = self.capacity
        self.elements = cap if elements is None else int(elements)
        self.next_ind = int(next_ind)
        self.indices = np.zeros(0, dtype=np.int64)
        self.total_count = self.elements
        self.pp = None
        if prioritized:
            if "weight" not in data:
                data["weight"] = np.ones((cap, 1), dtype=F32)
            self.pp = PriorityParams(priorities=np.zeros(cap, dtype=F32), alpha=F32(alpha))
            if beta is not None:
                self.pp.beta = beta
            if self.elements:
                self.update_priorities(np.arange(1, self.elements + 1),
                                       self.pp.max_priority * np.ones(self.elements, dtype=F32))  # :72 ones(Float32, ...)

    @classmethod
    def create(cls, sdims, adims, capacity, extras=(), stype=F32, atype=F32, **kw):
        return cls(mdp_data(sdims, stype, adims, atype, capacity, extras), elements=0, **kw)

    @property
    def capacity(self):
        return next(iter(self.data.values())).shape[0]

    def __len__(self):
        return self.elements

    def __getitem__(self, k):
        return self.data[k][: self.elements]

    def keys(self):
        return self.data.keys()

    def clear(self):
        """experience_buffer.jl:97-104."""
        self.elements, self.next_ind, self.total_count = 0, 1, 0
        self.indices = np.zeros(0, dtype=np.int64)
        if self.pp is not None:
            self.pp = PriorityParams(priorities=np.zeros(self.capacity, dtype=F32), alpha=self.pp.alpha,
                                     beta=self.pp.beta, max_priority=self.pp.max_priority)

    def get_last_N_indices(self, N):
        """experience_buffer.jl:223-229 (1-based)."""
        N = min(len(self), N)
        C = self.capacity
        start = int(mod1(self.next_ind - N, C))
        return circ_inds(start, N, C)

    def push(self, data, ids=None):
        """experience_buffer.jl:232-259.  ``ids`` 1-based; returns 1-based ``I``."""
        src = data.data if isinstance(data, ExperienceBuffer) else data
        first = next(iter(src.values()))
        if ids is None:
            n_src = len(data) if isinstance(data, ExperienceBuffer) else first.shape[0]
            ids = np.arange(1, n_src + 1)
        ids = np.asarray(ids, dtype=np.int64)
        N, C = len(ids), self.capacity
        self.total_count += N
        I = circ_inds(self.next_ind, N, C)
        for k in self.data:
            if k not in src:
                continue
            v2 = np.array(src[k][ids - 1])  # collect(...) : copy first (push!(b, b) safe)
            assert self.data[k].shape[1:] == v2.shape[1:], k
            # copyto! over an index-vector view: later rows win on wrap
            for j in range(N):
                self.data[k][I[j] - 1] = v2[j]
        if self.pp is not None:
            self.update_priorities(I, np.float64(self.pp.max_priority) * np.ones(N))
        self.elements = min(C, self.elements + N)
        self.next_ind = int(mod1(self.next_ind + N, C))
        return I

    def update_priorities(self, I, v):
        """experience_buffer.jl:290-301.  The element type of ``v`` decides the arithmetic like in Julia:
        Float32 ``v`` (``cpu(td_error(...))``, off_policy.jl:83) -> ``val`` and ``val^α`` are Float32;
        Float64 ``v`` (``max_priority*ones(N)`` in push!, :254; the Float64 literals of the reference test)
        -> Float64, rounded when stored into the Float32 fields."""
        assert len(I) == len(v)
        pp = self.pp
        v = np.asarray(v)
        wide = v.dtype == np.float64
        for i in range(len(I)):
            if wide:
                val = np.float64(v[i]) + np.float64(EPS32)
                pp.priorities[I[i] - 1] = F32(val ** np.float64(pp.alpha))
            else:
                val = F32(F32(v[i]) + EPS32)
                pp.priorities[I[i] - 1] = pow_f32(val, pp.alpha)
            pp.max_priority = F32(max(val, pp.max_priority))
            pp.min_priority = F32(min(val, pp.min_priority))
            pp.cumsum_valid = False

    def episodes(self):
        """experience_buffer.jl:194-221 (1-based inclusive pairs)."""
        n = len(self)
        if "episode_end" in self.data:
            ends = list(np.flatnonzero(self["episode_end"][:, 0]) + 1)
            starts = [1] + [e + 1 for e in ends[:-1]]
        elif "t" in self.data:
            starts = list(np.flatnonzero(self["t"][:, 0] == 1) + 1)
            ends = [s - 1 for s in starts[1:]] + [n]
        else:
            raise ValueError("Need :episode_end flag or :t column to determine episodes")
        if n > 0 and (not ends or ends[-1] != n):
            starts.append((ends[-1] + 1) if ends else 1)
            ends.append(n)
        return list(zip(starts, ends))


def uniform_sample(target, source, ids):
    """experience_buffer.jl:317-321 with the draw ``ids = rand(1:length(source), B)``
    supplied by the caller (1-based)."""
    target.indices = np.asarray(ids, dtype=np.int64)
    return target.push(source, ids=ids)


def prioritized_indices(cumsum, B, rands):
    """experience_buffer.jl:333-341.  Returns 1-based ids (may be N+1 if the
    threshold rounds above ptot: ``searchsortedfirst`` semantics)."""
    ptot = F32(cumsum[-1])
    dp = F32(ptot / F32(B))  # Float32 / Int -> Float32
    x = (np.arange(1, B + 1, dtype=np.float64) + np.asarray(rands, dtype=np.float64) - 1.0) * np.float64(dp)
    return np.searchsorted(np.asarray(cumsum, dtype=np.float64), x, side="left").astype(np.int64) + 1


def prioritized_sample(target, source, rands, i=1, B=None, cumsum=None):
    """experience_buffer.jl:324-349 with ``rands = rand(B)`` supplied.
    ``cumsum`` may be injected to test index equality given the same prefix."""
    assert "weight" in source.data
    pp = source.pp
    B = target.capacity if B is None else B
    N = len(source)
    prs = pp.priorities[:N]
    if cumsum is not None:
        pp.cumsum, pp.cumsum_valid = np.asarray(cumsum, dtype=F32), True
    if not pp.cumsum_valid:
        pp.cumsum, pp.cumsum_valid = cumsum_f32(prs), True
    ptot = F32(pp.cumsum[-1])
    ids = prioritized_indices(pp.cumsum, B, rands)
    target.indices = ids
    pmin = F32(pp.min_priority / ptot)
    beta = F32(pp.beta(i))
    max_w = pow_f32(F32(pmin * F32(N)), -beta)
    for id_ in ids:
        w = pow_f32(F32(F32(F32(N) * prs[id_ - 1]) / ptot), beta)
        source.data["weight"][id_ - 1, 0] = F32(w / max_w)
    return target.push(source, ids=ids)


def rand_b(target, sources, draws, i=1, fracs=None):
    """``rand!`` experience_buffer.jl:303-315.  ``draws[k]`` is the ids (uniform)
    or U(0,1) draws (prioritized) for source k."""
    fracs = np.ones(len(sources)) / len(sources) if fracs is None else np.array(fracs, dtype=np.float64)
    lens = np.array([len(s) for s in sources])
    if np.any(lens == 0):
        fracs[lens == 0] = 0
        fracs = fracs / fracs.sum()
    batches = split_batches(target.capacity, fracs)
    for b, B, dr in zip(sources, batches, draws):
        if B == 0:
            continue
        if b.pp is not None:
            prioritized_sample(target, b, dr, i=i, B=int(B))
        else:
            uniform_sample(target, b, dr)
    return batches


# --------------------------------------------------------------------------- #
# advantage / returns  (src/sampler.jl:262-281, :232-238)
# --------------------------------------------------------------------------- #
def fill_gae(r, done, v_s, v_sp, lam, gamma, rng=None, out=None):
    """sampler.jl:262-273 over one episode range (0-based slice ``rng``)."""
    n = len(r)
    rng = range(n) if rng is None else rng
    out = np.zeros(n, dtype=F32) if out is None else out
    A, c = F32(0), F32(F32(lam) * F32(gamma))
    gamma = F32(gamma)
    for i in reversed(rng):
        A = F32(F32(F32(c * A) + r[i]) + F32(F32(F32(1) - F32(done[i])) * gamma) * v_sp[i]) - v_s[i]
        A = F32(A)
        assert not np.isnan(A)
        out[i] = A
    return out


def fill_returns(r, gamma, rng=None, out=None):
    """sampler.jl:275-281 (never bootstraps)."""
    n = len(r)
    rng = range(n) if rng is None else rng
    out = np.zeros(n, dtype=F32) if out is None else out
    R, gamma = F32(0), F32(gamma)
    for i in reversed(rng):
        R = F32(r[i] + F32(gamma * R))
        out[i] = R
    return out


def discounted_return(r, gamma):
    """sampler.jl:232-238."""
    acc = F32(0)
    for x in reversed(list(r)):
        acc = F32(F32(x) + F32(gamma) * acc)
    return acc


def gae_returns_TN(r, done, episode_end, v_s, v_sp, gamma, lam):
    """The vector-env semantics of the new engine: the single-env
    recurrences above applied independently to every env's strided stream of a
    ``[T, N]`` rollout; a trace is cut wherever ``episode_end`` is set (that is what
    ``terminate_episode!`` sampler.jl:53-57 does by calling them per episode range).
    Vectorised over N, sequential over T, float32 with the reference's op order."""
    T, N = r.shape
    adv = np.zeros((T, N), dtype=F32)
    ret = np.zeros((T, N), dtype=F32)
    A = np.zeros(N, dtype=F32)
    R = np.zeros(N, dtype=F32)
    c = F32(F32(lam) * F32(gamma))
    g = F32(gamma)
    one = F32(1)
    for t in range(T - 1, -1, -1):
        cut = episode_end[t].astype(bool)
        A = np.where(cut, F32(0), A)
        R = np.where(cut, F32(0), R)
        nd = (one - done[t].astype(F32)) * g
        A = (((c * A) + r[t]) + nd * v_sp[t]) - v_s[t]
        R = r[t] + g * R
        adv[t], ret[t] = A, R
    return adv, ret


# --------------------------------------------------------------------------- #
# networks (Flux Dense/Chain restated; policies.jl:94-98,120-157,333-400)
# --------------------------------------------------------------------------- #
import torch  # noqa: E402  (CPU float32 autograd twin; never used on the product path)

ACT_IDENTITY, ACT_TANH, ACT_RELU = 0, 1, 2
_ACTS = {ACT_IDENTITY: lambda x: x, ACT_TANH: torch.tanh, ACT_RELU: torch.relu}
LOG_SQRT_2PI = 0.9189385332046727
ENT_CONST = 1.4189385332046727


def glorot_uniform(rng, out, inp):
    """Flux default init: ``(rand(Float32,out,in) .- 0.5f0) .* sqrt(24f0/(in+out))`` [3P]."""
    return ((rng.random((out, inp), dtype=F32) - F32(0.5)) * F32(math.sqrt(24.0 / (inp + out)))).astype(F32)


class MLP:
    """``Chain(Dense(in,out,act)...)``.  ``W[l]`` is ``[out,in]`` (Flux), ``b[l]`` ``[out]``."""

    def __init__(self, dims, acts, rng=None, Ws=None, bs=None):
        self.dims, self.acts = list(dims), list(acts)
        rng = np.random.default_rng(0) if rng is None else rng
        if Ws is None:
            Ws = [glorot_uniform(rng, dims[l + 1], dims[l]) for l in range(len(acts))]
            bs = [np.zeros(dims[l + 1], dtype=F32) for l in range(len(acts))]
        self.W = [torch.tensor(np.asarray(w, dtype=F32), requires_grad=True) for w in Ws]
        self.b = [torch.tensor(np.asarray(b, dtype=F32), requires_grad=True) for b in bs]

    def params(self):
        """``Flux.params`` order: W1, b1, W2, b2, ..."""
        out = []
        for w, b in zip(self.W, self.b):
            out += [w, b]
        return out

    def flat(self):
        """Flat parameter vector in the C-ABI order: per layer W in Julia memory
        order (column-major ``[out,in]`` == row-major ``[in,out]``) then b."""
        return np.concatenate([np.concatenate([w.detach().numpy().T.reshape(-1), b.detach().numpy()])
                               for w, b in zip(self.W, self.b)]).astype(F32)

    def set_flat(self, flat):
        off = 0
        with torch.no_grad():
            for l in range(len(self.acts)):
                i, o = self.dims[l], self.dims[l + 1]
                self.W[l].copy_(torch.from_numpy(np.array(flat[off:off + i * o]).reshape(i, o).T.copy()))
                off += i * o
                self.b[l].copy_(torch.from_numpy(np.array(flat[off:off + o])))
                off += o

    def __call__(self, x):
        x = torch.as_tensor(x, dtype=torch.float32)
        for w, b, a in zip(self.W, self.b, self.acts):
            x = _ACTS[a](x @ w.T + b)
        return x

    def clone(self):
        return MLP(self.dims, self.acts, Ws=[w.detach().numpy().copy() for w in self.W],
                   bs=[b.detach().numpy().copy() for b in self.b])


def flat_grads(params):
    """Gradient in the same flat C-ABI order as ``MLP.flat`` (+ trailing vectors)."""
    out = []
    for p in params:
        g = torch.zeros_like(p) if p.grad is None else p.grad
        out.append(g.detach().numpy().T.reshape(-1) if g.ndim == 2 else g.detach().numpy().reshape(-1))
    return np.concatenate(out).astype(F32)


def flat_values(params):
    return np.concatenate([(p.detach().numpy().T.reshape(-1) if p.ndim == 2 else p.detach().numpy().reshape(-1))
                           for p in params]).astype(F32)


def softplus(x):
    """NNlib ``softplus(x) = log1p(exp(-|x|)) + relu(x)`` [3P]."""
    return torch.log1p(torch.exp(-torch.abs(x))) + torch.relu(x)


class GaussianPolicy:
    """policies.jl:315-350 with a state-independent ``logΣ`` vector (ConstantLayer)."""

    def __init__(self, mu: MLP, log_sigma):
        self.mu = mu
        self.log_sigma = torch.tensor(np.asarray(log_sigma, dtype=F32), requires_grad=True)

    def params(self):
        return self.mu.params() + [self.log_sigma]

    def logpdf(self, s, a):
        """``gaussian_logpdf`` policies.jl:333-336."""
        mu = self.mu(s)
        a = torch.as_tensor(a, dtype=torch.float32)
        var = torch.exp(self.log_sigma) ** 2
        return torch.sum(-((a - mu) ** 2) / (2 * var) - LOG_SQRT_2PI - self.log_sigma, dim=1, keepdim=True)

    def exploration(self, s, eps):
        """policies.jl:338-344 with the noise ``eps`` supplied."""
        mu = self.mu(s)
        sigma = torch.exp(self.log_sigma)
        a = torch.as_tensor(eps, dtype=torch.float32) * sigma + mu
        var = torch.exp(self.log_sigma) ** 2
        logp = torch.sum(-((a - mu) ** 2) / (2 * var) - LOG_SQRT_2PI - self.log_sigma, dim=1, keepdim=True)
        return a, logp

    def action(self, s):
        return self.mu(s)

    def entropy(self, s=None):
        """policies.jl:348: a scalar; the constant is NOT scaled by the action dim."""
        return ENT_CONST + torch.sum(self.log_sigma)


class SquashedGaussianPolicy:
    """policies.jl:355-400.  ``mu``/``log_sigma`` are callables s -> [B,A]
    (they may share a trunk, examples/rl/half_cheetah_mujoco.jl:37-43)."""

    def __init__(self, mu, log_sigma, ascale=1.0, params=()):
        self.mu, self.log_sigma, self.ascale, self._params = mu, log_sigma, F32(ascale), list(params)

    def params(self):
        return self._params

    @staticmethod
    def sigma(log_sigma):
        """policies.jl:374-380."""
        return torch.exp(torch.clamp(log_sigma, -5, 2))

    @staticmethod
    def logprob(mu, log_sigma, a):
        """policies.jl:383-386: sigma from the CLAMPED logΣ, ``- logΣ`` UNclamped."""
        var = SquashedGaussianPolicy.sigma(log_sigma) ** 2
        return torch.sum(-((a - mu) ** 2) / (2 * var) - LOG_SQRT_2PI - log_sigma
                         - 2 * (math.log(2.0) - a - softplus(-2 * a)), dim=1, keepdim=True)

    def exploration(self, s, eps):
        """policies.jl:388-394."""
        mu, ls = self.mu(s), self.log_sigma(s)
        a_pre = torch.as_tensor(eps, dtype=torch.float32) * self.sigma(ls) + mu
        return float(self.ascale) * torch.tanh(a_pre), self.logprob(mu, ls, a_pre)

    def logpdf(self, s, a):
        """policies.jl:396."""
        a = torch.as_tensor(a, dtype=torch.float32)
        x = torch.clamp(a / float(self.ascale), -1.0 + 1e-5, 1.0 - 1e-5)
        return self.logprob(self.mu(s), self.log_sigma(s), torch.atanh(x))

    def action(self, s):
        return float(self.ascale) * torch.tanh(self.mu(s))

    def entropy(self, s):
        """policies.jl:398 ([B,1]; ignores the squash)."""
        return ENT_CONST + torch.sum(self.log_sigma(s), dim=1, keepdim=True)


class DiscreteNetwork:
    """policies.jl:104-157.  Actions cross as one-hot rows."""

    def __init__(self, net: MLP, outputs):
        self.net, self.outputs = net, list(outputs)

    def params(self):
        return self.net.params()

    def value(self, s, a_oh=None):
        q = self.net(s)
        if a_oh is None:
            return q
        return torch.sum(q * torch.as_tensor(np.asarray(a_oh, dtype=F32)), dim=1, keepdim=True)

    def logits(self, s):
        return torch.softmax(self.net(s), dim=1)

    def action_index(self, s):
        """policies.jl:124 argmax (first max wins, like Julia ``argmax``)."""
        return torch.argmax(self.net(s), dim=1)

    def logpdf(self, s, a_oh):
        """policies.jl:135,144-150."""
        return torch.log(torch.sum(self.logits(s) * torch.as_tensor(np.asarray(a_oh, dtype=F32)), dim=1, keepdim=True))

    def entropy(self, s):
        """policies.jl:152-155."""
        ps = self.logits(s)
        return -torch.sum(ps * torch.log(ps + float(EPS32)), dim=1, keepdim=True)

    def exploration(self, s, u):
        """policies.jl:137-142 with the Categorical draw replaced by inverse-CDF of the
        supplied uniforms ``u`` (Distributions' sampler stream is not reproducible here)."""
        ps = self.logits(s)
        cdf = torch.cumsum(ps.double(), dim=1)
        ai = torch.sum((cdf < torch.as_tensor(np.asarray(u, dtype=np.float64)).reshape(-1, 1)).long(), dim=1)
        ai = torch.clamp(ai, max=ps.shape[1] - 1)
        oh = torch.nn.functional.one_hot(ai, ps.shape[1]).float()
        return ai, torch.log(torch.sum(ps * oh, dim=1, keepdim=True))


def eps_greedy_logprob(eps, n_actions):
    """``exploration(::MixedPolicy)`` policies.jl:485-493 for a non-stochastic on-policy:
    ``log(ε·(1/n) + (1-ε))`` (Float64 like the reference's schedule)."""
    return math.log(eps * math.exp(math.log(1.0 / n_actions)) + (1.0 - eps))


# --------------------------------------------------------------------------- #
# losses  (rl/ppo.jl:4-21, rl/a2c.jl:4-16, utils.jl:76-96,112, rl/dqn.jl:4-6, rl/sac.jl)
# --------------------------------------------------------------------------- #
def ppo_loss(pi, P, D, info=None):
    """rl/ppo.jl:4-21."""
    info = {} if info is None else info
    new_probs = pi.logpdf(D["s"], D["a"])
    old = torch.as_tensor(D["logprob"], dtype=torch.float32).reshape(-1, 1)
    A = torch.as_tensor(D["advantage"], dtype=torch.float32).reshape(-1, 1)
    r = torch.exp(new_probs - old)
    lo, hi = F32(1) - F32(P["eps"]), F32(1) + F32(P["eps"])
    p_loss = -torch.mean(torch.minimum(r * A, torch.clamp(r, float(lo), float(hi)) * A))
    e_loss = -torch.mean(torch.as_tensor(pi.entropy(D["s"])))
    with torch.no_grad():
        info["entropy"] = float(-e_loss)
        info["kl"] = float(torch.mean(old - new_probs))
        info["clip_fraction"] = float(torch.sum((r > float(hi)) | (r < float(lo)))) / r.numel()
        info["avg_advantage"] = float(torch.mean(A))
        if "return" in D:
            info["avg_return"] = float(np.mean(np.asarray(D["return"], dtype=F32)))
    return float(P["lp"]) * p_loss + float(P["le"]) * e_loss


def a2c_loss(pi, P, D, info=None):
    """rl/a2c.jl:4-16."""
    info = {} if info is None else info
    new_probs = pi.logpdf(D["s"], D["a"])
    A = torch.as_tensor(D["advantage"], dtype=torch.float32).reshape(-1, 1)
    p_loss = -torch.mean(new_probs * A)
    e_loss = -torch.mean(torch.as_tensor(pi.entropy(D["s"])))
    with torch.no_grad():
        info["entropy"] = float(-e_loss)
        info["kl"] = float(torch.mean(torch.as_tensor(D["logprob"], dtype=torch.float32).reshape(-1, 1) - new_probs))
    return float(P["lp"]) * p_loss + float(P["le"]) * e_loss


def lagrange_params(eps=0.2, lp=1.0, le=0.1, target_cost=0.025, penalty_max=math.inf, Ki_max=10.0, Ki=1e-3, Kp=1, Kd=0, ema_alpha=0.95):
    """The 𝒫 NamedTuple of ``LagrangePPO`` (rl/ppo.jl:185-202): Float32 scalars and 1-element Float32 state arrays;
    ``ema_α`` stays Float64 like the reference's literal ``0.95``."""
    return {"eps": F32(eps), "lp": F32(lp), "le": F32(le), "target_cost": F32(target_cost), "penalty_max": F32(penalty_max),
            "Ki_max": F32(Ki_max), "Ki": F32(Ki), "Kp": Kp, "Kd": Kd, "ema_alpha": float(ema_alpha),
            "I": F32(0), "Jc_prev": F32(0), "smooth_D": F32(0), "smooth_Jc": F32(0)}


def lagrange_ppo_loss(pi, P, D, info=None):
    """rl/ppo.jl:70-131.  The PID penalty update (:79-106) runs inside the loss, once per evaluation, and mutates ``P``."""
    info = {} if info is None else info
    new_probs = pi.logpdf(D["s"], D["a"])
    old = torch.as_tensor(D["logprob"], dtype=torch.float32).reshape(-1, 1)
    r = torch.exp(new_probs - old)
    A = torch.as_tensor(D["advantage"], dtype=torch.float32).reshape(-1, 1)
    Ac = torch.as_tensor(D["cost_advantage"], dtype=torch.float32).reshape(-1, 1)
    lo, hi = float(F32(1) - F32(P["eps"])), float(F32(1) + F32(P["eps"]))
    p_loss = -torch.mean(torch.minimum(r * A, torch.clamp(r, lo, hi) * A))
    e_loss = -torch.mean(torch.as_tensor(pi.entropy(D["s"])))
    # ---- penalty (ignore_derivatives block)
    Jc = F32(np.sum(np.asarray(D["cost"], dtype=F32), dtype=np.float64)) / F32(np.sum(np.asarray(D["episode_end"]).astype(np.int64)))
    d = F32(Jc - P["target_cost"])
    P["I"] = F32(min(max(F32(P["I"] + P["Ki"] * d), F32(0)), P["Ki_max"]))
    al = P["ema_alpha"]
    P["smooth_D"] = F32(al * float(P["smooth_D"]) + (1.0 - al) * float(d))
    P["smooth_Jc"] = F32(al * float(P["smooth_Jc"]) + (1.0 - al) * float(Jc))
    der = F32(max(F32(0), F32(P["smooth_Jc"] - P["Jc_prev"])))
    P["Jc_prev"] = P["smooth_Jc"]
    penalty = F32(min(max(F32(F32(P["Kp"]) * P["smooth_D"] + P["I"] + F32(P["Kd"]) * der), F32(0)), P["penalty_max"]))
    info.update({"penalty": float(penalty), "cur_cost": float(Jc), "prop_term": float(F32(P["Kp"]) * P["smooth_D"]),
                 "deriv_term": float(der), "integral term": float(P["I"])})
    cost_loss = float(penalty) * torch.mean(torch.maximum(r * Ac, torch.clamp(r, lo, hi) * Ac))
    with torch.no_grad():
        info["entropy"] = float(-e_loss)
        info["kl"] = float(torch.mean(old - new_probs))
        info["clip_fraction"] = float(torch.sum((r > hi) | (r < lo))) / r.numel()
        info["p_loss"] = float(float(P["lp"]) * p_loss)
        info["cost_loss"] = float(cost_loss)
        info["avg_advantage"] = float(torch.mean(A))
        info["avg_return"] = float(np.mean(np.asarray(D["return"], dtype=F32)))
    return (float(P["lp"]) * p_loss + float(P["le"]) * e_loss + cost_loss) / (1.0 + float(penalty))


def reinforce_loss(pi, P, D, info=None):
    """rl/reinforce.jl:4-13: ``-mean(logpdf(π, s, a) .* return)``; info: entropy, kl."""
    info = {} if info is None else info
    new_probs = pi.logpdf(D["s"], D["a"])
    R = torch.as_tensor(D["return"], dtype=torch.float32).reshape(-1, 1)
    with torch.no_grad():
        info["entropy"] = float(torch.mean(torch.as_tensor(pi.entropy(D["s"]))))
        info["kl"] = float(torch.mean(torch.as_tensor(D["logprob"], dtype=torch.float32).reshape(-1, 1) - new_probs))
    return -torch.mean(new_probs * R)


def value_mse_loss(V, D):
    """PPO/A2C critic loss ``Flux.mse(value(π, D[:s]), D[:return])`` (ppo.jl:60, a2c.jl:47)."""
    ret = torch.as_tensor(D["return"], dtype=torch.float32).reshape(-1, 1)
    return torch.mean((V(D["s"]) - ret) ** 2)


def td_loss(q_sa, y, weight=None, info=None, name="Qavg"):
    """utils.jl:76-87 with Q(s,a) already evaluated; ``weighted_mean`` utils.jl:47."""
    y = torch.as_tensor(y, dtype=torch.float32).reshape(-1, 1)
    if info is not None:
        info[name] = float(torch.mean(q_sa).detach())
    e = (q_sa - y) ** 2
    if weight is None:
        return torch.mean(e)
    return torch.mean(e * torch.as_tensor(weight, dtype=torch.float32).reshape(-1, 1))


def dqn_target(q_target_sp, r, done, gamma):
    """rl/dqn.jl:4-6."""
    q = torch.as_tensor(q_target_sp, dtype=torch.float32)
    r = torch.as_tensor(r, dtype=torch.float32).reshape(-1, 1)
    nd = 1.0 - torch.as_tensor(np.asarray(done, dtype=F32)).reshape(-1, 1)
    return (r + float(F32(gamma)) * nd * torch.max(q, dim=1, keepdim=True).values).detach()


def soft_value(q, alpha=1.0):
    """rl/softq.jl:8: ``α .* logsumexp(value(π, s) ./ α, dims=1)`` (NNlib logsumexp = max + log Σ exp(x - max))."""
    x = torch.as_tensor(q, dtype=torch.float32) / float(F32(alpha))
    m = torch.max(x, dim=1, keepdim=True).values
    return float(F32(alpha)) * (m + torch.log(torch.sum(torch.exp(x - m), dim=1, keepdim=True)))


def softq_target(q_target_sp, r, done, gamma, alpha=1.0):
    """rl/softq.jl:13-17."""
    r = torch.as_tensor(r, dtype=torch.float32).reshape(-1, 1)
    nd = 1.0 - torch.as_tensor(np.asarray(done, dtype=F32)).reshape(-1, 1)
    return (r + float(F32(gamma)) * nd * soft_value(q_target_sp, alpha)).detach()


def softq_logits(q, alpha=1.0):
    """The logit_conversion SoftQ installs: ``softmax(value(π, s) ./ α)`` (rl/softq.jl:48)."""
    return torch.softmax(torch.as_tensor(q, dtype=torch.float32) / float(F32(alpha)), dim=1)


def td_error(q_sa, y):
    """utils.jl:112."""
    return torch.abs(q_sa - torch.as_tensor(y, dtype=torch.float32).reshape(-1, 1)).detach()


def sac_target(actor, q1_t, q2_t, D, gamma, log_alpha, eps):
    """rl/sac.jl:4-9: next action from the ONLINE actor, TARGET critics."""
    with torch.no_grad():
        ap, logp = actor.exploration(D["sp"], eps)
        sp = torch.as_tensor(D["sp"], dtype=torch.float32)
        x = torch.cat([sp, ap], dim=1)
        qmin = torch.minimum(q1_t(x), q2_t(x))
        r = torch.as_tensor(D["r"], dtype=torch.float32).reshape(-1, 1)
        nd = 1.0 - torch.as_tensor(np.asarray(D["done"], dtype=F32)).reshape(-1, 1)
        return r + float(F32(gamma)) * nd * (qmin - math.exp(float(log_alpha)) * logp)


def gaussian_noise_exploration(a, eps, sigma, eps_min=-math.inf, eps_max=math.inf, a_min=-math.inf, a_max=math.inf):
    """``exploration(::GaussianNoiseExplorationPolicy)`` policies.jl:510-514 on a = action(π_on, s):
    ``clamp.(a .+ clamp.(randn .* σ(i), ϵ_min, ϵ_max), a_min, a_max)``."""
    a = torch.as_tensor(a, dtype=torch.float32)
    n = torch.clamp(torch.as_tensor(eps, dtype=torch.float32) * float(F32(sigma)), float(eps_min), float(eps_max))
    lo = torch.as_tensor(np.asarray(a_min, dtype=F32)).reshape(1, -1)
    hi = torch.as_tensor(np.asarray(a_max, dtype=F32)).reshape(1, -1)
    return torch.minimum(torch.maximum(a + n, lo), hi)


def ddpg_target(actor_t, critics_t, D, gamma, smooth=None):
    """rl/ddpg.jl:6-8 (one critic, no noise), smoothed_ddpg_target :14-17, td3_target rl/td3.jl:4-7 (min over two critics).
    ``smooth`` = dict(eps, sigma, eps_min, eps_max, a_min, a_max) or None; networks are the TARGET copies (off_policy.jl:80)."""
    with torch.no_grad():
        ap = actor_t(D["sp"])
        if smooth is not None:
            ap = gaussian_noise_exploration(ap, **smooth)
        x = torch.cat([torch.as_tensor(D["sp"], dtype=torch.float32), ap], dim=1)
        q = critics_t[0](x)
        for c in critics_t[1:]:
            q = torch.minimum(q, c(x))
        r = torch.as_tensor(D["r"], dtype=torch.float32).reshape(-1, 1)
        nd = 1.0 - torch.as_tensor(np.asarray(D["done"], dtype=F32)).reshape(-1, 1)
        return r + float(F32(gamma)) * nd * q


def ddpg_actor_loss(actor, q1, D):
    """rl/ddpg.jl:25 / td3_actor_loss rl/td3.jl:12: ``-mean(value(Q1, s, action(π, s)))``."""
    s = torch.as_tensor(D["s"], dtype=torch.float32)
    return -torch.mean(q1(torch.cat([s, actor(D["s"])], dim=1)))


def sac_actor_loss(actor, q1, q2, D, log_alpha, eps, info=None):
    """rl/sac.jl:34-40 (fresh noise ``eps``)."""
    a, logp = actor.exploration(D["s"], eps)
    if info is not None:
        info["entropy"] = float(-torch.mean(logp))
    x = torch.cat([torch.as_tensor(D["s"], dtype=torch.float32), a], dim=1)
    return torch.mean(math.exp(float(log_alpha)) * logp - torch.minimum(q1(x), q2(x)))


def sac_temp_loss(actor, D, log_alpha_t, H_target, eps):
    """rl/sac.jl:45-52; ``log_alpha_t`` is a torch scalar parameter."""
    with torch.no_grad():
        _, logp = actor.exploration(D["s"], eps)
    return -torch.mean(torch.exp(log_alpha_t) * (logp + float(H_target)))


# --------------------------------------------------------------------------- #
# optimiser / training loop  (src/training.jl, Flux.Optimise.Adam [3P])
# --------------------------------------------------------------------------- #
class Adam:
    """Flux 0.14 ``Optimise.Adam`` [3P]: η, β, ϵ are Float64 scalars; per-array state
    ``(m, v, βp)`` is float32/float32/Float64; element math promotes to Float64 and each
    broadcast assignment rounds to float32:

        m  = β1 m + (1-β1) g ;  v = β2 v + (1-β2) g²
        Δ  = m / (1-βp1) / (sqrt(v / (1-βp2)) + ϵ) * η ;  βp .*= β ;  x .-= Δ
    """

    def __init__(self, eta=F32(3e-4), beta=(0.9, 0.999), eps=1e-8):
        self.eta, self.beta, self.eps = float(eta), (float(beta[0]), float(beta[1])), float(eps)
        self.state = {}

    def step(self, params):
        for p in params:
            if p.grad is None:
                continue
            g = p.grad.detach().numpy().astype(np.float64)
            if id(p) not in self.state:
                self.state[id(p)] = [np.zeros(p.shape, dtype=F32), np.zeros(p.shape, dtype=F32), list(self.beta)]
            m, v, bp = self.state[id(p)]
            m[...] = (self.beta[0] * m.astype(np.float64) + (1 - self.beta[0]) * g).astype(F32)
            v[...] = (self.beta[1] * v.astype(np.float64) + (1 - self.beta[1]) * g * g).astype(F32)
            delta = (m.astype(np.float64) / (1 - bp[0]) / (np.sqrt(v.astype(np.float64) / (1 - bp[1])) + self.eps)
                     * self.eta).astype(F32)
            bp[0] *= self.beta[0]
            bp[1] *= self.beta[1]
            with torch.no_grad():
                p -= torch.from_numpy(delta)


def grad_norm(params):
    """utils.jl:49-55: 2-norm of the per-array 2-norms."""
    v = [float(torch.linalg.vector_norm(p.grad)) for p in params if p.grad is not None]
    return float(np.linalg.norm(np.asarray(v, dtype=F32)))


def train_step(params, loss_fn, opt, info, name=""):
    """``train!`` training.jl:15-25."""
    for p in params:
        p.grad = None
    l = loss_fn(info)
    l.backward()
    gn = grad_norm(params)
    if math.isnan(gn):
        raise FloatingPointError(f"NaN detected! Loss: {float(l.detach())}")
    opt.step(params)
    info[name + "loss"] = float(l.detach())
    info[name + "grad_norm"] = gn
    return info


def batch_train(params, loss_on_mb, opt, D, n, batch_size, epochs, perms, info=None, name="",
                early_stopping=None, max_batches=math.inf):
    """``batch_train!`` training.jl:28-55 for one buffer.

    ``shuffle!`` permutes the buffer in place every epoch (cumulatively,
    experience_buffer.jl:118-124); ``perms[e]`` (0-based) is that epoch's
    ``shuffle(1:n)``.  The same ``info`` dict is pushed for every minibatch
    (training.jl:43), so aggregates equal the LAST minibatch's values."""
    info = {} if info is None else info
    infos, total = [], 0
    order = np.arange(n)
    for e in range(epochs):
        order = order[np.asarray(perms[e])]
        mb_infos, stop = [], False
        for start in range(0, n, batch_size):
            idx = order[start:start + batch_size]
            mb = {k: np.asarray(v)[idx] for k, v in D.items()}
            mb_infos.append(train_step(params, lambda inf, mb=mb: loss_on_mb(mb, inf), opt, info, name))
            total += 1
            if total >= max_batches:
                break
            if early_stopping is not None and early_stopping(infos + [aggregate_info(mb_infos)]):
                stop = True
                break
        infos.append(aggregate_info(mb_infos))
        if stop:  # `break` in training.jl:46 only leaves the minibatch loop ...
            pass
        if early_stopping is not None and early_stopping(infos):  # ... :49 leaves the epoch loop
            break
        if total >= max_batches:
            break
    info[name + "batches_trained"] = total
    info.update(aggregate_info(infos))
    return info, order


def aggregate_info(infos):
    """logging.jl:60-66."""
    keys = []
    for inf in infos:
        for k in inf:
            if k not in keys:
                keys.append(k)
    return {k: float(np.mean([inf[k] for inf in infos if k in inf])) for k in keys}


def polyak_average(to_params, from_params, tau):
    """policies.jl:48-59."""
    tau = F32(tau)
    with torch.no_grad():
        for t, f in zip(to_params, from_params):
            t.copy_(float(tau) * f + float(F32(1) - tau) * t)


# --------------------------------------------------------------------------- #
# synthetic MDPs used by the benchmark configs -- not part of the
# reference; they stand in for a user's POMDPs.jl model
# --------------------------------------------------------------------------- #
class LinQuadSpec:
    """"LinQuad-17x6": s' = clip(A s + B tanh(a) + 0.01 ξ, -10, 10); r = 1 - |s'|²/17 - 0.1|a|²/6;
    terminal if |s'_1| > 5; s0 ~ U(-0.1, 0.1)^17; γ = 0.99; max_steps = 1000."""

    def __init__(self, obs_dim=17, act_dim=6, seed=0):
        rng = np.random.default_rng(seed)
        self.obs_dim, self.act_dim = obs_dim, act_dim
        G1 = rng.standard_normal((obs_dim, obs_dim)).astype(F32)
        G2 = rng.standard_normal((obs_dim, act_dim)).astype(F32)
        self.A = (F32(0.95) * np.eye(obs_dim, dtype=F32) + F32(0.02) * G1).astype(F32)
        self.B = (F32(0.1) * G2).astype(F32)
        self.gamma = F32(0.99)
        self.max_steps = 1000

    def step(self, s, a, xi):
        """s [N,17], a [N,6], xi [N,17] standard normal -> sp, r, done (float32)."""
        sp = np.clip(s @ self.A.T + np.tanh(a) @ self.B.T + F32(0.01) * xi, F32(-10), F32(10)).astype(F32)
        r = (F32(1) - (sp * sp).sum(1) / F32(self.obs_dim) - F32(0.1) * (a * a).sum(1) / F32(self.act_dim)).astype(F32)
        done = np.abs(sp[:, 0]) > F32(5)
        return sp, r, done
