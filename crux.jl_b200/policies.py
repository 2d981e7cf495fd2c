"""Host-side mirror of ``src/policies.jl`` driving libcrux_cuda.so.

Names follow the reference (``ContinuousNetwork``, ``DiscreteNetwork``, ``GaussianPolicy``,
``SquashedGaussianPolicy``, ``ActorCritic``, ``DoubleNetwork``, ``value``/``action``/``exploration``/
``logpdf``/``entropy``, ``polyak_average_``); Julia's ``f!`` is spelled ``f_`` here.  Arrays are
batch-major ``[B, features]`` torch CUDA tensors (the memory order of the reference's ``[features, B]``).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _abi
from .device import default_context, ptr

identity, tanh, relu = _abi.ACT_IDENTITY, _abi.ACT_TANH, _abi.ACT_RELU
_ACT = {"identity": identity, "tanh": tanh, "relu": relu, None: identity}


def glorot_uniform(rng, out, inp):
    """Flux's default Dense init: (rand(Float32,out,in) .- 0.5f0) .* sqrt(24f0/(in+out))."""
    return ((rng.random((out, inp), dtype=np.float32) - np.float32(0.5)) * np.float32(math.sqrt(24.0 / (inp + out)))).astype(np.float32)


class Dense:
    """``Flux.Dense(in, out, σ)``: weight ``[out, in]``, bias ``[out]`` (zeros)."""

    def __init__(self, inp, out, act=identity, weight=None, bias=None, rng=None):
        self.inp, self.out = int(inp), int(out)
        self.act = _ACT.get(act, act)
        rng = rng if rng is not None else np.random.default_rng()
        self.weight = glorot_uniform(rng, out, inp) if weight is None else np.asarray(weight, dtype=np.float32).reshape(out, inp)
        self.bias = np.zeros(out, dtype=np.float32) if bias is None else np.asarray(bias, dtype=np.float32).reshape(out)


class Conv:
    """``Flux.Conv((k, k), cin => cout, σ; stride)``: weight ``[kw, kh, cin, cout]`` (held here as its memory ``[cout, cin, kh, kw]``), bias
    ``[cout]``; a true convolution (flipped kernel), no padding (the defaults of examples/rl/atari.jl:8)."""

    def __init__(self, k, cin, cout, act=identity, stride=1, weight=None, bias=None, rng=None):
        k = k[0] if isinstance(k, (tuple, list)) else k
        self.k, self.cin, self.cout, self.stride = int(k), int(cin), int(cout), int(stride)
        self.act = _ACT.get(act, act)
        rng = rng if rng is not None else np.random.default_rng()
        lim = math.sqrt(6.0 / (self.k * self.k * (self.cin + self.cout)))     # Flux.glorot_uniform on the 4-d array
        self.weight = (((rng.random((cout, cin, self.k, self.k), dtype=np.float32) * 2 - 1) * lim).astype(np.float32) if weight is None
                       else np.asarray(weight, dtype=np.float32).reshape(cout, cin, self.k, self.k))
        self.bias = np.zeros(cout, dtype=np.float32) if bias is None else np.asarray(bias, dtype=np.float32).reshape(cout)


class _Marker:
    def __init__(self, name):
        self.name = name

    def __repr__(self):
        return self.name


flatten = _Marker("flatten")      # Flux.flatten
scale255 = _Marker("x -> x ./ 255f0")


class Chain:
    def __init__(self, *layers):
        self.layers = list(layers)
        for a, b in zip(self.layers[:-1], self.layers[1:]):
            if isinstance(a, Dense) and isinstance(b, Dense):
                assert a.out == b.inp, "Chain: layer widths do not match"

    @property
    def is_conv(self):
        return any(isinstance(l, Conv) for l in self.layers)

    def flat(self):
        """Flux.params order, W in Julia memory order (Dense: row-major [in][out]; Conv: [cout][cin][kh][kw])."""
        parts = []
        for l in self.layers:
            if isinstance(l, Dense):
                parts += [l.weight.T.reshape(-1), l.bias]
            elif isinstance(l, Conv):
                parts += [l.weight.reshape(-1), l.bias]
        return np.concatenate(parts).astype(np.float32)


class Policy:
    pass


class NetworkPolicy(Policy):
    ctx = None

    def __call__(self, *x):  # policies.jl:35
        return value(self, *x)


class _MLP:
    """Owner of one crux_mlp handle."""

    def __init__(self, chain, ctx=None):
        self.ctx = ctx or default_context()
        self.chain = chain
        self.dims = [chain.layers[0].inp] + [l.out for l in chain.layers]
        self.acts = [l.act for l in chain.layers]
        n = len(self.acts)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_mlp_create(self.ctx.h, n, (C.c_int32 * (n + 1))(*self.dims), (C.c_int32 * n)(*self.acts), C.byref(h)))
        self.h = h
        self.set_flat(chain.flat())

    @property
    def n_params(self):
        n = C.c_int64()
        self.ctx.check(self.ctx.lib.crux_mlp_num_params(self.h, C.byref(n)))
        return n.value

    def set_flat(self, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.n_params
        self.ctx.check(self.ctx.lib.crux_mlp_set_params(self.h, ptr(flat)))

    def get_flat(self):
        out = np.empty(self.n_params, dtype=np.float32)
        self.ctx.check(self.ctx.lib.crux_mlp_get_params(self.h, ptr(out)))
        return out

    def grads(self):
        from .device import view
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_mlp_grads_ptr(self.h, C.byref(p)))
        return view(p.value, (self.n_params,), _abi.F32, self.ctx.device)

    def set_adam(self, eta=np.float32(3e-4), beta=(0.9, 0.999), eps=1e-8):
        self.ctx.check(self.ctx.lib.crux_mlp_set_adam(self.h, float(eta), float(beta[0]), float(beta[1]), float(eps)))

    def forward(self, x, out=None):
        x = _as_dev(self.ctx, x)
        B = x.shape[0]
        assert x.shape[1] == self.dims[0], f"input width {x.shape[1]} != {self.dims[0]}"
        out = self.ctx.empty((B, self.dims[-1])) if out is None else out
        self.ctx.check(self.ctx.lib.crux_mlp_forward(self.h, ptr(x), B, ptr(out)))
        return out

    def value_next(self, sp, s, v_s, T, N, out):
        """value(V, sp) over a [T][N] rollout given v_s = value(V, s): rows whose sp equals the next step's s bit for bit reuse v_s."""
        self.ctx.check(self.ctx.lib.crux_value_next(self.h, ptr(sp), ptr(s), ptr(v_s), T, N, ptr(out)))
        return out

    def train_dqn(self, s, a_onehot, y, weight, B, info=None):
        """one DQN critic ``train!`` (off_policy.jl:91-93 with td_loss)"""
        self.ctx.check(self.ctx.lib.crux_dqn_train(self.h, ptr(s), ptr(a_onehot), ptr(y), ptr(weight), B, None if info is None else ptr(info)))

    def polyak_from(self, frm, tau):
        self.ctx.check(self.ctx.lib.crux_mlp_polyak(self.h, frm.h, float(tau)))

    def copy_from(self, frm):
        self.ctx.check(self.ctx.lib.crux_mlp_copy(self.h, frm.h))

    def forward_sa(self, s, a):
        s, a = _as_dev(self.ctx, s), _as_dev(self.ctx, a)
        B = s.shape[0]
        out = self.ctx.empty((B, self.dims[-1]))
        self.ctx.check(self.ctx.lib.crux_mlp_forward_sa(self.h, ptr(s), s.shape[1], ptr(a), a.shape[1], B, ptr(out)))
        return out

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.ctx.lib.crux_mlp_destroy(self.h)
                self.h = None
        except Exception:
            pass


class _ConvNet:
    """Owner of one crux_convq handle: ``Chain([scale255,] Conv, Conv, flatten, Dense, Dense)`` over ``[C, H, W]`` observations (u8 or f32).
    Same surface as ``_MLP`` where the DQN paths use it (forward, flat parameters, Adam, train_dqn, polyak / copy)."""

    def __init__(self, chain, input_dims, ctx=None):
        self.ctx = ctx or default_context()
        self.chain = chain
        ls = [l for l in chain.layers if l is not flatten and l is not scale255]
        assert (len(ls) == 4 and isinstance(ls[0], Conv) and isinstance(ls[1], Conv) and isinstance(ls[2], Dense) and isinstance(ls[3], Dense)
                and ls[0].act == relu and ls[1].act == relu and ls[2].act == relu and ls[3].act == identity), \
            "the pixel network is Chain([x -> x ./ 255f0,] Conv(relu), Conv(relu), flatten, Dense(relu), Dense) (examples/rl/atari.jl:8)"
        W, H, Cin = (int(d) for d in input_dims)                 # Julia (w, h, c) like ContinuousSpace((84, 84, 4))
        c1, c2, d1, d2 = ls
        assert c1.cin == Cin and c2.cin == c1.cout and d2.inp == d1.out
        self.input_shape = (Cin, H, W)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_convq_create(self.ctx.h, Cin, H, W, 1 if scale255 in chain.layers else 0, c1.k, c1.stride, c1.cout,
                                                      c2.k, c2.stride, c2.cout, d1.out, d2.out, C.byref(h)))
        self.h = h
        F = C.c_int32()
        self.ctx.check(self.ctx.lib.crux_convq_shape(self.h, C.byref(F), None, None, None, None))
        assert d1.inp == F.value, f"Dense after flatten expects {d1.inp} inputs, the convolutions produce {F.value}"
        self.dims = [Cin * H * W, d2.out]
        self.acts = []
        self.set_flat(chain.flat())

    @property
    def n_params(self):
        n = C.c_int64()
        self.ctx.check(self.ctx.lib.crux_convq_num_params(self.h, C.byref(n)))
        return n.value

    def set_flat(self, flat):
        flat = np.ascontiguousarray(flat, dtype=np.float32)
        assert flat.size == self.n_params
        self.ctx.check(self.ctx.lib.crux_convq_set_params(self.h, ptr(flat)))

    def get_flat(self):
        out = np.empty(self.n_params, dtype=np.float32)
        self.ctx.check(self.ctx.lib.crux_convq_get_params(self.h, ptr(out)))
        return out

    def grads(self):
        out = np.empty(self.n_params, dtype=np.float32)
        self.ctx.check(self.ctx.lib.crux_convq_grads(self.h, ptr(out)))
        return out

    def set_adam(self, eta=np.float32(3e-4), beta=(0.9, 0.999), eps=1e-8, clip_value=0.0):
        self.ctx.check(self.ctx.lib.crux_convq_set_adam(self.h, float(eta), float(beta[0]), float(beta[1]), float(eps), float(clip_value)))

    def _obs(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.ascontiguousarray(x))
        x = x.to(self.ctx.device)
        if x.dtype != torch.uint8:
            x = x.to(torch.float32)
        x = x.reshape(-1, self.dims[0]).contiguous()
        return x, 1 if x.dtype == torch.uint8 else 0

    def forward(self, x, out=None):
        x, u8 = self._obs(x)
        out = self.ctx.empty((x.shape[0], self.dims[-1])) if out is None else out
        self.ctx.check(self.ctx.lib.crux_convq_forward(self.h, ptr(x), u8, x.shape[0], ptr(out)))
        return out

    def train_dqn(self, s, a_onehot, y, weight, B, info=None):
        s, u8 = self._obs(s)
        self.ctx.check(self.ctx.lib.crux_convq_dqn_train(self.h, ptr(s), u8, ptr(a_onehot), ptr(y), ptr(weight), B, None if info is None else ptr(info)))

    def polyak_from(self, frm, tau):
        self.ctx.check(self.ctx.lib.crux_convq_polyak(self.h, frm.h, float(tau)))

    def copy_from(self, frm):
        self.ctx.check(self.ctx.lib.crux_convq_copy(self.h, frm.h))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.ctx.lib.crux_convq_destroy(self.h)
                self.h = None
        except Exception:
            pass


def fusable(mlp):
    """Shapes served by the fused kernels (csrc/ppo_fused.cu): Chain(Dense(I,64,act), Dense(64,64,act), Dense(64,O)), I <= 32, O <= 8."""
    return (len(mlp.acts) == 3 and mlp.dims[1] == 64 and mlp.dims[2] == 64 and 1 <= mlp.dims[0] <= 32 and 1 <= mlp.dims[3] <= 8
            and mlp.acts[0] == mlp.acts[1] and mlp.acts[0] in (tanh, relu) and mlp.acts[2] == identity)


def _as_dev(ctx, x, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        t = x if x.device == ctx.device else x.to(ctx.device)
        if t.dtype != dtype:
            t = t.to(dtype)
    else:
        t = torch.as_tensor(np.ascontiguousarray(x)).to(dtype).to(ctx.device)
    if t.dim() == 1:
        t = t.reshape(1, -1)
    return t.contiguous()


class ContinuousNetwork(NetworkPolicy):
    """policies.jl:68-98."""

    def __init__(self, network, output_dim=None, ctx=None):
        self.network = network
        self.mlp = _MLP(network, ctx)
        self.ctx = self.mlp.ctx
        self.output_dim = network.layers[-1].out if output_dim is None else output_dim
        self.device = self.ctx.device


class DiscreteNetwork(NetworkPolicy):
    """policies.jl:104-157.  Actions cross as one-hot rows (``a_oh``)."""

    def __init__(self, network, outputs, always_stochastic=False, ctx=None, temperature=1.0, input_dims=None):
        self.network = network
        self.input_dims = input_dims
        if getattr(network, "is_conv", False):
            assert input_dims is not None, "a convolutional DiscreteNetwork needs input_dims = (w, h, c) (Flux infers it from the first input)"
            self.mlp = _ConvNet(network, input_dims, ctx)          # pixel DQN (examples/rl/atari.jl:8)
        else:
            self.mlp = _MLP(network, ctx)
        self.ctx = self.mlp.ctx
        self.outputs = list(outputs)
        self.always_stochastic = always_stochastic
        # logit_conversion = (π, s) -> softmax(value(π, s) ./ α): α = 1 is the default (policies.jl:108), SoftQ sets its α (rl/softq.jl:48)
        self.temperature = float(np.float32(temperature))
        self.device = self.ctx.device
        assert network.layers[-1].out == len(self.outputs)
        self._cat_h = None

    @property
    def h(self):
        """Handle of this network as a categorical ACTOR (on-policy updates: ppo_loss / a2c_loss / reinforce_loss through
        categorical_logpdf and entropy, policies.jl:135,152-155), created on first use."""
        if self._cat_h is None:
            assert isinstance(self.mlp, _MLP), "a categorical actor needs a Chain(Dense...) network"
            assert self.temperature == 1.0, "the on-policy categorical head implements the default logit_conversion softmax(value(π, s))"
            h = C.c_void_p()
            self.ctx.check(self.ctx.lib.crux_categorical_create(self.ctx.h, self.mlp.h, len(self.outputs), C.byref(h)))
            self._cat_h = h
        return self._cat_h

    def __del__(self):
        try:
            if getattr(self, "_cat_h", None) is not None:
                self.ctx.lib.crux_gaussian_destroy(self._cat_h)
                self._cat_h = None
        except Exception:
            pass


class DoubleNetwork(NetworkPolicy):
    """policies.jl:162-187."""

    def __init__(self, N1, N2):
        self.N1, self.N2 = N1, N2
        self.ctx = N1.ctx


class ActorCritic(NetworkPolicy):
    """policies.jl:246-276."""

    def __init__(self, A, C_):
        self.A, self.C = A, C_
        self.ctx = A.ctx


class GaussianPolicy(NetworkPolicy):
    """policies.jl:315-350 with a state-independent ``logΣ`` vector (``ConstantLayer``)."""
    squashed = False

    def __init__(self, mu, log_sigma, always_stochastic=False, ascale=1.0):
        assert isinstance(mu, ContinuousNetwork)
        self.mu, self.ctx = mu, mu.ctx
        self.always_stochastic = always_stochastic
        self.ascale = float(ascale)
        self.adim = mu.output_dim if log_sigma is not None else mu.output_dim // 2
        ls = None if log_sigma is None else np.ascontiguousarray(log_sigma, dtype=np.float32).reshape(-1)
        if ls is not None:
            assert ls.size == self.adim
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_gaussian_create(self.ctx.h, mu.mlp.h, self.adim, ptr(ls), 1 if self.squashed else 0,
                                                         self.ascale, C.byref(h)))
        self.h = h

    @property
    def log_sigma(self):
        """The trainable logΣ vector (device view) or None in head mode."""
        from .device import view
        p = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_gaussian_log_sigma_ptr(self.h, C.byref(p)))
        return view(p.value, (self.adim,), _abi.F32, self.ctx.device) if p.value else None

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.ctx.lib.crux_gaussian_destroy(self.h)
                self.h = None
        except Exception:
            pass


class SquashedGaussianPolicy(GaussianPolicy):
    """policies.jl:355-400.  ``log_sigma=None``: ``mu`` outputs ``[μ | logΣ]`` heads on a shared trunk
    (examples/rl/half_cheetah_mujoco.jl:37-43)."""
    squashed = True

    def __init__(self, mu, log_sigma=None, ascale=1.0, always_stochastic=False):
        super().__init__(mu, log_sigma, always_stochastic, ascale)


# ------------------------------------------------------------------------------------------- generic functions
def device(pi):
    return pi.ctx.device


def actor(pi):
    return pi.A if isinstance(pi, ActorCritic) else pi


def critic(pi):
    return pi.C if isinstance(pi, ActorCritic) else pi


def value(pi, s, a=None):
    """``POMDPs.value`` (policies.jl:94-96,120-122,175-177,259-261)."""
    if isinstance(pi, ActorCritic):
        return value(pi.C, s, a)
    if isinstance(pi, DoubleNetwork):
        return value(pi.N1, s, a), value(pi.N2, s, a)
    if isinstance(pi, DiscreteNetwork):
        q = pi.mlp.forward(s)
        if a is None:
            return q
        a = _as_dev(pi.ctx, a)
        out = pi.ctx.empty((q.shape[0], 1))
        pi.ctx.check(pi.ctx.lib.crux_discrete_q_sa(pi.ctx.h, ptr(q), ptr(a), q.shape[0], q.shape[1], ptr(out)))
        return out
    if isinstance(pi, ContinuousNetwork):
        return pi.mlp.forward(s) if a is None else pi.mlp.forward_sa(s, a)
    raise TypeError(f"value: unsupported policy {type(pi).__name__}")


def action(pi, s):
    """``POMDPs.action`` (policies.jl:92,124,263,331,372).  Discrete: returns action indices ``[B]`` (int32)."""
    if isinstance(pi, ActorCritic):
        return action(pi.A, s)
    if isinstance(pi, GaussianPolicy):
        if pi.always_stochastic:
            return exploration(pi, s)[0]
        s = _as_dev(pi.ctx, s)
        out = pi.ctx.empty((s.shape[0], pi.adim))
        pi.ctx.check(pi.ctx.lib.crux_gaussian_action(pi.h, ptr(s), s.shape[0], ptr(out)))
        return out
    if isinstance(pi, DiscreteNetwork):
        if pi.always_stochastic:
            return exploration(pi, s)[0]
        q = pi.mlp.forward(s)
        idx = pi.ctx.empty((q.shape[0],), torch.int32)
        pi.ctx.check(pi.ctx.lib.crux_discrete_argmax(pi.ctx.h, ptr(q), q.shape[0], q.shape[1], ptr(idx)))
        return idx
    if isinstance(pi, ContinuousNetwork):
        return pi.mlp.forward(s)
    raise TypeError(f"action: unsupported policy {type(pi).__name__}")


def exploration(pi, s, eps=None, seed=0, ctr=0, **_):
    """``exploration(π, s)`` (policies.jl:137-142,265,338-344,388-394) -> (a, logprob).
    ``eps`` injects the noise (Gaussian: ``[B,A]`` normals; discrete: ``[B]`` uniforms); default is device Philox."""
    if isinstance(pi, ActorCritic):
        return exploration(pi.A, s, eps, seed, ctr)
    ctx = pi.ctx
    if isinstance(pi, GaussianPolicy):
        s = _as_dev(ctx, s)
        B = s.shape[0]
        e = None if eps is None else _as_dev(ctx, eps)
        a, lp = ctx.empty((B, pi.adim)), ctx.empty((B, 1))
        ctx.check(ctx.lib.crux_gaussian_explore(pi.h, ptr(s), B, ptr(e), seed, ctr, ptr(a), ptr(lp)))
        return a, lp
    if isinstance(pi, DiscreteNetwork):
        q = pi.mlp.forward(s)
        B, nA = q.shape
        u = None if eps is None else _as_dev(ctx, np.asarray(eps, dtype=np.float64).reshape(-1, 1), torch.float64)
        idx, lp = ctx.empty((B,), torch.int32), ctx.empty((B, 1))
        ctx.check(ctx.lib.crux_discrete_explore_t(ctx.h, ptr(q), B, nA, pi.temperature, ptr(u), seed, ctr, ptr(idx), ptr(lp)))
        return idx, lp
    raise TypeError(f"exploration: unsupported policy {type(pi).__name__}")


def logpdf(pi, s, a):
    """``Distributions.logpdf`` (policies.jl:144-150,267,346,396) -> ``[B,1]``."""
    if isinstance(pi, ActorCritic):
        return logpdf(pi.A, s, a)
    ctx = pi.ctx
    s, a = _as_dev(ctx, s), _as_dev(ctx, a)
    out = ctx.empty((s.shape[0], 1))
    if isinstance(pi, GaussianPolicy):
        ctx.check(ctx.lib.crux_gaussian_logpdf(pi.h, ptr(s), ptr(a), s.shape[0], ptr(out)))
        return out
    if isinstance(pi, DiscreteNetwork):
        q = pi.mlp.forward(s)
        ctx.check(ctx.lib.crux_discrete_logpdf_t(ctx.h, ptr(q), ptr(a), q.shape[0], q.shape[1], pi.temperature, ptr(out)))
        return out
    raise TypeError(f"logpdf: unsupported policy {type(pi).__name__}")


def entropy(pi, s):
    """``Distributions.entropy``: GaussianPolicy -> 0-dim scalar tensor (policies.jl:348; test/policy_tests.jl:271-272);
    squashed / discrete -> ``[B,1]`` (:398, :152-155)."""
    if isinstance(pi, ActorCritic):
        return entropy(pi.A, s)
    ctx = pi.ctx
    s = _as_dev(ctx, s)
    B = s.shape[0]
    if isinstance(pi, GaussianPolicy):
        scalar = pi.log_sigma is not None
        out = ctx.empty((1,)) if scalar else ctx.empty((B, 1))
        ctx.check(ctx.lib.crux_gaussian_entropy(pi.h, ptr(s), B, ptr(out)))
        return out.reshape(()) if scalar else out
    if isinstance(pi, DiscreteNetwork):
        q = pi.mlp.forward(s)
        out = ctx.empty((B, 1))
        ctx.check(ctx.lib.crux_discrete_entropy_t(ctx.h, ptr(q), B, q.shape[1], pi.temperature, ptr(out)))
        return out
    raise TypeError(f"entropy: unsupported policy {type(pi).__name__}")


def _mlps(pi):
    if isinstance(pi, ActorCritic):
        return _mlps(pi.A) + _mlps(pi.C)
    if isinstance(pi, DoubleNetwork):
        return _mlps(pi.N1) + _mlps(pi.N2)
    if isinstance(pi, GaussianPolicy):
        return [pi.mu.mlp]
    return [pi.mlp]


def polyak_average_(to, frm, tau=1.0):
    """``polyak_average!(to, from, τ)`` policies.jl:48-59: to ← τ·from + (1-τ)·to for every parameter array."""
    for t, f in zip(_mlps(to), _mlps(frm)):
        t.polyak_from(f, tau)
    if isinstance(actor(to), GaussianPolicy) and actor(to).log_sigma is not None:
        lt, lf = actor(to).log_sigma, actor(frm).log_sigma
        lt.copy_(np.float32(tau) * lf + (np.float32(1) - np.float32(tau)) * lt)


def copyto_(to, frm):
    """``copyto!(to, from)`` policies.jl:61-65."""
    for t, f in zip(_mlps(to), _mlps(frm)):
        t.copy_from(f)
    if isinstance(actor(to), GaussianPolicy) and actor(to).log_sigma is not None:
        actor(to).log_sigma.copy_(actor(frm).log_sigma)


def deepcopy(pi):
    """``Base.deepcopy(::NetworkPolicy)`` policies.jl:24-26: an independent copy on the same device."""
    def chain_of(mlp):
        flat, layers, off = mlp.get_flat(), [], 0
        for l in range(len(mlp.acts)):
            i, o = mlp.dims[l], mlp.dims[l + 1]
            W = flat[off:off + i * o].reshape(i, o).T.copy(); off += i * o
            b = flat[off:off + o].copy(); off += o
            layers.append(Dense(i, o, mlp.acts[l], W, b))
        return Chain(*layers)
    if isinstance(pi, ActorCritic):
        return ActorCritic(deepcopy(pi.A), deepcopy(pi.C))
    if isinstance(pi, DoubleNetwork):
        return DoubleNetwork(deepcopy(pi.N1), deepcopy(pi.N2))
    if isinstance(pi, SquashedGaussianPolicy):
        ls = None if pi.log_sigma is None else pi.log_sigma.cpu().numpy()
        return SquashedGaussianPolicy(deepcopy(pi.mu), ls, pi.ascale, pi.always_stochastic)
    if isinstance(pi, GaussianPolicy):
        return GaussianPolicy(deepcopy(pi.mu), pi.log_sigma.cpu().numpy(), pi.always_stochastic)
    if isinstance(pi, DiscreteNetwork):
        if isinstance(pi.mlp, _ConvNet):
            cp = DiscreteNetwork(pi.network, pi.outputs, pi.always_stochastic, pi.ctx, pi.temperature, pi.input_dims)
            cp.mlp.copy_from(pi.mlp)
            return cp
        return DiscreteNetwork(chain_of(pi.mlp), pi.outputs, pi.always_stochastic, pi.ctx, pi.temperature)
    if isinstance(pi, ContinuousNetwork):
        return ContinuousNetwork(chain_of(pi.mlp), pi.output_dim, pi.ctx)
    raise TypeError(type(pi))


def action_space(pi):
    from .spaces import ContinuousSpace, DiscreteSpace
    if isinstance(pi, ActorCritic):
        return action_space(pi.A)
    if isinstance(pi, DoubleNetwork):
        return action_space(pi.N1)
    if isinstance(pi, GaussianPolicy):
        return ContinuousSpace(pi.adim)
    if isinstance(pi, DiscreteNetwork):
        return DiscreteSpace(len(pi.outputs), pi.outputs)
    return ContinuousSpace(pi.output_dim)


class PolicyParams:
    """policies.jl:12-19."""

    def __init__(self, pi, space=None, pi_explore=None, pi_target=None, pa=None):
        self.pi = pi
        self.space = action_space(pi) if space is None else space
        self.pi_explore = pi if pi_explore is None else pi_explore
        self.pi_target = pi_target
        self.pa = pa


# ------------------------------------------------------------------------------------------- exploration policies
class LinearDecaySchedule:
    """utils.jl:116-126."""

    def __init__(self, start, stop, steps):
        self.start, self.stop, self.steps = start, stop, int(steps)

    def __call__(self, i):
        rate = (self.start - self.stop) / self.steps
        return max(self.stop, self.start - i * rate)


class MixedPolicy(Policy):
    """ϵ-greedy (policies.jl:466-494) over N env streams at once: one (coin, pick) pair per stream."""

    def __init__(self, eps, n_actions):
        self.eps = eps if callable(eps) else (lambda i, e=eps: e)
        self.n_actions = n_actions

    def exploration(self, s, pi_on, i, u=None, seed=0, ctr=0):
        ctx = pi_on.ctx
        net = actor(pi_on)
        q = net.mlp.forward(s)
        B, nA = q.shape
        idx, oh, lp = ctx.empty((B,), torch.int32), ctx.empty((B, nA)), ctx.empty((B, 1))
        uu = None if u is None else _as_dev(ctx, np.asarray(u, dtype=np.float64).reshape(B, 2), torch.float64)
        ctx.check(ctx.lib.crux_discrete_eps_greedy(ctx.h, ptr(q), B, nA, float(self.eps(i)), ptr(uu), seed, ctr, ptr(idx), ptr(oh), ptr(lp)))
        return idx, oh, lp


def eps_greedy_policy(eps, actions):
    """``ϵGreedyPolicy(ϵ, actions)`` policies.jl:472."""
    return MixedPolicy(eps, len(actions))


class GaussianNoiseExplorationPolicy(Policy):
    """policies.jl:499-514: a = clamp(π(s) + clamp(σ(i)·ε, ε_min, ε_max), a_min, a_max); logprob = NaN."""

    def __init__(self, sigma=0.01, a_min=-math.inf, a_max=math.inf, eps_min=-math.inf, eps_max=math.inf):
        self.sigma = sigma if callable(sigma) else (lambda i, s=sigma: s)
        self.a_min, self.a_max, self.eps_min, self.eps_max = a_min, a_max, eps_min, eps_max

    def bounds(self):
        """(a_min, a_max) as float32 host vectors, or None for an unbounded side (n = 0 across the ABI)."""
        f = lambda v, inf: None if np.all(np.asarray(v, dtype=np.float64) == inf) else np.ascontiguousarray(np.atleast_1d(v), dtype=np.float32)
        return f(self.a_min, -math.inf), f(self.a_max, math.inf)

    def exploration(self, s, pi_on, i, eps=None, seed=0, ctr=0):
        ctx = pi_on.ctx
        a = action(pi_on, s).contiguous().clone()   # the noise is applied in place, never on the network's own output buffer
        B, A = a.shape
        e = None if eps is None else _as_dev(ctx, np.ascontiguousarray(eps, dtype=np.float32))
        lo, hi = self.bounds()
        ctx.check(ctx.lib.crux_noise_explore(ctx.h, ptr(a), B, A, float(np.float32(self.sigma(i))), float(self.eps_min), float(self.eps_max),
                                             ptr(lo), 0 if lo is None else lo.size, ptr(hi), 0 if hi is None else hi.size, ptr(e), seed, ctr))
        return a, float("nan")


class FirstExplorePolicy(Policy):
    """policies.jl:518-534."""

    def __init__(self, N, initial_policy, after_policy=None):
        self.N, self.initial_policy, self.after_policy = N, initial_policy, after_policy

    def resolve(self, i):
        """The policy that acts at step ``i``: (policy, True) = take its plain ``action`` (no exploration noise, logprob NaN),
        (policy, False) = call its own ``exploration``; ``policy is None`` stands for the on-policy network."""
        if i < self.N:
            return self.initial_policy, True
        if self.after_policy is None:
            return None, True
        return self.after_policy, False

    def exploration(self, s, pi_on, i, **kw):
        pol, plain = self.resolve(i)
        if plain:   # policies.jl:527-530: action(π.initial_policy, s) / action(π_on, s), NaN
            return action(pi_on if pol is None else pol, s), float("nan")
        return pol.exploration(s, pi_on, i, **kw)   # a MixedPolicy returns (index, one-hot, logprob): Sampler._act dispatches on the resolved policy
