"""ORACLE (test infrastructure only) -- CPU restatement of the pixel-DQN network of the reference's Atari example and of its
optimiser chain.  Parity unpinned by the reference (it holds no vectors for this path and Julia cannot run here); the
restatement is anchored on the published semantics of the third-party pieces it calls:

* ``examples/rl/atari.jl:8``: ``Chain(x -> x ./ 255f0, Conv((8,8), 4=>16, relu, stride=4), Conv((4,4), 16=>32, relu, stride=2), flatten,
  Dense(2048, 256, relu), Dense(256, nA))`` wrapped in a ``DiscreteNetwork`` (``src/policies.jl:104-157``).
* Flux 0.14 ``Conv`` [3P] = ``NNlib.conv`` with ``flipped = false``: a TRUE convolution (the kernel is flipped with respect to
  cross-correlation), weight array ``[kw, kh, cin, cout]``, input ``[w, h, c, batch]`` (column-major: memory ``[b][c][h][w]``),
  no padding, dilation 1.  torch's ``conv2d`` is a cross-correlation, so the restatement flips the kernel in both spatial dims.
* ``Flux.flatten`` [3P]: ``reshape(x, :, size(x)[end])`` of the column-major array = features in ``w + W (h + H c)`` order = torch's
  ``flatten(1)`` of an NCHW tensor.
* ``examples/rl/atari.jl:10``: ``Flux.Optimiser(ClipValue(1f0), Adam(1f-3))`` [3P]: every gradient entry is clamped to ``[-1, 1]``
  (``clamp!(Δ, -thresh, thresh)``) before the Adam rule; ``train!`` (``src/training.jl:13-25``) reports ``norm`` of the RAW gradient.
* the loss is ``td_loss`` (``src/utils.jl:76-87``) on ``value(π, s, a) = sum(value(π, s) .* a, dims=1)`` (``policies.jl:122``).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import crux_oracle as o

F32 = np.float32


class ConvQ:
    """Parameters are kept in Flux memory layout: conv ``W`` as torch ``[cout, cin, kh, kw]`` (the memory of Julia's ``[kw, kh, cin, cout]``),
    Dense ``W`` as ``[in, out]`` row-major (the memory of Julia's ``[out, in]``), exactly like ``o.MLP``."""

    def __init__(self, C, H, W, k1, s1, c1, k2, s2, c2, hidden, nA, rng, scale255=True):
        self.shape = (C, H, W)
        self.s1, self.s2, self.scale255 = s1, s2, scale255
        oh1, ow1 = (H - k1) // s1 + 1, (W - k1) // s1 + 1
        oh2, ow2 = (oh1 - k2) // s2 + 1, (ow1 - k2) // s2 + 1
        self.F = c2 * oh2 * ow2

        def conv_init(co, ci, k):   # Flux.glorot_uniform over the [kw, kh, ci, co] array: fan_in = k*k*ci, fan_out = k*k*co
            lim = np.sqrt(6.0 / (k * k * ci + k * k * co))
            return torch.tensor(((rng.random((co, ci, k, k), dtype=F32) * 2 - 1) * lim).astype(F32), requires_grad=True)
        self.W1, self.b1 = conv_init(c1, C, k1), torch.tensor((rng.standard_normal(c1) * 0.05).astype(F32), requires_grad=True)
        self.W2, self.b2 = conv_init(c2, c1, k2), torch.tensor((rng.standard_normal(c2) * 0.05).astype(F32), requires_grad=True)
        self.head = o.MLP([self.F, hidden, nA], [2, 0], rng)

    def params(self):
        return [self.W1, self.b1, self.W2, self.b2] + self.head.params()

    def _flat(self, get):
        """C-ABI order: conv arrays as they lie in memory, Dense W ([out, in] here like o.MLP) in Julia memory order = row-major [in][out]."""
        conv = [get(p).reshape(-1) for p in (self.W1, self.b1, self.W2, self.b2)]
        head = [np.concatenate([get(w).T.reshape(-1), get(b)]) for w, b in zip(self.head.W, self.head.b)]
        return np.concatenate(conv + head).astype(F32)

    def flat(self):
        return self._flat(lambda p: p.detach().numpy())

    def flat_grads(self):
        return self._flat(lambda p: p.grad.detach().numpy())

    def __call__(self, s):
        """``value(π, s)``: ``s`` is ``[B, C, H, W]`` (uint8 or float32) -> ``[B, nA]``."""
        x = torch.as_tensor(np.asarray(s), dtype=torch.float32).reshape(-1, *self.shape)
        if self.scale255:
            x = x / F32(255.0)
        x = F.relu(F.conv2d(x, torch.flip(self.W1, dims=(2, 3)), self.b1, stride=self.s1))
        x = F.relu(F.conv2d(x, torch.flip(self.W2, dims=(2, 3)), self.b2, stride=self.s2))
        return self.head(x.flatten(1))


class ClipAdam(o.Adam):
    """``Flux.Optimiser(ClipValue(thresh), Adam(η, β, ϵ))``: clamp each gradient entry, then the Adam rule of ``o.Adam``."""

    def __init__(self, thresh, *a, **kw):
        super().__init__(*a, **kw)
        self.thresh = None if thresh is None else float(thresh)

    def step(self, params):
        if self.thresh is not None:
            for p in params:
                if p.grad is not None:
                    p.grad = p.grad.clamp(-self.thresh, self.thresh)
        super().step(params)


def dqn_train_step(net, opt, s, a_onehot, y, weight=None):
    """One ``train!`` (training.jl:15-25) of ``td_loss`` on the pixel network -> info dict (loss, grad_norm of the raw gradient, Qavg)."""
    info = {}
    a = torch.as_tensor(a_onehot, dtype=torch.float32)

    def loss(info_):
        q_sa = torch.sum(net(s) * a, dim=1, keepdim=True)
        return o.td_loss(q_sa, y, weight, info_)
    o.train_step(net.params(), loss, opt, info)
    return info
