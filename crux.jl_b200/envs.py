"""Vectorised environments driven by ``Sampler`` (the role POMDPs.jl models play for the reference, src/sampler.jl:39-50,89-97).

The reference steps ONE ``mdp`` through ``@gen(:sp,:r)`` / ``isterminal`` / ``initialstate``.  Here an environment object
steps N independent copies at once (each copy is one reference ``Sampler`` stream):

    env.n_envs, env.obs_dim, env.gamma (= POMDPs.discount), env.action_space
    env.reset(idx | None) -> obs[len(idx), obs_dim]         # rand(initialstate(mdp)) + convert_s for those streams
    env.step(a)           -> (sp, r, done)                  # @gen(:sp,:r)(mdp, s, a), convert_s, isterminal

Host environments exchange numpy arrays; a device environment (``on_device = True``) exchanges device pointers and its
step never leaves the GPU.  These synthetic MDPs are NOT part of the reference: they are the benchmark workloads of
SURVEY 8d ("LinQuad-17x6") and the restated ``SimpleGridWorld`` of the README example.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .device import default_context, ptr
from .spaces import ContinuousSpace, DiscreteSpace

F32 = np.float32


def linquad_matrices(obs_dim=17, act_dim=6, seed=0):
    """A = 0.95 I + 0.02 G1, B = 0.1 G2, G ~ default_rng(seed).standard_normal (float32)  (SURVEY 8d)."""
    rng = np.random.default_rng(seed)
    G1 = rng.standard_normal((obs_dim, obs_dim)).astype(F32)
    G2 = rng.standard_normal((obs_dim, act_dim)).astype(F32)
    A = (F32(0.95) * np.eye(obs_dim, dtype=F32) + F32(0.02) * G1).astype(F32)
    B = (F32(0.1) * G2).astype(F32)
    return np.ascontiguousarray(A), np.ascontiguousarray(B)


class HostLinQuad:
    """LinQuad on the host (numpy): s' = clip(A s + B tanh(a) + 0.01 ξ, -10, 10); r = 1 - |s'|²/S - 0.1|a|²/A;
    terminal if |s'_1| > 5; s0 ~ U(-0.1, 0.1)^S; γ = 0.99."""
    on_device = False

    def __init__(self, n_envs, obs_dim=17, act_dim=6, seed=0, gamma=0.99, cost_threshold=None):
        self.n_envs, self.obs_dim, self.act_dim = int(n_envs), obs_dim, act_dim
        # optional safety signal like the reference's safety-gym envs (sampler.jl:76-78,114): info["cost"] = 1 when |s'_2| exceeds
        # the threshold; read by the Sampler when the rollout has a :cost column (LagrangePPO)
        self.cost_threshold = cost_threshold
        self.last_info = {}
        self.A, self.B = linquad_matrices(obs_dim, act_dim, 0)
        self.AT, self.BT = np.ascontiguousarray(self.A.T), np.ascontiguousarray(self.B.T)
        self.gamma = F32(gamma)
        self.rng = np.random.default_rng(seed)
        self.action_space = ContinuousSpace(act_dim)
        self.state = np.zeros((self.n_envs, obs_dim), dtype=F32)

    def reset(self, idx=None):
        n = self.n_envs if idx is None else len(idx)
        s0 = ((self.rng.random((n, self.obs_dim), dtype=F32) * F32(2) - F32(1)) * F32(0.1)).astype(F32)
        if idx is None:
            self.state[:] = s0
        else:
            self.state[idx] = s0
        return s0

    def step(self, a):
        a = np.asarray(a, dtype=F32)
        xi = self.rng.standard_normal((self.n_envs, self.obs_dim), dtype=F32)
        sp = self.state @ self.AT
        sp += np.tanh(a) @ self.BT
        sp += F32(0.01) * xi
        np.clip(sp, F32(-10), F32(10), out=sp)
        r = (F32(1) - np.einsum("ij,ij->i", sp, sp) / F32(self.obs_dim) - F32(0.1) * np.einsum("ij,ij->i", a, a) / F32(self.act_dim)).astype(F32)
        done = np.abs(sp[:, 0]) > F32(5)
        self.state = sp
        if self.cost_threshold is not None:
            self.last_info = {"cost": (np.abs(sp[:, 1]) > F32(self.cost_threshold)).astype(F32)}
        return sp, r, done


class NativeHostLinQuad:
    """LinQuad on the host in C++ (csrc/host/linquad_host.cpp -> lib/libcrux_hostenv.so): N streams stepped by a pool of
    worker threads, same Philox noise streams as ``DeviceLinQuad``.  This is the "vectorised CPU env step of a synthetic
    MDP" of the north-star's e2e path; ``step_into`` writes straight into the sampler's pinned staging buffers."""
    on_device = False
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            import os
            libdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib")
            path = os.path.join(libdir, "libcrux_hostenv.so")
            try:  # the widest libmvec-vectorised build the CPU supports (CRUX_HOSTENV_ISA=baseline|avx2|avx512 overrides)
                flags = open("/proc/cpuinfo").read()
                want = os.environ.get("CRUX_HOSTENV_ISA", "baseline" if os.environ.get("CRUX_HOSTENV_BASELINE") else "")
                fast2 = os.path.join(libdir, "libcrux_hostenv_avx2.so")
                fast5 = os.path.join(libdir, "libcrux_hostenv_avx512.so")
                has5 = all(f" {f}" in flags for f in ("avx512f", "avx512dq", "avx512vl", "avx512bw")) and " fma" in flags
                has2 = " avx2" in flags and " fma" in flags
                if want in ("", "avx512") and has5 and os.path.exists(fast5):
                    path = fast5
                elif want in ("", "avx2", "avx512") and has2 and os.path.exists(fast2):
                    path = fast2
            except OSError:
                pass
            if not os.path.exists(path):
                raise ImportError(f"{path} is missing: run `python crux.jl_b200/build.py`")
            L = C.CDLL(path)
            L.crux_hostenv_create.restype = C.c_void_p
            L.crux_hostenv_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int]
            L.crux_hostenv_destroy.argtypes = [C.c_void_p]
            L.crux_hostenv_threads.argtypes = [C.c_void_p]
            L.crux_hostenv_reset.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
            L.crux_hostenv_step.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            cls._lib = L
        return cls._lib

    def __init__(self, n_envs, obs_dim=17, act_dim=6, seed=0, gamma=0.99, n_threads=0):
        self.n_envs, self.obs_dim, self.act_dim = int(n_envs), obs_dim, act_dim
        self.gamma = F32(gamma)
        self.action_space = ContinuousSpace(act_dim)
        A, B = linquad_matrices(obs_dim, act_dim, 0)
        self.h = self.lib().crux_hostenv_create(self.n_envs, obs_dim, act_dim, A.ctypes.data, B.ctypes.data, int(seed), int(n_threads))
        assert self.h, "crux_hostenv_create failed"
        self.n_threads = self.lib().crux_hostenv_threads(self.h)
        self._sp = np.empty((self.n_envs, obs_dim), F32)
        self._r = np.empty(self.n_envs, F32)
        self._done = np.empty(self.n_envs, np.uint8)

    def reset(self, idx=None):
        if idx is None:
            out = np.empty((self.n_envs, self.obs_dim), F32)
            self.lib().crux_hostenv_reset(self.h, None, 0, out.ctypes.data)
            return out
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        out = np.empty((len(idx), self.obs_dim), F32)
        self.lib().crux_hostenv_reset(self.h, idx.ctypes.data, len(idx), out.ctypes.data)
        return out

    def c_callbacks(self):
        """(step_fn, reset_fn, user) C pointers matching crux_env_step_fn / crux_env_reset_fn (include/crux_cuda.h): the
        rollout loop of ``crux_rollout_host`` calls the environment without going through Python."""
        L = self.lib()
        return (C.cast(L.crux_hostenv_step_range, C.c_void_p), C.cast(L.crux_hostenv_reset, C.c_void_p), C.c_void_p(self.h))

    def step_into(self, a, sp, r, done):
        """a [N, act] f32, sp [N, obs] f32, r [N] f32, done [N] u8: contiguous numpy arrays (may be pinned memory)."""
        self.lib().crux_hostenv_step(self.h, a.ctypes.data, sp.ctypes.data, r.ctypes.data, done.ctypes.data)

    def step(self, a):
        a = np.ascontiguousarray(a, dtype=F32)
        self.step_into(a, self._sp, self._r, self._done)
        return self._sp.copy(), self._r.copy(), self._done.astype(bool)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib().crux_hostenv_destroy(self.h)
                self.h = None
        except Exception:
            pass


class DeviceLinQuad:
    """The same MDP stepped on the GPU (``crux_linquad_*``): observations, actions and transitions never leave HBM.
    Episode bookkeeping (episode_length, max_steps, reset) is done inside the step kernel like ``step!`` does
    (sampler.jl:130-136)."""
    on_device = True

    def __init__(self, n_envs, obs_dim=17, act_dim=6, seed=0, gamma=0.99, max_steps=1000, ctx=None):
        self.ctx = ctx or default_context()
        self.n_envs, self.obs_dim, self.act_dim = int(n_envs), obs_dim, act_dim
        self.gamma, self.max_steps = F32(gamma), int(max_steps)
        self.action_space = ContinuousSpace(act_dim)
        A, B = linquad_matrices(obs_dim, act_dim, 0)
        h = C.c_void_p()
        self.ctx.check(self.ctx.lib.crux_linquad_create(self.ctx.h, obs_dim, act_dim, ptr(A), ptr(B), self.n_envs, self.max_steps,
                                                        int(seed), C.byref(h)))
        self.h = h

    def reset_into(self, obs_out):
        self.ctx.check(self.ctx.lib.crux_linquad_reset(self.h, ptr(obs_out)))

    def step_into(self, obs, a, sp, r, done, episode_end, next_obs, force_end=False):
        self.ctx.check(self.ctx.lib.crux_linquad_step(self.h, ptr(obs), ptr(a), ptr(sp), ptr(r), ptr(done), ptr(episode_end),
                                                      ptr(next_obs), 1 if force_end else 0))

    def rollout_into(self, actor_pi, T, force_end_last, obs_io, data, seed, ctr0):
        """T vector steps (policy forward + sample + transition + bookkeeping) in one persistent launch."""
        from . import _abi
        lp = data.get("logprob")
        cols = _abi.RolloutCols(data["s"].data_ptr(), data["a"].data_ptr(), data["sp"].data_ptr(), data["r"].data_ptr(), data["done"].data_ptr(),
                                data["episode_end"].data_ptr(), lp.data_ptr() if lp is not None else None)
        self.ctx.check(self.ctx.lib.crux_linquad_rollout(self.h, actor_pi.h, int(T), 1 if force_end_last else 0, ptr(obs_io), C.byref(cols),
                                                         int(seed), int(ctr0)))

    def __del__(self):
        try:
            if getattr(self, "h", None) and self.ctx.h:
                self.ctx.lib.crux_linquad_destroy(self.h)
                self.h = None
        except Exception:
            pass


class SimpleGridWorld:
    """POMDPModels.SimpleGridWorld restated [3P]: 10x10 grid, actions up/down/left/right, rewards
    {(4,3): -10, (4,6): -5, (9,3): +10, (8,8): +3} (acting from a reward cell pays it and ends the episode), the move
    succeeds with probability 0.7 (otherwise one of the other three directions, uniformly), walls keep the agent in
    place, γ = 0.95, initial state uniform over the grid.  ``convert_s`` gives the 2-vector (x, y)
    (test/spaces_tests.jl:42-43).  N independent copies."""
    on_device = False
    MOVES = np.array([[0, 1], [0, -1], [-1, 0], [1, 0]])  # :up, :down, :left, :right

    def __init__(self, n_envs=1, size=(10, 10), tprob=0.7, gamma=0.95, seed=0, rewards=None):
        self.n_envs, self.size, self.tprob = int(n_envs), size, tprob
        self.obs_dim, self.gamma = 2, F32(gamma)
        self.rewards = rewards or {(4, 3): -10.0, (4, 6): -5.0, (9, 3): 10.0, (8, 8): 3.0}
        self.rmap = np.zeros((size[0] + 1, size[1] + 1), dtype=F32)
        for (x, y), v in self.rewards.items():
            self.rmap[x, y] = v
        self.action_space = DiscreteSpace(4, [0, 1, 2, 3])
        self.rng = np.random.default_rng(seed)
        self.state = np.ones((self.n_envs, 2), dtype=np.int64)

    def reset(self, idx=None):
        n = self.n_envs if idx is None else len(idx)
        s0 = np.stack([self.rng.integers(1, self.size[0] + 1, n), self.rng.integers(1, self.size[1] + 1, n)], 1)
        if idx is None:
            self.state[:] = s0
        else:
            self.state[idx] = s0
        return s0.astype(F32)

    def step(self, a):
        a = np.asarray(a).astype(np.int64).reshape(-1)
        s = self.state
        r = self.rmap[s[:, 0], s[:, 1]].copy()
        done = r != 0  # acting from a reward cell moves to the terminal state
        ok = self.rng.random(self.n_envs) < self.tprob
        other = (a + self.rng.integers(1, 4, self.n_envs)) % 4
        d = np.where(ok, a, other)
        sp = s + self.MOVES[d]
        sp[:, 0] = np.clip(sp[:, 0], 1, self.size[0])
        sp[:, 1] = np.clip(sp[:, 1], 1, self.size[1])
        sp[done] = -1  # GWPos(-1,-1): the terminal state
        self.state = sp
        return sp.astype(F32), r, done
