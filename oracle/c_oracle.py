"""ctypes loader of the plain-C oracle (oracle/gae_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_build", "libcrux_oracle.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "gae_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return LIB


def load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_set_threads.argtypes = [C.c_int]
        _lib.oracle_get_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def set_threads(n):
    load().oracle_set_threads(int(n))


def get_threads():
    return load().oracle_get_threads()


def gae_returns(r, done, episode_end, v_s, v_sp, gamma, lam, want_adv=True, want_ret=True):
    """fill_gae!/fill_returns! (sampler.jl:262-281) over a [T, N] rollout."""
    lib = load()
    r = np.ascontiguousarray(r, dtype=np.float32)
    T, N = r.shape
    done = np.ascontiguousarray(done, dtype=np.uint8)
    ee = np.ascontiguousarray(episode_end, dtype=np.uint8)
    v_s = np.ascontiguousarray(v_s, dtype=np.float32)
    v_sp = np.ascontiguousarray(v_sp, dtype=np.float32)
    adv = np.empty((T, N), dtype=np.float32) if want_adv else None
    ret = np.empty((T, N), dtype=np.float32) if want_ret else None
    lib.oracle_gae_returns(_p(r), _p(done), _p(ee), _p(v_s), _p(v_sp), C.c_int64(T), C.c_int64(N),
                           C.c_float(gamma), C.c_float(lam), _p(adv), _p(ret))
    return adv, ret


def per_indices(cumsum, B, rands):
    lib = load()
    cumsum = np.ascontiguousarray(cumsum, dtype=np.float32)
    rands = np.ascontiguousarray(rands, dtype=np.float64)
    ids = np.empty(B, dtype=np.int64)
    lib.oracle_per_indices(_p(cumsum), C.c_int64(cumsum.size), C.c_int64(B), _p(rands), _p(ids))
    return ids


def ring_indices(next0, n, cap):
    lib = load()
    out = np.empty(n, dtype=np.int64)
    lib.oracle_ring_indices(C.c_int64(next0), C.c_int64(n), C.c_int64(cap), _p(out))
    return out
