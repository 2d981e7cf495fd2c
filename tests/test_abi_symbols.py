"""CPU: libcrux_cuda.so loads and exports every symbol include/crux_cuda.h declares (no compute calls)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "crux_cuda.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(crux_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_entry_points():
    syms = _header_symbols()
    assert len(syms) >= 80
    for must in ("crux_rollout_step", "crux_fill_gae_returns", "crux_ppo_update", "crux_buffer_push", "crux_nccl_init"):
        assert must in syms


def test_library_exports_every_declared_symbol(crux):
    lib = crux._abi.load()
    missing = [s for s in _header_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in include/crux_cuda.h but not exported: {missing}"
    assert lib.crux_abi_version() == 1


def test_python_binding_covers_header(crux):
    assert set(_header_symbols()) == set(crux._abi.declared_symbols())


def test_no_cpu_fallback(crux):
    """Without a GPU a context cannot be created: the product path fails loudly (no oracle / CPU routing)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError):
        crux.Context()
    h = C.c_void_p()
    rc = crux._abi.load().crux_ctx_create(0, None, C.byref(h))
    assert rc != 0
    assert b"CUDA" in crux._abi.load().crux_last_error(None)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "crux.jl_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), fn


def test_split_batches_host_abi(crux):
    # crux_split_batches is host integer arithmetic (experience_buffer.jl:126-131): callable without a GPU
    lib = crux._abi.load()
    for N, fr, want in [(100, [0.5, 0.5], [50, 50]), (100, [1.0], [100]), (100, [1 / 3] * 3, [34, 33, 33]), (10, [0.4, 0.3, 0.3], [4, 3, 3])]:
        out = (C.c_int64 * len(fr))()
        assert lib.crux_split_batches(N, (C.c_double * len(fr))(*fr), len(fr), out) == 0
        assert list(out) == want
    out = (C.c_int64 * 1)()
    assert lib.crux_split_batches(100, (C.c_double * 1)(0.4), 1, out) != 0  # @assert sum(fracs) ≈ 1
