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


def test_c_client_struct_layouts_match_ctypes():
    """tests/abi_smoke.c (plain C, built by build()) prints sizeof / offsetof of every struct that crosses the ABI: they must equal the
    ctypes Structures' and the constants the Julia package asserts in julia/CruxB200.jl/test/runtests.jl."""
    import ctypes as C
    import subprocess
    import __graft_entry__ as g
    from crux_b200 import _abi
    exe = g.build_abi_smoke()
    out = subprocess.run([exe, "layout"], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    seen = {}
    for line in out:
        tok = line.split()
        seen[tok[0]] = (int(tok[1]), {tok[i]: int(tok[i + 1]) for i in range(2, len(tok), 2)})
    for name, cls in (("crux_ppo_hp", _abi.PPOHp), ("crux_lagrange_hp", _abi.LagrangeHp), ("crux_rollout_cols", _abi.RolloutCols),
                      ("crux_col_desc", _abi.ColDesc)):
        size, offs = seen[name]
        assert size == C.sizeof(cls), name
        for f, o in offs.items():
            assert getattr(cls, f).offset == o, (name, f)
    assert seen["abi_version"][0] == 1
    # the same numbers, as julia/CruxB200.jl/test/runtests.jl states them
    jl = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "julia", "CruxB200.jl", "test", "runtests.jl")).read()
    for name in ("PPOHp", "LagrangeHp", "RolloutCols", "ColDesc"):
        c_name = {"PPOHp": "crux_ppo_hp", "LagrangeHp": "crux_lagrange_hp", "RolloutCols": "crux_rollout_cols", "ColDesc": "crux_col_desc"}[name]
        assert f"sizeof(CruxB200.{name}) == {seen[c_name][0]}" in jl, name
