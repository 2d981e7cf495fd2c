"""Build libcrux_cuda.so (sm_100a) in-tree with nvcc.  No GPU needed: nvcc cross-compiles."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libcrux_cuda.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# no --use_fast_math: parity is 1e-5 rtol fp32 against the oracle (tanhf/expf/logf stay accurate)
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(LIBDIR, exist_ok=True)
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(HERE, "..", "include", "crux_cuda.h"))
    jobs = []
    objs = []
    for src in _sources():
        s = os.path.join(CSRC, src)
        o = os.path.join(objdir, src[:-3] + ".o")
        objs.append(o)
        if force or _stale(o, [s] + headers):
            jobs.append([NVCC, *FLAGS, "-c", s, "-o", o] + (["-Xptxas", "-v"] if verbose else []))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stderr

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for out in ex.map(run, jobs):
            if verbose and out:
                print(out)
    if jobs or force or _stale(LIB, objs):
        run([NVCC, "-shared", *objs, "-o", LIB, "-lcudart", "-ldl",
             "-gencode", "arch=compute_100a,code=sm_100a"])
    build_host(force)
    return LIB


HOST_LIB = os.path.join(LIBDIR, "libcrux_hostenv.so")
HOST_LIB_AVX2 = os.path.join(LIBDIR, "libcrux_hostenv_avx2.so")
HOST_LIB_AVX512 = os.path.join(LIBDIR, "libcrux_hostenv_avx512.so")


def build_host(force: bool = False) -> str:
    """g++ builds of the host-side synthetic env stepper (no CUDA): a baseline x86-64 variant that runs anywhere, an AVX2+FMA
    variant and an AVX-512 variant (libmvec-vectorised); envs.py selects the widest one /proc/cpuinfo advertises."""
    src = os.path.join(CSRC, "host", "linquad_host.cpp")
    cxx = os.environ.get("CXX", "g++")
    for out, extra in ((HOST_LIB, []), (HOST_LIB_AVX2, ["-mavx2", "-mfma"]),
                       (HOST_LIB_AVX512, ["-mavx512f", "-mavx512dq", "-mavx512vl", "-mavx512bw", "-mfma", "-mprefer-vector-width=512"])):
        if force or _stale(out, [src]):
            r = subprocess.run([cxx, "-O3", "-ffast-math", "-fopenmp-simd", "-std=c++17", "-fPIC", "-shared", "-pthread", *extra, src, "-o", out, "-lm"],
                               capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return HOST_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
