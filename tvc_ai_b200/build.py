"""In-tree build of libtvc_b200.so (nvcc, sm_100a only)."""
from __future__ import annotations

import os
import shutil
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB = os.path.join(_HERE, "libtvc_b200.so")
SOURCES = ("tvc_abi.cu", "tvc_rollout.cu", "tvc_replay.cu", "tvc_curiosity.cu")
HEADERS = ("tvc_device.cuh", "tvc_internal.h", "tvc_umma.cuh", os.path.join("..", "..", "include", "tvc_b200.h"))
# -cudart shared: the library imports only the runtime symbols it uses (libcudart.so.12, the same SONAME torch loads)
# instead of embedding the whole static runtime; --threads 2: the two translation units compile side by side
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "shared", "--threads", "4"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libtvc_b200.so cannot be built (there is no fallback path)")


def stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not stale():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + \
          [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)   # the image's CC wrapper is not the nvcc host compiler
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    return LIB


def build_phase_profiler() -> str:
    """Diagnostic build with clock64 phase counters in the legacy CTA-exchange integrate() (tools/phase_prof.py)."""
    out = os.path.join(_HERE, "libtvc_b200_prof.so")
    cmd = [_nvcc()] + NVCC_FLAGS + ["-DTVC_PHASE_PROF", "-o", out] + [os.path.join(CSRC, s) for s in SOURCES]
    env = dict(os.environ)
    env.pop("CC", None), env.pop("CXX", None)
    res = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    return out
