"""Builds libf2q.so (the CUDA library behind include/f2q.h) in-tree for sm_100a with nvcc."""
from __future__ import annotations

import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libf2q.so")
SOURCES = ["f2q_api.cu"]
HEADERS = ["f2q_dev.cuh", "generic.cuh", "resolve.cuh", "stream.cuh", "tile.cuh", "spec.cuh", "flex.cuh", "flex_core.h", "synth_gen.h", "inflate_core.h",
           os.path.join("..", "..", "include", "f2q.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr", "--extended-lambda",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "-shared", "-cudart", "static", "-lz",
]


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    for f in SOURCES + HEADERS:
        if os.path.getmtime(os.path.join(CSRC, f)) > t:
            return True
    return os.path.getmtime(__file__) > t


def build(force: bool = False, verbose: bool = False) -> str:
    if force or _stale():
        nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
        # F2Q_FAST_BUILD=1: -split-compile halves the build time but costs the streaming kernel ~5 % (measured); development only
        fast = ["-split-compile", "0"] if os.environ.get("F2Q_FAST_BUILD") else []
        cmd = [nvcc] + NVCC_FLAGS + fast + (["-Xptxas", "-v"] if verbose else []) + \
              [os.path.join(CSRC, s) for s in SOURCES] + ["-o", LIB]
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
