"""builds tests/native/hostcheck.cpp (g++) — the host-compiled halves of the device headers, for CPU-side bit-exact checks"""
import ctypes as C
import glob
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "hostcheck.cpp")
LIB = os.path.join(HERE, "native", "libhostcheck.so")
DEPS = [SRC] + glob.glob(os.path.join(HERE, "..", "2fast2q_b200", "csrc", "*.h")) + [os.path.join(HERE, "..", "include", "f2q.h")]
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-o", LIB, SRC])
        _lib = C.CDLL(LIB)
    return _lib
