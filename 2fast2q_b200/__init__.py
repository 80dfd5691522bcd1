"""
2fast2q_b200 — B200-native (sm_100a) read -> feature -> count engine with 2FAST2Q's drop-in surface.

    _lib.py     ctypes binding of libf2q.so (include/f2q.h) + Engine wrapper
    fast2q.py   host-side mirror of the reference's interface for this path (reads_counter, features_loader, CLI, csv)
    csrc/       CUDA kernels and the C-ABI
    multi.py    rank sharding and merge (torch.distributed)
    synth.py    numpy restatement of the synthetic generator K0 (bench / tests)
    build.py    nvcc build of libf2q.so

The directory name starts with a digit (it is the project's name); import it with
importlib.import_module("2fast2q_b200").
"""
from . import _lib  # noqa: F401
from ._lib import Engine, F2QError, make_config  # noqa: F401

__version__ = "0.1.0"
