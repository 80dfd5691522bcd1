"""
TEST INFRASTRUCTURE ONLY — ctypes front end of oracle/f2q_oracle.c, the CPU restatement of the
reference's read -> feature -> count path (fast2q/fast2q.py:215-285, 306-409, 514-582, 601-750).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module.  The product (2fast2q_b200/) never does; it fails loudly without its CUDA library.

Parity status: pinned — see tests/golden/make_golden.py and tests/test_oracle_golden.py.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "f2q_oracle.c")
_LIB = os.path.join(_HERE, "libf2q_oracle.so")

MAX_ITER = 8
MAX_DELIM = 64
N_STATS = 5
STAT_NAMES = ("reads", "perfect_counter", "imperfect_counter", "non_aligned_counter", "quality_failed")


class Config(C.Structure):
    """Mirror of struct f2q_config (include/f2q.h)."""
    _fields_ = [
        ("mode", C.c_int32), ("miss", C.c_int32), ("phred", C.c_int32), ("qual_up", C.c_int32),
        ("qual_down", C.c_int32), ("miss_up", C.c_int32), ("miss_down", C.c_int32), ("length", C.c_int32),
        ("n_iter", C.c_int32), ("has_up", C.c_int32), ("has_down", C.c_int32),
        ("starts", C.c_int32 * MAX_ITER), ("up_len", C.c_int32 * MAX_ITER), ("down_len", C.c_int32 * MAX_ITER),
        ("up", (C.c_uint8 * MAX_DELIM) * MAX_ITER), ("down", (C.c_uint8 * MAX_DELIM) * MAX_ITER),
    ]


def build(force: bool = False) -> str:
    """gcc the restatement into oracle/libf2q_oracle.so (git-ignored; travels to the GPU box)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-fvisibility=hidden",
                               "-o", _LIB, _SRC])
    return _LIB


_lib = None
_lock = threading.Lock()


def lib():
    global _lib
    with _lock:
        if _lib is None:
            build()
            L = C.CDLL(_LIB)
            u8p, u64p, i32p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
            L.f2qo_count.argtypes = [C.POINTER(Config), C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64,
                                     C.c_void_p, C.c_void_p]
            L.f2qo_count.restype = C.c_int
            L.f2qo_ec_run.argtypes = [C.POINTER(Config), C.c_void_p, C.c_uint64, C.c_void_p]
            L.f2qo_ec_run.restype = C.c_void_p
            L.f2qo_ec_size.argtypes = [C.c_void_p, u64p, u64p]
            L.f2qo_ec_get.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            L.f2qo_ec_free.argtypes = [C.c_void_p]
            L.f2qo_border_finder.argtypes = [C.c_char_p, C.c_int32, C.c_char_p, C.c_int32, C.c_int32, C.c_int32]
            L.f2qo_border_finder.restype = C.c_int
            L.f2qo_sequence_tinder.argtypes = [C.POINTER(Config), C.c_int32, C.c_char_p, C.c_int32, C.c_char_p,
                                               C.c_int32, C.c_void_p, C.c_void_p, i32p, i32p]
            L.f2qo_sequence_tinder.restype = C.c_int
            L.f2qo_build_key.argtypes = [C.POINTER(Config), C.c_char_p, C.c_int32, C.c_char_p, C.c_int32,
                                         C.c_char_p, i32p]
            L.f2qo_build_key.restype = C.c_int
            _lib = L
    return _lib


def make_config(mode="C", miss=1, phred=30, length=20, start="0", upstream=None, downstream=None,
                miss_search_up=0, miss_search_down=0, qual_up=30, qual_down=30) -> Config:
    """Derive the hot-path parameters the way reads_counter does (fast2q.py:538-558)."""
    c = Config()
    c.mode = 1 if "EC" in str(mode).upper() else 0
    c.miss, c.phred, c.length = int(miss), int(phred), int(length)
    c.qual_up, c.qual_down = int(qual_up), int(qual_down)
    c.miss_up, c.miss_down = int(miss_search_up), int(miss_search_down)
    if upstream is None and downstream is None:
        st = [int(n) for n in str(start).split(",")]
        if len(st) > MAX_ITER:
            raise ValueError("too many --st items")
        c.n_iter = len(st)
        for i, s in enumerate(st):
            c.starts[i] = s
    else:
        ups = [u.upper().encode() for u in upstream.split(",")] if upstream is not None else []
        downs = [d.upper().encode() for d in downstream.split(",")] if downstream is not None else []
        if upstream is not None and downstream is not None and len(ups) != len(downs):
            raise ValueError("Up and Downstream sequences must be submitted in concurrent pairs")
        c.has_up, c.has_down = int(upstream is not None), int(downstream is not None)
        c.n_iter = max(len(ups), len(downs))
        if c.n_iter > MAX_ITER:
            raise ValueError("too many search sequences")
        for i, u in enumerate(ups):
            if len(u) > MAX_DELIM:
                raise ValueError("search sequence too long")
            c.up_len[i] = len(u)
            for j, b in enumerate(u):
                c.up[i][j] = b
        for i, d in enumerate(downs):
            if len(d) > MAX_DELIM:
                raise ValueError("search sequence too long")
            c.down_len[i] = len(d)
            for j, b in enumerate(d):
                c.down[i][j] = b
    return c


def pack_keys(keys):
    """list of bytes/str -> (uint8 blob, uint64 offsets[n+1])"""
    bs = [k.encode() if isinstance(k, str) else bytes(k) for k in keys]
    off = np.zeros(len(bs) + 1, dtype=np.uint64)
    if bs:
        off[1:] = np.cumsum([len(b) for b in bs], dtype=np.uint64)
    blob = np.frombuffer(b"".join(bs) + b"\0", dtype=np.uint8).copy()
    return blob, off


def _as_u8(data):
    if isinstance(data, np.ndarray):
        a = np.ascontiguousarray(data, dtype=np.uint8)
    else:
        a = np.frombuffer(bytes(data), dtype=np.uint8)
    return a


def count(cfg: Config, keys, data):
    """Counter mode over one whole uncompressed FASTQ byte string.
    returns (counts uint64[n_keys], stats dict) — the fastq_parser result (fast2q.py:409)."""
    blob, off = pack_keys(keys)
    a = _as_u8(data)
    counts = np.zeros(max(len(keys), 1), dtype=np.uint64)
    stats = np.zeros(N_STATS, dtype=np.uint64)
    rc = lib().f2qo_count(C.byref(cfg), blob.ctypes.data, off.ctypes.data, len(keys),
                          a.ctypes.data if a.size else None, a.size, counts.ctypes.data, stats.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"oracle f2qo_count failed: {rc}")
    return counts[:len(keys)], dict(zip(STAT_NAMES, (int(x) for x in stats)))


def extract_count(cfg: Config, data):
    """Extract+Count mode.  returns (dict key(bytes)->count in insertion order, stats dict)."""
    a = _as_u8(data)
    stats = np.zeros(N_STATS, dtype=np.uint64)
    L = lib()
    h = L.f2qo_ec_run(C.byref(cfg), a.ctypes.data if a.size else None, a.size, stats.ctypes.data)
    if not h:
        raise MemoryError("oracle f2qo_ec_run failed")
    try:
        n, nb = C.c_uint64(), C.c_uint64()
        L.f2qo_ec_size(h, C.byref(n), C.byref(nb))
        kb = np.zeros(nb.value + 1, dtype=np.uint8)
        ko = np.zeros(n.value + 1, dtype=np.uint64)
        cnt = np.zeros(n.value + 1, dtype=np.uint64)
        L.f2qo_ec_get(h, kb.ctypes.data, ko.ctypes.data, cnt.ctypes.data)
    finally:
        L.f2qo_ec_free(h)
    raw = kb.tobytes()
    out = {}
    for j in range(n.value):
        out[raw[int(ko[j]):int(ko[j + 1])]] = int(cnt[j])
    return out, dict(zip(STAT_NAMES, (int(x) for x in stats)))


def border_finder(seq: bytes, read: bytes, mismatch: int, start_place: int = 0):
    p = lib().f2qo_border_finder(seq, len(seq), read, len(read), mismatch, start_place)
    return None if p < 0 else p


def byteset_mask(chars):
    """set of characters/bytes -> uint64[4] 256-bit mask"""
    m = np.zeros(4, dtype=np.uint64)
    for ch in chars:
        b = ord(ch) if isinstance(ch, str) else int(ch)
        m[b >> 6] |= np.uint64(1) << np.uint64(b & 63)
    return m


def sequence_tinder(cfg: Config, read: bytes, qual: bytes, i: int = 0, set_up=None, set_down=None):
    """set_up / set_down: explicit fail sets (iterables of characters) overriding --qsu/--qsd"""
    s, e = C.c_int32(), C.c_int32()
    mu = byteset_mask(set_up) if set_up is not None else None
    md = byteset_mask(set_down) if set_down is not None else None
    ok = lib().f2qo_sequence_tinder(C.byref(cfg), i, read, len(read), qual, len(qual),
                                    mu.ctypes.data if mu is not None else None,
                                    md.ctypes.data if md is not None else None, C.byref(s), C.byref(e))
    return (s.value, e.value) if ok else (None, None)


def build_key(cfg: Config, read: bytes, qual: bytes):
    buf = C.create_string_buffer((len(read) + 2) * max(1, cfg.n_iter) + 16)
    kl = C.c_int32()
    ok = lib().f2qo_build_key(C.byref(cfg), read, len(read), qual, len(qual), buf, C.byref(kl))
    return buf.raw[:kl.value] if ok else None
