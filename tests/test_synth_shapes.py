"""
CPU: the numpy restatement of the synthetic generator K0 (2fast2q_b200/synth.py) against the C++ function the CUDA kernel
k_synth runs (csrc/synth_gen.h, compiled here with g++ by tests/hostcheck.py) — every shape, bit for bit.  The -m gpu test
test_synth_generator_matches_numpy repeats the comparison against the kernel's own output.
"""
import ctypes as C
import importlib

import numpy as np
import pytest

import hostcheck

synth = importlib.import_module("2fast2q_b200.synth")
lib = importlib.import_module("2fast2q_b200._lib")


def host_generate(guides, first, n, spec):
    s = lib.make_synth_spec(len(guides), first, n, **spec)
    g = np.frombuffer(b"".join(guides), dtype=np.uint8)
    out = np.zeros(n * (2 * spec["read_len"] + 18), dtype=np.uint8)
    H = hostcheck.lib()
    H.hc_synth.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    H.hc_synth(C.byref(s), g.ctypes.data, first, n, out.ctypes.data)
    return out


def workload(config):
    if config in ("2", "3"):
        spec = synth.default_spec(int(config))
        _, keys = synth.make_library(int(config), 700, 20)
        return keys, spec, lambda f, n: synth.fixed_reads(keys, f, n, **spec)
    spec = synth.shape_spec(config)
    if config == "4":
        _, guides = synth.make_library(44, 5000, 20)
    else:
        _, _, xs, ys = synth.dual_keys(400)
        guides = xs + ys
    return guides, spec, lambda f, n: synth.shaped_reads(guides, f, n, **spec)


@pytest.mark.parametrize("config", ["2", "3", "4", "5a", "5b"])
def test_numpy_generator_equals_device_function(config):
    guides, spec, gen = workload(config)
    for first, n in ((0, 4000), (123_456_789_012, 1500)):
        want = host_generate(guides, first, n, spec)
        got = gen(first, n)
        assert got.shape == want.shape
        bad = np.nonzero(got != want)[0]
        assert bad.size == 0, (config, first, int(bad[0]) // (2 * spec["read_len"] + 18))


def test_shapes_look_like_their_configs():
    """the class mixes are what SURVEY.md §8d names: delimiters present in ~97 % of Bar-seq reads, barcodes heavy-tailed"""
    guides, spec, gen = workload("4")
    data = gen(0, 20000).reshape(-1, 168)
    seqs = [bytes(r[14:89]) for r in data]
    both = sum(1 for s in seqs if synth.BARSEQ_US in s and synth.BARSEQ_DS in s)
    assert 0.88 < both / len(seqs) < 0.96          # 5 % carry one substitution, 3 % lack one delimiter
    _, spec5, gen5 = workload("5b")
    d5 = gen5(0, 5000).reshape(-1, 168)
    u1, d1, u2, d2 = synth.DUAL_DELIMS
    ok = sum(1 for r in d5 if u1 in bytes(r[14:89]) and d2 in bytes(r[14:89]))
    assert ok / 5000 > 0.95
