"""
CPU: the DEFLATE decoder the GPU inflate kernel runs per BGZF block (2fast2q_b200/csrc/inflate_core.h, compiled here with g++)
against zlib: stored, fixed-Huffman and dynamic-Huffman blocks, every compression level, FASTQ text, random bytes, long
matches, empty input, and corrupted / truncated streams (which must be refused, never mis-decoded silently).
"""
import ctypes as C
import importlib
import random
import zlib

import numpy as np
import pytest

import hostcheck

synth = importlib.import_module("2fast2q_b200.synth")
DECODERS = ("hc_inflate_raw", "hc_inflate_lane", "hc_inflate_lane8")   # free-running; lock-step with 9+7 / 8+6 table index bits


def infl(comp: bytes, out_len: int, fn="hc_inflate_raw"):
    H = hostcheck.lib()
    f = getattr(H, fn)
    f.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32]
    out = C.create_string_buffer(max(out_len, 1) + 16)
    rc = f(comp + b"\xAA" * 16, len(comp), out, out_len)              # (readable slack behind the stream, as in the staging buffer)
    return rc, out.raw[:out_len]


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def payloads():
    rnd = random.Random(3)
    spec = synth.default_spec(3)
    _, keys = synth.make_library(3, 300, 20)
    fastq = synth.fixed_reads(keys, 0, 380, **spec).tobytes()          # ~64 KB of FASTQ text: one BGZF block's worth
    yield "fastq", fastq
    yield "fastq_short", fastq[:777]
    yield "empty", b""
    yield "one_byte", b"A"
    yield "random", bytes(rnd.randrange(256) for _ in range(30000))
    yield "runs", b"A" * 40000 + b"CG" * 10000 + b"\n" * 300
    yield "text", (b"the quick brown fox jumps over the lazy dog\n" * 1400)[:65280]


@pytest.mark.parametrize("name,data", list(payloads()), ids=[n for n, _ in payloads()])
def test_decoder_equals_zlib(name, data):
    for level in (0, 1, 4, 6, 9):
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            comp = raw_deflate(data, level, strategy)
            for fn in DECODERS:
                rc, out = infl(comp, len(data), fn)
                assert rc == 0 and out == data, (name, level, strategy, fn)


def test_corrupted_and_truncated_streams_are_refused():
    data = next(d for n, d in payloads() if n == "fastq")
    comp = raw_deflate(data, 6)
    for fn in DECODERS:
        assert infl(comp, len(data), fn)[0] == 0
        assert infl(comp, len(data) - 1, fn)[0] != 0         # produces more than the trailer promised
        assert infl(comp, len(data) + 1, fn)[0] != 0         # produces less
        for cut in (0, 1, 10, len(comp) // 2, len(comp) - 1):
            assert infl(comp[:cut], len(data), fn)[0] != 0, (fn, cut)
    rnd = random.Random(9)
    bad = 0
    for _ in range(200):
        b = bytearray(comp)
        p = rnd.randrange(len(b))
        b[p] ^= 1 << rnd.randrange(8)
        rc, out = infl(bytes(b), len(data))
        for fn in DECODERS[1:]:
            rc2, out2 = infl(bytes(b), len(data), fn)
            assert (rc == 0) == (rc2 == 0) and (rc != 0 or out == out2), fn  # all decoders accept / refuse the same streams
        if rc != 0 or out != data:
            bad += 1
        assert rc != 0 or len(out) == len(data)
    assert bad > 150                                         # (a flipped bit may leave a valid stream of the same length; that is what the CRC is for)
