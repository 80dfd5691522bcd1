"""
-m gpu: the native file ingest f2q_submit_file (read / zlib inflate / BGZF block-parallel inflate straight into pinned ring
buffers) against what the reference's `for line in gzip.open(raw, 'rb')` delivers: plain files, single- and multi-member
gzip, zero padding, empty members, bgzip, every kind of truncation (the reference's EOFError semantics, fast2q.py:405-407:
all complete lines before the break are parsed, the partial one is dropped), the preprocess line limit, corrupted data.
"""
import gzip
import importlib
import random

import numpy as np
import pytest

from test_host_mirror import _bgzf, _gz, _py_gzip_lines

pytestmark = pytest.mark.gpu

f2q = importlib.import_module("2fast2q_b200")
lib = f2q._lib
synth = importlib.import_module("2fast2q_b200.synth")


@pytest.fixture(scope="module")
def world(oracle):
    spec = synth.default_spec(2)
    names, keys = synth.make_library(2, 400, 20)
    body = synth.fixed_reads(keys, 0, 60_000, **spec).tobytes()
    eng = lib.Engine(lib.make_config(miss=1), 0)
    eng.set_library(keys)
    ocfg = oracle.make_config(miss=1)

    def check(path, expect_bytes, gz, complete=True, limit=0, threads=4):
        eng.begin()
        ok, nbytes = eng.submit_file(str(path), gz, limit, threads)
        counts, stats = eng.end()
        want_c, want_s = oracle.count(ocfg, keys, expect_bytes)
        assert ok == complete, path
        assert stats == want_s and np.array_equal(counts, want_c), path
        return nbytes

    yield body, check
    eng.close()


def test_plain_and_gzip_variants(world, tmp_path):
    body, check = world
    p = tmp_path / "a.fastq"
    p.write_bytes(body)
    assert check(p, body, False) == len(body)
    p.write_bytes(body + b"@last\nACGTACGTACGTACGTACGTAAAA\n+\nIIIIIIIIIIIIIIIIIIIIIIII")       # unterminated final line still counts
    check(p, body + b"@last\nACGTACGTACGTACGTACGTAAAA\n+\nIIIIIIIIIIIIIIIIIIIIIIII", False)
    p.write_bytes(b"")
    check(p, b"", False)
    variants = {
        "plain": _gz(body),
        "multi_member": _gz(body[:100000]) + _gz(body[100000:3000000], 1) + _gz(body[3000000:], 9),
        "zero_padded": _gz(body[:5000]) + b"\0" * 37 + _gz(body[5000:]) + b"\0" * 5,
        "empty_member": _gz(b"") + _gz(body[:99999]),
        "empty_file": b"",
        "unterminated": _gz(body + b"@last\nACGT"),
        "bgzf": _bgzf(body),
        "bgzf_then_gzip": _bgzf(body[:2_000_000]) + _gz(body[2_000_000:]),
    }
    for name, blob in variants.items():
        g = tmp_path / (name + ".fastq.gz")
        g.write_bytes(blob)
        want, ok = _py_gzip_lines(g)
        assert ok
        for threads in (1, 6):
            check(g, want, True, threads=threads)


def test_truncated_streams_end_like_the_reference_iterator(world, tmp_path):
    body, check = world
    for name, whole in (("gz", _gz(body[:100000]) + _gz(body[100000:2_500_000], 1) + _gz(body[2_500_000:], 9)), ("bgzf", _bgzf(body))):
        rnd = random.Random(11)
        cuts = [9, 10, 500, len(whole) // 3, len(whole) // 2 + 777, len(whole) - 9, len(whole) - 1] + [rnd.randrange(20, len(whole)) for _ in range(6)]
        for cut in cuts:
            g = tmp_path / "cut.fastq.gz"
            g.write_bytes(whole[:cut])
            want, ok = _py_gzip_lines(g)                    # the lines the reference's loop sees before EOFError
            check(g, want, complete=ok, gz=True, threads=5)


def test_line_limit_and_corruption(world, tmp_path):
    body, check = world
    p = tmp_path / "a.fastq"
    p.write_bytes(body)
    lines = body.split(b"\n")
    first = b"\n".join(lines[:40000]) + b"\n"
    check(p, first, False, limit=40000)
    g = tmp_path / "a.fastq.gz"
    g.write_bytes(_gz(body))
    check(g, first, True, limit=40000)
    g.write_bytes(_bgzf(body))
    check(g, first, True, limit=40000, threads=8)
    g.write_bytes(b"this is not gzip at all")
    e = lib.Engine(lib.make_config(mode="EC", upstream="ACGT", downstream="TTTT"), 0)
    e.begin()
    with pytest.raises(lib.F2QError) as err:
        e.submit_file(str(g), True)
    assert "corrupted gzip" in str(err.value)
    with pytest.raises(lib.F2QError):
        e.submit_file(str(tmp_path / "missing.fastq"), False)
    e.close()


def test_bgzf_inflated_on_the_device(world, tmp_path, oracle):
    """bgzip input is inflated by k_inflate_bgzf (one thread per block, only compressed bytes cross PCIe): same counts as the
    host path and the oracle; a foreign gzip member behind the blocks is picked up by the host reader at the right offset; a
    corrupted block is reported, never mis-counted"""
    body, check = world
    blob = _bgzf(body, level=1)
    g = tmp_path / "x.fastq.gz"
    g.write_bytes(blob)
    check(g, body, True, threads=6)                                   # the fixture engine: host threads (the default)
    spec = synth.default_spec(2)
    names, keys = synth.make_library(2, 400, 20)
    want_c, want_s = oracle.count(oracle.make_config(miss=1), keys, body)
    for opt in (1, 2, 0):
        with lib.Engine(lib.make_config(miss=1), 0, None, gpu_inflate=opt) as e:
            e.set_library(keys)
            e.begin(); ok, nb = e.submit_file(str(g), True, 0, 6); c, s = e.end()
            assert ok and nb == len(body) and s == want_s and np.array_equal(c, want_c), opt
    # blocks + an ordinary gzip member + the bgzip end-of-file block
    a, b = body[:3_000_000], body[3_000_000:]
    g.write_bytes(_bgzf(a)[:-28] + _gz(b) + _bgzf(b"")[-28:])
    want, ok = _py_gzip_lines(g)
    assert ok and want == body
    check(g, body, True, threads=6)
    with lib.Engine(lib.make_config(miss=1), 0, None, gpu_inflate=1) as e:
        e.set_library(keys)
        e.begin(); ok, nb = e.submit_file(str(g), True, 0, 6); c, s = e.end()
        assert ok and nb == len(body) and s == want_s and np.array_equal(c, want_c)
    # a flipped byte inside a block's payload
    bad = bytearray(blob)
    bad[len(bad) // 2] ^= 0x5A
    g.write_bytes(bytes(bad))
    with lib.Engine(lib.make_config(miss=1), 0, None, gpu_inflate=1) as e:
        e.set_library(keys)
        e.begin()
        with pytest.raises(lib.F2QError) as err:
            e.submit_file(str(g), True, 0, 6)
            e.end()
        assert "corrupted gzip" in str(err.value)
