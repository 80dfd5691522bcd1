"""
-m gpu: the speculative streaming kernel (csrc/spec.cuh) and its fallback.  Whatever the bytes are, the results must be
the reference's: ordinary FASTQ must be committed by the speculation (so that the fast path is the one the other tests
measure), hostile FASTQ must fall back to the exact look-back kernel and still be bit-exact.
"""
import numpy as np
import importlib
import pytest

import cases
import golden_io as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gu():
    import gpu_util
    assert gpu_util.lib.device_count() >= 1, "no B200 visible"
    return gpu_util


def counter_cases():
    return [c for c in G.kat() + G.fuzz() if c["params"]["mode"] == "C"]


@pytest.mark.parametrize("opts", [dict(spec=0), dict(spec=0, tile_threads=128), dict(spec_range_tiles=1), dict(spec_range_tiles=2, spec_warps=12),
                                  dict(spec_range_tiles=1, force_generic=1)], ids=str)
def test_golden_cases_every_path(gu, opts):
    """exact kernel only / one-tile ranges (every tile speculates) / 12-warp CTAs / generic per-read code"""
    for c in counter_cases():
        gu.check_case(c, c["fastq"], **opts)
    for c in counter_cases()[::3]:
        gu.check_case(c, c["fastq"], 1000, **opts)


@pytest.mark.parametrize("name", cases.SHAPED)
@pytest.mark.parametrize("opts", [dict(spec=0), dict(spec=0, tile_threads=128), dict(spec_range_tiles=1), dict(spec_range_tiles=3, row_chunks=5)], ids=str)
def test_shaped_cases_every_path(gu, name, opts):
    c = [x for x in G.shaped() if x["name"] == name][0]
    params, lib, data = cases.shaped_inputs(name)
    c = dict(c, library=lib)
    gu.check_case(c, data, **opts)
    gu.check_case(c, data, 100003, **opts)


def _run(gu, params, keys, data, chunk=None, **options):
    cfg = gu.lib.make_config(**params)
    with gu.lib.Engine(cfg, 0, None, **options) as e:
        e.set_library(keys)
        counts, stats = e.run(data, chunk)
        return counts, stats, e.spec_counts()


def test_ordinary_fastq_is_committed_by_the_speculation(gu, oracle):
    """config-2 shape: many ranges, every one must guess its phase right; nothing may reach the exact kernel"""
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.default_spec(2)
    names, keys = synth.make_library(2, 2000, 20)
    data = synth.fixed_reads(keys, 0, 300_000, **spec).tobytes()
    want_c, want_s = oracle.count(oracle.make_config(miss=1), keys, data)
    for opts in (dict(), dict(spec_range_tiles=1), dict(spec_range_tiles=5, spec_warps=12)):
        c, s, (ok, fb) = _run(gu, dict(miss=1), keys, data, **opts)
        assert s == want_s and np.array_equal(c, want_c), opts
        assert ok == 1 and fb == 0, (opts, ok, fb)
    c, s, (ok, fb) = _run(gu, dict(miss=1), keys, data, 1_000_003, spec_range_tiles=2)
    assert s == want_s and np.array_equal(c, want_c)
    assert fb == 0 and ok >= len(data) // 1_000_003
    c, s, (ok, fb) = _run(gu, dict(miss=1), keys, data, spec=0)
    assert s == want_s and np.array_equal(c, want_c) and (ok, fb) == (0, 0)


def _hostile(kind, n=40_000):
    """FASTQ-shaped bytes that defeat the local phase heuristic"""
    synth = importlib.import_module("2fast2q_b200.synth")
    r = synth.SM64(1234)
    guides = [r.dna(20) for _ in range(64)]
    out = []
    for i in range(n):
        g = r.choice(guides)
        if r.below(5) == 0:
            g = synth.mutate(r, g, 1)
        s = g + r.dna(30)
        q = synth.qual_line(r, len(s), 0.05)
        if kind == "at_plus_everywhere":           # every line starts with '@' or '+': four alignments look alike
            out.append(b"@" + s[1:] + b"\n" + b"+" + s[1:] + b"\n+" + q[1:] + b"\n@" + q[1:] + b"\n")
        elif kind == "blank_lines":                # blank lines shift the phase: records are still 'every 4 lines'
            out.append(b"@r%d\n" % i + s + b"\n+\n" + q + b"\n" + (b"\n" if i % 7 == 0 else b""))
        elif kind == "quality_is_header":          # the '+' line is missing: the true phase walks through all alignments
            out.append(b"@r%d\n" % i + s + b"\n" + q + b"\n")
        elif kind == "unequal_lengths":            # sequence and quality lengths differ: the heuristic never accepts
            out.append(b"@r%d\n" % i + s + b"\n+\n" + q[:-3] + b"\n")
        else:
            raise ValueError(kind)
    return guides, b"".join(out)


@pytest.mark.parametrize("kind", ["at_plus_everywhere", "blank_lines", "quality_is_header", "unequal_lengths"])
def test_hostile_fastq_falls_back_and_stays_exact(gu, oracle, kind):
    guides, data = _hostile(kind)
    keys = list(dict.fromkeys(guides))
    want_c, want_s = oracle.count(oracle.make_config(miss=1), keys, data)
    for opts in (dict(), dict(spec_range_tiles=1), dict(spec_range_tiles=1, force_generic=1)):
        c, s, (ok, fb) = _run(gu, dict(miss=1), keys, data, **opts)
        assert s == want_s and np.array_equal(c, want_c), (kind, opts)
        assert ok + fb == 1
        if opts:
            assert fb == 1, (kind, opts)              # one-tile ranges: some range must have mis-speculated or declined
    # chunked: after the first failure the sample stays on the exact kernel (sticky), results unchanged
    c, s, (ok, fb) = _run(gu, dict(miss=1), keys, data, 300_007, spec_range_tiles=1)
    assert s == want_s and np.array_equal(c, want_c), kind
    assert fb >= 1


def test_wrong_guess_is_caught_by_the_verification(gu, oracle):
    """a clean-looking stretch whose true phase differs from its looks: a 5-line oddity in front shifts every record by
    one line, but each range after it still looks like perfectly aligned FASTQ.  The verification must reject it."""
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.default_spec(2)
    names, keys = synth.make_library(2, 500, 20)
    body = synth.fixed_reads(keys, 0, 60_000, **spec).tobytes()
    data = b"@odd\n" + body
    want_c, want_s = oracle.count(oracle.make_config(miss=1), keys, data)
    c, s, (ok, fb) = _run(gu, dict(miss=1), keys, data, spec_range_tiles=4)
    assert s == want_s and np.array_equal(c, want_c)
    assert (ok, fb) == (0, 1)


def test_extract_count_never_speculates(gu):
    c = [x for x in G.kat() + G.fuzz() if x["params"]["mode"] == "EC"][0]
    cfg = gu.lib.make_config(**c["params"])
    with gu.lib.Engine(cfg, 0, None, spec_range_tiles=1) as e:
        e.run(c["fastq"])
        assert e.spec_counts() == (0, 0)


@pytest.mark.parametrize("length,start", [(20, 0), (13, 5), (24, 2), (17, 30)])
def test_ordinary_read_code_with_whitespace_tails_and_short_lines(gu, oracle, length, start):
    """the compile-time-window code of the streaming kernel must hand the reads it cannot decide to the general code:
    CRLF / blank tails (rstrip), N inside the window, lines that end inside the window, empty lines"""
    synth = importlib.import_module("2fast2q_b200.synth")
    r = synth.SM64(99 + length)
    guides = [r.dna(length) for _ in range(300)]
    out = []
    for i in range(20_000):
        g = r.choice(guides)
        k = r.below(10)
        if k == 0:
            g = synth.mutate(r, g, 1)
        s = r.dna(start) + g + r.dna(r.below(12))
        if k == 1:
            s = s[:start + r.below(length + 1)]                  # the line ends inside (or right at the start of) the window
        if k == 2:
            p = start + r.below(length)
            s = s[:p] + b"N" + s[p + 1:]
        q = synth.qual_line(r, len(s), 0.05)
        tail = [b"", b"\r", b"  ", b"\t \r", b""][r.below(5)]
        if k == 3:
            q = q[:start + r.below(length + 1)]
        if k == 4 and len(s) > start + 2:
            s = s[:start + 2] + b" " * (len(s) - start - 2)        # whitespace reaching back into the window
        out.append(b"@r%d%s\n" % (i, tail) + s + tail + b"\n+" + tail + b"\n" + q + tail + b"\n")
    data = b"".join(out)
    keys = list(dict.fromkeys(guides))
    params = dict(miss=1, length=length, start=str(start))
    want_c, want_s = oracle.count(oracle.make_config(**params), keys, data)
    for opts in (dict(), dict(spec_range_tiles=2), dict(spec=0)):
        c, s, _ = _run(gu, params, keys, data, **opts)
        assert s == want_s and np.array_equal(c, want_c), (length, start, opts)
