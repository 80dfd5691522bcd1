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
    return G.kat() + G.fuzz()        # (Extract+Count cases too: their inserts go through the insert log of the flex policy)


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


def test_extract_count_speculates_through_the_insert_log(gu):
    """Extract+Count runs on the streaming kernel too (flex policy): its inserts wait in a log until the chunk verified"""
    c = [x for x in G.kat() + G.fuzz() if x["params"]["mode"] == "EC"][0]
    cfg = gu.lib.make_config(**c["params"])
    with gu.lib.Engine(cfg, 0, None, spec_range_tiles=1) as e:
        e.run(c["fastq"])
        assert sum(e.spec_counts()) >= 1
    with gu.lib.Engine(cfg, 0, None, flex=0) as e:                      # the byte-wise generic code cannot take its inserts back
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


# ------------------------------------------------------------------------------------------------------------------
# the bit-parallel (flex) policies of the streaming kernel: search sequences with mismatches, several windows per read,
# Extract+Count through the insert log.  Every golden / fuzz case that is not "one fixed window, Counter" runs on them
# by default; here every side path: the byte-wise generic kernels instead, overflowing queues, every resolver group size
# ------------------------------------------------------------------------------------------------------------------
def flex_cases():
    out = []
    for c in G.kat() + G.fuzz():
        p = c["params"]
        single_fixed = p["upstream"] is None and p["downstream"] is None and "," not in str(p["start"])
        if p["mode"] == "EC" or not single_fixed:
            out.append(c)
    return out


@pytest.mark.parametrize("opts", [dict(flex=0), dict(spec=0), dict(spec_range_tiles=1), dict(queue_entries=16), dict(queue_entries=16, generic_entries=3),
                                  dict(resolve_group=1), dict(resolve_group=8), dict(resolve_group=32), dict(row_chunks=5, spec_range_tiles=2)], ids=str)
def test_flex_cases_every_path(gu, opts):
    cs = flex_cases()
    assert len(cs) > 150
    for c in cs:
        gu.check_case(c, c["fastq"], **opts)
    for c in cs[::4]:
        gu.check_case(c, c["fastq"], 777, **opts)


def _shaped_workload(config, n_reads, first=0):
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.shape_spec(config)
    if config == "4":
        guides = synth.random_kmers(4, 20_000, 20)
        params = dict(mode="EC", upstream=synth.BARSEQ_US.decode(), downstream=synth.BARSEQ_DS.decode(), miss_search_up=1, miss_search_down=1)
        return params, None, synth.shaped_reads(guides, first, n_reads, **spec)
    names, keys, xs, ys = synth.dual_keys(3000)
    params = dict(mode="C", miss=1, start="0,30") if config == "5a" else dict(mode="C", miss=1, upstream="ACCGGT,GGATCC", downstream="TTGACA,CAATTG")
    return params, keys, synth.shaped_reads(xs + ys, first, n_reads, **spec)


@pytest.mark.parametrize("config", ["4", "5a", "5b"])
def test_flex_workloads_are_committed_and_exact(gu, oracle, config):
    """the bench's config-4 / 5a / 5b streams (K0 shapes 1-3), 300k reads: the streaming kernel must commit every chunk (this
    is the path the bench measures) and equal the oracle; then the same through odd host chunks, tiny queues and flex off"""
    params, keys, data = _shaped_workload(config, 300_000)
    ocfg = oracle.make_config(**params)
    if keys is None:
        want, want_s = oracle.extract_count(ocfg, data)
    else:
        want, want_s = oracle.count(ocfg, keys, data)
    cfg = gu.lib.make_config(**params)
    for chunk, opts in ((None, {}), (5_000_011, {}), (None, dict(queue_entries=1000)), (None, dict(flex=0)), (1_000_003, dict(spec=0))):
        with gu.lib.Engine(cfg, 0, None, **opts) as e:
            if keys is not None:
                e.set_library(keys)
            counts, stats = e.run(data, chunk)
            got = e.ec_items() if keys is None else counts
            committed, fell_back = e.spec_counts()
        assert stats == want_s, (config, chunk, opts)
        if keys is None:
            assert got == want, (config, chunk, opts)
        else:
            assert np.array_equal(got, want), (config, chunk, opts)
        if opts.get("spec", 1) and opts.get("flex", 1) and "queue_entries" not in opts:      # (a tiny queue may legitimately end in a fallback)
            assert fell_back == 0 and committed >= 1, (config, chunk, opts, committed, fell_back)


def test_extract_count_tables_grow_on_the_device(gu, oracle):
    """every read a new key: the packed table starts small and is rehashed on the device several times; a second sample on
    the same context starts from empty tables again"""
    synth = importlib.import_module("2fast2q_b200.synth")
    r = synth.SM64(4242)
    recs = []
    for i in range(120_000):
        bc = r.dna(24) if i % 7 else r.dna(23) + b"N"                   # every 7th key goes to the byte-arena table
        s = b"GGATCCAA" + bc + b"TTGACACC" + r.dna(6)
        recs.append(b"@g%d\n" % i + s + b"\n+\n" + b"I" * len(s) + b"\n")
    data = b"".join(recs)
    params = dict(mode="EC", upstream="GGATCCAA", downstream="TTGACACC")
    want, want_s = oracle.extract_count(oracle.make_config(**params), data)
    cfg = gu.lib.make_config(**params)
    with gu.lib.Engine(cfg, 0, None, stage_bytes=1 << 20) as e:
        for rep in range(2):
            counts, stats = e.run(data, 3_000_000)
            assert stats == want_s and e.ec_items() == want, rep


def test_extract_count_large_resident_chunk_is_sub_chunked(gu, oracle):
    """a multi-GB f2q_submit_device in Extract+Count mode must not reserve worst-case tables; 1.3 GB here"""
    params, _, piece = _shaped_workload("4", 1_000_000)
    want, want_s = oracle.extract_count(oracle.make_config(**params), piece)
    reps = 8
    cfg = gu.lib.make_config(**params)
    with gu.lib.Engine(cfg) as e:
        d = e.device_alloc(piece.size * reps)
        for k in range(reps):
            e.h2d(d + k * piece.size, piece)
        e.begin(); e.submit_device(d, piece.size * reps, True); counts, stats = e.end()
        got = e.ec_items()
        e.device_free(d)
    assert stats == {k: v * reps for k, v in want_s.items()}
    assert got == {k: v * reps for k, v in want.items()}
