"""
-m gpu: the CUDA path (through the C-ABI) against the golden vectors of the unmodified reference and against the
oracle.  Bit-exact: integer counts and the five statistics.
"""
import importlib
import os

import numpy as np
import pytest

import cases
import golden_io as G

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gu():
    import gpu_util
    assert gpu_util.lib.device_count() >= 1, "no B200 visible"
    return gpu_util


def test_kat_cases(gu):
    for c in G.kat():
        gu.check_case(c, c["fastq"])


def test_kat_cases_generic_kernels(gu):
    for c in G.kat():
        gu.check_case(c, c["fastq"], force_generic=1)


@pytest.mark.parametrize("chunk", [1, 3, 16, 127])
def test_kat_cases_chunked(gu, chunk):
    """records cut at arbitrary byte positions: the carry/stitch logic must give the same answer"""
    for c in G.kat():
        gu.check_case(c, c["fastq"], chunk)


def test_fuzz_cases(gu):
    for c in G.fuzz():
        gu.check_case(c, c["fastq"])


@pytest.mark.parametrize("chunk", [5, 64, 1000])
def test_fuzz_cases_chunked(gu, chunk):
    for c in G.fuzz()[::3]:
        gu.check_case(c, c["fastq"], chunk)


@pytest.mark.parametrize("ch", [3, 5, 7])
def test_row_sizes(gu, ch):
    """every row size of the tile kernel (16*ch bytes per thread) on inputs whose records are shorter and longer than it"""
    for c in G.kat():
        gu.check_case(c, c["fastq"], row_chunks=ch)
    for c in G.fuzz()[::4]:
        gu.check_case(c, c["fastq"], row_chunks=ch)
    for name in ("config2_slice", "config3_m3", "config4_barseq"):
        c = [x for x in G.shaped() if x["name"] == name][0]
        params, lib, data = cases.shaped_inputs(name)
        gu.check_case(dict(c, library=lib), data, row_chunks=ch)
        gu.check_case(dict(c, library=lib), data, 250007, row_chunks=ch)


def test_fuzz_cases_generic_kernels(gu):
    for c in G.fuzz()[::2]:
        gu.check_case(c, c["fastq"], force_generic=1)


@pytest.mark.parametrize("name", cases.SHAPED)
def test_shaped_cases(gu, name):
    c = [x for x in G.shaped() if x["name"] == name][0]
    params, lib, data = cases.shaped_inputs(name)
    assert cases.sha(data) == c["sha256"]
    c = dict(c, library=lib)
    gu.check_case(c, data)
    gu.check_case(c, data, 100003)        # odd chunk size: every submit cuts a record


@pytest.mark.parametrize("name,resolver", [("config2_slice", 1), ("config2_slice", 2), ("config2_slice", 3), ("config3_slice", 2),
                                           ("config3_slice", 3), ("config3_m3", 2), ("config3_m3", 3)])
def test_resolvers_agree(gu, name, resolver):
    c = [x for x in G.shaped() if x["name"] == name][0]
    params, lib, data = cases.shaped_inputs(name)
    gu.check_case(dict(c, library=lib), data, resolver=resolver)


@pytest.mark.parametrize("resolver", [1, 2, 3])
def test_resolvers_on_fuzz(gu, resolver):
    """ties, N symbols, mixed key lengths, m up to 3: every resolver must reproduce the reference's answers"""
    for c in G.fuzz()[::2]:
        if c["params"]["mode"] == "C":
            gu.check_case(c, c["fastq"], resolver=resolver)
    for c in G.kat():
        if c["params"]["mode"] == "C":
            gu.check_case(c, c["fastq"], resolver=resolver)


@pytest.mark.parametrize("parts", [2, 3, 4, 6])
def test_seed_plans_agree(gu, parts):
    """the resolver's seed plan (P segments, seeds = every choice of P - m of them): every P the option accepts gives the
    reference's answers — ties, N symbols, mixed key lengths, m = 1..3; a P that does not fit m falls back to m + 1 one-segment
    seeds (the classic plan), which is covered here too"""
    for c in G.fuzz()[::2] + G.kat():
        if c["params"]["mode"] == "C":
            gu.check_case(c, c["fastq"], seed_parts=parts)
    for name in ("config3_slice", "config3_m3"):
        c = [x for x in G.shaped() if x["name"] == name][0]
        params, lib, data = cases.shaped_inputs(name)
        gu.check_case(dict(c, library=lib), data, seed_parts=parts)
        gu.check_case(dict(c, library=lib), data, seed_parts=parts, queue_entries=16)       # resolved inside the streaming kernel


@pytest.mark.parametrize("miss,n_keys", [(9, 300), (5, 3000), (2, 100000), (3, 100000)])
def test_seed_plans_on_large_libraries_and_many_mismatches(gu, oracle, miss, n_keys):
    """what the automatic plan picks: two-segment seeds for 100 000 guides at m = 2, three-segment seeds at m = 3, the
    classic one-segment seeds for m = 9 (more segments than a seed mask holds); against the oracle, keys resolved by the
    resolver kernel and (16-entry queue) inside the streaming kernel"""
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.default_spec(3)
    names, keys = synth.make_library(3, n_keys, 20)
    data = synth.fixed_reads(keys, 0, 20000 if n_keys < 50000 else 6000, **spec)
    want_c, want_s = oracle.count(oracle.make_config(miss=miss), keys, data)
    cfg = gu.lib.make_config(miss=miss)
    for opts in ({}, dict(queue_entries=16), dict(resolve_group=8)):
        with gu.lib.Engine(cfg, 0, None, **opts) as e:
            e.set_library(keys)
            e.begin(); e.submit(data, True); c, s = e.end()
        assert np.array_equal(c, want_c) and s == want_s, (miss, n_keys, opts)


def test_queue_overflow_resolves_in_place(gu):
    """a 16-entry queue overflows immediately; results must not change"""
    c = [x for x in G.shaped() if x["name"] == "config2_slice"][0]
    params, lib, data = cases.shaped_inputs("config2_slice")
    gu.check_case(dict(c, library=lib), data, queue_entries=16)


def test_config1_surrogate_equals_reference_compiled_csv(gu):
    g = G.config1()
    lib, want, data = cases.config1_surrogate(os.path.join(G.HERE, "D39V_guides.csv"), os.path.join(G.HERE, "ref_compiled.csv"))
    assert cases.sha(data) == g["sha256"]
    got, stats = gu.run_case(cases.P(), lib, data)
    assert {n: c for (n, _), c in zip(lib, got)} == want
    assert stats == g["stats"]


def test_primitives(gu):
    p = G.primitives()
    for v in p["border_finder"][::4]:
        assert gu.lib.border_finder_device(v["seq"].encode(), v["read"].encode(), v["mismatch"], v["start_place"]) == v["expect"], v
    for v in p["sequence_tinder"][::4] + p["sequence_tinder"][:5]:
        cfg = gu.lib.make_config(upstream=v["upstream"], downstream=v["downstream"], miss_search_up=v["msu"],
                                 miss_search_down=v["msd"], qual_up=v.get("qsu", 30), qual_down=v.get("qsd", 30), length=v["length"])
        got = gu.lib.sequence_tinder_device(cfg, v["read"].encode(), v["qual"].encode(), v["i"], v.get("set_up"), v.get("set_down"))
        assert list(got) == v["expect"], v


def test_long_records_spill_path(gu, oracle):
    """records longer than the tile halo (1 KiB) and than a whole tile take the global-memory path"""
    synth = importlib.import_module("2fast2q_b200.synth")
    r = synth.SM64(77)
    guides = [r.dna(20) for _ in range(50)]
    recs = []
    for i in range(300):
        L = r.choice([30, 600, 1500, 5000, 40000]) if i % 3 == 0 else 60
        g = r.choice(guides)
        if r.below(4) == 0:
            g = synth.mutate(r, g, 1)
        s = g + r.dna(L)
        recs.append(b"@x%d\n" % i + s + b"\n+\n" + synth.qual_line(r, len(s), 0.1) + b"\n")
    data = b"".join(recs)
    params = cases.P()
    lib = [("g%d" % i, g.decode()) for i, g in enumerate(dict.fromkeys(guides))]
    want_c, want_s = oracle.count(oracle.make_config(**params), [s for _, s in lib], data)
    for chunk in (None, 70001):
        got, stats = gu.run_case(params, lib, data, chunk)
        assert stats == want_s and got == [int(x) for x in want_c]


def test_short_lines_many_newlines(gu, oracle):
    """more newlines per tile than one rank window holds (lines of 0-3 bytes)"""
    synth = importlib.import_module("2fast2q_b200.synth")
    r = synth.SM64(5)
    data = b"".join(r.dna(r.below(4)) + b"\n" for _ in range(60000))
    params = cases.P(mode="EC", length=3)
    want, want_s = oracle.extract_count(oracle.make_config(**params), data)
    got, stats = gu.run_case(params, None, data)
    assert stats == want_s and got == want
    params = cases.P(length=2, miss=1)
    lib = [("a", "AC"), ("b", "GT"), ("c", "A"), ("d", "")]
    want_c, want_s = oracle.count(oracle.make_config(**params), [s for _, s in lib], data)
    got, stats = gu.run_case(params, lib, data)
    assert stats == want_s and got == [int(x) for x in want_c]


def test_synth_generator_matches_numpy(gu):
    """K0 on the device == its numpy restatement, every shape (configs 2, 3, 4, 5a, 5b); the second call re-uses the
    uploaded guide table asynchronously, as the per-chunk generation of the bench does"""
    synth = importlib.import_module("2fast2q_b200.synth")
    work = []
    for config in (2, 3):
        spec = synth.default_spec(config)
        names, keys = synth.make_library(config, 500, 20)
        work.append((keys, spec, lambda f, n, keys=keys, spec=spec: synth.fixed_reads(keys, f, n, **spec)))
    _, pool = synth.make_library(44, 3000, 20)
    _, _, xs, ys = synth.dual_keys(300)
    for config, guides in (("4", pool), ("5a", xs + ys), ("5b", xs + ys)):
        spec = synth.shape_spec(config)
        work.append((guides, spec, lambda f, n, guides=guides, spec=spec: synth.shaped_reads(guides, f, n, **spec)))
    cfg = gu.lib.make_config()
    for guides, spec, gen in work:
        want = gen(12345, 3000)
        want2 = gen(999_000_000_000, 1000)
        with gu.lib.Engine(cfg) as e:
            d = e.device_alloc(want.size)
            e.synth(d, guides, 12345, 3000, **spec)
            got = e.d2h(d, want.size)
            e.synth(d, guides, 999_000_000_000, 1000, reuse_guides=True, **spec)
            got2 = e.d2h(d, want2.size)
            e.device_free(d)
        assert np.array_equal(got, want), spec
        assert np.array_equal(got2, want2), spec


def test_resident_device_submit_large(gu, oracle):
    """config-2 shape, 2M reads generated on the device: one submit vs many odd-sized submits vs the oracle"""
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.default_spec(2)
    names, keys = synth.make_library(2, 2000, 20)
    n = 2_000_000
    rec = 2 * spec["read_len"] + 18
    cfg = gu.lib.make_config(miss=1)
    with gu.lib.Engine(cfg) as e:
        e.set_library(keys)
        d = e.device_alloc(n * rec)
        e.synth(d, keys, 0, n, **spec)
        e.begin(); e.submit_device(d, n * rec, True); c1, s1 = e.end()
        e.begin()
        step, o = 7_777_777, 0
        while o < n * rec:
            m = min(step, n * rec - o)
            e.submit_device(d + o, m, o + m >= n * rec)
            o += m
        c2, s2 = e.end()
        host = e.d2h(d, n * rec)
        e.device_free(d)
    assert s1 == s2 and np.array_equal(c1, c2)
    assert s1["reads"] == n and s1["reads"] == s1["perfect_counter"] + s1["imperfect_counter"] + s1["non_aligned_counter"] + s1["quality_failed"]
    assert int(c1.sum()) == s1["perfect_counter"] + s1["imperfect_counter"]
    want_c, want_s = oracle.count(oracle.make_config(miss=1), keys, host)
    assert want_s == s1 and np.array_equal(want_c, c1)


def test_end_sample_async_back_to_back(gu, oracle):
    """three samples without a host round trip in between: every pinned result equals the blocking path's"""
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.default_spec(2)
    names, keys = synth.make_library(2, 700, 20)
    datas = [synth.fixed_reads(keys, 1000 * k, 40_000 + 777 * k, **spec).tobytes() for k in range(3)]
    cfg = gu.lib.make_config(miss=1)
    with gu.lib.Engine(cfg) as e:
        e.set_library(keys)
        want = [e.run(d) for d in datas]
        bufs = [gu.lib.PinnedBuffer(8 * (len(keys) + 6)) for _ in datas]
        for d, b in zip(datas, bufs):
            e.begin()
            e.submit(d[:len(d) // 2], False)
            e.submit(d[len(d) // 2:], False)          # not closed: end_sample_async flushes the carried record itself
            e.end_async(b)
        e.sync()
        for (wc, ws), b, d in zip(want, bufs, datas):
            c, s = e.read_async_result(b)
            assert s == ws and np.array_equal(c, wc)
            oc, os_ = oracle.count(oracle.make_config(miss=1), keys, d)
            assert s == os_ and np.array_equal(c, oc)
            b.free()


@pytest.mark.parametrize("name", ["config2_slice", "config3_slice", "config3_m3", "config5a_dual_fixed", "config5b_dual_delim"])
@pytest.mark.parametrize("opts", [dict(memo_entries=0), dict(memo_entries=1 << 12), dict(memo_entries=1 << 20, resolve_group=8),
                                  dict(resolve_group=32), dict(resolve_group=1, memo_entries=1 << 16)], ids=str)
def test_resolvers_agree_with_memo_and_groups(gu, name, opts):
    """the seed-index resolvers with 1 / 8 / 32 lanes per key, the device memo of resolved keys off, tiny (constant eviction)
    and large: always the reference's counts (fast2q.py:692-750 — the memo never changes an outcome)"""
    c = [x for x in G.shaped() if x["name"] == name][0]
    params, lib, data = cases.shaped_inputs(name)
    gu.check_case(dict(c, library=lib), data, **opts)
    gu.check_case(dict(c, library=lib), data, 300007, **opts)


def test_memo_is_reused_across_chunks_and_samples(gu, oracle):
    """the same non-exact keys again and again (what real screens do): the first sample fills the memo, the
    second one — another f2q_begin_sample on the same context — resolves nothing itself, and both equal the oracle"""
    synth = importlib.import_module("2fast2q_b200.synth")
    spec = synth.default_spec(3)
    names, keys = synth.make_library(3, 20_000, 20)
    block = synth.fixed_reads(keys, 0, 20_000, **spec)
    data = np.tile(block, 6)
    want_c, want_s = oracle.count(oracle.make_config(miss=2), keys, data)
    cfg = gu.lib.make_config(miss=2)
    with gu.lib.Engine(cfg, 0, None, memo_entries=1 << 20) as e:
        e.set_library(keys)
        c1, s1 = e.run(data, block.size)                 # one chunk per repetition
        look1, hit1 = e.memo_counts()
        c2, s2 = e.run(data)
        look2, hit2 = e.memo_counts()
    assert s1 == want_s and s2 == want_s and np.array_equal(c1, want_c) and np.array_equal(c2, want_c)
    assert look1 > 10_000 and hit1 >= look1 * 4 // 6      # repetitions 2..6 of the first sample hit
    assert look2 >= look1 and hit2 >= look2 * 85 // 100   # the second sample finds most keys: entries survive f2q_begin_sample
    # (look1 < look2 and not every key is found: the first sample's small chunks overflow their queue segments, those keys
    # are resolved in place without the memo; the memo is direct-mapped, so a few entries evict each other)
