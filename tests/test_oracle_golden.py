"""
The oracle (oracle/f2q_oracle.c) against every golden vector produced by the unmodified reference
(tests/golden/make_golden.py) and the reference's own unit tests (tests/test_mainfunctions.py:4-78).
CPU only.
"""
import os

import numpy as np
import pytest

import cases
import golden_io as G


def _run(O, c, fastq):
    cfg = O.make_config(**c["params"])
    if c.get("library") is not None or "counts" in c:
        lib = c["library"]
        counts, stats = O.count(cfg, [s for _, s in lib], fastq)
        return [int(x) for x in counts], stats
    ec, stats = O.extract_count(cfg, fastq)
    return [[k.decode("latin-1"), v] for k, v in ec.items()], stats


def _check(O, c, fastq):
    got, stats = _run(O, c, fastq)
    assert stats == c["stats"], c["name"]
    if "counts" in c:
        assert got == c["counts"], c["name"]
    else:
        assert got == c["ec"], c["name"]          # same keys, same counts, same (insertion) order


def test_kat_cases(oracle):
    cs = G.kat()
    assert len(cs) >= 70
    for c in cs:
        _check(oracle, c, c["fastq"])


def test_fuzz_cases(oracle):
    cs = G.fuzz()
    assert len(cs) == cases.N_FUZZ
    for c in cs:
        _check(oracle, c, c["fastq"])


@pytest.mark.parametrize("name", cases.SHAPED)
def test_shaped_cases(oracle, name):
    c = [x for x in G.shaped() if x["name"] == name][0]
    params, lib, data = cases.shaped_inputs(name)
    assert cases.sha(data) == c["sha256"], "synthetic generator drifted: regenerate the golden fixtures"
    c = dict(c, library=lib)
    _check(oracle, c, data)


def test_config1_surrogate_equals_reference_compiled_csv(oracle):
    """config 1: the oracle on the example.fastq.gz surrogate reproduces the reference's tests/compiled.csv"""
    g = G.config1()
    lib, want, data = cases.config1_surrogate(os.path.join(G.HERE, "D39V_guides.csv"), os.path.join(G.HERE, "ref_compiled.csv"))
    assert cases.sha(data) == g["sha256"]
    assert len(lib) == 1498 and sum(want.values()) == 60916
    counts, stats = oracle.count(oracle.make_config(**cases.P()), [s for _, s in lib], data)
    assert {n: int(c) for (n, _), c in zip(lib, counts)} == want
    assert stats == g["stats"]


def test_primitives(oracle):
    p = G.primitives()
    for v in p["border_finder"]:
        assert oracle.border_finder(v["seq"].encode(), v["read"].encode(), v["mismatch"], v["start_place"]) == v["expect"], v
    for v in p["sequence_tinder"]:
        cfg = oracle.make_config(upstream=v["upstream"], downstream=v["downstream"], miss_search_up=v["msu"],
                                 miss_search_down=v["msd"], qual_up=v.get("qsu", 30), qual_down=v.get("qsd", 30),
                                 length=v["length"])
        got = oracle.sequence_tinder(cfg, v["read"].encode(), v["qual"].encode(), v["i"], v.get("set_up"), v.get("set_down"))
        assert list(got) == v["expect"], v
    assert p["seq2bin"]["expect"] == [71, 65, 84, 84, 65, 67, 65]


def test_reference_unit_vectors(oracle):
    """tests/test_mainfunctions.py:10-78 restated on the oracle API"""
    assert oracle.border_finder(b"GATTACA", b"TACTGATTACAGCAC", 1) == 4
    cfg = oracle.make_config(upstream="TACT", downstream="GCAC", miss_search_up=1, miss_search_down=1)
    r, q = b"TACTGATTACAGCAC", b"AAII$%&#III/(&/"
    assert oracle.sequence_tinder(cfg, r, q, 0, "", "") == (4, 11)
    assert oracle.sequence_tinder(cfg, r, q, 0, "", "/") == (None, None)
    cfg.miss_down = 2
    assert oracle.sequence_tinder(cfg, r, q, 0, "", "/") == (4, 6)
    read = b"AAAAAACACACACACACACACATTCAGGGGGGCCAAAAATAGAGAGAGAGAGACCGAGAGGGGGTTAGCATCG"
    cfg = oracle.make_config(upstream="CACACATT,GAGACCGA", downstream="TAGAGAGA,TAGCATCG")
    out = []
    for i in range(2):
        s, e = oracle.sequence_tinder(cfg, read, b"B" * 90, i, "", "")
        out.append(read[s:e])
    assert out == [b"CAGGGGGGCCAAAAA", b"GAGGGGGT"]
