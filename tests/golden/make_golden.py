#!/usr/bin/env python
"""
Generates the committed golden fixtures by running the UNMODIFIED reference (/root/reference, via
refharness.py) on the inputs defined in cases.py.  Run in the build container only:

    python tests/golden/make_golden.py

Outputs (all under tests/golden/):
    kat_cases.json          handcrafted cases: inputs + reference answers
    fuzz_cases.json.gz      seeded fuzz cases: inputs + reference answers
    shaped_cases.json       config-shaped workloads: sha256 of the regenerated input + reference answers
    primitive_cases.json    border_finder / sequence_tinder vectors incl. the reference's own unit tests
    config1_surrogate.json  sha256 + statistics of the example.fastq.gz surrogate; its per-guide counts
                            are REQUIRED to equal the reference's tests/compiled.csv (copied as ref_compiled.csv)
    D39V_guides.csv, ref_compiled.csv   data fixtures copied from the reference (fast2q/data, tests/)
"""
import gzip
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cases  # noqa: E402
import refharness  # noqa: E402
from oracle import synth  # noqa: E402


def run_case(c):
    counts, stats = refharness.run_reads_counter(c["fastq"], c["library"], **c["params"])
    out = dict(name=c["name"], params=c["params"], library=c["library"], stats=stats)
    if c["library"] is not None:
        seqs = [s for _, s in c["library"]]
        assert list(counts.keys()) == seqs, "reference dict order differs from library order"
        out["counts"] = [counts[s] for s in seqs]
    else:
        out["ec"] = [[k, v] for k, v in counts.items()]
    return out


def primitives():
    ref = refharness.load()
    out = dict(border_finder=[], sequence_tinder=[])
    # the reference's own unit vectors (tests/test_mainfunctions.py:10-78)
    out["border_finder"].append(dict(seq="GATTACA", read="TACTGATTACAGCAC", mismatch=1, start_place=0,
                                     expect=ref.border_finder(ref.seq2bin("GATTACA"), ref.seq2bin("TACTGATTACAGCAC"), 1)))
    r = synth.SM64(0xB0BDE8)
    for _ in range(600):
        read = r.dna(r.below(40), b"ACGTacgtN")
        if r.below(100) < 70 and len(read) > 4:
            a = r.below(len(read) - 2)
            seq = synth.mutate(r, read[a:a + 2 + r.below(8)].upper(), r.below(3))
        else:
            seq = r.dna(1 + r.below(9))
        k = r.below(4)
        sp = r.below(len(read) + 3)
        if len(read) == 0:
            continue  # numba cannot type an empty int8 array built from an empty str in the same way; skip
        e = ref.border_finder(ref.seq2bin(seq.decode()), ref.seq2bin(read.decode()), k, sp)
        out["border_finder"].append(dict(seq=seq.decode(), read=read.decode(), mismatch=k, start_place=sp,
                                         expect=None if e is None else int(e)))
    # sequence_tinder: reference unit test (explicit quality sets) ...
    read, qual = "TACTGATTACAGCAC", "AAII$%&#III/(&/"
    for msd, sdown in ((1, ""), (1, "/"), (2, "/")):
        info = dict(upstream="TACT", downstream="GCAC", upstream_bin=[ref.seq2bin("TACT")], downstream_bin=[ref.seq2bin("GCAC")],
                    miss_search_up=1, miss_search_down=msd, quality_set_up=set(""), quality_set_down=set(sdown))
        s, e = ref.sequence_tinder(ref.seq2bin(read), qual.encode(), info)
        out["sequence_tinder"].append(dict(read=read, qual=qual, upstream="TACT", downstream="GCAC", msu=1, msd=msd,
                                           set_up="", set_down=sdown, length=20, i=0,
                                           expect=[None if s is None else int(s), None if e is None else int(e)]))
    read = "AAAAAACACACACACACACACATTCAGGGGGGCCAAAAATAGAGAGAGAGAGACCGAGAGGGGGTTAGCATCG"
    qual = "B" * 90
    for i in range(2):
        info = dict(upstream="CACACATT", downstream="TAGAGAGA",
                    upstream_bin=[ref.seq2bin("CACACATT"), ref.seq2bin("GAGACCGA")],
                    downstream_bin=[ref.seq2bin("TAGAGAGA"), ref.seq2bin("TAGCATCG")],
                    miss_search_up=0, miss_search_down=0, quality_set_up=set(""), quality_set_down=set(""))
        s, e = ref.sequence_tinder(ref.seq2bin(read), qual.encode(), info, i)
        out["sequence_tinder"].append(dict(read=read, qual=qual, upstream="CACACATT,GAGACCGA", downstream="TAGAGAGA,TAGCATCG",
                                           msu=0, msd=0, set_up="", set_down="", length=20, i=i, expect=[int(s), int(e)]))
    # ... and random vectors in the three modes with phred-derived sets
    for _ in range(500):
        us, ds = r.dna(2 + r.below(6)), r.dna(2 + r.below(6))
        body = r.dna(r.below(14))
        read = r.dna(r.below(5)) + synth.mutate(r, us, r.below(3)) + body + synth.mutate(r, ds, r.below(3)) + r.dna(r.below(5))
        if r.below(100) < 10:
            read = read[: r.below(len(read)) + 1]
        q = bytearray(63 + r.below(11) for _ in range(len(read)))
        for _k in range(r.below(3)):
            q[r.below(len(q))] = 33 + r.below(40)
        if r.below(100) < 10:
            q = q[: r.below(len(q) + 1)]
        mode = r.choice(["both", "up", "down"])
        msu, msd, qsu, qsd, length = r.below(3), r.below(3), r.choice([0, 20, 30, 41]), r.choice([0, 20, 30, 41]), r.choice([3, 8, 20])
        p = refharness.make_param(upstream=us.decode() if mode != "down" else None,
                                  downstream=ds.decode() if mode != "up" else None,
                                  miss_search_up=msu, miss_search_down=msd, qual_up=qsu, qual_down=qsd, length=length)
        if p["upstream"] is not None:
            p["upstream_bin"] = [ref.seq2bin(us.decode())]
        if p["downstream"] is not None:
            p["downstream_bin"] = [ref.seq2bin(ds.decode())]
        s, e = ref.sequence_tinder(ref.seq2bin(read.decode()), bytes(q), p, 0)
        out["sequence_tinder"].append(dict(read=read.decode(), qual=bytes(q).decode(), upstream=p["upstream"], downstream=p["downstream"],
                                           msu=msu, msd=msd, qsu=qsu, qsd=qsd, length=length, i=0,
                                           expect=[None if s is None else int(s), None if e is None else int(e)]))
    out["seq2bin"] = dict(seq="GATTACA", expect=[int(x) for x in ref.seq2bin("GATTACA")])
    return out


def enc(c):
    c = dict(c)
    c.pop("fastq", None)
    return c


def main():
    assert refharness.available(), "reference tree not found"
    t0 = time.time()
    # 1. handcrafted
    kat = []
    for c in cases.kat_cases():
        o = run_case(c)
        o["fastq"] = c["fastq"].decode("latin-1")
        kat.append(o)
    json.dump(kat, open(os.path.join(HERE, "kat_cases.json"), "w"), indent=0)
    print("kat", len(kat), time.time() - t0)
    # 2. fuzz
    fz = []
    for s in range(cases.N_FUZZ):
        c = cases.fuzz_case(s)
        try:
            o = run_case(c)
        except Exception as ex:  # the reference itself raised (e.g. a UnicodeDecodeError): not a usable vector
            print("skip", c["name"], type(ex).__name__, ex)
            continue
        o["fastq"] = c["fastq"].decode("latin-1")
        fz.append(o)
    with gzip.GzipFile(os.path.join(HERE, "fuzz_cases.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(fz).encode())
    print("fuzz", len(fz), time.time() - t0)
    # 3. shaped
    sh = []
    for name in cases.SHAPED:
        params, lib, data = cases.shaped_inputs(name)
        o = run_case(dict(name=name, params=params, library=lib, fastq=data))
        o.pop("library")
        o["sha256"] = cases.sha(data)
        o["nbytes"] = len(data)
        sh.append(o)
        print("shaped", name, o["stats"], time.time() - t0)
    json.dump(sh, open(os.path.join(HERE, "shaped_cases.json"), "w"))
    # 4. primitives
    json.dump(primitives(), open(os.path.join(HERE, "primitive_cases.json"), "w"))
    # 5. config 1 surrogate
    shutil.copyfile(os.path.join(refharness.REF_ROOT, "fast2q", "data", "D39V_guides.csv"), os.path.join(HERE, "D39V_guides.csv"))
    shutil.copyfile(os.path.join(refharness.REF_ROOT, "tests", "compiled.csv"), os.path.join(HERE, "ref_compiled.csv"))
    lib, want, data = cases.config1_surrogate(os.path.join(HERE, "D39V_guides.csv"), os.path.join(HERE, "ref_compiled.csv"))
    counts, stats = refharness.run_reads_counter(data, lib, **cases.P())
    got = {name: counts[seq] for name, seq in lib}
    assert got == want, "reference(surrogate) != tests/compiled.csv"
    json.dump(dict(sha256=cases.sha(data), nbytes=len(data), stats=stats, n_guides=len(lib), total=sum(want.values())),
              open(os.path.join(HERE, "config1_surrogate.json"), "w"))
    print("config1 surrogate ok", stats, time.time() - t0)


if __name__ == "__main__":
    main()
