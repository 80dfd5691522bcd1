"""
Input side of the golden fixtures: handcrafted known-answer cases, a seeded fuzzer and
the config-shaped workloads.  Pure Python/numpy, no reference needed — make_golden.py runs the reference on
these inputs and stores its answers; the tests re-create the same inputs and compare.
"""
from __future__ import annotations

import importlib
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
synth = importlib.import_module("2fast2q_b200.synth")
testdata = importlib.import_module("2fast2q_b200.testdata")

PARAM_KEYS = ("mode", "miss", "phred", "length", "start", "upstream", "downstream", "miss_search_up",
              "miss_search_down", "qual_up", "qual_down")


def P(**kw):
    d = dict(mode="C", miss=1, phred=30, length=20, start="0", upstream=None, downstream=None,
             miss_search_up=0, miss_search_down=0, qual_up=30, qual_down=30)
    d.update(kw)
    return d


def rec(seq: bytes, qual: bytes | None = None, name: bytes = b"@r", eol: bytes = b"\n") -> bytes:
    if qual is None:
        qual = b"I" * len(seq)
    return name + eol + seq + eol + b"+" + eol + qual + eol


G1 = b"AATAGCATAGAAATCATACA"
G2 = b"AGTGTTGATTTACCAACGTT"
G3 = b"TTTCAAGTCCGTTGAACTTT"
LIB3 = [("g1", G1.decode()), ("g2", G2.decode()), ("g3", G3.decode())]
TAIL = b"GGGCCCGGGCCCGGGCCCGGGCCCGGGCCC"


def kat_cases():
    """handcrafted cases; every rule of SURVEY.md Appendix A has at least one"""
    I20 = b"I" * 20
    out = []

    def add(name, fastq, lib=LIB3, **kw):
        if kw.get("mode") == "EC":
            lib = None          # main() loads no library in Extract+Count mode (fast2q.py:1700-1702)
        out.append(dict(name=name, params=P(**kw), library=lib, fastq=fastq))

    add("exact_and_1mm", rec(G1 + TAIL) + rec(G1[:5] + b"T" + G1[6:] + TAIL) + rec(G2 + TAIL))
    # Q29 ('>') passes, Q28 ('=') fails at --ph 30 (A2)
    add("phred_q29_passes", rec(G1 + TAIL, b">" * 50))
    add("phred_q28_fails", rec(G1 + TAIL, b"I" * 7 + b"=" + b"I" * 42))
    add("lowq_outside_window_ignored", rec(G1 + TAIL, I20 + b"!" * 30))
    add("phred_1_filters_nothing", rec(G1 + TAIL, b"!" * 50), phred=1)
    add("phred_0_clamps_to_1", rec(G1 + TAIL, b"!" * 50), phred=0)
    add("phred_neg_clamps_to_1", rec(G1 + TAIL, b"!" * 50), phred=-7)
    add("phred_94", rec(G1 + TAIL, b"}" * 50) + rec(G1 + TAIL, b"~" * 50), phred=94)
    add("phred_95_all_fail", rec(G1 + TAIL, b"~" * 50), phred=95)
    add("phred_200", rec(G1 + TAIL, b"~" * 50) + rec(G2 + TAIL, b" " * 50), phred=200)
    # ties / unique at minimum distance (A6)
    tie_lib = [("a", "AAAAAAAAAAAAAAAAAAAA"), ("b", "AAAAAAAAAAAAAAAAAACC"), ("c", "GGGGGGGGGGGGGGGGGGGG")]
    add("tie_dist1_two_guides", rec(b"AAAAAAAAAAAAAAAAAAAC" + TAIL), lib=tie_lib)
    add("unique_min_dist_wins_m2", rec(b"AAAAAAAAAAAAAAAAAAAT" + TAIL) + rec(b"AAAAAAAAAAAAAAAAAATT" + TAIL)
        + rec(b"AAAAAAAAAAAAAAAAATTT" + TAIL), lib=tie_lib, miss=2)
    add("m2_tie_at_2", rec(b"AAAAAAAAAAAAAAAAAACA" + TAIL) + rec(b"AAAAAAAAAAAAAAAAAAGG" + TAIL), lib=tie_lib, miss=2)
    add("m0_no_search", rec(G1[:5] + b"T" + G1[6:] + TAIL), miss=0)
    add("m3", rec(b"TTTAGCATAGAAATCATACA" + TAIL) + rec(b"TTTTGCATAGAAATCATACA" + TAIL), miss=3)
    add("N_is_one_mismatch", rec(G1[:9] + b"N" + G1[10:] + TAIL) + rec(b"NN" + G1[2:] + TAIL))
    add("N_two_m2", rec(b"NN" + G1[2:] + TAIL) + rec(b"NNN" + G1[3:] + TAIL), miss=2)
    add("lowercase_fixed_matches", rec(G1.lower() + TAIL.lower()))
    add("mixed_case", rec(G1[:10].lower() + G1[10:] + TAIL))
    # short reads (A3/A4 slice clamping)
    add("short_read_15bp_counter", rec(G1[:15]))
    add("short_read_15bp_ec", rec(G1[:15]) + rec(G1[:15]) + rec(G1), mode="EC")
    add("empty_seq_line", b"@r\n\n+\n\n" + rec(G1 + TAIL))
    add("empty_key_in_library", b"@r\n\n+\n\n", lib=LIB3 + [("empty", "")])
    add("start_beyond_read", rec(G1), start="30")
    add("negative_start", rec(TAIL + G1, None), start="-20")
    add("negative_start_long", rec(G1), start="-50")
    add("length_0", rec(G1 + TAIL), length=0, mode="EC")
    add("length_neg", rec(G1 + TAIL), length=-2, mode="EC")
    add("length_30_longer_lib", rec(G1 + TAIL), length=30, lib=LIB3 + [("long", (G1 + TAIL[:10]).decode())])
    # quality line shorter / longer than the sequence line: slices clamp independently
    add("qual_shorter_than_seq", rec(G1 + TAIL, b"I" * 10) + rec(G1 + TAIL, b"!" * 3) + rec(G1 + TAIL, b""))
    add("qual_longer_than_seq", rec(G1, b"I" * 20 + b"!" * 10))
    # line structure (A1)
    add("crlf", rec(G1 + TAIL, eol=b"\r\n") + rec(G2 + TAIL, eol=b"\r\n"))
    add("trailing_spaces_tabs", b"@r \n" + G1 + b" \t\n+\n" + I20 + b"  \n")
    add("whitespace_inside_window_after_strip", b"@r\n" + G1[:18] + b" \t\n+\n" + I20 + b"\n", mode="EC")
    add("trailing_3_line_partial", rec(G1 + TAIL) + b"@r\n" + G2 + b"\n+\n")
    add("last_line_no_newline", rec(G1 + TAIL) + b"@r\n" + G2 + b"\n+\n" + I20)
    add("last_qual_empty_after_newline", rec(G1 + TAIL) + b"@r\n" + G2 + b"\n+\n")
    add("fourth_line_blank", b"@r\n" + G2 + b"\n+\n\n")
    add("blank_line_shifts_phase", b"\n" + rec(G1 + TAIL) + rec(G2 + TAIL) + b"x\ny\nz\n")
    add("only_newlines", b"\n" * 9)
    add("only_newlines_ec", b"\n" * 9, mode="EC")
    add("empty_file", b"")
    add("no_newline_at_all", b"ACGT")
    add("at_sign_in_quality", rec(G1 + TAIL, b"@" * 50) + rec(G2 + TAIL, b"+" * 50), phred=10)
    add("header_with_high_bytes", b"@r\xff\xfe\xc3\xa9\n" + G1 + b"\n+\xff\n" + I20 + b"\n")
    add("vertical_tab_formfeed", b"@r\n" + G1 + b"\x0b\x0c\n+\n" + I20 + b"\x0c\n")
    # multi feature (A5)
    dual = [("d1", (G1 + b":" + G2).decode()), ("s1", G1.decode()), ("s2", G2.decode())]
    pad10 = b"CCCCCCCCCC"
    add("dual_both_pass", rec(G1 + pad10 + G2), lib=dual, start="0,30")
    add("dual_second_lowq", rec(G1 + pad10 + G2, I20 + b"I" * 10 + b"!" + b"I" * 19), lib=dual, start="0,30")
    add("dual_first_lowq", rec(G1 + pad10 + G2, b"!" + b"I" * 49), lib=dual, start="0,30")
    add("dual_both_lowq", rec(G1 + pad10 + G2, b"!" + b"I" * 29 + b"!" + b"I" * 19), lib=dual, start="0,30")
    add("dual_1mm_in_second", rec(G1 + pad10 + G2[:3] + b"A" + G2[4:]), lib=dual, start="0,30")
    add("dual_colon_compare", rec(G1 + pad10 + G2[:19]), lib=dual + [("w", (G1 + b":" + G2[:19]).decode())], start="0,30")
    add("dual_ec", rec(G1 + pad10 + G2) + rec(G1 + pad10 + G2, b"!" + b"I" * 49), start="0,30", mode="EC")
    add("dual_short_read_second_empty", rec(G1 + pad10), lib=dual + [("e", G1.decode() + ":")], start="0,30")
    add("triple", rec(G1 + G2 + G3), lib=[("t", (G1 + b":" + G2 + b":" + G3).decode())], start="0,20,40")
    add("overlapping_windows", rec(G1 + TAIL), start="0,5", mode="EC")
    # delimiter modes (A4)
    US, DS = b"GATTACA", b"GCACGGG"
    bc = b"ACGTACGTAC"
    add("delim_both_ec", rec(b"TT" + US + bc + DS + b"TTTT") + rec(US + b"AAA" + DS), mode="EC", upstream="GATTACA",
        downstream="GCACGGG")
    add("delim_adjacent_empty_feature", rec(b"CC" + US + DS + b"CC"), mode="EC", upstream="GATTACA", downstream="GCACGGG")
    add("delim_case_sensitive", rec((b"TT" + US + bc + DS).lower()) + rec(b"TT" + US + bc.lower() + DS), mode="EC",
        upstream="GATTACA", downstream="GCACGGG")
    add("delim_lowercase_cli_arg", rec(b"TT" + US + bc + DS), mode="EC", upstream="gattaca", downstream="gcacggg")
    add("delim_up_only", rec(b"TT" + US + bc + b"CCCCCCCCCCCCCC") + rec(b"TT" + US + bc[:4]), mode="EC",
        upstream="GATTACA", length=10)
    add("delim_down_only", rec(bc + bc + DS) + rec(bc[:4] + DS + bc) + rec(DS + bc), mode="EC",
        downstream="GCACGGG", length=10)
    add("delim_down_only_negative_wrap", rec(b"ACGT" + DS + b"TTTTTTTTTTTTTTTTTTTTTTTTT"), mode="EC", downstream="GCACGGG",
        length=20)
    add("delim_down_only_wrap_both_lines", rec(b"ACGT" + DS + b"TTTTTTTTT", b"I" * 30), mode="EC", downstream="GCACGGG",
        length=20)
    add("delim_mismatch_allowed", rec(b"TT" + b"GATTTCA" + bc + b"GCACGGC"), mode="EC", upstream="GATTACA",
        downstream="GCACGGG", miss_search_up=1, miss_search_down=1)
    add("delim_mismatch_not_allowed", rec(b"TT" + b"GATTTCA" + bc + DS), mode="EC", upstream="GATTACA", downstream="GCACGGG")
    add("delim_first_occurrence_only", rec(US + b"AAAA" + US + bc + DS, b"!" * 7 + b"I" * 35), mode="EC", upstream="GATTACA",
        downstream="GCACGGG")
    add("delim_quality_in_delims", rec(b"TT" + US + bc + DS, b"II" + b"I" * 7 + b"I" * 10 + b"5" * 7), mode="EC",
        upstream="GATTACA", downstream="GCACGGG", qual_down=21)
    add("delim_quality_in_delims_pass", rec(b"TT" + US + bc + DS, b"II" + b"I" * 7 + b"I" * 10 + b"5" * 7), mode="EC",
        upstream="GATTACA", downstream="GCACGGG", qual_down=20)
    add("delim_feature_quality", rec(b"TT" + US + bc + DS, b"II" + b"I" * 7 + b"I" * 4 + b"#" + b"I" * 5 + b"I" * 7), mode="EC",
        upstream="GATTACA", downstream="GCACGGG")
    add("delim_no_downstream_after_up", rec(DS + b"TT" + US + bc), mode="EC", upstream="GATTACA", downstream="GCACGGG")
    add("delim_counter_mode", rec(b"TT" + US + G1 + DS) + rec(b"T" + US + G1[:7] + b"C" + G1[8:] + DS), upstream="GATTACA",
        downstream="GCACGGG")
    add("delim_two_pairs", rec(b"AAAAAACACACACACACACACATTCAGGGGGGCCAAAAATAGAGAGAGAGAGACCGAGAGGGGGTTAGCATCG", b"B" * 73),
        mode="EC", upstream="CACACATT,GAGACCGA", downstream="TAGAGAGA,TAGCATCG")
    add("delim_two_pairs_one_missing", rec(b"AAAAAACACACACACACACACATTCAGGGGGGCCAAAAATAGAGAGAGAGAGACCGAGAGGGGGTTAGCATCC", b"B" * 73),
        mode="EC", upstream="CACACATT,GAGACCGA", downstream="TAGAGAGA,TAGCATCG")
    add("delim_two_ups_one_down", rec(b"AAAAAACACACACACACACACATTCAGGGGGGCC", b"B" * 34), mode="EC", upstream="CACACATT,GAGACCGA",
        length=5)
    add("delim_longer_than_read", rec(b"ACGT"), mode="EC", upstream="GATTACAGATTACA", downstream="GCACGGG")
    add("delim_msu_large", rec(b"ACGTACGTACGTACGTACGTACGTA"), mode="EC", upstream="GATTACA", downstream="GCACGGG",
        miss_search_up=7, miss_search_down=7)
    add("delim_qual_shorter", rec(b"TT" + US + bc + DS, b"I" * 5), mode="EC", upstream="GATTACA", downstream="GCACGGG")
    return out


# ---------------------------------------------------------------------------------------------------
# seeded fuzzer
# ---------------------------------------------------------------------------------------------------
def fuzz_case(seed: int):
    r = synth.SM64(0xF00D0000 + seed)
    style = r.choice(["fixed", "fixed", "fixed", "multi", "both", "both", "up", "down"])
    mode = "EC" if r.below(100) < 30 else "C"
    l = r.choice([4, 6, 8, 8, 10, 12, 20])
    params = P(mode=mode, miss=r.choice([0, 1, 1, 2, 2, 3]), phred=r.choice([0, 1, 2, 15, 30, 30, 30, 31, 42, 94, 95, 200]),
               length=l)
    alpha = b"ACGT" if r.below(100) < 85 else b"ACGTN"
    nlib = 3 + r.below(40)
    keys = []
    for _ in range(nlib):
        t = r.below(100)
        if keys and t < 40:
            k = synth.mutate(r, r.choice(keys)[:l].ljust(l, b"A"), 1 + r.below(2), alpha)
        elif t < 50:
            k = r.dna(max(1, l + r.below(5) - 2), alpha)
        else:
            k = r.dna(l, alpha)
        keys.append(k)
    us = ds = us2 = ds2 = None
    n_iter = 1
    if style == "multi":
        n_iter = 2 + r.below(2)
        gap = r.below(4)
        starts = [i * (l + gap) for i in range(n_iter)]
        if r.below(100) < 15:
            starts[0] = -r.below(3 * l) - 1
        params["start"] = ",".join(str(s) for s in starts)
        multi = []
        for _ in range(nlib):
            parts = [r.choice(keys)[:l].ljust(l, b"C") for _ in range(n_iter)]
            multi.append(b":".join(parts))
        keys = [k[:l].ljust(l, b"C") for k in keys[: nlib // 2]] + multi
    elif style == "fixed":
        params["start"] = str(r.choice([0, 0, 0, 1, 3, 7, -5, -l, 40]))
    else:
        us, ds = r.dna(3 + r.below(6)), r.dna(3 + r.below(6))
        params["miss_search_up"], params["miss_search_down"] = r.choice([0, 0, 1, 2]), r.choice([0, 0, 1, 2])
        params["qual_up"], params["qual_down"] = r.choice([0, 10, 30, 30, 41]), r.choice([0, 10, 30, 30, 41])
        if r.below(100) < 25:
            us2, ds2 = r.dna(4 + r.below(4)), r.dna(4 + r.below(4))
            n_iter = 2
        if style in ("both", "up"):
            params["upstream"] = (us + (b"," + us2 if us2 else b"")).decode()
        if style in ("both", "down"):
            params["downstream"] = (ds + (b"," + ds2 if ds2 else b"")).decode()
        if r.below(100) < 10 and params["upstream"]:
            params["upstream"] = params["upstream"].lower()
        if n_iter == 2:
            keys = keys[: nlib // 2] + [r.choice(keys) + b":" + r.choice(keys) for _ in range(nlib // 2)]
    # de-duplicate keeping first (features_loader)
    seen, lib = set(), []
    for i, k in enumerate(keys):
        if k not in seen:
            seen.add(k)
            lib.append(("f%d" % i, k.decode()))
    keys = [k.encode() for _, k in lib]
    eol = b"\r\n" if r.below(100) < 10 else b"\n"
    nreads = 20 + r.below(100)
    chunks = []
    for i in range(nreads):
        def feat():
            k = r.choice(keys).split(b":")[0]
            t = r.below(100)
            if t < 45:
                return k
            if t < 70:
                return synth.mutate(r, k, 1, alpha)
            if t < 82:
                return synth.mutate(r, k, 2, alpha)
            if t < 88:
                return synth.mutate(r, k, 3, alpha)
            if t < 94 and len(k):
                p = r.below(len(k))
                return k[:p] + b"N" + k[p + 1:]
            return r.dna(len(k) + r.below(3) - 1 if len(k) > 1 else 1)
        if style in ("fixed", "multi"):
            st = [int(s) for s in params["start"].split(",")]
            s = bytearray()
            for j, a in enumerate(st):
                a = max(a, 0)
                while len(s) < a:
                    s += r.dna(1)
                s += feat()
            s = bytes(s) + r.dna(r.below(12))
            if r.below(100) < 8:
                s = s[: r.below(len(s) + 1)]
        else:
            s = r.dna(r.below(6))
            u_, d_ = us, ds
            if r.below(100) < 20:
                u_ = synth.mutate(r, us, 1 + r.below(2))
            if r.below(100) < 20:
                d_ = synth.mutate(r, ds, 1 + r.below(2))
            if r.below(100) < 6:
                u_ = b""
            if r.below(100) < 6:
                d_ = b""
            s += u_ + feat() + d_ + r.dna(r.below(5))
            if n_iter == 2:
                s += (us2 if r.below(100) < 85 else r.dna(len(us2))) + feat() + (ds2 if r.below(100) < 85 else b"") + r.dna(r.below(4))
            if r.below(100) < 5:
                s = s[: r.below(len(s) + 1)]
        if r.below(100) < 10:
            a = r.below(len(s) + 1)
            b_ = a + r.below(len(s) - a + 1)
            s = s[:a] + s[a:b_].lower() + s[b_:]
        q = bytearray(63 + r.below(11) for _ in range(len(s)))
        for _ in range(r.choice([0, 0, 0, 0, 1, 1, 2, 5])):
            if q:
                q[r.below(len(q))] = r.choice([33, 35, 47, 61, 62, 63, 34 + r.below(60), 125, 126])
        q = bytes(q)
        t = r.below(100)
        if t < 4:
            q = q[: r.below(len(q) + 1)]
        elif t < 7:
            q = q + bytes(63 + r.below(11) for _ in range(1 + r.below(5)))
        hdr = b"@f%d" % i + (b" x~{:" if r.below(100) < 5 else b"")
        pad = r.choice([b"", b"", b"", b" ", b"\t", b" \r"]) if r.below(100) < 10 else b""
        chunks.append(hdr + eol + s + pad + eol + b"+" + eol + q + pad + eol)
        if r.below(1000) < 6:
            chunks.append(eol)            # stray blank line: shifts the 4-line phase for everything after it
    data = b"".join(chunks)
    t = r.below(100)
    if t < 10 and data.endswith(eol):
        data = data[: -len(eol)]          # final line unterminated
    elif t < 18:
        data = data[: len(data) - r.below(min(len(data), 60) + 1)]   # truncated tail
    return dict(name="fuzz%04d" % seed, params=params, library=None if mode == "EC" else lib, fastq=data)


N_FUZZ = 400


# ---------------------------------------------------------------------------------------------------
# config-shaped workloads: inputs are regenerated from the spec, the fixture stores sha256 + answers
# ---------------------------------------------------------------------------------------------------
def shaped_inputs(name: str):
    """returns (params, library [(name, seq)] or None, fastq bytes)"""
    if name == "config2_slice":       # 2k guides, 50 bp, m=1  (BASELINE.json configs[1]), first 60k reads
        names, keys = synth.make_library(2, 2000, 20)
        spec = synth.default_spec(2)
        data = synth.fixed_reads(keys, 0, 60000, **spec).tobytes()
        return P(miss=1), list(zip(names, [k.decode() for k in keys])), data
    if name == "config3_slice":       # 75 bp, m=2, library cut to 8k guides so the reference finishes; 12k reads
        names, keys = synth.make_library(3, 8000, 20)
        spec = synth.default_spec(3)
        data = synth.fixed_reads(keys, 0, 12000, **spec).tobytes()
        return P(miss=2), list(zip(names, [k.decode() for k in keys])), data
    if name == "config3_m3":          # same generator, m=3, 1k guides
        names, keys = synth.make_library(33, 1000, 20)
        spec = synth.default_spec(3)
        data = synth.fixed_reads(keys, 500, 8000, **spec).tobytes()
        return P(miss=3), list(zip(names, [k.decode() for k in keys])), data
    if name == "config4_barseq":      # EC, both delimiters with 1 mismatch each
        data = synth.barseq_reads(4, 15000)
        return P(mode="EC", upstream="GTTCAGAGTTCT", downstream="CTGAATAGGCCA", miss_search_up=1, miss_search_down=1), None, data
    if name == "config4_up_only":
        data = synth.barseq_reads(44, 6000)
        return P(mode="EC", upstream="GTTCAGAGTTCT", miss_search_up=1, length=20), None, data
    if name == "config4_down_only":
        data = synth.barseq_reads(45, 6000)
        return P(mode="EC", downstream="CTGAATAGGCCA", miss_search_down=1, length=20), None, data
    if name == "config5a_dual_fixed":
        names, keys, xs, ys = synth.dual_library(5, 1500)
        data = synth.dual_reads(5, 12000, xs, ys, mode="fixed")
        return P(miss=1, start="0,30"), list(zip(names, [k.decode() for k in keys])), data
    if name == "config5b_dual_delim":
        names, keys, xs, ys = synth.dual_library(5, 1500)
        data = synth.dual_reads(5, 12000, xs, ys, mode="delim")
        return P(miss=1, upstream="ACCGGT,GGATCC", downstream="TTGACA,CAATTG"), list(zip(names, [k.decode() for k in keys])), data
    raise KeyError(name)


SHAPED = ["config2_slice", "config3_slice", "config3_m3", "config4_barseq", "config4_up_only", "config4_down_only",
          "config5a_dual_fixed", "config5b_dual_delim"]


def sha(data: bytes) -> str:
    return hashlib.sha256(data).hexdigest()


# ---------------------------------------------------------------------------------------------------
# config 1 surrogate: example.fastq.gz is missing from the reference checkout, so rebuild a FASTQ whose
# reference answer is exactly tests/compiled.csv
# ---------------------------------------------------------------------------------------------------
def load_guides_csv(path):
    return testdata.load_guides_csv(path)


def config1_surrogate(guides_csv, compiled_csv):
    """the generator lives in the package (2fast2q_b200/testdata.py) because `-c -t` needs it too"""
    return testdata.config1_surrogate(guides_csv, compiled_csv)
