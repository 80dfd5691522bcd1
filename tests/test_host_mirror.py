"""
CPU tests of the host-side mirror (2fast2q_b200/fast2q.py) against fixtures produced by the UNMODIFIED reference
(tests/golden/make_cli_golden.py): features .csv loader, command-line parser, per-sample csv, compiled.csv and
compiled_stats.csv writers, gzip streaming, record-aligned sharding.  No kernel is called here.
"""
import base64
import gzip
import importlib
import io
import os
import random
import re
import zlib

import numpy as np
import pytest

import cli_golden as CG

fq = importlib.import_module("2fast2q_b200.fast2q")


@pytest.mark.parametrize("name", sorted(CG.load()["loader"]))
def test_features_loader_matches_reference(name, tmp_path, capsys):
    c = CG.load()["loader"][name]
    p = tmp_path / "lib.csv"
    p.write_text(c["text"], newline="")
    if c["expect"] == "FATAL":
        with pytest.raises(SystemExit):
            fq.features_loader(str(p))
        return
    feats = fq.features_loader(str(p))
    assert [[seq, f.name] for seq, f in feats.items()] == c["expect"]
    assert all(f.counts == 0 for f in feats.values())


def test_features_loader_missing_file_is_fatal(tmp_path):
    with pytest.raises(SystemExit):
        fq.features_loader(str(tmp_path / "nope.csv"))


@pytest.mark.parametrize("k", range(len(CG.load()["parser"])))
def test_input_parser_matches_reference(k, tmp_path, monkeypatch):
    c = CG.load()["parser"][k]
    monkeypatch.chdir(tmp_path)
    got = fq.input_parser(c["argv"])
    for key, want in c["expect"].items():
        if isinstance(want, str):
            want = want.replace("<CWD>", str(tmp_path))
        assert got[key] == want, key
    assert set(got) - set(c["expect"]) == {"gpus"}          # the one documented addition


def test_input_parser_gui_mode_and_version(capsys):
    assert fq.input_parser([]) is None                       # no -c: the reference opens its GUI (out of scope here)
    with pytest.raises(SystemExit):
        fq.initializer(None)
    with pytest.raises(SystemExit):
        fq.input_parser(["-v"])
    assert "Version: 2.8.1" in capsys.readouterr().out


def test_initializer_clamps_and_sets(tmp_path):
    p = fq.input_parser(["-c", "--s", "/d", "--g", "/l.csv", "--o", str(tmp_path), "--ph", "0", "--qsu", "-3", "--qsd", "31", "--cp", "1"])
    p = fq.initializer(p)
    assert p["phred"] == 1 and p["qual_up"] == 1 and p["quality_set"] == set() and p["quality_set_up"] == set()
    assert p["quality_set_down"] == {chr(33 + i) for i in range(30)}          # Q0..Q29 fail at 31
    assert re.fullmatch(r"2FAST2Q_output_\d{4}(_\d\d){5}", os.path.basename(p["directory"]))
    assert p["cpu"] == 1


def _param_for(case, tmp_path, directory):
    args = ["-c", "--s", str(tmp_path / "in"), "--o", str(tmp_path / "out")] + case["args"]
    if case["library"] is not None:
        args += ["--g", str(tmp_path / "library.csv")]
    p = fq.initializer(fq.input_parser(args))
    p["directory"] = str(directory)
    p["used_cmd"] = p["used_cmd"].replace(str(tmp_path), "<TMP>")
    return p


@pytest.mark.parametrize("name", [c["name"] for c in CG.load()["cli"] if "--k" in c["args"]])
def test_compiling_reproduces_reference_bytes(name, tmp_path):
    """golden *_reads.csv in -> compiled.csv and compiled_stats.csv out, byte for byte"""
    c = CG.case(name)
    d = tmp_path / "res"
    d.mkdir()
    for fn, text in c["outputs"].items():
        if fn.endswith("_reads.csv"):
            (d / fn).write_text(text, newline="")
    p = _param_for(c, tmp_path, d)
    p["delete"] = True
    fq.compiling(p)
    assert (d / "compiled.csv").read_bytes().decode() == c["outputs"]["compiled.csv"]
    assert (d / "compiled_stats.csv").read_bytes().decode() == c["outputs"]["compiled_stats.csv"]
    # intermediates removed (fast2q.py:1375-1377); the four summary plots stand beside the two csv files
    assert sorted(f for f in os.listdir(d) if not f.endswith(".png")) == ["compiled.csv", "compiled_stats.csv"]


@pytest.mark.parametrize("name", ["cli_counter", "cli_single_split", "cli_dual_delim"])
def test_write_sample_reproduces_reference_bytes(name, tmp_path):
    c = CG.case(name)
    (tmp_path / "library.csv").write_text(c["library"], newline="")
    feats = fq.features_loader(str(tmp_path / "library.csv"))
    d = tmp_path / "res"
    d.mkdir()
    p = _param_for(c, tmp_path, d)
    for fn, text in c["outputs"].items():
        if not fn.endswith("_reads.csv"):
            continue
        lines = text.split("\r\n")
        t = lines[0].split()
        secs = float(t[3])
        stats = dict(reads=int(t[12]), perfect_counter=int(t[15]), imperfect_counter=int(t[19]), non_aligned_counter=int(t[24]),
                     quality_failed=int(t[32]))
        by_name = dict(l.split(",") for l in lines[2:] if l)
        sample = {seq: fq.Features(f.name, int(by_name[f.name])) for seq, f in feats.items()}
        raw = [k for k in c["files"] if fq.sample_name(k) == fn[:-len("_reads.csv")]][0]
        fq.write_sample(str(tmp_path / "in" / raw), sample, stats, p, secs)
        assert (d / fn).read_bytes().decode() == text
        # the command line's writer (straight from the count vector) produces the same bytes
        (d / fn).unlink()
        fq._write_sample_counts(str(tmp_path / "in" / raw), feats, [int(by_name[f.name]) for f in feats.values()], stats, p, secs)
        assert (d / fn).read_bytes().decode() == text


@pytest.mark.parametrize("names", [["10", "9", "-3", "007", "9"], ["b", "a,1", 'q"uote', "B", "a"], ["1", "x", "2"], ["3"]])
def test_count_vector_writer_equals_write_sample(names, tmp_path):
    """integer / text / mixed names, duplicates (stable order), names the csv module has to quote"""
    feats = {f"SEQ{k}": fq.Features(n, 0) for k, n in enumerate(names)}
    counts = [7 * k + 1 for k in range(len(names))]
    stats = dict(reads=100, perfect_counter=50, imperfect_counter=10, non_aligned_counter=30, quality_failed=10)
    outs = []
    for k, writer in enumerate(("object", "vector")):
        d = tmp_path / writer
        d.mkdir()
        p = {"directory": str(d), "Progress bar": True}
        if writer == "object":
            fq.write_sample("/x/s1.fastq.gz", {q: fq.Features(f.name, c) for (q, f), c in zip(feats.items(), counts)}, stats, p, 1.5)
        else:
            fq._write_sample_counts("/x/s1.fastq.gz", feats, counts, stats, p, 1.5)
        outs.append((d / "s1_reads.csv").read_bytes())
    assert outs[0] == outs[1]


def test_input_kind_sniffs_bgzf_and_gzip(tmp_path):
    text = b"@r\nACGT\n+\nIIII\n" * 50
    eof = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
    c = zlib.compressobj(6, zlib.DEFLATED, -15)
    cd = c.compress(text) + c.flush()
    import struct
    blk = b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", 18 + len(cd) + 8 - 1) + cd + struct.pack("<II", zlib.crc32(text), len(text))
    (tmp_path / "a.fastq").write_bytes(text)
    (tmp_path / "b.fastq.gz").write_bytes(blk + eof)
    (tmp_path / "c.fastq.gz").write_bytes(gzip.compress(text))
    (tmp_path / "d.fastq.gz").write_bytes(blk)                          # bgzip blocks without the end-of-file block
    f = lambda *n: fq._input_kind([str(tmp_path / x) for x in n])
    assert f("a.fastq") == "plain" and f("a.fastq", "b.fastq.gz") == "bgzf" and f("b.fastq.gz", "c.fastq.gz") == "gzip"
    assert f("d.fastq.gz") == "gzip" and f("missing.fastq.gz") == "gzip"


def test_sample_name_rules():
    assert fq.sample_name("/x/a.fastq.gz") == "a"
    assert fq.sample_name("/x/a.fastq") == "a"
    assert fq.sample_name("/x/a.fq.gz") == "a.fq"
    assert fq.sample_name("/x/a.b.fastq.gz") == "a.b"
    assert fq._timing_text(3.14159) == "3.14 seconds" and fq._timing_text(61) == "1.02 minutes" and fq._timing_text(7200) == "2.0 hours"


def _py_gzip_lines(path):
    """what the reference's `for line in gzip.open(raw, 'rb')` delivers before it ends or raises EOFError"""
    out, ok = [], True
    try:
        with gzip.open(path, "rb") as f:
            for line in f:
                out.append(line)
    except EOFError:
        ok = False
    return b"".join(out), ok


def _gz(data, level=6):
    b = io.BytesIO()
    with gzip.GzipFile(fileobj=b, mode="wb", compresslevel=level, mtime=0) as f:
        f.write(data)
    return b.getvalue()


def test_inflate_blocks_equals_gzip_module(tmp_path):
    rnd = random.Random(5)
    body = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(50)), b"I" * 50) for i in range(4000))
    variants = {
        "plain": _gz(body),
        "multi_member": _gz(body[:100000]) + _gz(body[100000:300000], 1) + _gz(body[300000:], 9),
        "zero_padded": _gz(body[:5000]) + b"\0" * 37 + _gz(body[5000:]) + b"\0" * 5,
        "empty_member": _gz(b"") + _gz(body[:999]),
        "empty_file": b"",
        "unterminated": _gz(body + b"@last\nACGT"),
    }
    for name, blob in variants.items():
        p = tmp_path / (name + ".fastq.gz")
        p.write_bytes(blob)
        want, ok = _py_gzip_lines(p)
        assert ok
        assert b"".join(fq._inflate_blocks(str(p), want=70001)) == want, name
    # truncation anywhere: every decodable byte is delivered, then TruncatedGzip
    whole = variants["multi_member"]
    for cut in [9, 10, 500, len(whole) // 3, len(whole) // 2, len(whole) - 9, len(whole) - 1]:
        p = tmp_path / "cut.fastq.gz"
        p.write_bytes(whole[:cut])
        want, ok = _py_gzip_lines(p)
        got, trunc = [], False
        try:
            for b in fq._inflate_blocks(str(p), want=4096):
                got.append(b)
        except fq.TruncatedGzip:
            trunc = True
        got = b"".join(got)
        assert trunc == (not ok), cut
        assert got[:got.rfind(b"\n") + 1] == want, cut      # the reference's iterator never yields the unfinished line
    p = tmp_path / "garbage.fastq.gz"
    p.write_bytes(b"this is not gzip")
    with pytest.raises(zlib.error):
        list(fq._inflate_blocks(str(p)))


def test_record_aligned_shards():
    rnd = random.Random(11)
    data = b"".join(b"@r%d\n" % i + bytes(rnd.choice(b"ACGT") for _ in range(rnd.randint(0, 30))) + b"\n+\n" +
                    b"I" * rnd.randint(0, 30) + b"\n" for i in range(3000)) + b"@x\nAC"

    def blocks(n):
        for o in range(0, len(data), n):
            yield data[o:o + n]
    for bs in (7, 100, 1000, 4096, 1 << 20):
        for sb in (1, 50, 1000, 5000, 1 << 22):
            out = list(fq.record_aligned_shards(blocks(bs), sb))
            assert b"".join(x for x, _ in out) == data
            assert [f for _, f in out] == [False] * (len(out) - 1) + [True]
            for x, _ in out[:-1]:
                assert x and x.endswith(b"\n") and x.count(b"\n") % 4 == 0
    assert list(fq.record_aligned_shards(iter([]))) == [(b"", True)]
    assert list(fq.record_aligned_shards(iter([b"\n\n\n"]), 1)) == [(b"\n\n\n", True)]


def test_merge_sample_adds_counts_and_stats():
    lib = {"AAAA": fq.Features("a", 0), "CCCC": fq.Features("b", 0)}
    s1 = dict(reads=5, perfect_counter=3, imperfect_counter=1, non_aligned_counter=1, quality_failed=0)
    s2 = dict(reads=2, perfect_counter=1, imperfect_counter=0, non_aligned_counter=0, quality_failed=1)
    r1 = ({"AAAA": fq.Features("a", 3), "CCCC": fq.Features("b", 1)}, s1)
    r2 = ({"AAAA": fq.Features("a", 0), "CCCC": fq.Features("b", 1)}, s2)
    m, s = fq._merge_sample({"Running Mode": "C"}, lib, [r1, r2])
    assert [(k, v.name, v.counts) for k, v in m.items()] == [("AAAA", "a", 3), ("CCCC", "b", 2)]
    assert s == dict(reads=7, perfect_counter=4, imperfect_counter=1, non_aligned_counter=1, quality_failed=1)
    e1 = ({"GG": fq.Features("GG", 2)}, s1)
    e2 = ({"TT": fq.Features("TT", 1), "GG": fq.Features("GG", 5)}, s2)
    m, s = fq._merge_sample({"Running Mode": "EC"}, {}, [e1, e2])
    assert {k: v.counts for k, v in m.items()} == {"GG": 7, "TT": 1}


def test_seq2bin_known_answer():
    assert fq.seq2bin("GATTACA").tolist() == [71, 65, 84, 84, 65, 67, 65]      # tests/test_mainfunctions.py:4-8
    assert fq.seq2bin("GATTACA").dtype.name == "int8"


def _bgzf(data, bs=65280, level=4):
    """a BGZF (bgzip) file of `data`: independent gzip members with a 'BC' extra field holding their size + the EOF block"""
    import struct
    import zlib
    out = []
    for o in list(range(0, len(data), bs)) + [None]:
        chunk = b"" if o is None else data[o:o + bs]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cd = c.compress(chunk) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", 18 + len(cd) + 8 - 1) + cd
                   + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def test_bgzf_block_parallel_inflate(tmp_path):
    """bgzip input is inflated block-parallel, in order; truncation and a foreign trailing member behave like the serial reader"""
    import gzip
    import random
    host = importlib.import_module("2fast2q_b200.fast2q")
    rnd = random.Random(7)
    data = b"".join(b"@r%d\n%s\n+\n%s\n" % (i, bytes(rnd.choice(b"ACGT") for _ in range(40)), b"I" * 40) for i in range(30000))
    blob = _bgzf(data)
    p = tmp_path / "x.fastq.gz"
    p.write_bytes(blob)
    assert gzip.open(p).read() == data                                  # (it IS a valid gzip file)
    pieces = list(host._inflate_blocks(str(p), want=1 << 20))
    assert b"".join(pieces) == data and max(map(len, pieces)) <= 1 << 20
    assert b"".join(host._inflate_blocks(str(p), parallel=False)) == data
    # cut inside a block: both readers give the same decodable prefix, then TruncatedGzip
    p.write_bytes(blob[:len(blob) // 2 + 777])
    got = []
    for par in (True, False):
        acc = []
        with pytest.raises(host.TruncatedGzip):
            for b in host._inflate_blocks(str(p), parallel=par):
                acc.append(b)
        got.append(b"".join(acc))
    assert got[0] == got[1] and data.startswith(got[0]) and len(got[0]) > len(data) // 3
    # bgzip blocks followed by an ordinary gzip member, and an ordinary file
    p.write_bytes(blob + gzip.compress(b"@t\nACGT\n+\nIIII\n"))
    assert b"".join(host._inflate_blocks(str(p))) == data + b"@t\nACGT\n+\nIIII\n"
    p.write_bytes(gzip.compress(data))
    assert b"".join(host._inflate_blocks(str(p))) == data


def test_test_mode_data_is_bundled(tmp_path, monkeypatch):
    """`-c -t` must find its data (the data directory was not shipped): the surrogate example.fastq.gz is
    generated on first use and is the stream whose sha256 the golden fixture pins"""
    import gzip
    import hashlib
    import golden_io as G
    td = importlib.import_module("2fast2q_b200.testdata")
    monkeypatch.setattr(td, "EXAMPLE", str(tmp_path / "example.fastq.gz"))
    path = td.ensure_example()
    assert os.path.exists(td.GUIDES) and os.path.exists(td.EXPECTED)
    data = gzip.open(path, "rb").read()
    assert hashlib.sha256(data).hexdigest() == G.config1()["sha256"]
    assert open(td.EXPECTED).read() == open(os.path.join(G.HERE, "ref_compiled.csv")).read()
    fq = importlib.import_module("2fast2q_b200.fast2q")
    monkeypatch.chdir(tmp_path)
    p = fq.input_parser(["-c", "-t"])
    assert p["test_mode"] and p["seq_files"] == path and p["feature"] == td.GUIDES and p["out"] == str(tmp_path)


def test_output_folder_holds_the_six_files(tmp_path):
    """the reference's tests/test_cli.py:17-25 counts six files in the output folder: compiled.csv, compiled_stats.csv and
    the four summary plots (drawn here by 2fast2q_b200/plots.py with Pillow)"""
    pytest.importorskip("PIL")
    from PIL import Image
    c = CG.case("cli_counter")
    d = tmp_path / "res"
    d.mkdir()
    for fn, text in c["outputs"].items():
        if fn.endswith("_reads.csv"):
            (d / fn).write_text(text, newline="")
    p = _param_for(c, tmp_path, d)
    p["delete"] = True
    fq.compiling(p)
    names = sorted(os.listdir(d))
    fnm = p["out_file_name"]
    assert names == sorted([f"{fnm}.csv", f"{fnm}_stats.csv", f"{fnm}_reads_plot.png", f"{fnm}_reads_plot_percentage.png",
                            f"{fnm}_distribution_plot.png", f"{fnm}_distribution_normalized_RPM_plot.png"])
    for n in names:
        if n.endswith(".png"):
            with Image.open(d / n) as im:
                im.verify()
            with Image.open(d / n) as im:
                assert im.size[0] == 1800 and im.size[1] >= 450 and len(np.unique(np.asarray(im))) >= 3      # axes, ink and at least one colour
    assert (d / f"{fnm}.csv").read_bytes().decode() == c["outputs"][f"{fnm}.csv"]


def test_plots_survive_degenerate_samples(tmp_path):
    """an empty sample (0 reads: the reference divides by zero here), one feature, identical counts"""
    pytest.importorskip("PIL")
    plots = importlib.import_module("2fast2q_b200.plots")
    table = [["#x"], ["#Sample name"] + ["h"] * 8, ["a", "1", "seconds", "0", "0", "0", "0", "0", "0"], ["b", "1", "seconds", "10", "5", "5", "0", "3", "2"]]
    out = plots.write_all(table, ["#Feature", "a", "b"], {"g1": [0, 5]}, str(tmp_path / "x"))
    assert len(out) == 4 and all(os.path.getsize(f) > 100 for f in out)
    out = plots.write_all(table, ["#Feature", "a", "b"], {"g1": [0, 5], "g2": [0, 5], "g3": [0, 5]}, str(tmp_path / "y"))
    assert all(os.path.getsize(f) > 100 for f in out)
