"""
-m gpu: the whole drop-in surface end to end on a B200 — `2fast2q -c ...` arguments in, compiled.csv /
compiled_stats.csv / *_reads.csv out — against the bytes the UNMODIFIED reference CLI wrote for the same input
folders (tests/golden/cli_cases.json.gz, made by tests/golden/make_cli_golden.py).  Wall-clock fields are masked.
"""
import glob
import importlib
import os

import pytest

import cli_golden as CG
import golden_io as G

pytestmark = pytest.mark.gpu

fq = importlib.import_module("2fast2q_b200.fast2q")


def run_cli(c, root, extra=()):
    argv = CG.write_inputs(c, str(root)) + list(extra)
    fq.main(argv)
    dirs = glob.glob(os.path.join(str(root), "out", "2FAST2Q_output_*"))
    assert len(dirs) == 1
    # (the golden runs of the reference drew on a stubbed matplotlib: the four .png files are not part of the comparison)
    return {f: open(os.path.join(dirs[0], f), newline="").read() for f in sorted(os.listdir(dirs[0])) if not f.endswith(".png")}


def check_outputs(c, got):
    want = c["outputs"]
    assert sorted(got) == sorted(want)                       # same file set (intermediates kept only with --k)
    for fn, text in want.items():
        if fn.endswith("_reads.csv"):
            assert CG.mask_reads_csv(got[fn]) == CG.mask_reads_csv(text), fn
        elif fn.endswith("_stats.csv"):
            assert CG.mask_stats_csv(got[fn]) == CG.mask_stats_csv(text), fn
        else:
            assert got[fn] == text, fn                       # compiled.csv: byte for byte


@pytest.mark.parametrize("name", [c["name"] for c in CG.load()["cli"]])
def test_cli_matches_reference_outputs(name, tmp_path):
    c = CG.case(name)
    check_outputs(c, run_cli(c, tmp_path))


@pytest.mark.parametrize("name", ["cli_counter", "cli_ec", "cli_truncated_gz"])
@pytest.mark.parametrize("native", [True, False])
def test_cli_file_split_mode_same_counts(name, native, tmp_path, monkeypatch):
    """--fs: every file is cut into record-aligned shards (one stream per GPU) — or, on one GPU, streamed through the native
    ingest; counts must not change"""
    monkeypatch.setattr(fq, "SPLIT_NATIVE_SINGLE_GPU", native)
    c = CG.case(name)
    got = run_cli(c, tmp_path, extra=["--fs"])
    want = dict(c["outputs"])
    for fn in want:
        if fn.endswith("_reads.csv"):
            assert CG.mask_reads_csv(got[fn]) == CG.mask_reads_csv(want[fn]), fn
    assert got["compiled.csv"] == want["compiled.csv"]


def test_small_shards_over_streams(tmp_path, monkeypatch):
    """force many tiny shards and chunks through the split path"""
    monkeypatch.setattr(fq, "CHUNK_BYTES", 50_000)
    monkeypatch.setattr(fq, "SPLIT_NATIVE_SINGLE_GPU", False)
    c = CG.case("cli_single_split")
    check_outputs(c, run_cli(c, tmp_path))


def test_reads_counter_boundary(tmp_path):
    """reads_counter(i, raw, features, param, reads_stats) -> (features, reads_stats, local_read_stats), fast2q.py:514"""
    c = [x for x in G.kat() if x["name"] == "exact_and_1mm"][0]
    raw = tmp_path / "x.fastq"
    raw.write_bytes(c["fastq"])
    feats = {seq: fq.Features(name, 0) for name, seq in c["library"]}
    p = fq.input_parser(["-c", "--s", str(tmp_path), "--g", "x", "--o", str(tmp_path), "--pb"])
    p = fq.initializer(p)
    rs = {"failed_reads": set(), "passed_reads": {}}
    out, rs2, stats = fq.reads_counter(0, str(raw), feats, p, rs)
    assert rs2 is rs and stats == c["stats"]
    assert [f.counts for f in out.values()] == c["counts"] and list(out) == list(feats)
    assert all(f.counts == 0 for f in feats.values())        # the caller's library dict is not accumulated into
    # preprocess=True: only the first 10 000 reads (fast2q.py:398-400)
    big = tmp_path / "big.fastq"
    big.write_bytes(c["fastq"] * 4000)
    _, _, st = fq.reads_counter(0, str(big), feats, p, rs, True)
    assert st["reads"] == 10000
    _, _, st = fq.reads_counter(0, str(big), feats, p, rs)
    assert st["reads"] == 12000
    fq.release_engines()


def test_unpaired_delimiters_are_fatal(tmp_path):
    p = fq.initializer(fq.input_parser(["-c", "--s", str(tmp_path), "--o", str(tmp_path), "--mo", "EC", "--us", "AC,GT", "--ds", "TT"]))
    (tmp_path / "x.fastq").write_bytes(b"@r\nACGT\n+\nIIII\n")
    with pytest.raises(SystemExit):
        fq.reads_counter(0, str(tmp_path / "x.fastq"), {}, p, {})


def test_public_helpers_known_answers():
    """tests/test_mainfunctions.py of the reference, through the same public names"""
    assert fq.border_finder(fq.seq2bin("GATTACA"), fq.seq2bin("TACTGATTACAGCAC"), 1) == 4
    read, qual = fq.seq2bin("TACTGATTACAGCAC"), b"AAII$%&#III/(&/"
    info = dict(upstream="TACT", downstream="GCAC", upstream_bin=[fq.seq2bin("TACT")], downstream_bin=[fq.seq2bin("GCAC")],
                miss_search_up=1, miss_search_down=1, quality_set_up=set(""), quality_set_down=set(""))
    assert fq.sequence_tinder(read, qual, info) == (4, 11)
    info["quality_set_down"] = set("/")
    assert fq.sequence_tinder(read, qual, info) == (None, None)
    info["miss_search_down"] = 2
    assert fq.sequence_tinder(read, qual, info) == (4, 6)


def test_test_mode_equals_reference_compiled_csv(tmp_path, monkeypatch):
    """`2fast2q -c -t` (BASELINE.json configs[0]; fast2q.py:1237-1240, reference tests/test_cli.py:5-28): the bundled
    example (a labelled surrogate of the missing example.fastq.gz, see 2fast2q_b200/testdata.py) and D39V_guides.csv give
    a compiled.csv whose counts equal the reference's own tests/compiled.csv."""
    monkeypatch.chdir(tmp_path)
    fq.main(["-c", "-t"])
    dirs = glob.glob(os.path.join(str(tmp_path), "2FAST2Q_output_*"))
    assert len(dirs) == 1
    files = sorted(os.listdir(dirs[0]))
    assert "compiled.csv" in files and "compiled_stats.csv" in files
    assert len(files) == 6 and sum(f.endswith(".png") for f in files) == 4      # the folder shape the reference's test_cli.py:17-25 pins
    got = open(os.path.join(dirs[0], "compiled.csv"), newline="").read().split("\r\n")
    want = open(os.path.join(G.HERE, "ref_compiled.csv"), newline="").read().splitlines()
    assert got[0] == "#Feature,example"
    assert [r for r in got[1:] if r] == [w.strip() for w in want[1:] if w.strip()]
