#!/usr/bin/env python
"""
Golden fixtures for the drop-in surface (CLI, features .csv loader, *_reads.csv / compiled.csv /
compiled_stats.csv): runs the UNMODIFIED reference CLI (`python -m fast2q -c ...` from /root/reference, with the
colorama / matplotlib stubs of tests/golden/_stubs) on small input folders and stores inputs + outputs in
tests/golden/cli_cases.json.gz.  Build container only:

    python tests/golden/make_cli_golden.py

Also stores the reference's features_loader() result for a few awkward library files and the reference's
input_parser() dict for a few command lines.
"""
import base64
import glob
import gzip
import io
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import cases  # noqa: E402
import refharness  # noqa: E402
from oracle import synth  # noqa: E402


def gz(data: bytes, level=6) -> bytes:
    b = io.BytesIO()
    with gzip.GzipFile(fileobj=b, mode="wb", compresslevel=level, mtime=0) as f:
        f.write(data)
    return b.getvalue()


def cli_cases():
    out = []
    # 1. Counter, fixed position, three files (.fastq, .fastq.gz, CRLF .fastq), numeric names, awkward library file
    names, keys = synth.make_library(2, 150, 20)
    spec = synth.default_spec(2)
    lib_lines = []
    for i, k in enumerate(keys):
        s = k.decode()
        if i == 7:
            s = s.lower()
        if i == 9:
            s = s[:10] + " " + s[10:]
        lib_lines.append(f"{i + 1},{s}")
    lib_lines.insert(20, f"999,{keys[3].decode()}")            # shares its sequence with entry 4: ignored with a warning
    # (a repeated NAME makes the reference itself crash in run_stats, fast2q.py:1475, so none is used here)
    reads = synth.fixed_reads(keys, 0, 5400, **spec).tobytes()
    rec = 118
    c = reads[5000 * rec:5400 * rec].replace(b"\n", b"\r\n")
    out.append(dict(name="cli_counter", library="\n".join(lib_lines) + "\n",
                    files={"sampleA.fastq": reads[:3000 * rec], "sampleB.fastq.gz": gz(reads[3000 * rec:5000 * rec]),
                           "sampleC.fastq": c},
                    args=["--m", "1", "--ph", "30", "--k", "--pb", "--cp", "2"]))
    # 2. Extract + Count between delimiters with mismatches, two files
    bs = synth.barseq_reads(4, 3000)
    cut = bs.find(b"\n@", len(bs) // 2) + 1
    out.append(dict(name="cli_ec", library=None, files={"bar1.fastq": bs[:cut], "bar2.fastq.gz": gz(bs[cut:])},
                    args=["--mo", "EC", "--us", "GTTCAGAGTTCT", "--ds", "CTGAATAGGCCA", "--msu", "1", "--msd", "1", "--pb", "--k",
                          "--cp", "2"]))
    # 3. dual fixed windows, ';' separated library, alphabetical names, custom file name, intermediates deleted
    dn, dk, xs, ys = synth.dual_library(5, 120)
    dr = synth.dual_reads(5, 2400, xs, ys, mode="fixed")
    cut = dr.find(b"\n@", len(dr) // 3) + 1
    out.append(dict(name="cli_dual", library="".join(f"{n};{k.decode()}\n" for n, k in zip(dn, dk)),
                    files={"d_one.fastq": dr[:cut], "d_two.fastq": dr[cut:]},
                    args=["--st", "0,30", "--l", "20", "--m", "1", "--pb", "--cp", "2", "--fn", "mycounts"]))
    # 4. a single file: File-Split mode is forced (fast2q.py:1671-1672); --cp 1 keeps the reference's chunking exact
    out.append(dict(name="cli_single_split", library="".join(f"{n},{k.decode()}\n" for n, k in zip(names, keys)),
                    files={"only.fastq.gz": gz(reads[:2000 * rec])}, args=["--m", "2", "--pb", "--k", "--cp", "1"]))
    # 5. a gzip file cut off mid-stream: warning + partial counts (fast2q.py:405-407)
    whole = gz(reads[:2500 * rec])
    out.append(dict(name="cli_truncated_gz", library="".join(f"{n},{k.decode()}\n" for n, k in zip(names, keys)),
                    files={"good.fastq": reads[2500 * rec:3500 * rec], "cut.fastq.gz": whole[: len(whole) * 6 // 10]},
                    args=["--pb", "--k", "--cp", "2"]))
    # 6. delimiter pairs in Counter mode, tab separated library
    n5, k5, xs5, ys5 = synth.dual_library(5, 100)
    d5 = synth.dual_reads(5, 1500, xs5, ys5, mode="delim")
    cut = d5.find(b"\n@", len(d5) // 2) + 1
    out.append(dict(name="cli_dual_delim", library="".join(f"{n}\t{k.decode()}\n" for n, k in zip(n5, k5)),
                    files={"p1.fastq": d5[:cut], "p2.fastq": d5[cut:]},
                    args=["--us", "ACCGGT,GGATCC", "--ds", "TTGACA,CAATTG", "--m", "1", "--pb", "--k", "--cp", "2"]))
    return out


def run_reference_cli(case):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(HERE, "_stubs"), refharness.REF_ROOT]), PYTHONWARNINGS="ignore")
    with tempfile.TemporaryDirectory() as td:
        src, outd = os.path.join(td, "in"), os.path.join(td, "out")
        os.makedirs(src); os.makedirs(outd)
        for fn, data in case["files"].items():
            open(os.path.join(src, fn), "wb").write(data)
        args = ["-c", "--s", src, "--o", outd] + case["args"]
        if case["library"] is not None:
            lp = os.path.join(td, "library.csv")
            open(lp, "w").write(case["library"])
            args += ["--g", lp]
        p = subprocess.run([sys.executable, "-m", "fast2q"] + args, env=env, cwd=td, capture_output=True, text=True, timeout=900)
        assert p.returncode == 0, p.stderr[-2000:]
        dirs = glob.glob(os.path.join(outd, "2FAST2Q_output_*"))
        assert len(dirs) == 1
        outs = {}
        for f in sorted(os.listdir(dirs[0])):
            if f.endswith(".csv"):
                outs[f] = open(os.path.join(dirs[0], f), newline="").read().replace(td, "<TMP>")
        return outs, p.stdout.replace(td, "<TMP>")


def loader_cases():
    """library files that exercise features_loader's quirks (fast2q.py:125-186)"""
    return {
        "comma": "a,ACGT\nb,acgt t\nc,GGGG\nb,TTTT\n",
        "semicolon": "x;AAAA\ny;CCCC\n",
        "tab": "x\tAAAA\ny\tCCCC\n",
        "comma_then_short_line": "a,ACGT\nb,CCCC\nbroken\nc,GGGG\n",
        "mixed_separators": "a,ACGT\nb;CCCC\nc\tGGGG\n",
        "three_columns": "a,ACGT,extra\nb,CCCC,more\n",
        "semicolon_with_commas_in_name": "n,1;ACGT\nm,2;CCCC\n",
        "crlf": "a,ACGT\r\nb,CCCC\r\n",
        "blank_last_line": "a,ACGT\nb,CCCC\n\n",
        "header_like": "name,sequence\ng1,ACGTAC\n",
        "trailing_space_after_seq": "a,ACGT \nb, CCCC\n",
    }


def parser_cases():
    return [
        ["-c"],
        ["-c", "--s", "/data/in", "--g", "/data/lib.csv", "--o", "/data/out"],
        ["-c", "--s", "/d", "--g", "/l.csv", "--o", "/o", "--m", "2", "--ph", "25", "--st", "3,40", "--l", "18", "--pb", "--k", "--fs",
         "--cp", "3", "--fn", "xyz"],
        ["-c", "--s", "/d", "--o", "/o", "--mo", "ec", "--us", "acgt", "--ds", "TTGA", "--msu", "1", "--msd", "2", "--qsu", "10",
         "--qsd", "0"],
        ["-c", "--s", "/d", "--g", "/l.csv", "--o", "/o", "--mo", "Counter", "--fn"],
    ]


def main():
    assert refharness.available()
    ref = refharness.load()
    out = dict(cli=[], loader={}, parser=[])
    for c in cli_cases():
        outs, stdout = run_reference_cli(c)
        out["cli"].append(dict(name=c["name"], args=c["args"], library=c["library"],
                               files={k: base64.b64encode(v).decode() for k, v in c["files"].items()}, outputs=outs))
        print(c["name"], {k: len(v) for k, v in outs.items()})
    for name, text in loader_cases().items():
        with tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False) as f:
            f.write(text)
        try:
            feats = ref.features_loader(f.name)
            res = [[seq, v.name] for seq, v in feats.items()]
        except SystemExit:
            res = "FATAL"
        os.unlink(f.name)
        out["loader"][name] = dict(text=text, expect=res)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.chdir(td)                                     # an empty cwd: no stray .csv for the --g default
        for argv in parser_cases():
            old = sys.argv
            sys.argv = ["2fast2q"] + argv
            try:
                p = ref.input_parser()
            finally:
                sys.argv = old
            out["parser"].append(dict(argv=argv, expect={k: (v.replace(td, "<CWD>") if isinstance(v, str) else v) for k, v in p.items()}))
        os.chdir(cwd)
    with gzip.GzipFile(os.path.join(HERE, "cli_cases.json.gz"), "wb", mtime=0) as f:
        f.write(json.dumps(out).encode())
    print("written", os.path.getsize(os.path.join(HERE, "cli_cases.json.gz")))


if __name__ == "__main__":
    main()
