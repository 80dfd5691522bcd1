#!/usr/bin/env python
"""
File -> compiled.csv throughput of the drop-in command line (`python -m 2fast2q_b200 -c ...`) on the GPU box: what a user of
the reference sees, file reads / gzip inflate included (task 8).

    python tools/cli_ingest_bench.py --config 5a --files 48 --reads 1500000 --gz gzip|bgzf|none [--gpus N]

Writes the synthetic sample files (K0 generator's numpy restatement, one process per file), runs the CLI in this process,
checks the first sample's column of compiled.csv against the oracle, prints one JSON line.
"""
import argparse, csv, glob, gzip, importlib, json, os, shutil, struct, sys, tempfile, time, zlib
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bgzf_bytes(data, level=1, bs=65280):
    out = []
    for o in list(range(0, len(data), bs)) + [None]:
        chunk = b"" if o is None else data[o:o + bs]
        c = zlib.compressobj(level, zlib.DEFLATED, -15)
        cd = c.compress(chunk) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", 18 + len(cd) + 8 - 1) + cd
                   + struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
    return b"".join(out)


def workload(config):
    synth = importlib.import_module("2fast2q_b200.synth")
    if config == "2":
        names, keys = synth.make_library(2, 2000, 20)
        spec = synth.default_spec(2)
        return names, keys, keys, spec, ["--st", "0", "--l", "20", "--m", "1", "--ph", "30"], "fixed"
    names, keys, xs, ys = synth.dual_keys(10_000)
    spec = synth.shape_spec(config)
    flags = ["--st", "0,30", "--l", "20"] if config == "5a" else ["--us", "ACCGGT,GGATCC", "--ds", "TTGACA,CAATTG"]
    return names, keys, xs + ys, spec, flags + ["--m", "1", "--ph", "30"], "shaped"


def write_file(job):
    config, k, reads, gz, path = job
    synth = importlib.import_module("2fast2q_b200.synth")
    names, keys, guides, spec, flags, kind = workload(config)
    gen = synth.fixed_reads if kind == "fixed" else synth.shaped_reads
    with open(path, "wb") as f:
        if gz == "none":
            for o in range(0, reads, 500_000):
                gen(guides, k * reads + o, min(500_000, reads - o), **spec).tofile(f)
        else:
            blob = b"".join(gen(guides, k * reads + o, min(500_000, reads - o), **spec).tobytes() for o in range(0, reads, 500_000))
            f.write(bgzf_bytes(blob) if gz == "bgzf" else gzip.compress(blob, compresslevel=1))
    return os.path.getsize(path)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="5a")
    ap.add_argument("--files", type=int, default=48)
    ap.add_argument("--reads", type=int, default=1_500_000)
    ap.add_argument("--gz", default="gzip", choices=["none", "gzip", "bgzf"])
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--cp", type=int, default=0)
    ap.add_argument("--dir", default=None)
    a = ap.parse_args()
    names, keys, guides, spec, flags, kind = workload(a.config)
    root = a.dir or tempfile.mkdtemp(prefix="f2q_cli_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    src, out = os.path.join(root, "in"), os.path.join(root, "out")
    os.makedirs(src, exist_ok=True); os.makedirs(out, exist_ok=True)
    ext = ".fastq" if a.gz == "none" else ".fastq.gz"
    jobs = [(a.config, k, a.reads, a.gz, os.path.join(src, "s%03d%s" % (k, ext))) for k in range(a.files)]
    t0 = time.perf_counter()
    with ProcessPoolExecutor(min(os.cpu_count() or 1, a.files)) as ex:
        sizes = list(ex.map(write_file, jobs))
    t_gen = time.perf_counter() - t0
    lib_csv = os.path.join(root, "lib.csv")
    with open(lib_csv, "w") as f:
        f.write("".join(f"{n},{k.decode()}\n" for n, k in zip(names, keys)))
    fq = importlib.import_module("2fast2q_b200.fast2q")
    argv = ["-c", "--s", src, "--g", lib_csv, "--o", out, "--pb"] + flags
    if a.gpus:
        argv += ["--gpus", str(a.gpus)]
    if a.cp:
        argv += ["--cp", str(a.cp)]
    t0 = time.perf_counter()
    fq.main(argv)
    wall = time.perf_counter() - t0
    # parity of the first sample against the oracle
    from oracle import oracle as O
    synth = importlib.import_module("2fast2q_b200.synth")
    gen = synth.fixed_reads if kind == "fixed" else synth.shaped_reads
    n_chk = min(a.reads, 300_000)
    comp = glob.glob(os.path.join(out, "2FAST2Q_output_*", "compiled.csv"))[0]
    rows = list(csv.reader(open(comp, newline="")))
    col = rows[0].index("s000")
    got_total = sum(int(r[col]) for r in rows[1:])
    params = dict(mode="C", miss=1, phred=30, length=20, start="0")
    if a.config == "5a":
        params["start"] = "0,30"
    if a.config == "5b":
        params.update(upstream="ACCGGT,GGATCC", downstream="TTGACA,CAATTG")
    parity = None
    if n_chk == a.reads:
        want_c, want_s = O.count(O.make_config(**params), keys, gen(guides, 0, a.reads, **spec))
        by_name = {r[0]: int(r[col]) for r in rows[1:]}
        parity = all(by_name[n] == int(c) for n, c in zip(names, want_c))
        assert parity, "compiled.csv column s000 differs from the oracle"
    total_reads = a.files * a.reads
    unc = total_reads * (2 * spec["read_len"] + 18)
    print(json.dumps({"tool": "cli_ingest_bench", "config": a.config, "files": a.files, "reads_per_file": a.reads, "compression": a.gz,
                      "gpus": a.gpus or "all", "cli_wall_s": round(wall, 3), "M_reads_per_s": round(total_reads / wall / 1e6, 2),
                      "uncompressed_GB_per_s": round(unc / wall / 1e9, 3), "compressed_GB": round(sum(sizes) / 1e9, 3),
                      "host_threads_inflate": fq._INFLATE_WORKERS, "cpus": os.cpu_count(), "aligned_reads_s000": got_total,
                      "parity_first_sample_vs_oracle": parity, "generation_s": round(t_gen, 1)}), flush=True)
    if not a.dir:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    main()
