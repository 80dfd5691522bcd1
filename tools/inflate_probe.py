#!/usr/bin/env python
"""developer probe: bgzip file -> counts through f2q_submit_file with the inflate on the device (k_inflate_bgzf) and on host
threads; prints uncompressed GB/s end to end (file in the page cache)"""
import importlib, os, struct, sys, tempfile, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
synth = importlib.import_module("2fast2q_b200.synth")
sys.path.insert(0, os.path.join(ROOT, "tools"))
from cli_ingest_bench import bgzf_bytes
from concurrent.futures import ProcessPoolExecutor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
level = int(sys.argv[2]) if len(sys.argv) > 2 else 1        # zlib level of the blocks (bgzip's default is 6)
spec = synth.default_spec(2)
names, keys = synth.make_library(2, 2000, 20)


def part(k):
    return bgzf_bytes(synth.fixed_reads(keys, k * 500_000, 500_000, **spec).tobytes(), level)[:-28]


if __name__ == "__main__":
    t0 = time.perf_counter()
    with ProcessPoolExecutor(os.cpu_count()) as ex:
        parts = list(ex.map(part, range(n // 500_000)))
    d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    path = os.path.join(d, "x.fastq.gz")
    with open(path, "wb") as f:
        for p in parts:
            f.write(p)
        f.write(bgzf_bytes(b""))
    comp = os.path.getsize(path)
    unc = n * 118
    print(f"file: {comp / 1e9:.2f} GB compressed, {unc / 1e9:.2f} GB uncompressed (ratio {unc / comp:.2f}, zlib level {level}), written in {time.perf_counter() - t0:.0f} s", flush=True)
    for opt, bits, threads in ((1, 8, 16), (1, 9, 16), (1, 0, 16), (0, 0, 16)):
        with lib.Engine(lib.make_config(miss=1), 0, None, gpu_inflate=opt, gpu_inflate_bits=bits) as e:
            e.set_library(keys)
            for it in range(2):
                e.begin(); t = time.perf_counter(); ok, nb = e.submit_file(path, True, 0, threads); c, s = e.end(); dt = time.perf_counter() - t
            print(f"gpu_inflate={opt} bits={bits} threads={threads}: {dt * 1e3:8.1f} ms  {unc / dt / 1e9:6.2f} GB/s uncompressed  {n / dt / 1e6:7.1f} M reads/s  reads {s['reads']}", flush=True)
    os.remove(path)
