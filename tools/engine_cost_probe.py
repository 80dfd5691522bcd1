#!/usr/bin/env python
"""developer probe: what creating a context costs (create / set_library / first file), alone and from 8 threads at once"""
import importlib, os, sys, tempfile, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
synth = importlib.import_module("2fast2q_b200.synth")
names, keys, xs, ys = synth.dual_keys(10_000)
spec = synth.shape_spec("5a")
d = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
path = os.path.join(d, "x.fastq")
synth.shaped_reads(xs + ys, 0, 200_000, **spec).tofile(path)
cfg = lib.make_config(mode="C", miss=1, phred=30, length=20, start="0,30")


def one(tag, out):
    t0 = time.perf_counter(); e = lib.Engine(cfg, 0, memo_entries=1 << 20)
    t1 = time.perf_counter(); e.set_library(keys)
    t2 = time.perf_counter(); e.begin(); e.submit_file(path, False, 0, 8); e.end()
    t3 = time.perf_counter(); e.begin(); e.submit_file(path, False, 0, 8); e.end()
    t4 = time.perf_counter(); e.close()
    t5 = time.perf_counter()
    out.append(f"{tag}: create {t1 - t0:.3f}  set_library {t2 - t1:.3f}  first file {t3 - t2:.3f}  second file {t4 - t3:.3f}  close {t5 - t4:.3f} s")


res = []
one("warm-up (CUDA context)", res); one("alone", res)
print("\n".join(res), flush=True)
for n in (4, 16):
    res = []; t0 = time.perf_counter()
    th = [threading.Thread(target=one, args=(f"{n} threads #{k}", res)) for k in range(n)]
    [t.start() for t in th]; [t.join() for t in th]
    print(f"-- {n} threads at once: {time.perf_counter() - t0:.3f} s wall"); print("\n".join(sorted(res)[:3]), flush=True)
os.remove(path)
