#!/usr/bin/env python
"""developer micro-bench: resident config-2 pass, prints the tile-kernel time; no correctness asserts (used with debug switches)"""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
synth = importlib.import_module("2fast2q_b200.synth")
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
opts = dict(kv.split("=") for kv in sys.argv[2:])
cfgno = int(opts.pop("config", 2))
spec = synth.default_spec(cfgno)
nguides = int(opts.pop("guides", 2000))
miss = int(opts.pop("miss", 1))
names, keys = synth.make_library(cfgno, nguides, 20)
rec = 2 * spec["read_len"] + 18
cfg = lib.make_config(mode="C", miss=miss, phred=30, length=20, start="0")
with lib.Engine(cfg, 0, None, time_kernels=1, **{k: int(v) for k, v in opts.items()}) as e:
    e.set_library(keys)
    d = e.device_alloc(reads * rec)
    e.synth(d, keys, 0, reads, **spec)
    for it in range(5):
        e.begin(); e.submit_device(d, reads * rec, True); c, s = e.end()
        kt = e.kernel_times()
    print("stats", s)
    print("tile ms %.3f  resolve ms %.3f  aux ms %.3f  spec %s -> %.1f GB/s, %.2f G reads/s (tile only)" % (kt["tile"][0], kt["resolve"][0], kt["aux"][0], e.spec_counts(), reads * rec / kt["tile"][0] / 1e6, reads / kt["tile"][0] / 1e6))
