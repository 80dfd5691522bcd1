#!/usr/bin/env python
"""developer probe: cost of f2q_ec_merge (one-rank communicator: the whole code path without a second GPU) and of the
Extract+Count sample end, 25 M Bar-seq reads, 1 M-barcode pool"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
synth = importlib.import_module("2fast2q_b200.synth")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 25_000_000
spec = synth.shape_spec("4")
guides = synth.random_kmers(4, 1_000_000, 20)
params = dict(mode="EC", upstream=synth.BARSEQ_US.decode(), downstream=synth.BARSEQ_DS.decode(), miss_search_up=1, miss_search_down=1)
with lib.Engine(lib.make_config(**params), 0, None, time_kernels=1) as e:
    lib.comm_init([e])
    d = e.device_alloc(n * 168)
    e.synth(d, guides, 0, n, **spec)
    for it in range(3):
        t0 = time.perf_counter(); e.begin(); t1 = time.perf_counter()
        e.submit_device(d, n * 168, True); t2 = time.perf_counter()
        e.allreduce_counts(); c, s = e.end(); t3 = time.perf_counter()
        e.ec_merge(); e.sync(); t4 = time.perf_counter()
        items = e.ec_items(); t5 = time.perf_counter()
        print(f"begin {1e3*(t1-t0):.2f} submit(enqueue) {1e3*(t2-t1):.2f} allreduce+end {1e3*(t3-t2):.2f} ec_merge {1e3*(t4-t3):.2f} drain {1e3*(t5-t4):.1f} ms  keys {len(items)} kernels {e.kernel_times()}", flush=True)
    e.device_free(d)
