#!/usr/bin/env python
"""developer probe: resident throughput of the generic-policy configurations (BASELINE configs 4 and 5) on repeated shaped inputs"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests", "golden")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import cases
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
only = sys.argv[2].split(",") if len(sys.argv) > 2 else ("config4_barseq", "config5a_dual_fixed", "config5b_dual_delim", "config3_slice")
opts = {k: int(v) for k, v in (kv.split("=") for kv in sys.argv[3:])}
for name in only:
    params, library, data = cases.shaped_inputs(name)
    n_reads = data.count(b"\n") // 4
    blob = np.frombuffer(data * reps, dtype=np.uint8)
    cfg = lib.make_config(**params)
    with lib.Engine(cfg, 0, None, time_kernels=1, **opts) as e:
        if library is not None:
            e.set_library([s for _, s in library])
        d = e.device_alloc(blob.size)
        e.h2d(d, blob)
        best = 1e9
        for it in range(4):
            e.begin(); t0 = time.perf_counter(); e.submit_device(d, blob.size, True); c, s = e.end(); dt = time.perf_counter() - t0
            best = min(best, dt)
        kt = e.kernel_times()
        print(f"{name:22s} reads {n_reads * reps:9d}  bytes/read {blob.size / (n_reads * reps):6.1f}  best pass {best * 1e3:8.2f} ms  "
              f"{n_reads * reps / best / 1e6:8.1f} M reads/s  {blob.size / best / 1e9:7.1f} GB/s  kernels {dict((k, round(v[0], 2)) for k, v in kt.items())}  spec {e.spec_counts()}  reads {s['reads']}")
        e.device_free(d)
