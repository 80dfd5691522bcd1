#!/usr/bin/env python
"""developer probe (torchrun, N ranks): ms per resident pass with blocking / non-blocking sample ends, with and without the all-reduce"""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
f2q = importlib.import_module("2fast2q_b200"); lib = f2q._lib
synth = importlib.import_module("2fast2q_b200.synth")
rank, lr, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1: dist.init_process_group("nccl", device_id=dev)
reads = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
spec = synth.default_spec(2); names, keys = synth.make_library(2, 2000, 20); rec = 118
stream = torch.cuda.Stream(device=dev)
eng = lib.Engine(lib.make_config(miss=1), lr, stream.cuda_stream)
eng.set_library(keys)
data = torch.empty(reads * rec, dtype=torch.uint8, device=dev)
eng.synth(data.data_ptr(), keys, rank * reads, reads, **spec)
class A:
    def __init__(s, p, n): s.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (p, False), "version": 2}
rp, rw = eng.result_device(); rt = torch.as_tensor(A(rp, rw), device=dev)
bufs = [lib.PinnedBuffer(8 * (len(keys) + 6)) for _ in range(8)]
def run(mode, K=6):
    def one(k):
        eng.begin(); eng.submit_device(data.data_ptr(), reads * rec, True)
        if "ar" in mode and world > 1:
            with torch.cuda.stream(stream): dist.all_reduce(rt)
        if "async" in mode: eng.end_async(bufs[k % 8])
        else: eng.end()
    for k in range(3): one(k)
    eng.sync()
    if world > 1: dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record(stream)
    for k in range(K): one(k)
    e1.record(stream)
    if world > 1: dist.barrier()
    torch.cuda.synchronize(dev); eng.sync()
    ms = torch.tensor([e0.elapsed_time(e1) / K], device=dev, dtype=torch.float64)
    if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0: print(f"{mode:14s} {ms.item():7.3f} ms/pass (max over {world} ranks)   host wall {1e3 * (time.perf_counter() - t0) / K:7.3f}", flush=True)
for mode in ("sync", "sync+ar", "async", "async+ar", "sync+ar", "async+ar"):
    run(mode)
eng.close()
if world > 1: dist.destroy_process_group()
