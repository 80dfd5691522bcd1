#!/usr/bin/env python
"""
bench.py — headline benchmark of the read -> feature -> count path (BASELINE.json configs[1]):
synthetic 100 M x 50 bp reads vs a 2 000-guide library, --st 0 --l 20 --m 1 --ph 30.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--reads R]

One "step" = one pass of the hot path over the whole synthetic sample (R reads per GPU, 2L+18 = 118 bytes each).
  value   M reads/s with the sample resident in HBM (f2q_submit_device); device-timed, max over ranks
  e2e     the same through the C-ABI call a user makes with HOST buffers (f2q_submit from pinned memory):
          H2D of every byte + D2H of the counts inside the timed region
  roofline achieved HBM GB/s of the fused streaming kernel k_spec = algorithmic bytes (118 B/read) / its CUDA-event time
  cpu_baseline  the oracle port (oracle/f2q_oracle.c, the reference's algorithm in C) on a bounded sample, 1 core
--impl reference times that port on all host threads (file-parallel, like the reference's multiprocessing mode).
Data are synthetic (K0 generator, bit-identical to 2fast2q_b200/synth.py); weak scaling: every rank owns its own R reads.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

_JSON_OUT = None


def emit(obj):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(obj) + "\n")
    out.flush()


ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

READ_LEN, FEAT_LEN, N_GUIDES, CONFIG = 50, 20, 2000, 2
REC = 2 * READ_LEN + 18


def env_int(k, d):
    try:
        return int(os.environ.get(k, d))
    except ValueError:
        return d


class ClockSampler:
    """samples SM clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe) with an `nvidia-smi`
    subprocess polling every 20 ms.  Deliberately NOT in-process NVML: measured on this pool, one NVML query from a thread
    of the benchmark process blocks its CUDA launches for ~15 ms, three times the length of a whole resident pass."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index, uuid=None):
        self.gpu, self.uuid, self.rows, self.p = gpu_index, uuid, [], None

    def start(self):
        sel = f"GPU-{self.uuid}" if self.uuid and not str(self.uuid).startswith("GPU-") else (self.uuid or str(self.gpu))
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                       "-i", str(sel)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.5)                                        # let it come up before the timed region starts
        except OSError:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([time.perf_counter()] + [x.strip() for x in line.split(",")])

    def stop(self, windows=()):
        """windows: (t0, t1) perf_counter intervals of the timed regions; samples inside them are counted separately"""
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.p.kill()
        sm, mx, reasons, inside = [], [], set(), [0] * len(windows)
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[2])); mx.append(float(r[3]))
            except ValueError:
                continue
            for k, (a, b) in enumerate(windows):
                if a <= r[0] <= b + 0.02:
                    inside[k] += 1
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_in_timed_regions": inside,
                "source": "nvidia-smi -lms 20 from just before the resident timed region to the end of the end-to-end timed region"}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def committed_traffic():
    """dram bytes per read of the dominant kernel from the committed ncu --set full capture (profiles/), scaled to this launch"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "tile_kernel_traffic.json")))
        return t
    except Exception:
        return None


def library():
    synth = importlib.import_module("2fast2q_b200.synth")        # generator spec + library builder (not the oracle)
    return synth.make_library(CONFIG, N_GUIDES, FEAT_LEN), synth.default_spec(CONFIG)


# ------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """the reference's CPU algorithm (oracle port; the Python reference cannot travel to the GPU box) on all host threads"""
    if rank != 0:
        return
    from oracle import oracle as O                               # the reference arm IS the oracle port (task statement ④)
    synth = importlib.import_module("2fast2q_b200.synth")
    (names, keys), spec = library()
    cores = os.cpu_count() or 1
    threads = max(1, min(cores, 64))
    per = args.ref_reads_per_thread
    cfg = O.make_config(miss=1, phred=30, length=FEAT_LEN, start="0")
    O.lib()
    shards = [synth.fixed_reads(keys, t * per, per, **spec) for t in range(threads)]      # distinct reads per thread
    from concurrent.futures import ThreadPoolExecutor

    def work(a):
        return O.count(cfg, keys, a)

    def step():
        with ThreadPoolExecutor(threads) as ex:
            res = list(ex.map(work, shards))                       # ctypes releases the GIL: real thread parallelism
        tot = np.zeros(len(keys), dtype=np.uint64)
        for c, _ in res:
            tot += c                                               # merge_feature_dicts, fast2q.py:439-445
        return tot

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = threads * per / dt / 1e6
    sample = f"{threads} threads x {per} reads of the config-2 stream per step (one shard per thread, counts merged by addition)"
    emit({
        "impl": "reference", "metric": "M reads/s", "value": v, "unit": "M reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": "config2: 50bp reads vs 2000-guide library, --st 0 --l 20 --m 1 --ph 30", "reads_per_step": threads * per},
        "cpu_baseline": {"value": v, "unit": "M reads/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "M reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


# ------------------------------------------------------------------------------------------------------------
class _DevArr:
    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i8", "data": (ptr, False), "version": 2}


def run_gpu(args, rank, local_rank, world):
    import torch
    import torch.distributed as dist
    f2q = importlib.import_module("2fast2q_b200")
    lib = f2q._lib
    multi = importlib.import_module("2fast2q_b200.multi")
    lib.load()                                                     # raises if the CUDA library is missing
    torch.cuda.set_device(local_rank)
    try:
        gpu_uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
    except Exception:
        gpu_uuid = None
    placement = numa_bind(gpu_uuid, local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    (names, keys), spec = library()
    n_reads = args.reads
    nbytes = n_reads * REC
    stream = torch.cuda.Stream(device=dev)
    cfg = lib.make_config(mode="C", miss=1, phred=30, length=FEAT_LEN, start="0")
    opts = {"time_kernels": 1}
    if args.tile_threads:
        opts["tile_threads"] = args.tile_threads
    eng = lib.Engine(cfg, local_rank, stream.cuda_stream, **opts)
    eng.set_library(keys)
    data = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    eng.synth(data.data_ptr(), keys, rank * n_reads, n_reads, **spec)      # every rank owns its own contiguous read range
    rptr, rwords = eng.result_device()
    result_t = torch.as_tensor(_DevArr(rptr, rwords), device=dev) if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def merge():
        if world > 1:
            with torch.cuda.stream(stream):
                multi.merge_results(result_t)        # ncclAllReduce(sum) of [counts | stats] over NVLink (merge_feature_dicts)

    def step_resident():
        eng.begin()
        eng.submit_device(data.data_ptr(), nbytes, True)
        merge()
        return eng.end()

    # ---- resident leg ----
    for _ in range(args.warmup):
        counts, stats = step_resident()
    assert stats["reads"] == n_reads * world, stats
    clocks = ClockSampler(local_rank, gpu_uuid)
    tile_ms = []
    if rank == 0 and not os.environ.get("F2Q_BENCH_NO_CLOCKS"):
        clocks.start()
    barrier()
    l0 = eng.launches
    # the K timed passes run back to back: each one ends with f2q_end_sample_async, i.e. its [counts | stats] vector is
    # copied into its own pinned host buffer, stream-ordered, and checked after the timed region (no host round trip
    # between passes; the end-to-end leg below does the blocking read every step)
    res_bufs = [lib.PinnedBuffer(8 * (len(keys) + 6)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w_res0 = time.perf_counter()
    e0.record(stream)
    for k in range(args.steps):
        eng.begin()
        eng.submit_device(data.data_ptr(), nbytes, True)
        merge()
        eng.end_async(res_bufs[k])
    e1.record(stream)
    barrier()
    eng.sync()
    windows = [(w_res0, time.perf_counter())]
    tile_ms.append(eng.kernel_times()["tile"])
    for b in res_bufs:
        c_k, s_k = eng.read_async_result(b)
        assert s_k == stats and np.array_equal(c_k, counts), "a timed pass returned different counts"
        b.free()
    launches = eng.launches - l0
    ms = e0.elapsed_time(e1) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_res = float(t.item())
    ktimes = eng.kernel_times()
    spec_counts = eng.spec_counts()

    # ---- end-to-end leg: host (pinned) buffers through f2q_submit ----
    e2e = None
    if not args.no_e2e:
        pin = lib.PinnedBuffer(nbytes)
        eng._ck(eng.L.f2q_memcpy_d2h(eng.h, pin.ptr, data.data_ptr(), nbytes))

        def step_e2e():
            eng.begin()
            eng.submit_ptr(pin.ptr.value, nbytes, True)
            merge()
            return eng.end()

        for _ in range(max(1, min(args.warmup, 3))):
            c2, s2 = step_e2e()
        assert s2 == stats and np.array_equal(c2, counts)
        # the ingest roofline: a plain pinned -> device copy of the same bytes on the same link (best of 2)
        pcie = 0.0
        for _ in range(2):
            torch.cuda.synchronize(dev)
            tp = time.perf_counter()
            eng.h2d(data.data_ptr(), pin.array)
            pcie = max(pcie, nbytes / (time.perf_counter() - tp) / 1e9)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        f0.record(stream)
        for _ in range(args.steps):
            step_e2e()
        f1.record(stream)
        barrier()
        windows.append((t0, time.perf_counter()))
        wall = (time.perf_counter() - t0) / args.steps * 1e3
        ms2 = max(f0.elapsed_time(f1) / args.steps, 0.0)
        t = torch.tensor([ms2, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms2, wall = float(t[0].item()), float(t[1].item())
        e2e = {"value": world * n_reads / (ms2 / 1e3) / 1e6, "unit": "M reads/s", "h2d_bytes_per_step": nbytes,
               "d2h_bytes_per_step": (len(keys) + 5) * 8, "ms_per_step": ms2, "wall_ms_per_step": wall,
               "h2d_gbs_per_gpu": nbytes / (ms2 / 1e3) / 1e9, "pcie_copy_gbs_measured": pcie,
               "pcie_frac": (nbytes / (ms2 / 1e3) / 1e9) / pcie if pcie else None}
        pin.free()
    clk = clocks.stop(windows) if rank == 0 else None

    if rank == 0:
        peak, peak_src = measured_peak()
        tm = [m for m, n in tile_ms if n]
        tile_avg = sum(tm) / max(1, sum(n for m, n in tile_ms if n))            # ms per tile-kernel launch
        achieved = n_reads * REC / (tile_avg / 1e3) / 1e9 if tile_avg > 0 else None
        traffic = committed_traffic()
        out = {
            "metric": "M reads/s", "value": world * n_reads / (ms_res / 1e3) / 1e6, "unit": "M reads/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_res, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "config2: synthetic 50bp reads vs 2000-guide library, --st 0 --l 20 --m 1 --ph 30",
                       "reads_per_gpu": n_reads, "bytes_per_read": REC, "input_bytes_per_gpu": nbytes,
                       "l2": "inputs (%.1f GB/GPU) are far larger than the 126 MB L2; no flush needed" % (nbytes / 1e9),
                       "parallelism": f"reads sharded by contiguous range over {world} GPU(s); counts merged by NCCL all-reduce",
                       "host_placement": placement},
            "clocks": clk,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         "traffic": traffic["dram_bytes_per_read"] * n_reads if traffic else None,
                         "kernel": "k_spec<POLICY_FAST1, CH=7, W=16> (speculative streaming kernel, csrc/spec.cuh)",
                         "kernel_ms_per_launch": tile_avg, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": n_reads * REC,
                         "other_kernels_ms_per_step": {"resolve": ktimes["resolve"][0] / args.steps, "generic": ktimes["generic"][0] / args.steps,
                                                       "verify_commit_fallback": ktimes["aux"][0] / args.steps},
                         "traffic_note": traffic.get("note") if traffic else "no committed ncu --set full capture yet"},
            "speculation": {"chunks_committed": spec_counts[0], "chunks_parsed_by_exact_kernel": spec_counts[1]},
            "stats": stats,
        }
        if not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(eng, data.data_ptr(), args.cpu_reads, keys)
        emit(out)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def cpu_baseline(eng, dptr, n, keys):
    """oracle port on the first n reads of rank 0's stream, one core"""
    from oracle import oracle as O
    host = eng.d2h(dptr, n * REC)
    cfg = O.make_config(miss=1, phred=30, length=FEAT_LEN, start="0")
    O.lib()
    t0 = time.perf_counter()
    O.count(cfg, keys, host)
    dt = time.perf_counter() - t0
    return {"value": n / dt / 1e6, "unit": "M reads/s", "cores": 1, "kind": "port",
            "sample": f"first {n} reads of the same synthetic stream, oracle/f2q_oracle.c single thread, {dt:.1f} s"}


def numa_bind(gpu_uuid, gpu_index):
    """run this rank (and first-touch its pinned buffers) on the CPUs nearest to its GPU, when the container allows it"""
    try:
        import pynvml as nv
        nv.nvmlInit()
        try:
            h = nv.nvmlDeviceGetHandleByUUID(f"GPU-{gpu_uuid}".encode()) if gpu_uuid else nv.nvmlDeviceGetHandleByIndex(gpu_index)
        except Exception:
            h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
        allowed = os.sched_getaffinity(0)
        words = nv.nvmlDeviceGetCpuAffinity(h, (max(allowed) // 64) + 1)
        ideal = {64 * w + b for w, x in enumerate(words) for b in range(64) if (int(x) >> b) & 1}
        both = allowed & ideal
        if both and both != allowed:
            os.sched_setaffinity(0, both)
            return f"bound to {len(both)} of {len(allowed)} allowed CPUs near the GPU"
        return "not bound (all allowed CPUs are equally near, or none is)"
    except Exception as e:                                        # noqa: BLE001 — placement is an optimisation only
        return f"not bound ({type(e).__name__})"


def main():
    # anything a library prints on stdout (NCCL's version banner) must not precede the ONE JSON line: fd 1 -> stderr,
    # the line itself goes to the saved descriptor
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="f2q")
    ap.add_argument("--reads", type=int, default=100_000_000, help="reads per GPU (configs[1] = 100 M)")
    ap.add_argument("--cpu-reads", type=int, default=6_000_000)
    ap.add_argument("--ref-reads-per-thread", type=int, default=400_000)
    ap.add_argument("--tile-threads", type=int, default=0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_gpu(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
