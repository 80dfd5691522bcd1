"""
The N>1 path on CPU: world_size-2 gloo.  Each rank owns part of the reads (contiguous read range, as bench.py shards them,
or every 2nd record-aligned shard, as the CLI's File-Split mode does), counts them — here with the oracle standing in for
the device, it is only the checker of the sharding/merge logic — and the [counts | stats] vectors are merged by ONE
all-reduce.  The merged vector must equal the count of the whole stream.
"""
import importlib
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STATS = ("reads", "perfect_counter", "imperfect_counter", "non_aligned_counter", "quality_failed")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mode, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        synth = importlib.import_module("2fast2q_b200.synth")
        multi = importlib.import_module("2fast2q_b200.multi")
        host = importlib.import_module("2fast2q_b200.fast2q")
        spec = synth.default_spec(2)
        names, keys = synth.make_library(2, 300, 20)
        cfg = O.make_config(miss=1)
        n = 30_001
        vec = torch.zeros(len(keys) + 5, dtype=torch.int64)

        def add(data):
            c, s = O.count(cfg, keys, data)
            vec[:len(keys)] += torch.from_numpy(c.astype(np.int64))
            vec[len(keys):] += torch.tensor([s[k] for k in STATS], dtype=torch.int64)

        if mode == "read_range":
            first, count = multi.rank_read_range(n, rank, world)
            add(synth.fixed_reads(keys, first, count, **spec))
        else:
            whole = synth.fixed_reads(keys, 0, n, **spec).tobytes() + b"@tail\nACGT\n"      # left-over lines stay in the last shard
            blocks = [whole[o:o + 70_001] for o in range(0, len(whole), 70_001)]
            for shard, final in multi.rank_shards(host.record_aligned_shards(blocks, 200_000), rank, world):
                add(np.frombuffer(shard, dtype=np.uint8))
        multi.merge_results(vec)
        if rank == 0:
            whole = synth.fixed_reads(keys, 0, n, **spec)
            c, s = O.count(cfg, keys, whole)
            ok = bool(np.array_equal(vec[:len(keys)].numpy(), c.astype(np.int64))) and [int(x) for x in vec[len(keys):]] == [s[k] for k in STATS]
            out.put((mode, ok, int(vec[len(keys)])))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["read_range", "record_shards"])
def test_world_size_2_gloo_merge(mode):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, mode, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    m, ok, reads = out.get(timeout=5)
    assert m == mode and ok and reads == 30_001


def test_rank_read_range_partitions_exactly():
    multi = importlib.import_module("2fast2q_b200.multi")
    for n in (0, 1, 7, 100, 1_000_003):
        for world in (1, 2, 3, 8):
            parts = [multi.rank_read_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[r][0] + parts[r][1] == parts[r + 1][0] for r in range(world - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1


def _ec_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        synth = importlib.import_module("2fast2q_b200.synth")
        multi = importlib.import_module("2fast2q_b200.multi")
        host = importlib.import_module("2fast2q_b200.fast2q")
        cfg = O.make_config(mode="EC", upstream="GTTCAGAGTTCT", downstream="CTGAATAGGCCA", miss_search_up=1, miss_search_down=1)
        whole = synth.barseq_reads(4, 9000) + b"@a\n\n+\n\n"                 # (an empty read: the empty key must survive the merge)
        blocks = [whole[o:o + 50_001] for o in range(0, len(whole), 50_001)]
        mine = {}
        for shard, final in multi.rank_shards(host.record_aligned_shards(blocks, 120_000), rank, world):
            d, _ = O.extract_count(cfg, np.frombuffer(shard, dtype=np.uint8))
            for k, v in d.items():
                mine[k] = mine.get(k, 0) + v
        keys = list(mine)
        kb = np.frombuffer(b"".join(keys), dtype=np.uint8) if keys else np.zeros(0, dtype=np.uint8)
        ko = np.cumsum([0] + [len(k) for k in keys]).astype(np.uint64)
        cn = np.array([mine[k] for k in keys], dtype=np.uint64)
        mk, mc = multi.merge_ec_tables(kb, ko, cn)
        want, _ = O.extract_count(cfg, np.frombuffer(whole, dtype=np.uint8))
        ok = dict(zip(mk, mc)) == want and mk == sorted(mk)
        allok = torch.tensor([1 if ok else 0])
        dist.all_reduce(allok, op=dist.ReduceOp.MIN)
        if rank == 0:
            out.put((bool(allok.item()), len(mk), len(mine)))
    finally:
        dist.destroy_process_group()


def test_extract_count_allgather_sort_merge_world_size_2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ec_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(240)
        assert p.exitcode == 0
    ok, n_merged, n_mine = out.get(timeout=5)
    assert ok and n_merged >= n_mine > 0


def test_merge_ec_tables_single_process():
    multi = importlib.import_module("2fast2q_b200.multi")
    keys = [b"ACGT", b"", b"A", b"A\x00", b"ACGT"]
    kb = np.frombuffer(b"".join(keys), dtype=np.uint8)
    ko = np.cumsum([0] + [len(k) for k in keys]).astype(np.uint64)
    mk, mc = multi.merge_ec_tables(kb, ko, np.array([3, 1, 2, 5, 4], dtype=np.uint64))
    assert dict(zip(mk, mc)) == {b"ACGT": 7, b"": 1, b"A": 2, b"A\x00": 5} and mk == sorted(mk)
