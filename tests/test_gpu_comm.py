"""
-m gpu: libf2q's multi-GPU entry points (f2q_comm_* / f2q_allreduce_counts / f2q_ec_merge, NCCL loaded at run time) in
ONE process driving one context per visible B200: record-aligned shards of a stream round-robin over the contexts (the
--fs layout), merged through the C-ABI, against the oracle on the whole stream.  With a single GPU the communicator has
one rank and the same code runs (the merge of a rank with itself must change nothing).
"""
import importlib

import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

f2q = importlib.import_module("2fast2q_b200")
lib = f2q._lib
host = importlib.import_module("2fast2q_b200.fast2q")
synth = importlib.import_module("2fast2q_b200.synth")


def engines(params, keys, n):
    es = [lib.Engine(lib.make_config(**params), d) for d in range(n)]
    if keys is not None:
        for e in es:
            e.set_library(keys)
    lib.comm_init(es)
    return es


def shards_of(data, n_bytes=1 << 20):
    blocks = (bytes(data[o:o + 300_000]) for o in range(0, len(data), 300_000))
    return [s for s, _ in host.record_aligned_shards(blocks, shard_bytes=n_bytes)]


def test_counter_allreduce_over_all_gpus(oracle):
    n = lib.device_count()
    spec = synth.default_spec(3)
    names, keys = synth.make_library(3, 5000, 20)
    data = synth.fixed_reads(keys, 0, 150_000, **spec)
    want_c, want_s = oracle.count(oracle.make_config(miss=2), keys, data)
    es = engines(dict(miss=2), keys, n)
    try:
        for rep in range(2):                                    # twice: the communicator is reused by the next sample
            for e in es:
                e.begin()
            for k, s in enumerate(shards_of(data)):
                es[k % n].submit(s, False)
            lib.allreduce_counts(es)                           # (flushes every context's carried record first)
            for e in es:
                c, st = e.end()
                assert st == want_s and np.array_equal(c, want_c), (n, rep)
    finally:
        for e in es:
            e.close()


def test_extract_count_merge_over_all_gpus(oracle):
    n = lib.device_count()
    spec = synth.shape_spec("4")
    guides = synth.random_kmers(4, 30_000, 20)
    data = bytearray(synth.shaped_reads(guides, 0, 120_000, **spec).tobytes())
    # some keys the packed table cannot hold (an N inside the barcode): they travel through the byte-arena tables
    rec = 168
    for i in range(0, 120_000, 97):
        data[i * rec + 14 + 30] = ord("N")
    data = bytes(data)
    params = dict(mode="EC", upstream=synth.BARSEQ_US.decode(), downstream=synth.BARSEQ_DS.decode(), miss_search_up=1, miss_search_down=1)
    want, want_s = oracle.extract_count(oracle.make_config(**params), data)
    assert any(b"N" in k for k in want)
    es = engines(params, None, n)
    try:
        for rep in range(2):
            for e in es:
                e.begin()
            for k, s in enumerate(shards_of(data)):
                es[k % n].submit(s, False)
            lib.allreduce_counts(es)                           # the five statistics
            stats = [e.end()[1] for e in es]
            lib.ec_merge(es)
            for e, st in zip(es, stats):
                assert st == want_s, (n, rep)
                assert e.ec_items() == want, (n, rep)
    finally:
        for e in es:
            e.close()


def test_comm_errors():
    e = lib.Engine(lib.make_config(), 0)
    try:
        with pytest.raises(lib.F2QError):
            lib.allreduce_counts([e])                          # no communicator yet
        lib.comm_init([e])
        with pytest.raises(lib.F2QError):
            lib.comm_init([e])                                 # already has one
        with pytest.raises(lib.F2QError):
            lib.ec_merge([e])                                  # Counter mode context
    finally:
        e.close()


def test_host_alloc_near_device():
    """pinned memory on the GPU's NUMA node where the platform shows several nodes, plain pinned memory otherwise (VMs)"""
    for wc in (False, True):
        b = lib.PinnedBuffer(1 << 20, device=0, write_combined=wc)
        assert b.numa_node >= -1
        b.array[:] = 7
        with lib.Engine(lib.make_config(mode="EC", upstream="ACGT", downstream="TTTT"), 0) as e:
            e.begin()
            b.array[:22] = np.frombuffer(b"@r\nACGTGGTTTT\n+\nIIIIIIIIII\n"[:22], dtype=np.uint8)
            e.submit_ptr(b.ptr.value, 0, True)
            e.end()
        b.free()
