"""One process per GPU (torch.distributed): how reads are sharded over ranks and how the per-rank results are merged.

The path shards by reads with no data-path collective: a rank parses its own reads into its own
uint64 vector [counts[n_keys] | stats[5]]; the vectors are added once per sample — merge_feature_dicts and the statistics
additions of the reference (fast2q.py:439-445, 487-495) — by one all-reduce (NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations


def rank_read_range(n_reads: int, rank: int, world: int) -> tuple[int, int]:
    """strong-scaling split of n_reads over `world` ranks: contiguous [first, first + count), sizes differ by at most 1"""
    base, extra = divmod(n_reads, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def rank_shards(shards, rank: int, world: int):
    """round-robin ownership of record-aligned shards (fast2q.record_aligned_shards): rank r parses shards r, r+world, ..."""
    for k, item in enumerate(shards):
        if k % world == rank:
            yield item


def merge_results(result, group=None):
    """in-place sum over ranks of the [counts | stats] vector (an int64 view of the device vector, or a CPU tensor)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(result, op=dist.ReduceOp.SUM, group=group)
    return result
