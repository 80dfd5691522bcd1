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


def merge_ec_tables(key_bytes, key_offsets, counts, group=None):
    """Extract+Count across ranks: all-gather of every rank's (keys, counts) table followed by a sort-merge.

    key_bytes uint8[total], key_offsets uint64[n+1], counts uint64[n] are one rank's drained table (f2q_ec_drain).  Every rank
    gets the merged table back as (list of key bytes, list of counts), keys in byte order.  The tables are padded to the
    longest key, gathered with two collectives (sizes, then rows), and merged by sorting the rows and summing equal
    neighbours — torch.unique(dim=0) is that sort + segmented reduce — on the device of the backend (CUDA for NCCL).
    Keys may be empty and may contain any byte; a length column keeps b"A" and b"A\\x00" apart."""
    import numpy as np
    import torch
    import torch.distributed as dist
    kb = np.ascontiguousarray(key_bytes, dtype=np.uint8)
    ko = np.ascontiguousarray(key_offsets, dtype=np.uint64).astype(np.int64)
    cn = np.ascontiguousarray(counts, dtype=np.uint64).astype(np.int64)
    n = len(cn)
    lens = (ko[1:n + 1] - ko[:n]) if n else np.zeros(0, dtype=np.int64)
    distributed = dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    dev = torch.device("cuda", torch.cuda.current_device()) if distributed and dist.get_backend(group) == "nccl" else torch.device("cpu")
    world = dist.get_world_size(group) if distributed else 1
    sizes = torch.tensor([n, int(lens.max()) if n else 0], dtype=torch.int64, device=dev)
    if distributed:
        all_sizes = [torch.zeros_like(sizes) for _ in range(world)]
        dist.all_gather(all_sizes, sizes, group=group)
        n_max, l_max = max(int(s[0]) for s in all_sizes), max(int(s[1]) for s in all_sizes)
        ns = [int(s[0]) for s in all_sizes]
    else:
        n_max, l_max, ns = n, int(sizes[1]), [n]
    width = l_max + 2                                           # [length hi, length lo... ] see below: 2 length bytes + key bytes
    rows = np.zeros((n_max, width + 8), dtype=np.uint8)         # + 8 bytes of count (little endian)
    if n:
        rows[:n, 0] = (lens >> 8).astype(np.uint8)
        rows[:n, 1] = (lens & 0xFF).astype(np.uint8)
        # scatter the key bytes into their rows
        row_of = np.repeat(np.arange(n), lens)
        col_of = np.arange(int(lens.sum())) - np.repeat(ko[:n] - ko[0], lens) + 2
        rows[row_of, col_of] = kb[int(ko[0]):int(ko[0]) + int(lens.sum())]
        rows[:n, width:] = cn[:n].astype("<i8").view(np.uint8).reshape(n, 8)
    t = torch.from_numpy(rows).to(dev)
    if distributed:
        gathered = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(gathered, t, group=group)
        t = torch.cat([g[:k] for g, k in zip(gathered, ns)], dim=0)
    else:
        t = t[:n]
    if t.shape[0] == 0:
        return [], []
    keys_t = t[:, :width].contiguous()
    cnt_t = t[:, width:].contiguous().cpu().numpy().view("<i8").reshape(-1)
    uniq, inverse = torch.unique(keys_t, dim=0, return_inverse=True)      # rows sorted; equal keys share an index
    total = torch.zeros(uniq.shape[0], dtype=torch.int64, device=dev)
    total.index_add_(0, inverse, torch.from_numpy(cnt_t.copy()).to(dev))
    u = uniq.cpu().numpy()
    tot = total.cpu().numpy()
    out_keys = []
    for r in range(u.shape[0]):
        ln = (int(u[r, 0]) << 8) | int(u[r, 1])
        out_keys.append(u[r, 2:2 + ln].tobytes())
    order = sorted(range(len(out_keys)), key=lambda i: out_keys[i])
    return [out_keys[i] for i in order], [int(tot[i]) for i in order]
