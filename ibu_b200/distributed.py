"""Multi-GPU plumbing: one process per GPU, records sharded by contiguous range, only the small
results merged (torch.distributed: NCCL on GPUs, gloo in the CPU tests).

The partition is the reference's thread partition with ranks in place of threads
(src/io/mmap.rs:297-307: len / world each, the last rank takes the remainder); the merge is
the reference processors' on_batch_complete merge (mmap.rs:365-372,
examples/parallel.rs:28-35): wrapping u64 adds, xor for the checksum.
"""
from __future__ import annotations

import numpy as np

from . import ROW_DTYPE, ReduceResult, shard_range

_FIELDS = ("n_records", "sum_barcode", "sum_umi", "sum_index", "xor_all", "n_bad_barcode", "n_bad_umi",
           "n_bad_records")
_XOR = _FIELDS.index("xor_all")


def my_shard(length: int, group=None) -> tuple[int, int]:
    """Record range of this rank."""
    import torch.distributed as dist

    return shard_range(length, dist.get_rank(group), dist.get_world_size(group))


def _u64_to_i64(values) -> np.ndarray:
    return np.asarray(values, dtype=np.uint64).view(np.int64)


def merge_results(local: ReduceResult, device=None, group=None) -> ReduceResult:
    """All-reduce of the 8-word result block.  Sums wrap mod 2^64 (two's-complement int64 adds on
    the wire are the same bits); xor_all cannot ride a sum, so the words are all-gathered and
    folded (world x 8 bytes)."""
    import torch
    import torch.distributed as dist

    words = torch.from_numpy(_u64_to_i64([local[k] for k in _FIELDS]).copy())
    if device is not None:
        words = words.to(device)
    summed = words.clone()
    dist.all_reduce(summed, op=dist.ReduceOp.SUM, group=group)
    gathered = [torch.empty_like(words) for _ in range(dist.get_world_size(group))]
    dist.all_gather(gathered, words, group=group)
    out = summed.cpu().numpy().view(np.uint64).copy()
    out[_XOR] = np.bitwise_xor.reduce(np.stack([g.cpu().numpy().view(np.uint64) for g in gathered])[:, _XOR])
    return ReduceResult(zip(_FIELDS, map(int, out)))


def merge_tables_host(tables: list[np.ndarray], boundaries: list[tuple] | None = None) -> np.ndarray:
    """Merge per-shard barcode tables (rank order) into one, on the host — the tables are
    O(#barcodes) rows, MBs at most.

    n_records adds by key.  n_distinct_umi adds by key as well when no (barcode, umi) pair
    occurs in two shards; for a file sorted by Record's Ord (record.rs:58) the only pairs that
    can straddle are the ones cut by a shard boundary, and `boundaries[r] = (first_record,
    last_record)` of every non-empty shard lets them be counted once.  For unsorted files the
    distinct count of a barcode seen by several shards is an upper bound (exact merging needs
    the de-duplicated pair tables)."""
    tables = [np.ascontiguousarray(t, ROW_DTYPE) for t in tables]
    nonempty = [t for t in tables if len(t)]
    if not nonempty:
        return np.zeros(0, ROW_DTYPE)
    cat = np.concatenate(nonempty)
    order = np.argsort(cat["barcode"], kind="stable")
    cat = cat[order]
    head = np.ones(len(cat), bool)
    head[1:] = cat["barcode"][1:] != cat["barcode"][:-1]
    seg = np.cumsum(head) - 1
    out = np.zeros(int(seg[-1]) + 1, ROW_DTYPE)
    out["barcode"] = cat["barcode"][head]
    np.add.at(out["n_records"], seg, cat["n_records"])
    np.add.at(out["n_distinct_umi"], seg, cat["n_distinct_umi"])
    if boundaries is not None:
        live = [b for b in boundaries if b is not None]
        for (_, last), (first, _) in zip(live[:-1], live[1:]):
            if last[0] == first[0] and last[1] == first[1]:  # the same (barcode, umi) run was cut
                out["n_distinct_umi"][np.searchsorted(out["barcode"], np.uint64(first[0]))] -= 1
    return out


def gather_tables(rows: np.ndarray, boundary=None, group=None):
    """All-gather the per-rank tables (and shard boundary records) as objects; every rank merges."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    got = [None] * world
    dist.all_gather_object(got, (np.ascontiguousarray(rows, ROW_DTYPE), boundary), group=group)
    tables = [g[0] for g in got]
    bounds = [g[1] for g in got]
    return merge_tables_host(tables, bounds if any(b is not None for b in bounds) else None)
