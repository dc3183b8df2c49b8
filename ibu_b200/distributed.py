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


def owner_of(barcode, world: int) -> np.ndarray:
    """owner(barcode) = splitmix64(barcode) % world — the rank on which a barcode's pairs meet
    (host restatement of the device function in barcode_count.cu, for tests and planning)."""
    x = np.asarray(barcode, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z % np.uint64(world)).astype(np.int64)


def exact_barcode_table(ctx, d_records, n: int, device, group=None) -> np.ndarray:
    """Per-barcode table of the WHOLE job from per-rank record shards, exact for any input order
    (the path's one real exchange step):

      1. each rank de-duplicates its shard into a (barcode, umi, count) pair table   [device]
      2. rows are grouped by owner(barcode) and exchanged with ONE all-to-all        [NCCL]
      3. each owner counts the pairs it received, weighted                           [device]
      4. the owners' disjoint row sets are all-gathered and ordered by barcode       [small]

    Every rank returns the same table."""
    import torch
    import torch.distributed as dist

    world, rank = dist.get_world_size(group), dist.get_rank(group)
    ptr, k = ctx.pair_table(d_records, n)
    send = torch.empty((max(k, 1), 3), dtype=torch.int64, device=device)
    try:
        counts = ctx.partition_by_owner(ptr, k, world, send) if k else [0] * world
    finally:
        if ptr:
            ctx.free(ptr)
    in_split = torch.tensor(counts, dtype=torch.int64, device=device)
    out_split = torch.empty_like(in_split)
    dist.all_to_all_single(out_split, in_split, group=group)  # how many rows each peer sends me
    recv_counts = [int(c) for c in out_split.cpu()]
    recv = torch.empty((max(sum(recv_counts), 1), 3), dtype=torch.int64, device=device)
    dist.all_to_all_single(recv[: sum(recv_counts)], send[:k], recv_counts, counts, group=group)
    torch.cuda.synchronize(device)
    m = sum(recv_counts)
    from . import COUNT_WEIGHTED

    # 3. the owner counts what it received; its rows stay on the device
    mine = torch.empty((0, 3), dtype=torch.int64, device=device)
    if m:
        table, _ = ctx.barcode_count_device(recv, m, COUNT_WEIGHTED)
        try:
            mine = torch.empty((int(table.n_rows), 3), dtype=torch.int64, device=device)
            if table.n_rows:
                ctx.memcpy(mine, int(table.d_rows), int(table.n_rows) * 24)
        finally:
            ctx.table_free(table)
    # 4. all-gather the owners' disjoint row sets (padded to the longest) and put them in barcode
    #    order with the device sort (rows are 24-byte records led by the barcode)
    sizes = torch.tensor([mine.shape[0]], dtype=torch.int64, device=device)
    all_sizes = [torch.empty_like(sizes) for _ in range(world)]
    dist.all_gather(all_sizes, sizes, group=group)
    counts_rows = [int(x.item()) for x in all_sizes]
    longest = max(max(counts_rows), 1)
    padded = torch.zeros((longest, 3), dtype=torch.int64, device=device)
    padded[: mine.shape[0]] = mine
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    cat = torch.cat([p[:c] for p, c in zip(parts, counts_rows)]) if sum(counts_rows) else mine
    total = int(cat.shape[0])
    out = np.zeros(total, ROW_DTYPE)
    if total:
        ordered = torch.empty_like(cat)
        torch.cuda.synchronize(device)
        ctx.sort_records(cat, total, ordered)
        ctx.d2h(out, ordered)
    return out
