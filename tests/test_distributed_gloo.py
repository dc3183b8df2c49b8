"""N>1 host logic on CPU: world_size-2/3 gloo groups exercise the range sharding
(mmap.rs:297-307 with ranks for threads) and the result/table merges of ibu_b200.distributed.
Per-shard results come from the oracle here (no GPU); on the GPU box the same merge code runs
over NCCL (bench.py, tests/test_gpu_multi.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import ibu_b200 as ibu
from ibu_b200 import distributed as ibd
from oracle import oracle_c as oc
from oracle import oracle_np as on


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, path, sorted_file, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        reader = ibu.MmapReader(path)
        n, h = reader.len(), reader.header()
        s, e = ibd.my_shard(n)
        assert (s, e) == ibu.shard_range(n, rank, world)
        shard = np.array(reader.slice(s, e)) if e > s else ibu.records(0)
        local = ibu.ReduceResult(oc.reduce_records(shard, h.bc_len, h.umi_len, 1))
        merged = ibd.merge_results(local)
        rows = on.barcode_table(shard)
        boundary = (tuple(shard[0]), tuple(shard[-1])) if sorted_file and len(shard) else None
        table = ibd.gather_tables(rows, boundary)
        np.save(os.path.join(out_dir, f"table{rank}.npy"), table)
        np.save(os.path.join(out_dir, f"res{rank}.npy"), np.array([merged[k] for k in ibd._FIELDS], np.uint64))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("sorted_file", [True, False])
def test_sharded_merge_equals_whole_file(tmp_path, world, sorted_file):
    n = 200_003
    recs = oc.generate_records(0, n, 16, 12, 3, (16 << 32) | 400, 17)
    recs["barcode"][::1000] |= np.uint64(1 << 40)  # a few invalid words; sums must still wrap-merge
    recs["index"] = np.uint64(2**64 - 1) - recs["index"]  # force u64 wrap-around in the sums
    if sorted_file:
        recs = recs[np.lexsort((recs["index"], recs["umi"], recs["barcode"]))]
    path = str(tmp_path / "f.ibu")
    with ibu.Writer(path, ibu.Header(16, 12)) as w:
        w.write_batch(recs)
    mp.spawn(_worker, args=(world, _free_port(), path, sorted_file, str(tmp_path)), nprocs=world, join=True)
    want = oc.reduce_records(recs, 16, 12)
    whole = on.barcode_table(recs)
    for r in range(world):
        got = dict(zip(ibd._FIELDS, map(int, np.load(tmp_path / f"res{r}.npy"))))
        assert got == want
        table = np.load(tmp_path / f"table{r}.npy")
        assert np.array_equal(table["barcode"], whole["barcode"])
        assert np.array_equal(table["n_records"], whole["n_records"])
        if sorted_file:  # exact: only boundary-cut pairs can repeat across shards
            assert np.array_equal(table["n_distinct_umi"], whole["n_distinct_umi"])
        else:            # documented upper bound for unsorted shards
            assert np.all(table["n_distinct_umi"] >= whole["n_distinct_umi"])


def test_merge_tables_host_boundary_cases():
    t = lambda *rows: np.array(list(rows), ibu.ROW_DTYPE)  # noqa: E731
    a, b = t((1, 5, 2), (7, 3, 3)), t((7, 4, 2), (9, 1, 1))
    m = ibd.merge_tables_host([a, np.zeros(0, ibu.ROW_DTYPE), b])
    assert [tuple(x) for x in m] == [(1, 5, 2), (7, 7, 5), (9, 1, 1)]
    # the (7, umi=4) run was cut by the boundary: counted once
    m = ibd.merge_tables_host([a, b], [((1, 0, 0), (7, 4, 10)), ((7, 4, 11), (9, 0, 0))])
    assert [tuple(x) for x in m] == [(1, 5, 2), (7, 7, 4), (9, 1, 1)]
    m = ibd.merge_tables_host([a, b], [((1, 0, 0), (7, 4, 10)), ((7, 5, 11), (9, 0, 0))])
    assert [tuple(x) for x in m] == [(1, 5, 2), (7, 7, 5), (9, 1, 1)]
    assert len(ibd.merge_tables_host([])) == 0


# ---- the exact table merge of the multi-GPU group (ibu_gpu_group_*, group.cu), rank for rank ----
# The library runs it with one host thread per GPU and device copies; here the same protocol runs
# on 8 gloo ranks with the numpy oracle standing in for the device steps:
#   1. de-duplicated (barcode, umi, multiplicity) pairs of the rank's shard   [ibu_gpu_pair_table]
#   2. grouped by owner(barcode) = splitmix64(barcode) % world and exchanged  [the one exchange step]
#   3. the owner counts what it received, weighted                            [IBU_COUNT_WEIGHTED]
#   4. the owners' disjoint row sets, gathered and put in barcode order
def _pairs_of(shard):
    key = np.stack([shard["barcode"], shard["umi"]], axis=1)
    uniq, counts = np.unique(key, axis=0, return_counts=True)
    return uniq[:, 0].copy(), uniq[:, 1].copy(), counts.astype(np.uint64)


def _weighted_table(bc, um, w):
    order = np.lexsort((um, bc))
    bc, um, w = bc[order], um[order], w[order]
    new_pair = np.ones(len(bc), bool)
    new_pair[1:] = (bc[1:] != bc[:-1]) | (um[1:] != um[:-1])
    head = np.ones(len(bc), bool)
    head[1:] = bc[1:] != bc[:-1]
    seg = np.cumsum(head) - 1
    out = np.zeros(int(seg[-1]) + 1 if len(bc) else 0, ibu.ROW_DTYPE)
    if len(bc):
        out["barcode"] = bc[head]
        np.add.at(out["n_records"], seg, w)
        np.add.at(out["n_distinct_umi"], seg, new_pair.astype(np.uint64))
    return out


def _exchange_worker(rank, world, port, n, out_dir):
    import torch

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        s, e = ibd.my_shard(n)
        shard = oc.generate_records(s, e - s, 16, 12, 5, (64 << 32) | 5_000, 91)  # Zipf barcodes, unsorted
        bc, um, w = _pairs_of(shard)
        owner = ibd.owner_of(bc, world)
        send = [torch.from_numpy(np.stack([bc[owner == r], um[owner == r], w[owner == r]], axis=1).view(np.int64).copy())
                for r in range(world)]
        sizes = torch.tensor([len(x) for x in send], dtype=torch.int64)
        got_sizes = torch.empty_like(sizes)
        dist.all_to_all_single(got_sizes, sizes)
        recv = torch.empty((int(got_sizes.sum()), 3), dtype=torch.int64)
        dist.all_to_all_single(recv, torch.cat(send), [int(c) for c in got_sizes], [int(c) for c in sizes])
        cat = recv.numpy().view(np.uint64)
        mine = _weighted_table(cat[:, 0].copy(), cat[:, 1].copy(), cat[:, 2].copy())
        assert np.all(ibd.owner_of(mine["barcode"], world) == rank)  # owners hold disjoint barcodes
        parts = [None] * world
        dist.all_gather_object(parts, mine)
        table = np.concatenate(parts)
        table = table[np.argsort(table["barcode"], kind="stable")]
        np.save(os.path.join(out_dir, f"exact{rank}.npy"), table)
    finally:
        dist.destroy_process_group()


def test_exact_table_exchange_protocol_on_8_ranks(tmp_path):
    n, world = 160_003, 8
    mp.spawn(_exchange_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    whole = on.barcode_table(oc.generate_records(0, n, 16, 12, 5, (64 << 32) | 5_000, 91))
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"exact{r}.npy"), whole), r
    # the owner function the library uses is the one restated for the host (barcode_count.cu, k_owner_*)
    assert [int(x) for x in ibd.owner_of(np.array([0, 1, 2**63, 2**64 - 1], np.uint64), 8)] == \
        [int(oc.splitmix64(int(v)) % 8) for v in (0, 1, 2**63, 2**64 - 1)]
