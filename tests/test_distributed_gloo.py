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
