#!/usr/bin/env python
"""multi_gpu_table.py — configs[3] across GPUs (run under torchrun): the 10^9-record example
pattern is range-sharded (mmap.rs:297-307 rule), every rank validates/reduces its shard (K1) and
the exact per-barcode record / distinct-UMI table of the whole job is built with one NCCL
all-to-all of de-duplicated pair tables.  Rank 0 prints JSON lines; results are checked against
the closed form (10^6 barcodes x N/10^6 records x 1 UMI)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import ibu_b200 as ibu  # noqa: E402
from ibu_b200 import distributed as ibd  # noqa: E402

n_total = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
ctx = ibu.GpuContext(local)
s, e = ibd.my_shard(n_total)
n = e - s
recs = torch.empty(n * 24, dtype=torch.uint8, device=dev)
res = torch.zeros(8, dtype=torch.int64, device=dev)
ctx.generate_records_async(recs, s, n, 16, 12, ibu.GEN_PATTERN, 0, 0)
ctx.synchronize()


def timed(fn):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize()
    dist.barrier()
    return out, time.perf_counter() - t0


def reduce_all():
    ctx.validate_reduce_async(recs, n, 16, 12, res)
    ctx.synchronize()
    return ibd.merge_results(ctx.read_result(res), device=dev)


for rep in range(2):  # second pass is warm (allocations cached)
    merged, t_red = timed(reduce_all)
    table, t_tab = timed(lambda: ibd.exact_barcode_table(ctx, recs, n, dev))
ok = (merged["n_records"] == n_total and merged["sum_index"] == (n_total * (n_total - 1) // 2) % 2**64
      and len(table) == min(n_total, 1_000_000) and int(table["n_records"].sum()) == n_total
      and bool((table["n_distinct_umi"] == 1).all()) and bool((table["n_records"] == n_total // 1_000_000).all()))
if rank == 0:
    print(json.dumps(dict(world=world, records=n_total, per_rank=n, k1_reduce_allreduce_ms=t_red * 1e3,
                          k1_grec_s=n_total / t_red / 1e9, exact_table_ms=t_tab * 1e3,
                          table_grec_s=n_total / t_tab / 1e9, rows=len(table), closed_form_ok=bool(ok))), flush=True)
# ---- sorted shards: host merge of the per-rank tables (boundary fix-up) vs the pair all-to-all ----
# GEN_SORTED is sorted by Record's Ord over the whole job, so every rank's contiguous range is a
# sorted shard and only runs cut by a shard boundary can repeat.
ctx.generate_records_async(recs, s, n, 16, 12, ibu.GEN_SORTED, (5 << 32) | 1000, 0)
ctx.synchronize()


def host_merge():
    rows, info = ctx.barcode_count(recs, n)
    assert info["input_was_sorted"]
    edge = ibu.records(2)
    ctx.d2h(edge[:1], recs.data_ptr())
    ctx.d2h(edge[1:], recs.data_ptr() + (n - 1) * 24)
    boundary = ((int(edge["barcode"][0]), int(edge["umi"][0])), (int(edge["barcode"][1]), int(edge["umi"][1])))
    return ibd.gather_tables(rows, boundary)


for rep in range(2):
    merged_rows, t_host = timed(host_merge)
    exact_rows, t_a2a = timed(lambda: ibd.exact_barcode_table(ctx, recs, n, dev))
same = bool(len(merged_rows) == len(exact_rows) and np.array_equal(merged_rows, exact_rows))
if rank == 0:
    print(json.dumps(dict(world=world, records=n_total, data="sorted (1000 records, 200 distinct UMIs per barcode)",
                          sorted_stream_tables_plus_host_merge_ms=t_host * 1e3, pair_all_to_all_ms=t_a2a * 1e3,
                          rows=len(merged_rows), tables_identical=same)), flush=True)
ctx.close()
dist.destroy_process_group()
