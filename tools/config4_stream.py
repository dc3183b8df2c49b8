#!/usr/bin/env python
"""config4_stream.py — configs[4] (and, with --records 1000000000, configs[3]) end to end across
GPUs: a record stream range-sharded over the ranks (mmap.rs:297-307 rule), every rank ingests its
shard from PINNED HOST memory through the double-buffered staging pipeline.

  phase A  ingest + validate + count: ibu_gpu_process_host (H2D chunks -> K1), 8-word results
           all-reduced — the streaming counterpart of process_parallel
  phase B  ingest into HBM + exact per-barcode record / distinct-UMI table of the whole job
           (hash aggregation per shard, one NCCL all-to-all of de-duplicated pairs)

The stream is the reference's example pattern (i % 10^6, 31 i % 10^6, i)
(examples/parallel.rs:65-69), generated on the device and copied into each rank's pinned shard
before the timed phases; results are checked against the closed form.  Run under torchrun;
rank 0 prints JSON lines.  Host memory needed: 24 bytes x records, pinned, over all ranks."""
import argparse
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


def mem_available() -> int:
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable:"):
            return int(line.split()[1]) * 1024
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=4_000_000_000)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--no-table", action="store_true")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)

    n_total = args.records
    budget = int(mem_available() * 0.6)  # pinned shards of all ranks must fit in host memory
    if 24 * n_total > budget:
        n_total = budget // 24 // (world * 1_000_000) * (world * 1_000_000)
    t = torch.tensor([n_total], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    n_total = int(t[0])
    s, e = ibd.my_shard(n_total)
    n = e - s
    ctx = ibu.GpuContext(local)

    # ---- the rank's shard of the stream in pinned host memory (untimed set-up) ----
    t0 = time.perf_counter()
    pin = ibu.PinnedBuffer(n * 24)
    h_recs = pin.array(ibu.RECORD_DTYPE, (n,))
    t_alloc = time.perf_counter() - t0
    gen_chunk = 64 << 20
    d_tmp = torch.empty(min(n, gen_chunk) * 24, dtype=torch.uint8, device=dev)
    t0 = time.perf_counter()
    for off in range(0, n, gen_chunk):
        cnt = min(gen_chunk, n - off)
        ctx.generate_records_async(d_tmp, s + off, cnt, 16, 12, ibu.GEN_PATTERN, 0, 0)
        ctx.synchronize()
        ctx.d2h(h_recs[off:off + cnt], d_tmp)
    t_fill = time.perf_counter() - t0
    del d_tmp
    torch.cuda.empty_cache()

    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        dist.barrier()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return out, float(dt[0])

    # ---- phase A: streaming ingest + validate + count ----
    def phase_a():
        return ibd.merge_results(ctx.process_host(h_recs, 16, 12), device=dev)

    best_a, merged = 1e30, None
    for _ in range(args.reps):
        merged, dt = timed(phase_a)
        best_a = min(best_a, dt)
    ok_a = (merged["n_records"] == n_total and merged["sum_index"] == (n_total * (n_total - 1) // 2) % 2**64
            and merged["n_bad_records"] == 0)
    if rank == 0:
        print(json.dumps(dict(phase="A ingest+validate+count (pinned host -> H2D chunks -> K1, all-reduce)", world=world,
                              records=n_total, per_rank=n, sec=best_a, grec_s=n_total / best_a / 1e9,
                              link_gbs_total=24 * n_total / best_a / 1e9, link_gbs_per_gpu=24 * n_total / best_a / 1e9 / world,
                              closed_form_ok=bool(ok_a), pinned_alloc_s=t_alloc, fill_s=t_fill)), flush=True)

    # ---- phase B: ingest into HBM + exact whole-job table ----
    if not args.no_table:
        recs = torch.empty(n * 24, dtype=torch.uint8, device=dev)

        def phase_b():
            ctx.h2d(recs, h_recs)
            return ibd.exact_barcode_table(ctx, recs, n, dev)

        best_b, table = 1e30, None
        for _ in range(args.reps):
            table, dt = timed(phase_b)
            best_b = min(best_b, dt)
        per = n_total // 1_000_000
        ok_b = (len(table) == min(n_total, 1_000_000) and int(table["n_records"].sum()) == n_total
                and bool((table["n_distinct_umi"] == 1).all())
                and (n_total % 1_000_000 != 0 or bool((table["n_records"] == per).all())))
        _, t_tab = timed(lambda: ibd.exact_barcode_table(ctx, recs, n, dev))
        if rank == 0:
            print(json.dumps(dict(phase="B ingest to HBM + exact per-barcode table (pair all-to-all)", world=world,
                                  records=n_total, sec=best_b, grec_s=n_total / best_b / 1e9,
                                  table_only_sec=t_tab, table_only_grec_s=n_total / t_tab / 1e9, rows=len(table),
                                  closed_form_ok=bool(ok_b))), flush=True)
        del recs

    # ---- the CPU restatement on the same host memory, rank 0 only, bounded sample ----
    if rank == 0:
        from oracle import oracle_c as oc
        sample = min(n, 500_000_000)
        best = 1e30
        for _ in range(2):
            t0 = time.perf_counter()
            oc.reduce_records(h_recs[:sample], 16, 12, 0)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps(dict(phase="cpu oracle process_parallel reduce (host memory, rank 0 alone)", cores=oc.num_cpus(),
                              sample_records=sample, sec=best, grec_s=sample / best / 1e9, gb_s=24 * sample / best / 1e9)),
              flush=True)
    dist.barrier()
    del h_recs
    pin.free()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
