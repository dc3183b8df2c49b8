#!/usr/bin/env python
"""group_bench.py — the exact multi-GPU table merge (ibu_gpu_group_barcode_count) over device-resident
shards, one JSON line per (shape, exchange): time split into local de-duplication, grouping +
exchange, owner count, gather.  Run under `gpurun --gpus N`; one process drives all N GPUs."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import ibu_b200 as ibu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=1_000_000_000, help="whole job")
    ap.add_argument("--gpus", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    world = args.gpus or ibu.device_count()
    n = args.records
    shapes = [("example pattern", ibu.GEN_PATTERN, 0),
              ("10x-like uniform whitelist 1M x umi 20", ibu.GEN_WHITELIST, (20 << 32) | 1_000_000),
              ("10x-like zipf 1M x umi 4096", ibu.GEN_ZIPF, (4096 << 32) | 1_000_000)]
    with ibu.GpuGroup(list(range(world))) as g:
        shards, lens = [], []
        for r in range(world):
            s, e = ibu.shard_range(n, r, world)
            c = g.ctx(r)
            shards.append(c.malloc(24 * (e - s)))
            lens.append(e - s)
        for name, gen, param in shapes:
            for r in range(world):
                s, _ = ibu.shard_range(n, r, world)
                c = g.ctx(r)
                c.generate_records_async(shards[r], s, lens[r], 16, 12, gen, param, 2024)
                c.synchronize()
            ref = None
            for ex_name, ex in (("p2p", ibu.EXCHANGE_P2P), ("host", ibu.EXCHANGE_HOST), ("nccl", ibu.EXCHANGE_NCCL)):
                if ex == ibu.EXCHANGE_NCCL and world < 2:
                    continue
                best = None
                try:
                    for _ in range(args.reps + 1):
                        t0 = time.perf_counter()
                        rows, info, tm = g.barcode_count(shards, lens, mode=ibu.count_lens(16, 12), exchange=ex)
                        sec = time.perf_counter() - t0
                        if best is None or sec < best[0]:
                            best = (sec, rows, info, tm)
                except ibu.IbuError as exc:
                    print(json.dumps(dict(shape=name, exchange=ex_name, error=str(exc))), flush=True)
                    continue
                sec, rows, info, tm = best
                if ref is None:
                    ref = rows
                print(json.dumps(dict(shape=name, n_gpus=world, records=n, exchange_name=ex_name, sec=sec, grec_s=n / sec / 1e9,
                                      rows=len(rows), distinct_pairs=info["n_distinct_pairs"],
                                      same_table_as_first_exchange=bool(np.array_equal(rows, ref)),
                                      records_sum_ok=bool(int(rows["n_records"].sum()) == n), **tm)), flush=True)
        for r in range(world):
            g.ctx(r).free(shards[r])


if __name__ == "__main__":
    main()
