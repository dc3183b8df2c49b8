#!/usr/bin/env python
"""k4bench.py — the per-barcode table (K4, ibu_gpu_barcode_count) on the shapes SURVEY §8(d) names.

Wall clock of the blocking C-ABI call with the rows left on the device (best / mean of `--iters`),
one JSON line per case.  Algorithmic bytes = 24 B/record (the records are read once); everything
else a path moves is overhead of the implementation, which is the point of the table.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ibu_b200 as ibu  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


CASES = {
    # name: (generator, param, mode bits)
    "sorted 1000/barcode 5/umi (streaming only)": (ibu.GEN_SORTED, (5 << 32) | 1000, 1),
    "sorted 1000/barcode 5/umi (auto)": (ibu.GEN_SORTED, (5 << 32) | 1000, 0),
    "10x-like uniform: 1M barcodes, umi space 20 (5x dup)": (ibu.GEN_WHITELIST, (20 << 32) | 1_000_000, 0),
    "10x-like zipf: 1M barcodes, umi space 4096": (ibu.GEN_ZIPF, (4096 << 32) | 1_000_000, 0),
    "example pattern (i%1e6, 31i%1e6)": (ibu.GEN_PATTERN, 0, 0),
    "whitelist 1M barcodes, umi space 4096 (near-distinct pairs)": (ibu.GEN_WHITELIST, (4096 << 32) | 1_000_000, 0),
    "near-distinct clean": (ibu.GEN_CLEAN, 0, 0),
    "near-distinct dirty 1%": (ibu.GEN_DIRTY, 10_000, 0),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--lens", type=int, default=1, help="pass the header's lengths as a hint")
    ap.add_argument("--path", default="", help="partition | legacy | sort: force one unsorted path")
    ap.add_argument("--ctx-stream", type=int, default=0, help="run the table on the context's own stream")
    ap.add_argument("--verbose", type=int, default=0, help="every call's time on stderr")
    ap.add_argument("--sort", type=int, default=0, help="also time ibu_gpu_sort_records on this many random bc16/umi12 records")
    args = ap.parse_args()
    n, pk = args.records, peak()
    dev = torch.device("cuda", 0)
    ctx = ibu.GpuContext(0)
    stream = torch.cuda.Stream(device=dev)
    extra = ibu.count_lens(16, 12) if args.lens else 0
    extra |= {"": 0, "partition": ibu.COUNT_PATH_PARTITION, "legacy": ibu.COUNT_PATH_LEGACY, "sort": ibu.COUNT_PATH_SORT}[args.path]
    with torch.cuda.stream(stream):
        recs = torch.empty(24 * n, dtype=torch.uint8, device=dev)
        for name, (gen, param, mode) in CASES.items():
            if args.only and not any(o in name for o in args.only.split(",")):
                continue
            ctx.generate_records_async(recs, 0, n, 16, 12, gen, param, 3, stream)
            stream.synchronize()
            ts, fs, info = [], [], None
            for _ in range(args.iters + 1):  # the first call sizes the scratch pools
                t0 = time.perf_counter()
                table, info = ctx.barcode_count_device(recs, n, mode | extra, None if args.ctx_stream else stream)
                t1 = time.perf_counter()
                ctx.table_free(table)
                ts.append((t1 - t0) * 1e3)
                fs.append((time.perf_counter() - t1) * 1e3)
            if args.verbose:
                print("# calls ms:", [round(t, 2) for t in ts], "table_free ms:", [round(t, 2) for t in fs], file=sys.stderr)
            first, ts = ts[0], sorted(ts[1:])
            print(json.dumps(dict(case=name, records=n, ms_best=ts[0], ms_mean=sum(ts) / len(ts), ms_first_call=first,
                                  grec_s=n / ts[0] / 1e6, achieved_gbs=24 * n / ts[0] / 1e6, frac=24 * n / ts[0] / 1e6 / pk,
                                  rows=info["n_rows"], distinct_pairs=info["n_distinct_pairs"],
                                  sorted_input=info["input_was_sorted"], lens_hint=bool(args.lens), path=args.path or "auto",
                                  timing="wall clock of the blocking ibu_gpu_barcode_count, rows left on the device")),
                  flush=True)
    if args.sort:
        m = args.sort
        with torch.cuda.stream(stream):
            src = torch.empty(24 * m, dtype=torch.uint8, device=dev)
            dst = torch.empty(24 * m, dtype=torch.uint8, device=dev)
            ctx.generate_records_async(src, 0, m, 16, 12, ibu.GEN_CLEAN, 0, 3, stream)
            stream.synchronize()
            ts = []
            for _ in range(args.iters + 1):
                t0 = time.perf_counter()
                ctx.sort_records(src, m, dst, stream)
                ts.append((time.perf_counter() - t0) * 1e3)
            first, ts = ts[0], sorted(ts[1:])
            # barcode 32 bits + umi 24 bits, 8 bits per pass; the generator's index is the record number,
            # i.e. in input order, so the index passes are skipped
            passes = 4 + 3
            print(json.dumps(dict(case="ibu_gpu_sort_records (random bc16/umi12, index = i)", records=m, ms_best=ts[0],
                                  ms_mean=sum(ts) / len(ts), ms_first_call=first, digit_passes=passes,
                                  gbs_of_48B_per_pass=48 * m * passes / ts[0] / 1e6,
                                  timing="wall clock of the blocking call")), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
