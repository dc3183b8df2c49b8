#!/usr/bin/env python
"""kbench.py — per-kernel roofline table (CUDA-event timed, device-resident inputs larger than L2).

Prints one JSON line per kernel: algorithmic bytes/record x records / mean launch
time vs the measured HBM copy peak (MEASURED_PEAKS.json).  Not the bench contract (bench.py is);
this is the tuning loop's view of every kernel on the path, committed under profiles/.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ibu_b200 as ibu  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timeit(fn, stream, iters, warmup=3):
    for _ in range(warmup):
        fn()
    stream.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(stream)
        fn()
        b.record(stream)
    stream.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return sum(ms) / len(ms), ms[0], ms[len(ms) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    n, pk = args.records, peak()
    dev = torch.device("cuda", 0)
    ctx = ibu.GpuContext(0)
    stream = torch.cuda.Stream(device=dev)
    u8 = lambda k: torch.empty(k, dtype=torch.uint8, device=dev)  # noqa: E731
    out = []

    def report(name, alg_bytes, fn, **extra):
        if args.only and not any(o in name for o in args.only.split(",")):
            return
        mean, best, med = timeit(fn, stream, args.iters)
        line = dict(kernel=name, records=n, alg_bytes_per_record=alg_bytes, ms_mean=mean, ms_best=best, ms_median=med,
                    grec_s=n / mean / 1e6, achieved_gbs=alg_bytes * n / mean / 1e6, peak_gbs=pk,
                    frac=alg_bytes * n / mean / 1e6 / pk, frac_of_8TBs=alg_bytes * n / mean / 1e6 / 8000.0, **extra)
        out.append(line)
        print(json.dumps(line), flush=True)

    with torch.cuda.stream(stream):
        res = torch.zeros(8, dtype=torch.int64, device=dev)
        shapes = [(16, 12), (32, 32), (16, 16), (16, 10), (20, 10), (15, 9)]
        if args.only:
            shapes = [sh for sh in shapes if any(f"bc{sh[0]}/umi{sh[1]}" in o or o == "K1" and sh == (16, 12) for o in args.only.split(","))] or shapes[:1]
        for bc, umi in shapes:
            recs, b, u = u8(24 * n), u8(bc * n), u8(umi * n)
            ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_DIRTY, 10_000, 1, stream)
            if (bc, umi) == (16, 12):
                report("K1 validate_reduce", 24, lambda: ctx.validate_reduce_async(recs, n, bc, umi, res, stream))
            report(f"K2 unpack bc{bc}/umi{umi}", 24 + bc + umi,
                   lambda: ctx.unpack_async(recs, n, bc, umi, b, u, None, res, stream))
            if (bc, umi) == (16, 12):
                report(f"K2 unpack bc{bc}/umi{umi} noresult", 24 + bc + umi,
                       lambda: ctx.unpack_async(recs, n, bc, umi, b, u, None, None, stream))
                fl = u8(n)
                report(f"K2 unpack bc{bc}/umi{umi} +flags", 25 + bc + umi,
                       lambda: ctx.unpack_async(recs, n, bc, umi, b, u, fl, res, stream))
                del fl
            back = u8(24 * n)
            report(f"K3 pack bc{bc}/umi{umi}", 24 + bc + umi,
                   lambda: ctx.pack_async(b, u, n, bc, umi, back, d_result=res, stream=stream))
            if (bc, umi) == (32, 32):
                idx = torch.arange(n, dtype=torch.int64, device=dev)
                report(f"K3 pack bc{bc}/umi{umi} +index", 32 + bc + umi,
                       lambda: ctx.pack_async(b, u, n, bc, umi, back, d_index=idx, d_result=res, stream=stream))
                del idx
            del recs, b, u, back
        # K4: sorted streaming path (blocking API: wall clock, includes scratch allocation and
        # the table's D2H) and the unsorted sort-then-segment path
        if not args.only or "K4" in args.only:  # noqa: E501
            import time
            recs = u8(24 * n)
            for label, mode, param, m in [("K4 barcode_count sorted (1000/barcode, 5/umi)", ibu.GEN_SORTED, (5 << 32) | 1000, 1),
                                          ("K4 barcode_count unsorted whitelist 1M barcodes", ibu.GEN_WHITELIST, (4096 << 32) | 1_000_000, 0),
                                          ("K4 barcode_count unsorted pattern", ibu.GEN_PATTERN, 0, 0)]:
                ctx.generate_records_async(recs, 0, n, 16, 12, mode, param, 3, stream)
                stream.synchronize()
                ts = []
                for _ in range(5):
                    t0 = time.perf_counter()
                    rows, info = ctx.barcode_count(recs, n, m, stream)
                    ts.append((time.perf_counter() - t0) * 1e3)
                ts.sort()
                print(json.dumps(dict(kernel=label, records=n, alg_bytes_per_record=24, ms_mean=sum(ts) / len(ts),
                                      ms_best=ts[0], achieved_gbs=24 * n / ts[0] / 1e6, frac=24 * n / ts[0] / 1e6 / pk,
                                      rows=len(rows), sorted=info["input_was_sorted"], timing="wall clock of the blocking call")),
                      flush=True)
            del recs
        # torch's own device copy of the same footprint as a same-run peak probe
        a, c = u8(2_600_000_000), u8(2_600_000_000)
        if not args.only:
            mean, best, med = timeit(lambda: c.copy_(a), stream, args.iters)
            print(json.dumps(dict(kernel="torch copy_ 2.6 GB (read+write)", ms_mean=mean, ms_best=best,
                                  achieved_gbs=5.2e9 / mean / 1e6, best_gbs=5.2e9 / best / 1e6)), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
