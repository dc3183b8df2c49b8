#!/usr/bin/env python
"""e2e_table.py — ingest -> validate -> per-barcode table from an mmap'ed .ibu file in ONE call
(ibu_gpu_process_mmap_ops, IBU_OP_TABLE), next to the plain ingest (K1 only) of the same file and
the H2D link rate: the table work must hide behind the link.  Prints JSON lines."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import ibu_b200 as ibu  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402


def best_of(fn, reps):
    best, out = 1e9, None
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--gen", type=int, default=5, help="generator: 5 zipf whitelist, 3 uniform whitelist, 2 example pattern")
    ap.add_argument("--param", type=int, default=(4096 << 32) | 1_000_000)
    ap.add_argument("--check", type=int, default=1)
    args = ap.parse_args()
    n = args.records
    path = os.path.join(args.dir, f"ibu_e2e_table_{n}.ibu")
    with ibu.Writer(path, ibu.Header(16, 12)) as w:
        step = 16_000_000
        for s in range(0, n, step):
            w.write_batch(oc.generate_records(s, min(step, n - s), 16, 12, args.gen, args.param, 7))
    reader = ibu.MmapReader(path)
    ctx = ibu.GpuContext(0)  # library defaults: 1 Mi-record chunks, 3 slots
    # link probe: pinned host -> device, what a plain copy of the file's bytes gets
    pin = ibu.PinnedBuffer(1 << 30)
    d = ctx.malloc(1 << 30)
    h = pin.array(np.uint8)
    link_s, _ = best_of(lambda: ctx.h2d(d, h), 3)
    ctx.free(d)
    del h
    pin.free()
    link = (1 << 30) / link_s / 1e9
    print(json.dumps(dict(stage="pinned H2D probe (1 GiB)", gb_s=link)), flush=True)

    def line(label, sec, extra):
        print(json.dumps(dict(stage=label, records=n, sec=sec, grec_s=n / sec / 1e9, gb_s=24 * n / sec / 1e9,
                              frac_of_link=24 * n / sec / 1e9 / link, **extra)), flush=True)

    t_red, red = best_of(lambda: reader.process_gpu(ctx), args.reps)
    line("staged: ingest + validate/reduce (K1)", t_red, {})
    def table_call():  # the C call alone: the rows stay on the device
        red, out = reader.process_gpu_ops(ctx, table=True, rows_on_device=True)
        ctx.table_free(out.table)
        return red, out

    t_call, _ = best_of(table_call, args.reps)
    line("staged: ingest + validate/reduce + per-barcode table (rows left on the device)", t_call, dict(table_overhead_ms=(t_call - t_red) * 1e3))
    t_tab, (red2, out) = best_of(lambda: reader.process_gpu_ops(ctx, table=True), args.reps)
    line("staged: ingest + validate/reduce + per-barcode table", t_tab,
         dict(rows=len(out.rows), pairs=out.table_info["n_distinct_pairs"], table_overhead_ms=(t_tab - t_red) * 1e3,
              reduce_equal=bool(red2 == red)))
    rows = out.rows
    try:
        reader.pin()
        t_red_p, _ = best_of(lambda: reader.process_gpu(ctx), args.reps)
        line("pinned mapping: ingest + validate/reduce (K1)", t_red_p, {})
        t_call_p, _ = best_of(table_call, args.reps)
        line("pinned mapping: ingest + validate/reduce + per-barcode table (rows left on the device)", t_call_p,
             dict(table_overhead_ms=(t_call_p - t_red_p) * 1e3))
        t_tab_p, (_, out) = best_of(lambda: reader.process_gpu_ops(ctx, table=True), args.reps)
        line("pinned mapping: ingest + validate/reduce + per-barcode table", t_tab_p,
             dict(rows=len(out.rows), table_overhead_ms=(t_tab_p - t_red_p) * 1e3, same_rows=bool(np.array_equal(out.rows, rows))))
        reader.unpin()
    except ibu.IbuError as e:
        print(json.dumps(dict(stage="pinned mapping", error=str(e))), flush=True)
    if args.check:  # the oracle's table of the same file (slow: numpy sort), on at most 2 x 10^7 records
        m = min(n, 20_000_000)
        recs = reader.slice(0, m)
        t0 = time.perf_counter()
        want, pairs = oc.barcode_table(np.array(recs))
        t_cpu = time.perf_counter() - t0
        _, sub = reader.process_gpu_ops(ctx, 0, m, table=True)
        print(json.dumps(dict(stage="parity vs oracle table", records=m, equal=bool(np.array_equal(sub.rows, want)),
                              cpu_oracle_sec=t_cpu, cpu_grec_s=m / t_cpu / 1e9)), flush=True)
    ctx.close()
    os.unlink(path)


if __name__ == "__main__":
    main()
