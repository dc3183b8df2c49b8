#!/usr/bin/env python
"""e2e_mmap.py — end-to-end ingest from an mmap'ed .ibu file (configs[3]/[4] shape):
Writer -> file -> MmapReader -> GPU validate/reduce, next to the CPU oracle's
process_parallel restatement on the same file (all host threads).  Prints JSON lines."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu", action="store_true")
    args = ap.parse_args()
    n = args.records
    path = os.path.join(args.dir, f"ibu_e2e_{n}.ibu")
    t0 = time.perf_counter()
    h = ibu.Header(16, 12)
    with ibu.Writer(path, h) as w:  # reference pattern (examples/parallel.rs:65-69), chunked
        step = 16_000_000
        for s in range(0, n, step):
            w.write_batch(oc.generate_records(s, min(step, n - s), 16, 12, 2, 0, 0))
    print(json.dumps(dict(stage="write", records=n, sec=time.perf_counter() - t0)), flush=True)
    reader = ibu.MmapReader(path)
    want = None
    if args.cpu:
        m = oc.MmapReader(path)
        best = 1e9
        for _ in range(2):
            t0 = time.perf_counter()
            want, _ = m.process_parallel_reduce(0)
            best = min(best, time.perf_counter() - t0)
        print(json.dumps(dict(stage="cpu process_parallel (oracle port)", cores=oc.num_cpus(), sec=best,
                              grec_s=n / best / 1e9, gb_s=24 * n / best / 1e9)), flush=True)
    for label, kw, pin in [("staged chunk4Mi x3 slots", dict(chunk_records=4 << 20, n_slots=3), False),
                           ("staged chunk1Mi x4 slots", dict(chunk_records=1 << 20, n_slots=4), False),
                           ("staged chunk4Mi x3 slots 16 copy threads", dict(chunk_records=4 << 20, n_slots=3, copy_threads=16), False),
                           ("staged chunk4Mi x3 slots 4 copy threads", dict(chunk_records=4 << 20, n_slots=3, copy_threads=4), False),
                           ("pinned mmap (cudaHostRegister) chunk4Mi x3", dict(chunk_records=4 << 20, n_slots=3), True)]:
        ctx = ibu.GpuContext(0, **kw)
        try:
            if pin:
                t0 = time.perf_counter()
                reader.pin()
                print(json.dumps(dict(stage="cudaHostRegister(mmap)", sec=time.perf_counter() - t0)), flush=True)
            best = 1e9
            for _ in range(args.reps):
                t0 = time.perf_counter()
                got = reader.process_gpu(ctx)
                best = min(best, time.perf_counter() - t0)
            ok = (got == want) if want is not None else (got["n_records"] == n)
            print(json.dumps(dict(stage="gpu process_mmap: " + label, sec=best, grec_s=n / best / 1e9,
                                  gb_s=24 * n / best / 1e9, parity=bool(ok))), flush=True)
        except ibu.IbuError as e:
            print(json.dumps(dict(stage=label, error=str(e))), flush=True)
        finally:
            if pin:
                reader.unpin()
            ctx.close()
    # load_to_device (device path of load_to_vec) vs load_to_vec
    ctx = ibu.GpuContext(0)
    t0 = time.perf_counter()
    hd, d = ibu.load_to_device(ctx, path)
    t_dev_first = time.perf_counter() - t0  # fresh context: allocates the pinned staging slots
    d.free()
    t0 = time.perf_counter()
    hd, d = ibu.load_to_device(ctx, path)
    t_dev = time.perf_counter() - t0
    for rep in range(2):  # configs[3]: per-barcode record / distinct-UMI table of the loaded file
        t0 = time.perf_counter()
        rows, info = ctx.barcode_count(d, len(d))
        t_tab = time.perf_counter() - t0
    print(json.dumps(dict(stage="gpu barcode_count of the loaded file", sec=t_tab, grec_s=n / t_tab / 1e9,
                          rows=len(rows), pairs=info["n_distinct_pairs"], sorted=info["input_was_sorted"],
                          closed_form_ok=bool(len(rows) == min(n, 1_000_000) and int(rows["n_records"].sum()) == n
                                              and bool((rows["n_distinct_umi"] == 1).all())))), flush=True)
    d.free()
    t_vec = float("nan")
    if n <= 200_000_000:
        t0 = time.perf_counter()
        hv, v = ibu.load_to_vec(path)
        t_vec = time.perf_counter() - t0
    print(json.dumps(dict(stage="load_to_device vs load_to_vec", dev_first_call_sec=t_dev_first, dev_sec=t_dev,
                          dev_gb_s=24 * n / t_dev / 1e9,
                          vec_sec=t_vec, vec_gb_s=24 * n / t_vec / 1e9)), flush=True)
    ctx.close()
    os.unlink(path)


if __name__ == "__main__":
    main()
