#!/usr/bin/env python
"""sortprobe.py — ibu_gpu_sort_records call by call (wall clock and CUDA events) for records whose
index word is ascending / descending / in random order.  Tuning tool; IBU_B200_SWEEP=1|2 picks the
form of the one-sweep pass, IBU_B200_TRACE_ALLOC=1 prints slow allocations."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ibu_b200 as ibu  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--iters", type=int, default=4)
    ap.add_argument("--orders", default="asc,desc,random")
    ap.add_argument("--check", action="store_true", help="verify the order and the column sums with torch")
    ap.add_argument("--gen", default="clean", choices=["clean", "zipf", "whitelist"],
                    help="random barcodes / Zipf over 10^6 barcodes (hot barcodes: the partition declines) / 10x-like whitelist")
    args = ap.parse_args()
    n = args.records
    dev = torch.device("cuda", 0)
    ctx = ibu.GpuContext(0)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        recs = torch.empty(24 * n, dtype=torch.uint8, device=dev)
        back = torch.empty(24 * n, dtype=torch.uint8, device=dev)
        for order in args.orders.split(","):
            gen, param = {"clean": (ibu.GEN_CLEAN, 0), "zipf": (ibu.GEN_ZIPF, (4096 << 32) | 1_000_000),
                          "whitelist": (ibu.GEN_WHITELIST, (20 << 32) | 1_000_000)}[args.gen]
            ctx.generate_records_async(recs, 0, n, 16, 12, gen, param, 5, stream)
            words = recs.view(torch.int64).view(-1, 3)
            if order == "desc":
                words[:, 2] = (n - 1) - words[:, 2]
            elif order == "random":
                words[:, 2] = torch.randperm(n, device=dev)
            stream.synchronize()
            wall, evms = [], []
            for _ in range(args.iters):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                t0 = time.perf_counter()
                ctx.sort_records(recs, n, back, stream)
                wall.append((time.perf_counter() - t0) * 1e3)
                b.record(stream)
                stream.synchronize()
                evms.append(a.elapsed_time(b))
            line = dict(order=order, records=n, gen=args.gen, msd=os.environ.get("IBU_B200_SORT_MSD", "1"),
                        wall_ms=[round(x, 3) for x in wall], event_ms=[round(x, 3) for x in evms])
            if args.check:
                out = back.view(torch.int64).view(-1, 3)
                # Record's Ord: (barcode, umi, index) ascending, compared as unsigned; words here are < 2^63
                k = out[1:] - out[:-1]
                ok = ((k[:, 0] > 0) | ((k[:, 0] == 0) & ((k[:, 1] > 0) | ((k[:, 1] == 0) & (k[:, 2] >= 0))))).all().item()
                line["sorted_ok"] = bool(ok)
                line["sum_ok"] = bool((out.sum(0) == words.sum(0)).all().item())
                del out, k
            print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
