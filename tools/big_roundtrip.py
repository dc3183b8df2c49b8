#!/usr/bin/env python
"""big_roundtrip.py — size-independent parity properties at 10^9 records (> 2^32-byte offsets):
generate -> K1, K2 unpack -> K3 pack -> K1 must equal the clean twin; sort -> K4 streaming table
must equal the hash/sort table of the unsorted input.  Prints JSON lines with timings."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ibu_b200 as ibu  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
bc, umi = 16, 12
ctx = ibu.GpuContext(0)
recs, b, u, back, res = (ctx.malloc(s) for s in (24 * n, bc * n, umi * n, 24 * n, 64))


def timed(label, fn):
    ctx.synchronize()
    t0 = time.perf_counter()
    out = fn()
    ctx.synchronize()
    dt = time.perf_counter() - t0
    print(json.dumps(dict(step=label, ms=dt * 1e3, grec_s=n / dt / 1e9)), flush=True)
    return out


ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_DIRTY, 10_000, 7)
timed("K1 dirty", lambda: ctx.validate_reduce_async(recs, n, bc, umi, res))
r0 = ctx.read_result(res)
timed("K2 unpack", lambda: ctx.unpack_async(recs, n, bc, umi, b, u, None, res))
r1 = ctx.read_result(res)
assert all(r0[k] == r1[k] for k in ("n_records", "n_bad_barcode", "n_bad_umi", "n_bad_records")), (r0, r1)
assert 0.9 * n / 100 < r0["n_bad_records"] < 1.1 * n / 100
timed("K3 pack", lambda: ctx.pack_async(b, u, n, bc, umi, back, d_result=res))
assert ctx.read_result(res)["n_bad_records"] == 0
ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_CLEAN, 0, 7)
ctx.validate_reduce_async(recs, n, bc, umi, res)
ctx.synchronize()
want = ctx.read_result(res)
ctx.validate_reduce_async(back, n, bc, umi, res)
ctx.synchronize()
got = ctx.read_result(res)
assert got == want and got["sum_index"] == (n * (n - 1) // 2) % 2**64, (got, want)
print(json.dumps(dict(step="roundtrip parity", ok=True, n=n)), flush=True)
ctx.free(b), ctx.free(u)

# whitelist data: unsorted table (hash path) vs sort + streaming table
ctx.generate_records_async(recs, 0, n, bc, umi, ibu.GEN_WHITELIST, (64 << 32) | 100_000, 9)
rows_a, info_a = timed("K4 unsorted (hash path)", lambda: ctx.barcode_count(recs, n))
timed("sort_records", lambda: ctx.sort_records(recs, n, back))
rows_b, info_b = timed("K4 sorted stream", lambda: ctx.barcode_count(back, n, 1))
assert info_b["input_was_sorted"] and not info_a["input_was_sorted"]
assert np.array_equal(rows_a, rows_b) and int(rows_a["n_records"].sum()) == n
print(json.dumps(dict(step="table parity", ok=True, rows=len(rows_a), pairs=info_a["n_distinct_pairs"])), flush=True)
