#!/usr/bin/env python
"""e2e_host.py — end-to-end (host buffer -> host buffer) rates of the three C-ABI host paths
with pinned memory, next to the CPU oracle on the same box: configs[1] unpack, configs[2] pack
(bc32/umi32), and the process_parallel counterpart.  JSON lines."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

import ibu_b200 as ibu  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
ctx = ibu.GpuContext(0)


def pinned(dtype, shape):
    buf = ibu.PinnedBuffer(int(np.prod(shape)) * np.dtype(dtype).itemsize)
    return buf, buf.array(dtype, shape)


def best(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


keep = []
for bc, umi in [(16, 12), (32, 32)]:
    pr, recs = pinned(ibu.RECORD_DTYPE, (n,))
    pb, b = pinned(np.uint8, (n, bc))
    pu, u = pinned(np.uint8, (n, umi))
    oc.generate_records(0, n, bc, umi, 1, 10_000, 3, 0, out=recs)
    t = best(lambda: ctx.unpack_host(recs, bc, umi, b, u))
    print(json.dumps(dict(path=f"unpack_host bc{bc}/umi{umi}", sec=t, grec_s=n / t / 1e9,
                          link_gbs=n * (24 + bc + umi) / t / 1e9)), flush=True)
    tc = best(lambda: oc.unpack_records(recs, bc, umi, 0, b, u), reps=2)
    print(json.dumps(dict(path=f"cpu oracle unpack bc{bc}/umi{umi}", cores=oc.num_cpus(), sec=tc, grec_s=n / tc / 1e9)), flush=True)
    out_p, out = pinned(ibu.RECORD_DTYPE, (n,))
    t = best(lambda: ctx.pack_host(b, u, index_base=0, out=out))
    print(json.dumps(dict(path=f"pack_host bc{bc}/umi{umi}", sec=t, grec_s=n / t / 1e9,
                          link_gbs=n * (24 + bc + umi) / t / 1e9)), flush=True)
    tc = best(lambda: oc.pack_records(b, u, None, 0, 0, out), reps=2)
    print(json.dumps(dict(path=f"cpu oracle pack bc{bc}/umi{umi}", cores=oc.num_cpus(), sec=tc, grec_s=n / tc / 1e9)), flush=True)
    if bc == 16:
        t = best(lambda: ctx.process_host(recs, bc, umi))
        print(json.dumps(dict(path="process_host (validate/reduce)", sec=t, grec_s=n / t / 1e9, link_gbs=n * 24 / t / 1e9)), flush=True)
        tc = best(lambda: oc.reduce_records(recs, bc, umi, 0), reps=2)
        print(json.dumps(dict(path="cpu oracle process_parallel reduce", cores=oc.num_cpus(), sec=tc, grec_s=n / tc / 1e9)), flush=True)
    del recs, b, u, out
    for p in (pr, pb, pu, out_p):
        p.free()
