#!/usr/bin/env python
"""stage_sweep.py — staged ingest (ibu_gpu_process_mmap: pread -> pinned slot -> H2D -> K1) of one file
over chunk size x slot count x store kind: does a ring of pinned slots small enough for the last-level
cache (ordinary stores, the DMA engine reading the staged bytes from cache) beat 96 MB chunks filled
with non-temporal stores (3 DRAM transfers per byte)?  One JSON line per setting."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ibu_b200 as ibu  # noqa: E402
from oracle import oracle_c as oc  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--records", type=int, default=100_000_000)
    ap.add_argument("--dir", default="/dev/shm")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n = args.records
    path = os.path.join(args.dir, f"ibu_stage_sweep_{n}.ibu")
    with ibu.Writer(path, ibu.Header(16, 12)) as w:
        step = 16_000_000
        for s in range(0, n, step):
            w.write_batch(oc.generate_records(s, min(step, n - s), 16, 12, 0, 0, 7))
    reader = ibu.MmapReader(path)
    want = None
    for chunk in (1 << 16, 1 << 17, 1 << 18, 1 << 19, 1 << 20, 1 << 22):
        for slots in (3, 6):
            for plain in (0, 1):
                os.environ["IBU_B200_STAGE_PLAIN"] = str(plain)
                ctx = ibu.GpuContext(0, chunk_records=chunk, n_slots=slots)
                best, red = 1e9, None
                for _ in range(args.reps + 1):
                    t0 = time.perf_counter()
                    red = reader.process_gpu(ctx)
                    best = min(best, time.perf_counter() - t0)
                ctx.close()
                want = want or red
                print(json.dumps(dict(chunk_records=chunk, chunk_mb=chunk * 24 / 1e6, n_slots=slots, ring_mb=chunk * 24 * slots / 1e6,
                                      stores="plain" if plain else "non-temporal", sec=best, gb_s=24 * n / best / 1e9,
                                      same_result=bool(red == want))), flush=True)
    os.environ.pop("IBU_B200_STAGE_PLAIN", None)
    reader.close()
    os.unlink(path)


if __name__ == "__main__":
    main()
