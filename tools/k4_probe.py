#!/usr/bin/env python
"""k4_probe.py N MODE GEN [PARAM] — time ibu_gpu_barcode_count on N generated records
(IBU_B200_TRACE=1 prints the phases and the k_segments kernel time).  Tuning tool."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ibu_b200 as ibu
n = int(sys.argv[1]); mode = int(sys.argv[2]); gen = int(sys.argv[3])
ctx = ibu.GpuContext(0)
d = ctx.malloc(24 * n)
ctx.generate_records_async(d, 0, n, 16, 12, gen, int(sys.argv[4]) if len(sys.argv) > 4 else 0, 0)
ctx.synchronize()
for _ in range(3):
    t0 = time.time()
    rows, info = ctx.barcode_count(d, n, mode)
print("n", n, "mode", mode, "gen", gen, "rows", len(rows), info, "sec %.3f" % (time.time() - t0), flush=True)
