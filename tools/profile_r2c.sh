#!/bin/bash
# ncu pass for the partition sort (profiles/r2_sort_by_partition_*): one `--set full` capture of its three kernels
# inside ibu_gpu_sort_records, after the same command has run once without ncu.
set -x
python tools/sortprobe.py --orders asc --iters 2 > gpurun_out/plain_sortmsd.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"k_part1|k_part2|k_bucket_sort_records" \
    --launch-skip 3 -c 3 -f -o gpurun_out/prof_sortmsd_r2 python tools/sortprobe.py --orders asc --iters 2 > gpurun_out/ncu_sortmsd_r2.log 2>&1
ncu -i gpurun_out/prof_sortmsd_r2.ncu-rep --page raw --csv > gpurun_out/prof_sortmsd_r2_raw.csv 2>/dev/null
ls -la gpurun_out/*sortmsd*
