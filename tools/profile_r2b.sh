#!/bin/bash
# ncu passes for the ordered K4 path (profiles/r2_k4_ordered_*): launch list of the near-distinct table
# call, then one `--set full` capture of its kernels.  The command runs once without ncu first.
set -x
python tools/k4bench.py --iters 2 --only "near-distinct dirty" > gpurun_out/plain_k4ord.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k4ord_r2.csv \
      python tools/k4bench.py --iters 2 --only "near-distinct dirty" > gpurun_out/ncu_list_k4ord.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_part1|k_part2|k_bucket_sort|k_bucket_emit" \
    --launch-skip 4 -c 4 -f -o gpurun_out/prof_k4ord_r2 python tools/k4bench.py --iters 2 --only "near-distinct clean" > gpurun_out/ncu_k4ord_r2.log 2>&1
ls -la gpurun_out/*ord_r2*
