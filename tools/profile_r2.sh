#!/bin/bash
# The ncu passes behind profiles/r2_*: launch lists (gpu__time_duration only) of the bench and of
# the K4 table call, then one `--set full` capture per kernel of interest.  Each command runs once
# without ncu first.  Run under gpurun from the repository root; outputs land in gpurun_out/.
set -x
python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/plain_r2.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r2.csv \
      python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/ncu_list_r2.log 2>&1
python tools/k4bench.py --iters 2 --only "10x-like uniform" > gpurun_out/plain_k4.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_k4_r2.csv \
      python tools/k4bench.py --iters 2 --only "10x-like uniform" > gpurun_out/ncu_list_k4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_part1|k_part2|k_bucket_dedup2|k_table_rows" \
    --launch-skip 4 -c 4 -o gpurun_out/prof_k4_r2 python tools/k4bench.py --iters 2 --only "10x-like uniform" > gpurun_out/ncu_k4_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_unpack" --launch-skip 3 -c 1 \
    -o gpurun_out/prof_k2_r2 python bench.py --steps 2 --warmup 3 --quick --no-cpu > gpurun_out/ncu_k2_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_pack_rt" --launch-skip 2 -c 1 \
    -o gpurun_out/prof_k3rt_r2 python tools/kbench.py --only "bc15/umi9" > gpurun_out/ncu_k3rt_r2.log 2>&1
ls -la gpurun_out/*_r2.ncu-rep
