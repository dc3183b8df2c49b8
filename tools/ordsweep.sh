# the ordered K4 path (k4_ordered.cuh) stage by stage: device ms per call of the near-distinct cases
IBU_B200_TRACE=1 timeout 200 python tools/k4bench.py --only "near-distinct" --iters 5 2>&1 | grep "k_bucket\|k_part\|wide list\|partition path total"
