#!/bin/bash
# round-2 final ncu evidence (one GPU, under gpurun), same commit as the final bench line: launch lists restricted to the
# library + harness kernels, full captures of the shipped scatter kernels, metric captures of the big ones.
set -u
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct"
BA="--steps 1 --warmup 3 --no-cpu --no-e2e --no-graph --no-extras --spin 0"
bash profiles/launches.sh r2f_launches_c2_forward --workload c2 > $O/r2f_launches_c2_forward.txt 2>&1
bash profiles/launches.sh r2f_launches_c2_central --workload c2 --fdtype central > $O/r2f_launches_c2_central.txt 2>&1
bash profiles/launches.sh r2f_launches_c4 --workload c4 > $O/r2f_launches_c4.txt 2>&1
bash profiles/launches.sh r2f_launches_c3 --workload c3 > $O/r2f_launches_c3.txt 2>&1
bash profiles/ncu_kernel.sh r2f_full_c2_forward_staged "diff_scatter_staged" 3 --workload c2 > $O/r2f_full_c2_forward_staged.sum 2>&1
bash profiles/ncu_kernel.sh r2f_full_c2_central_staged "diff_scatter_staged" 3 --workload c2 --fdtype central > $O/r2f_full_c2_central_staged.sum 2>&1
bash profiles/ncu_kernel.sh r2f_full_c4_cm_resident "diff_scatter_cm" 3 --workload c4 > $O/r2f_full_c4_cm_resident.sum 2>&1
ncu --metrics $M --clock-control none -k regex:"diff_columns" -s 20 -c 2 --csv --log-file $O/r2f_columns_c5.csv python bench.py --workload c5 $BA > $O/r2f_columns_c5.log 2>&1
ls -la $O | grep r2f_ | tail -30
