#!/bin/bash
# round-2 ncu evidence (one GPU, under gpurun): launch lists + full captures of every kernel bench.py reports a roofline
# for.  Per-launch times are cold-cache and serialised: compare SHARES with bench.py's live numbers, not absolutes.
# (record of the run as it was made: at that commit the side-stream overlap was ON for strategy 2 and FDB_NO_OVERLAP=1
#  switched it off; since then the overlap is off by default and FDB_FORCE_OVERLAP=1 switches it on)
set -u
O=gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,lts__t_sector_hit_rate.pct"
BA="--steps 1 --warmup 3 --no-cpu --no-e2e --no-graph --no-extras --spin 0"
# launch lists
bash profiles/launches.sh r2_launches_c2_forward --workload c2 > $O/r2_launches_c2_forward.txt 2>&1
bash profiles/launches.sh r2_launches_c2_central --workload c2 --fdtype central > $O/r2_launches_c2_central.txt 2>&1
bash profiles/launches.sh r2_launches_c4_fused --workload c4 --strategy 1 > $O/r2_launches_c4_fused.txt 2>&1
FDB_NO_OVERLAP=1 bash profiles/launches.sh r2_launches_c4_lists --workload c4 --strategy 2 > $O/r2_launches_c4_lists.txt 2>&1
# full captures of the scatter kernels
bash profiles/ncu_kernel.sh r2_full_c2_forward_scatter "diff_scatter" 3 --workload c2 > $O/r2_full_c2_forward_scatter.sum 2>&1
bash profiles/ncu_kernel.sh r2_full_c2_central_scatter "diff_scatter" 3 --workload c2 --fdtype central > $O/r2_full_c2_central_scatter.sum 2>&1
FDB_NO_STAGED=1 bash profiles/ncu_kernel.sh r2_full_c2_forward_gather "diff_scatter" 3 --workload c2 > $O/r2_full_c2_forward_gather.sum 2>&1
bash profiles/ncu_kernel.sh r2_full_c4_fused_scatter "diff_scatter" 3 --workload c4 --strategy 1 > $O/r2_full_c4_fused_scatter.sum 2>&1
FDB_NO_OVERLAP=1 bash profiles/ncu_kernel.sh r2_full_c4_cm_scatter "diff_scatter_cm" 70 --workload c4 --strategy 2 > $O/r2_full_c4_cm_scatter.sum 2>&1
bash profiles/ncu_kernel.sh r2_full_c4_eps_lists "color_sumsq_lists" 3 --workload c4 --strategy 1 > $O/r2_full_c4_eps_lists.sum 2>&1
# metric-only captures (big kernels): C3 band fill, C5 dense columns
ncu --metrics $M --clock-control none -k regex:"diff_scatter_band|diff_slabs" -s 6 -c 2 --csv --log-file $O/r2_band_c3.csv python bench.py --workload c3 $BA > $O/r2_band_c3.log 2>&1
ncu --metrics $M --clock-control none -k regex:"diff_columns" -s 20 -c 3 --csv --log-file $O/r2_columns_c5.csv python bench.py --workload c5 $BA > $O/r2_columns_c5.log 2>&1
ls -la $O | grep r2_ | tail -40
