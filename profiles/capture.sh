#!/bin/bash
# Profiling session used for profiles/ (run under gpurun on one B200).  Usage: bash profiles/capture.sh <tag>
# Produces, under gpurun_out/: the ncu launch list and full captures the summaries in profiles/ are made from.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
# every launch of one C2 forward bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 160 --csv --log-file $OUT/${TAG}_launches_c2_forward.csv \
    python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $OUT/${TAG}_launches_c2_forward.log 2>&1
# full capture of the library kernels of one C2 forward Jacobian
ncu --set full --clock-control none --import-source on -k regex:"diff_scatter|perturb_colors|color_sumsq" -s 6 -c 3 \
    -o $OUT/${TAG}_full_c2_forward python bench.py --workload c2 --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $OUT/${TAG}_full_c2_forward.log 2>&1
# central: the scatter kernel
ncu --set full --clock-control none --import-source on -k regex:"diff_scatter" -s 3 -c 1 \
    -o $OUT/${TAG}_full_c2_central python bench.py --workload c2 --fdtype central --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph > $OUT/${TAG}_full_c2_central.log 2>&1
# C3 banded: DRAM bytes + duration of the band kernel (16 GB per launch: a small metric set keeps the replay short)
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"diff_scatter_band" -s 3 -c 1 --csv --log-file $OUT/${TAG}_band_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-e2e --no-graph > $OUT/${TAG}_band_c3.log 2>&1
ls -la $OUT | tail -12
