#!/bin/bash
# C3 banded: bench line, then DRAM bytes + duration of the band kernels (small metric set: 16 GB per launch)
python bench.py --workload c3 --no-e2e 2>&1 | tail -1 > gpurun_out/r1_bench_c3.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:"diff_scatter_band|diff_slabs" -s 6 -c 2 --csv --log-file gpurun_out/r1_band_c3.csv \
    python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-e2e --no-graph > gpurun_out/r1_band_c3.log 2>&1
tail -4 gpurun_out/r1_band_c3.csv
