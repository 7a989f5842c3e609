#!/bin/bash
# Full ncu capture of ONE kernel of a bench configuration; prints the details page.
# Usage: bash profiles/ncu_kernel.sh <tag> <kernel-regex> <skip> <bench args...>   -> gpurun_out/<tag>.ncu-rep + .txt
TAG=$1; KRE=$2; SKIP=$3; shift 3
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s $SKIP -c 1 -f -o gpurun_out/${TAG} \
    python bench.py "$@" --steps 1 --warmup 3 --no-cpu --no-e2e --no-graph --spin 0 > gpurun_out/${TAG}.log 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page details > gpurun_out/${TAG}.txt 2>&1
ncu -i gpurun_out/${TAG}.ncu-rep --page source --csv > gpurun_out/${TAG}_source.csv 2>&1
grep -E "Duration|DRAM Throughput|Memory Throughput|Achieved Occupancy|Registers Per|Issue Slots Busy|No Eligible|Stall|Warp Cycles Per Issued|Theoretical Occ" gpurun_out/${TAG}.txt | head -30
