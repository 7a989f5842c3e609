#!/bin/bash
# ncu on the C3 band kernel (selected sections only: the kernel writes 16 GB per launch)
ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section WarpStateStats --section SchedulerStats --section Occupancy \
  --clock-control none -k regex:"band" -s 4 -c 1 python bench.py --workload c3 --steps 1 --warmup 3 --no-cpu --no-e2e --no-graph 2>&1 | grep -vE "^\s*$|==PROF==" | head -120
