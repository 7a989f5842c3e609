#!/bin/bash
# final-commit validation on one GPU: the GPU test suite, the default bench line, the reference arm, smoke, and the ncu
# captures the traffic.json entries of the graded kernel point at
set -u
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -q --durations=3 --timeout=500 -p no:cacheprovider > $O/r2_final_tests.log 2>&1; tail -3 $O/r2_final_tests.log
timeout 600 python bench.py > $O/r2_final_bench.json 2> $O/r2_final_bench.err; tail -c 200 $O/r2_final_bench.err
timeout 400 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2_final_reference.json 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_final_smoke.log 2>&1; tail -2 $O/r2_final_smoke.log
bash profiles/ncu_kernel.sh r2g_full_c2_forward_staged "diff_scatter_staged" 3 --workload c2 > $O/r2g_full_c2_forward_staged.sum 2>&1
bash profiles/ncu_kernel.sh r2g_full_c2_central_staged "diff_scatter_staged" 3 --workload c2 --fdtype central > $O/r2g_full_c2_central_staged.sum 2>&1
bash profiles/launches.sh r2g_launches_c2_forward --workload c2 > $O/r2g_launches_c2_forward.txt 2>&1
tail -5 $O/r2g_launches_c2_forward.txt
