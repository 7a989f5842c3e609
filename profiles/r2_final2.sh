#!/bin/bash
# HEAD validation after A/B 9 (reverse walks + evict-first slab gathers became defaults): the GPU test suite, smoke, the
# default bench line, the reference arm, a store-policy A/B of the C2 harness f!, and fresh ncu captures of the changed kernels
set -u
O=gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q --durations=3 --timeout=500 -p no:cacheprovider > $O/r2i_tests.log 2>&1; tail -2 $O/r2i_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2i_smoke.log 2>&1; tail -2 $O/r2i_smoke.log
timeout 600 python bench.py > $O/r2i_bench.json 2> $O/r2i_bench.err; tail -c 300 $O/r2i_bench.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2i_reference.json 2>&1
echo "bench done $(( $(date +%s) - T0 )) s"
B="python bench.py --no-cpu --no-e2e --no-extras"
for rep in a b; do for pl in 0 1; do for fd in forward central; do
  FDBS_TRI_PLAIN=$pl $B --workload c2 --fdtype $fd > $O/r2i_tri_plain${pl}_${fd}_$rep.json 2> $O/r2i_tri_plain${pl}_${fd}_$rep.err
done; done; done
FDBS_TRI_PLAIN=1 timeout 200 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "c2_full_size or cuda_graph_replay" > $O/r2i_tests_plain.log 2>&1; tail -1 $O/r2i_tests_plain.log
for f in $O/r2i_tri_plain*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2i_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done > $O/r2i_tri_plain_summary.txt
cat $O/r2i_tri_plain_summary.txt
echo "ab done $(( $(date +%s) - T0 )) s"
bash profiles/ncu_kernel.sh r2i_full_c2_forward_staged "diff_scatter_staged" 3 --workload c2 > $O/r2i_full_c2_forward_staged.sum 2>&1
bash profiles/ncu_kernel.sh r2i_full_c2_central_staged "diff_scatter_staged" 3 --workload c2 --fdtype central > $O/r2i_full_c2_central_staged.sum 2>&1
bash profiles/ncu_kernel.sh r2i_full_c4_cm_resident "diff_scatter_cm" 3 --workload c4 > $O/r2i_full_c4_cm_resident.sum 2>&1
bash profiles/launches.sh r2i_launches_c2_forward --workload c2 > $O/r2i_launches_c2_forward.txt 2>&1
tail -6 $O/r2i_launches_c2_forward.txt
echo "all done $(( $(date +%s) - T0 )) s"
