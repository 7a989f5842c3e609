#!/bin/bash
# round-2 A/B 9 (one GPU, same box): (1) walk direction of the perturbation pass and of the staged scatter (FDB_REVERSE:
# the tail of what the previous kernel streamed is still in L2), (2) L2 eviction hints of the random-pattern gathers
# (FDB_CM_HINT, FDB_HI_STREAM=2: f(x) evict-last, slab values evict-first), (3) the same hint in the C4 harness f!.
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; env "$@" > $O/r2_ab9_$tag.json 2> $O/r2_ab9_$tag.err || echo "FAILED $tag" >&2; }
T0=$(date +%s)
# bit-exactness of the new switches first (oracle compare at full size)
FDB_REVERSE=3 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider -k "c2_full_size or cuda_graph_replay" > $O/r2_ab9_tests_reverse.log 2>&1; tail -1 $O/r2_ab9_tests_reverse.log
FDB_CM_HINT=3 FDB_HI_STREAM=2 FDBS_ELL_HINT=1 timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -x -p no:cacheprovider -k "c4_full_size" > $O/r2_ab9_tests_hint.log 2>&1; tail -1 $O/r2_ab9_tests_hint.log
FDB_CM_HINT=2 timeout 300 python -m pytest tests/test_gpu_fullsize.py -q -x -p no:cacheprovider -k "c4_full_size" > $O/r2_ab9_tests_hint2.log 2>&1; tail -1 $O/r2_ab9_tests_hint2.log
echo "tests $(( $(date +%s) - T0 )) s"
for rep in a b; do
  for r in 0 1 2 3; do
    run c2f_rev${r}_$rep FDB_REVERSE=$r $B --workload c2 --fdtype forward
  done
done
for r in 0 1 2 3; do run c2c_rev${r}_a FDB_REVERSE=$r $B --workload c2 --fdtype central; done
echo "c2 $(( $(date +%s) - T0 )) s"
for h in 0 1 2 3; do run c4_cm_hint${h}_a FDB_CM_HINT=$h $B --workload c4; done
run c4_cm_hint0_b FDB_CM_HINT=0 $B --workload c4
run c4_cm_fxcm_a FDB_FORCE_FX_CM=1 $B --workload c4
run c4_fused_hs1_a FDB_HI_STREAM=1 $B --workload c4 --strategy 1
run c4_fused_hs2_a FDB_HI_STREAM=2 $B --workload c4 --strategy 1
run c4_cm_hint2_ell_a FDB_CM_HINT=2 FDBS_ELL_HINT=1 $B --workload c4
run c4_cm_hint0_ell_a FDB_CM_HINT=0 FDBS_ELL_HINT=1 $B --workload c4
echo "c4 $(( $(date +%s) - T0 )) s"
for f in $O/r2_ab9_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab9_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done > $O/r2_ab9_summary.txt
cat $O/r2_ab9_summary.txt
