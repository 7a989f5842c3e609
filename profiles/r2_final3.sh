#!/bin/bash
# last GPU call of round 2: the GPU suite against the oracle whose `norm` is now OpenBLAS-faithful, smoke, the default bench
# line (with parity.eps_vs_reference_norm), and a second-box repeat of A/B 9's two kept switches
set -u
O=gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q --durations=3 --timeout=500 -p no:cacheprovider > $O/r2j_tests.log 2>&1; tail -2 $O/r2j_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2j_smoke.log 2>&1; tail -2 $O/r2j_smoke.log
timeout 600 python bench.py > $O/r2j_bench.json 2> $O/r2j_bench.err; tail -c 300 $O/r2j_bench.err
echo "bench done $(( $(date +%s) - T0 )) s"
B="python bench.py --no-cpu --no-e2e --no-extras"
for rep in a b; do for r in 0 3; do
  FDB_REVERSE=$r $B --workload c2 --fdtype forward > $O/r2j_c2f_rev${r}_$rep.json 2> $O/r2j_c2f_rev${r}_$rep.err
done; done
for h in 0 1; do FDB_CM_HINT=$h $B --workload c4 > $O/r2j_c4_cm_hint$h.json 2> $O/r2j_c4_cm_hint$h.err; done
for f in $O/r2j_c2f_rev*.json $O/r2j_c4_cm_hint*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2j_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done > $O/r2j_ab_summary.txt
cat $O/r2j_ab_summary.txt
echo "all done $(( $(date +%s) - T0 )) s"
