#!/bin/bash
# round-2 A/B measurements (one GPU): staged vs gather scatter (C2 fwd/central), eps lists vs windows and overlap on/off (C4),
# C5 alone; then ncu launch lists + full captures of the kernels bench.py reports on.  Run under gpurun from the repo root.
# (record of the run as it was made: at that commit the side-stream overlap was ON for strategy 2 and FDB_NO_OVERLAP=1
#  switched it off; since then the overlap is off by default and FDB_FORCE_OVERLAP=1 switches it on)
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_ab_$tag.json 2> $O/r2_ab_$tag.err || echo "FAILED $tag" >&2; }
run c2f_staged      $B --workload c2 --fdtype forward
run c2f_gather      FDB_NO_STAGED=1 $B --workload c2 --fdtype forward
run c2c_staged      $B --workload c2 --fdtype central
run c2c_gather      FDB_NO_STAGED=1 $B --workload c2 --fdtype central
run c4_fused        $B --workload c4 --strategy 1
run c4_fused_win    FDB_NO_EPS_LISTS=1 $B --workload c4 --strategy 1
run c4_fused_nostream FDB_HI_STREAM=0 $B --workload c4 --strategy 1
run c4_fused_persist FDB_L2_PERSIST_FX=1 $B --workload c4 --strategy 1 --no-graph
run c4_fused_nograph $B --workload c4 --strategy 1 --no-graph
run c4_lists        $B --workload c4 --strategy 2
run c4_lists_noov   FDB_NO_OVERLAP=1 $B --workload c4 --strategy 2
run c5              $B --workload c5
run c3              $B --workload c3
for f in $O/r2_ab_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "launches=%g" % r["scatter_launches_per_jacobian"], "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
