#!/bin/bash
# round-2 A/B, second pass (one GPU): staged-scatter variants, eps pass beside f(x), colour-major lists with all slabs resident.
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_ab2_$tag.json 2> $O/r2_ab2_$tag.err || echo "FAILED $tag" >&2; }
run c2f_default     $B --workload c2 --fdtype forward
run c2f_minb8       FDB_STAGED_MINB=8 $B --workload c2 --fdtype forward
run c2f_stages3     FDB_STAGES=3 $B --workload c2 --fdtype forward
run c2f_noepsov     FDB_NO_EPS_OVERLAP=1 $B --workload c2 --fdtype forward
run c2f_gather      FDB_NO_STAGED=1 $B --workload c2 --fdtype forward
run c2c_default     $B --workload c2 --fdtype central
run c2c_minb8       FDB_STAGED_MINB=8 $B --workload c2 --fdtype central
run c4_fused        $B --workload c4 --strategy 1
run c4_lists_res    $B --workload c4 --strategy 3
run c4_lists        $B --workload c4 --strategy 2
run c4_lists_ov     FDB_FORCE_OVERLAP=1 $B --workload c4 --strategy 2
for f in $O/r2_ab2_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab2_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "launches=%g" % r["scatter_launches_per_jacobian"], "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
