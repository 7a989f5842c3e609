#!/bin/bash
# round-2 A/B, third pass (one GPU, every variant twice, interleaved): staged-scatter variants (resident blocks x index
# prefetch) and the colour-major single launch with / without the L2 prefetch of the next colour's slab.
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_ab3_$tag.json 2> $O/r2_ab3_$tag.err || echo "FAILED $tag" >&2; }
for rep in a b; do
  for v in 8n 6p 6n; do
    run c2f_${v}_$rep FDB_STAGED_VARIANT=$v $B --workload c2 --fdtype forward
    run c2c_${v}_$rep FDB_STAGED_VARIANT=$v $B --workload c2 --fdtype central
  done
  run c2f_gather_$rep FDB_NO_STAGED=1 $B --workload c2 --fdtype forward
  run c4_fused_$rep        $B --workload c4 --strategy 1
  run c4_listsres_pf_$rep  $B --workload c4 --strategy 3
  run c4_listsres_nopf_$rep FDB_CM_PREFETCH=0 $B --workload c4 --strategy 3
done
for f in $O/r2_ab3_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab3_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
