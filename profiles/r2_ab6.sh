#!/bin/bash
# round-2 A/B, sixth pass (one GPU): staged scatter handing stages back through an "empty" mbarrier instead of __syncthreads
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_ab6_$tag.json 2> $O/r2_ab6_$tag.err || echo "FAILED $tag" >&2; }
for rep in a b; do
  run c2f_sync_$rep     $B --workload c2 --fdtype forward
  run c2f_empty_$rep    FDB_STAGED_VARIANT=6e $B --workload c2 --fdtype forward
  run c2f_empty3_$rep   FDB_STAGED_VARIANT=6e FDB_STAGES=3 $B --workload c2 --fdtype forward
  run c2c_sync_$rep     $B --workload c2 --fdtype central
  run c2c_empty_$rep    FDB_STAGED_VARIANT=6e $B --workload c2 --fdtype central
done
for f in $O/r2_ab6_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab6_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
FDB_STAGED_VARIANT=6e timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_random_parity.py -m gpu -q --timeout=300 -p no:cacheprovider 2>&1 | tail -2
