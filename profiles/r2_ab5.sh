#!/bin/bash
# round-2 A/B, fifth pass (one GPU): colour packed into the 16-bit row offset of the staged scatter (no colour stream)
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_ab5_$tag.json 2> $O/r2_ab5_$tag.err || echo "FAILED $tag" >&2; }
for rep in a b; do
  run c2f_packed_$rep   $B --workload c2 --fdtype forward
  run c2f_nopack_$rep   FDB_NO_PACK=1 $B --workload c2 --fdtype forward
  run c2c_packed_$rep   $B --workload c2 --fdtype central
  run c2c_nopack_$rep   FDB_NO_PACK=1 $B --workload c2 --fdtype central
done
for f in $O/r2_ab5_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab5_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
timeout 600 python -m pytest tests -m gpu -q --timeout=500 -p no:cacheprovider 2>&1 | tail -3
