#!/bin/bash
# round-2 A/B, fourth pass (one GPU): f(x) gathered once into colour-major order (fx_cm) for the list scatters; staged
# scatter with 5 / 4 resident blocks; then the launch lists of the final kernels.
set -u
O=gpurun_out
B="python bench.py --no-cpu --no-e2e --no-extras"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_ab4_$tag.json 2> $O/r2_ab4_$tag.err || echo "FAILED $tag" >&2; }
for rep in a b; do
  run c4_lists_fxcm_$rep    $B --workload c4
  run c4_lists_nofxcm_$rep  FDB_NO_FX_CM=1 $B --workload c4
  run c4_percolor_fxcm_$rep   $B --workload c4 --strategy 2
  run c4_percolor_nofxcm_$rep FDB_NO_FX_CM=1 $B --workload c4 --strategy 2
  run c2f_6n_$rep FDB_STAGED_VARIANT=6n $B --workload c2
  run c2f_5n_$rep FDB_STAGED_VARIANT=5n $B --workload c2
  run c2f_4n_$rep FDB_STAGED_VARIANT=4n $B --workload c2
done
for f in $O/r2_ab4_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_ab4_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "launches=%g" % r["scatter_launches_per_jacobian"], "frac=%.3f" % (r["frac"] or 0), "parity", d["parity"]["ok"], "clk", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
bash profiles/launches.sh r2f_launches_c2_forward --workload c2 > $O/r2f_launches_c2_forward.txt 2>&1
bash profiles/launches.sh r2f_launches_c2_central --workload c2 --fdtype central > $O/r2f_launches_c2_central.txt 2>&1
bash profiles/launches.sh r2f_launches_c4 --workload c4 > $O/r2f_launches_c4.txt 2>&1
bash profiles/launches.sh r2f_launches_c3 --workload c3 > $O/r2f_launches_c3.txt 2>&1
timeout 300 python -m pytest tests/test_gpu_fullsize.py tests/test_group.py tests/test_gpu_parity.py -m gpu -q -k "not barrier" -p no:cacheprovider 2>&1 | tail -3
