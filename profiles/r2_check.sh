#!/bin/bash
# HEAD validation on one GPU: GPU test suite, smoke, the default bench line, the reference arm, and a max_batch A/B of C2
set -u
O=gpurun_out
T0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 --timeout=500 -p no:cacheprovider > $O/r2h_tests.log 2>&1; tail -3 $O/r2h_tests.log
echo "tests done $(( $(date +%s) - T0 )) s"
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2h_smoke.log 2>&1; tail -2 $O/r2h_smoke.log
echo "smoke done $(( $(date +%s) - T0 )) s"
timeout 600 python bench.py > $O/r2h_bench.json 2> $O/r2h_bench.err; tail -c 300 $O/r2h_bench.err
echo "bench done $(( $(date +%s) - T0 )) s"
timeout 300 python bench.py --impl reference --steps 5 --warmup 2 > $O/r2h_reference.json 2>&1
echo "reference done $(( $(date +%s) - T0 )) s"
for mb in 1 4; do
  timeout 200 python bench.py --workload c2 --max-batch $mb --no-cpu --no-e2e --no-extras > $O/r2h_c2_mb$mb.json 2> $O/r2h_c2_mb$mb.err
done
python - <<'PY'
import json
for mb in (1, 4):
    try:
        d = json.loads(open(f"gpurun_out/r2h_c2_mb{mb}.json").read().strip().splitlines()[-1])
        print("max_batch", mb, "ms_per_step", d["ms_per_step"], "value", d["value"])
    except Exception as e:
        print("mb", mb, "failed", e)
PY
echo "all done $(( $(date +%s) - T0 )) s"
