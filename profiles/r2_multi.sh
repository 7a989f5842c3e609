#!/bin/bash
# round-2 multi-GPU measurements: bash profiles/r2_multi.sh N   (under `gpurun --gpus N`; one process per GPU via torchrun)
set -u
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
run() { tag=$1; shift; echo "== $tag" >&2; env "$@" > $O/r2_scale_${N}_$tag.json 2> $O/r2_scale_${N}_$tag.err || { echo "FAILED $tag" >&2; tail -5 $O/r2_scale_${N}_$tag.err >&2; }; }
if [ "$N" = "2" ]; then
  timeout 1200 python -m pytest tests/test_gpu_multi.py tests/test_group.py -m gpu -q --durations=8 --timeout=900 -p no:cacheprovider > $O/r2_multi2_tests.log 2>&1
  tail -15 $O/r2_multi2_tests.log
fi
run c4            $TR bench.py --gpus $N --steps 20 --warmup 5
run c4_overlap    FDB_FORCE_OVERLAP=1 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e
if [ "$N" != "8" ]; then
run c4_ncclbar    $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --barrier nccl
run c4_onelaunch  $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e --strategy 3
fi
if [ "$N" = "4" ]; then
  run c5          $TR bench.py --gpus $N --workload c5 --steps 3 --warmup 3
  run c2cols      $TR bench.py --gpus $N --workload c2 --shard columns --gather none --steps 50
fi
for f in $O/r2_scale_${N}_*.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(sys.argv[1].split("r2_scale_")[1][:-5], "ms/step=%.4f" % d["ms_per_step"], "scatter_ms=%.4f" % r["scatter_ms_per_jacobian"],
          "launches=%g" % r["scatter_launches_per_jacobian"], "strong", d.get("strong_scaling"), "parity", (d.get("parity") or {}).get("ok"),
          (d.get("parity") or {}).get("sharded_equals_unsharded", {}).get("equal"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
