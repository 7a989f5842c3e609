#!/bin/bash
# final-commit confirmation on N GPUs: the 2-GPU tests (N=2) and ONE default `bench.py --gpus N` line, launched as the driver does
set -u
N=${1:-2}
O=gpurun_out
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_group.py -m gpu -q --timeout=600 -p no:cacheprovider > $O/r2q_multi2_tests.log 2>&1
  tail -3 $O/r2q_multi2_tests.log
fi
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2q_scale_$N.json 2> $O/r2q_scale_$N.err || tail -5 $O/r2q_scale_$N.err
python - $O/r2q_scale_$N.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r = d["roofline"]
print("N=%d ms/step=%.4f scatter_ms=%.4f launches=%g" % (d["n_gpus"], d["ms_per_step"], r["scatter_ms_per_jacobian"], r["scatter_launches_per_jacobian"]),
      "strong", d.get("strong_scaling"), "parity", d["parity"]["ok"], d["parity"].get("sharded_equals_unsharded", {}).get("equal"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
PY
