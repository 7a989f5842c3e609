#!/bin/bash
# C2 strong scaling with contiguous column blocks (slice-aware f!): J left column-sharded, and gathered on rank 0.
# Usage: bash profiles/colshard.sh <ngpus>
N=$1
for g in ${GATHERS:-none root}; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py \
      --gpus $N --workload c2 --shard columns --gather $g --steps 100 --warmup 5 2>&1 | tail -1 > gpurun_out/colshard_${N}_$g.json
  python - $N $g <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/colshard_{sys.argv[1]}_{sys.argv[2]}.json"))
    print("gpus", sys.argv[1], "gather", sys.argv[2], "ms/step", round(d["ms_per_step"], 4), "nnz/s %.3e" % d["value"])
except Exception as e:
    print("FAILED", sys.argv[1:], e, open(f"gpurun_out/colshard_{sys.argv[1]}_{sys.argv[2]}.json").read()[-1500:])
PY
done
