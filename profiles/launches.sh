#!/bin/bash
# ncu launch list (device time per launch, cold-cache + serialised: compare SHARES) of one bench configuration.
# Usage: bash profiles/launches.sh <tag> <bench args...>    -> gpurun_out/<tag>.csv
TAG=$1; shift
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"diff_|perturb_|color_sumsq|finalize_eps|gather_fx|zero_slots|replicate_x|set_components|component_eps|k_tridiag|k_ellrows|k_lap5|k_rank1|k_block_sums" -c 600 --csv --log-file gpurun_out/${TAG}.csv \
    python bench.py "$@" --steps 2 --warmup 3 --no-cpu --no-e2e --no-graph --spin 0 > gpurun_out/${TAG}.log 2>&1
python - "$TAG" <<'PY'
import csv, collections, sys
rows = list(csv.reader(l for l in open(f"gpurun_out/{sys.argv[1]}.csv") if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    agg.setdefault(r[ki][:90], []).append(float(r[vi].replace(",", "")))
for k, v in agg.items():
    print(f"{k:92s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us")
PY
