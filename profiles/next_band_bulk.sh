#!/bin/bash
# Round-2 lead: re-measure the TMA forms of the band fill with oversubscribed grids.
#   git apply profiles/r1_band_bulk_experiments.patch && python finitediff.jl_b200/build.py
#   gpurun --timeout 900 -- 'bash profiles/next_band_bulk.sh'
# FDB_TUNE_BANDBULK: 0 register flat stream (shipped), 1 chunked bulk copies, 2 row-stationary bulk copies;
# FDB_TUNE_BULKGRID: grid multiplier of the bulk kernels; FDB_TUNE_ROWS_T: rows per staged segment (form 2).
run() {
  timeout 120 python bench.py --workload c3 --no-cpu --no-e2e --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/c3_ab.json
  python - "$1" <<'PY'
import json, sys
try:
    d = json.load(open("gpurun_out/c3_ab.json")); r = d["roofline"]
    print(sys.argv[1], "ms/step", round(d["ms_per_step"], 4), "scatter ms", round(r["scatter_ms_per_jacobian"], 4))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
FDB_TUNE_BANDBULK=1 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band or lap5" 2>&1 | tail -1
FDB_TUNE_BANDBULK=2 timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "band or lap5" 2>&1 | tail -1
export FDB_TUNE_BANDBULK=0; run "shipped flat stream"
for g in 1 4 16; do
  export FDB_TUNE_BANDBULK=1 FDB_TUNE_BULKGRID=$g; run "chunked bulk, grid x$g"
  export FDB_TUNE_BANDBULK=2 FDB_TUNE_BULKGRID=$g FDB_TUNE_ROWS_T=2000; run "row-stationary bulk T=2000, grid x$g"
done
