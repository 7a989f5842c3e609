"""Writes tests/golden/dnrm2_openblas.json: results of OpenBLAS's OWN dnrm2 (the routine Julia's LinearAlgebra.norm calls
for Float64 vectors of 32 or more elements, jacobians.jl:560,601) on the deterministic vectors of dnrm2_vectors.py.

The binary is the OpenBLAS bundled with the scipy wheel in this image (scipy.linalg.blas.dnrm2); its version and kernel
target are recorded in the file.  The oracle's restatement (oracle/fd_oracle.c:openblas_dnrm2_x87) must reproduce every
value bit for bit — tests/test_oracle_norm.py.

Run:  python tests/golden/make_dnrm2_golden.py
"""
import json
import sys
from pathlib import Path

import numpy as np
from scipy.linalg import blas

sys.path.insert(0, str(Path(__file__).resolve().parent))
from dnrm2_vectors import CASES, vector  # noqa: E402


def main():
    info = {}
    try:
        from threadpoolctl import threadpool_info
        for lib in threadpool_info():
            if "scipy" in lib.get("filepath", "") and lib.get("internal_api") == "openblas":
                info = {"version": lib.get("version"), "architecture": lib.get("architecture"), "library": Path(lib["filepath"]).name}
    except Exception:
        pass
    out = {"routine": "dnrm2 (scipy.linalg.blas.dnrm2)", "openblas": info, "cases": []}
    for seed, n, kind, e in CASES:
        x = vector(seed, n, kind, e)
        r = float(blas.dnrm2(x))
        with np.errstate(over="ignore", under="ignore", invalid="ignore"):
            naive = float(np.sqrt(np.cumsum(x * x)[-1]))      # sqrt of the in-order Float64 sum: what a plain restatement gives
            ulps = (naive - r) / np.spacing(r)
        out["cases"].append({"seed": seed, "n": n, "kind": kind, "scale_exp": e, "dnrm2_hex": r.hex(),
                             "plain_double_sum_ulps": None if not np.isfinite(ulps) else round(float(ulps), 1)})
    p = Path(__file__).resolve().parent / "dnrm2_openblas.json"
    p.write_text(json.dumps(out, indent=1))
    print("wrote", p, len(out["cases"]), "cases", info)


if __name__ == "__main__":
    main()
