"""Writes tests/golden/kat_reference_tests.json.

The reference (FiniteDiff.jl, pure Julia) cannot run in this image (no `julia`), so the
golden vectors are the KNOWN-ANSWER fixtures its own tests hold for the hot path,
transcribed here with the file:line they come from.  Inputs the reference draws with an
unseeded `rand` are replaced by seeded draws — the expected answers are input-independent
(linear stencils) or closed-form in the input, exactly as in the reference tests.

Run:  python tests/golden/make_golden.py
"""
import json
from pathlib import Path

import numpy as np


def second_derivative_stencil(N):
    # test/coloring_tests.jl:19-26
    A = np.zeros((N, N))
    for i in range(N):
        for j in range(N):
            if abs(j - i) == 1:
                A[i, j] = 1.0
            if j == i:
                A[i, j] = -2.0
    return A


def main():
    rng = np.random.default_rng(20260924)
    out = {}
    # --- coloring_tests.jl:28-96: tridiagonal N=30, colorvec=repeat(1:3,10) ---
    N = 30
    out["tridiag30"] = {
        "cite": "test/coloring_tests.jl:5-13,19-49,51-96",
        "N": N,
        "x": rng.random(N).tolist(),
        "colorvec": (np.tile(np.arange(1, 4), 10)).tolist(),
        "J_expected": second_derivative_stencil(N).tolist(),
        "fcalls": {"forward": 4, "central": 6},
        "rtol": float(np.sqrt(np.finfo(float).eps)),  # Julia's `≈` default
    }
    # --- coloring_tests.jl:122-159: non-square 4x8, two colours ---
    n = 4
    x0 = np.concatenate([np.arange(1, n + 1) + 0.5, np.arange(1, n + 1) + 1.5])
    x1, x2 = x0[:n], x0[n:]
    # y = (x1-3)^2 + x1*x2 + (x2+4)^2 - 3 ; dy/dx1 = 2(x1-3)+x2 ; dy/dx2 = x1 + 2(x2+4)
    Jn = np.zeros((n, 2 * n))
    for i in range(n):
        Jn[i, i] = 2 * (x1[i] - 3) + x2[i]
        Jn[i, i + n] = x1[i] + 2 * (x2[i] + 4)
    out["nonsquare4x8"] = {
        "cite": "test/coloring_tests.jl:122-159",
        "n": n,
        "x0": x0.tolist(),
        "rows": list(range(1, n + 1)) * 2,
        "cols": list(range(1, n + 1)) + [i + n for i in range(1, n + 1)],
        "colorvec": [1] * n + [2] * n,
        "J_analytic": Jn.tolist(),
        "fcalls": {"forward": 3, "central": 4},
        "rtol": 1e-6,
    }
    # --- coloring_tests.jl:163-168: _findstructralnz ordering == findnz(sparse(a)) (column-major) ---
    protos = [[[1, 1], [0, 1]], [[1, 1, 1]], [[1.0, 1.0], [1.0, 1.0], [1.0, 1.0]], [[1, 1], [1, 1]]]
    exp = []
    for a in protos:
        A = np.array(a, dtype=float)
        r, c = np.nonzero(A.T)  # column-major order: iterate columns then rows
        exp.append({"A": a, "rows": (c + 1).tolist(), "cols": (r + 1).tolist()})
    out["findstructralnz"] = {"cite": "test/coloring_tests.jl:163-168; src/jacobians.jl:473-488", "cases": exp}
    # --- coloring_tests.jl:171-220: dense 0/1 prototypes with closed-form Jacobians ---
    out["dense_prototypes"] = {
        "cite": "test/coloring_tests.jl:171-220",
        "cases": [
            {"name": "_f", "theta": [5.0, 3.0], "m": 2, "sparsity": [[1, 1], [1, 1]], "J": [[10, 6], [1, 1]]},
            {"name": "_f2", "theta": [-3.0, 2.0], "m": 2, "sparsity": [[1, 1], [1, 0]], "J": [[-6, 4], [1, 0]]},
            {"name": "_f3", "theta": [-3.0, 2.0], "m": 1, "sparsity": [[1, 1]], "J": [[-7, 4]]},
            {"name": "_f4", "theta": [-3.0, 2.0, 13.3], "m": 4,
             "sparsity": [[1, 1, 0], [1, 1, 0], [1, 0, 1], [1, 0, 0]],
             "J": [[-7.0, 4.0, 0], [2.0, -3.0, 0.0], [13.3, 0.0, -3.0], [1.0, 0.0, 0.0]]},
            {"name": "_f5", "theta": [5.0, 3.0], "m": 1, "sparsity": [[1, 1]], "J": [[10.0, 6.0]]},
        ],
        "rtol": float(np.sqrt(np.finfo(float).eps)),
    }
    # --- cache_reuse_tests.jl:9-15,56-84: poisoned caches, x untouched in central ---
    out["cache_reuse"] = {
        "cite": "test/cache_reuse_tests.jl:9-15,56-84",
        "J_REF": [[2.0, 0.0], [0.0, 3.0], [4.0, 0.0]],
        "X_TEST": [1.0, 2.0],
        "poison": 1.0e10,
        "atol": 1e-6,
    }
    # --- finitedifftests.jl:398-463: analytic 2x2, bounds fwd<1e-6 central<1e-8, dir=-1, relstep, f_in ---
    x = rng.random(2)
    out["analytic2x2"] = {
        "cite": "test/finitedifftests.jl:398-463",
        "x": x.tolist(),
        "bounds": {"forward": 1e-6, "central": 1e-8},
    }
    # --- finitedifftests.jl:600-605: identity map at ones(2), cache-less in-place call, all three fdtypes, J ≈ I ---
    out["identity2"] = {
        "cite": "test/finitedifftests.jl:600-605",
        "x": [1.0, 1.0],
        "J_expected": [[1.0, 0.0], [0.0, 1.0]],
        "fdtypes": ["forward", "central", "complex"],
        "rtol": float(np.sqrt(np.finfo(float).eps)),       # Julia's isapprox default for Float64
    }
    # --- finitedifftests.jl:524-536: "Jacobian for non-vector inputs": x = rand(2,2), iipf(fx,x) = (fx .= x), J_ref = I(4),
    #     in-place cache-less call with f_in = iipf(similar(x), x); err < 1e-8 for forward / central / complex ---
    out["nonvector_identity"] = {
        "cite": "test/finitedifftests.jl:524-536",
        "x": rng.random((2, 2)).tolist(),
        "bound": 1e-8,
        "fdtypes": ["forward", "central", "complex"],
    }
    # --- epsilons.jl:134-144 defaults ---
    out["default_relstep"] = {
        "cite": "src/epsilons.jl:134-144",
        "forward": 1.4901161193847656e-08,
        "central": 6.0554544523933395e-06,
    }
    p = Path(__file__).resolve().parent / "kat_reference_tests.json"
    p.write_text(json.dumps(out, indent=1))
    print("wrote", p)


if __name__ == "__main__":
    main()
