"""In-tree build of the CUDA libraries for sm_100a (B200) — no other architecture is built.

    libfdjac_b200.so   the product: C ABI of include/fdjac_b200.h   (csrc/fdjac_abi.cu + kernels_*.cuh)
    libfdjac_synth.so  bench/test harness: synthetic f! device functions (include/fdjac_synth.h)

nvcc cross-compiles without a GPU.  -fmad=false: no FMA contraction, so the few multiply-adds on the path
(sum of squares, synthetic f!) round exactly like the CPU oracle compiled with -ffp-contract=off.
"""
from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
INCLUDE = HERE.parent / "include"

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "--shared", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default", "--extended-lambda",
]

TARGETS = {
    # every csrc/*.cuh is a dependency (globbed in build(): a header added later cannot be forgotten)
    "libfdjac_b200.so": (["fdjac_abi.cu"], ["*.cuh", "../../include/fdjac_b200.h"]),
    "libfdjac_synth.so": (["synth_fns.cu"], ["../../include/fdjac_synth.h"]),
}


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(p).exists():
        raise RuntimeError("nvcc not found: cannot build libfdjac_b200.so")
    return p


def lib_path(name: str) -> Path:
    return HERE / name


def _stale(out: Path, deps) -> bool:
    if not out.exists():
        return True
    t = out.stat().st_mtime
    return any(d.exists() and d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> dict:
    built = {}
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    for name, (srcs, deps) in TARGETS.items():
        out = lib_path(name)
        src_paths = [CSRC / s for s in srcs]
        dep_paths = src_paths + [Path(__file__)]
        for d in deps:
            dep_paths += sorted(CSRC.glob(d)) if "*" in d else [(CSRC / d).resolve()]
        if force or _stale(out, dep_paths):
            cmd = [nvcc_path(), *NVCC_FLAGS, "-I", str(INCLUDE), "-o", str(out), *map(str, src_paths)]
            if verbose:
                cmd.insert(1, "-Xptxas")
                cmd.insert(2, "-v")
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
            built[name] = (r.stdout + r.stderr).strip()
        else:
            built[name] = "up to date"
    return built


if __name__ == "__main__":
    import sys
    res = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for k, v in res.items():
        print(f"== {k}\n{v}")
