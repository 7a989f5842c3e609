#!/usr/bin/env python
"""bench.py — the coloured sparse-Jacobian hot path on B200, measured per the driver contract.

A "step" = ONE full finite_difference_jacobian! call (eps pass, f!(x), per-colour perturb + f!, fused diff+scatter)
over one synthetic problem.  Workloads (BASELINE.json configs):
    c2  N=10^7 tridiagonal f!, 3 colours, CSC J, forward (default at --gpus 1; `--fdtype central` for the central leg)
    c4  N=5*10^6 random sparse f! (8 nnz/row), 64 colours, CSC J — colours sharded over the ranks (default at --gpus>1,
        strong scaling: the problem is fixed, each rank evaluates its share of the colours and stores its Jacobian
        entries straight into every peer's nzval over NVLink)
    c1 / c3 / c5 are parity-test cases (tests/), selectable here for ad-hoc timing.

JSON line: metric/value = whole-job Jacobian nnz/s with inputs resident in HBM; e2e = same metric through the C-ABI
host-buffer entry point (H2D of x, D2H of nzval inside the timed region); roofline = the diff+scatter kernel
(SURVEY.md §8(d) algorithmic bytes / CUDA-event launch time, vs MEASURED_PEAKS.json hbm_gbs); cpu_baseline = the CPU
oracle (port of the reference; no Julia in this image) timed on the host cores in the same run.

`--impl reference` times the reference's own CPU algorithm (the oracle port, all host threads) on the same workload.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

SEED = 0x5EED


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ workloads
def tridiag_pattern_torch(n, device):
    """colptr/rowval (Int64, 1-based) of the n x n tridiagonal CSC pattern, built on `device`."""
    import torch
    c = torch.arange(n + 1, dtype=torch.int64, device=device)
    colptr = 3 * c
    colptr[0] = 1
    colptr[n] = 3 * n - 1
    p = torch.arange(3 * n - 2, dtype=torch.int64, device=device)
    # slot p belongs to column col = (p+1)//3, offset k = (p+1)%3 ; row (1-based) = col + k
    q = p + 1
    rowval = q // 3 + q % 3
    return colptr, rowval


def tridiag_pattern_numpy(n):
    c = np.arange(n + 1, dtype=np.int64)
    colptr = 3 * c
    colptr[0] = 1
    colptr[n] = 3 * n - 1
    q = np.arange(1, 3 * n - 1, dtype=np.int64)
    return colptr, q // 3 + q % 3


def ell_problem(n, K, C_, seed):
    """row i picks K distinct colours of C_ and one random column per colour (cyclic colouring)."""
    rng = np.random.default_rng(seed)
    cols = np.empty((n, K), np.int32)
    per_color = n // C_
    step = 1 << 18
    for i0 in range(0, n, step):
        i1 = min(n, i0 + step)
        colors = np.argsort(rng.random((i1 - i0, C_), dtype=np.float32), axis=1)[:, :K]
        which = rng.integers(0, per_color, size=(i1 - i0, K))
        cols[i0:i1] = (which * C_ + colors).astype(np.int32)
    coef = rng.uniform(-1, 1, size=(n, K))
    return cols, coef


def ell_csc(n, K, cols):
    import scipy.sparse as sps
    A = sps.csc_matrix((np.ones(n * K, np.int8), (np.repeat(np.arange(n, dtype=np.int32), K), cols.reshape(-1))), shape=(n, n))
    A.sort_indices()
    return A.indptr.astype(np.int64) + 1, A.indices.astype(np.int64) + 1


class Clocks:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(gpu_index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.split(", ") for r in Path(self.f.name).read_text().strip().splitlines() if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        for r in rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.strip().lower() == "active":
                        reasons.add(name)
            except Exception:
                pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def usable_cores() -> int:
    """Host threads this process can really use: scheduler affinity capped by the cgroup CPU quota (the GPU boxes report
    128 CPUs but run the container under a 16-CPU quota: 128 OpenMP threads there are ~300x SLOWER than 16)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, p = Path("/sys/fs/cgroup/cpu.max").read_text().split()[:2]
        if q != "max":
            n = min(n, max(1, int(int(q) / int(p))))
    except Exception:
        try:
            q = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read_text())
            p = int(Path("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read_text())
            if q > 0:
                n = min(n, max(1, q // p))
        except Exception:
            pass
    return max(1, n)


def cpu_jacobian_runner(workload, fdtype, nthreads, scale=1.0):
    """Returns (run_once, nnz, n_fcalls, description) for the oracle on `workload` (bounded size: scale<1 shrinks n)."""
    from oracle import fd_oracle as orc
    fd = 0 if fdtype == "forward" else 1
    if workload in ("c1", "c2"):
        n = 1000 if workload == "c1" else int(10_000_000 * scale)
        colptr, rowval = tridiag_pattern_numpy(n)
        cv = (np.arange(n, dtype=np.int64) % 3) + 1
        P = orc.Problem.csc_same(n, n, colptr, rowval)
        x = orc.fill_x(n, SEED + 2, nthreads)
        nz = np.zeros(len(rowval))
        ctx = orc.SynthTridiagCtx(n, nthreads)
        cache = dict(x1=np.zeros(n), x2=np.zeros(n), fx=np.zeros(n), fx1=np.zeros(n))
        fn = orc.native_fn("synth_tridiag")
        if fdtype == "complex":
            fnc = orc.native_fn("synth_tridiag_c")

            def run_c():
                return orc.jacobian_complex(P, nz, fnc, x, colorvec=cv, nthreads=nthreads, ctx=ctx)["fcalls"]
            return run_c, len(rowval), 3, f"N={n} tridiagonal, 3 colours, complex step"

        def run():
            return orc.jacobian(P, nz, fn, x, fdtype=fd, colorvec=cv, nthreads=nthreads, ctx=ctx, cache=cache)["fcalls"]
        return run, len(rowval), (4 if fd == 0 else 6), f"N={n} tridiagonal, 3 colours, {fdtype}"
    if workload == "c4":
        n = int(5_000_000 * scale) // 64 * 64
        cols, coef = ell_problem(n, 8, 64, 11)
        colptr, rowval = ell_csc(n, 8, cols)
        cv = (np.arange(n, dtype=np.int64) % 64) + 1
        P = orc.Problem.csc_same(n, n, colptr, rowval)
        x = orc.fill_x(n, SEED + 4, nthreads)
        nz = np.zeros(len(rowval))
        colsT, coefT = np.ascontiguousarray(cols.T), np.ascontiguousarray(coef.T)     # ELL layout [K][m]
        ctx = orc.SynthEllCtx(n, 8, colsT.ctypes.data_as(C.POINTER(C.c_int32)), coefT.ctypes.data_as(C.POINTER(C.c_double)), nthreads)
        cache = dict(x1=np.zeros(n), x2=np.zeros(n), fx=np.zeros(n), fx1=np.zeros(n))
        fn = orc.native_fn("synth_ellrows")
        keep = (colsT, coefT)

        def run(_keep=keep):
            return orc.jacobian(P, nz, fn, x, fdtype=fd, colorvec=cv, nthreads=nthreads, ctx=ctx, cache=cache)["fcalls"]
        return run, len(rowval), (65 if fd == 0 else 128), f"N={n} random sparse 8 nnz/row, 64 colours, {fdtype}"
    raise SystemExit(f"no CPU runner for workload {workload}")


def time_cpu(run, budget_s=12.0, max_reps=5):
    t0 = time.perf_counter()
    run()
    first = time.perf_counter() - t0
    ts = [first]
    while len(ts) < max_reps and sum(ts) + first < budget_s:
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    return statistics.median(ts), len(ts)


def reference_arm(args):
    """`--impl reference`: the reference's own CPU algorithm (oracle port; the reference is Julia and cannot run here)
    with all host threads, on the same workload/metric.  Each step = one bounded-size Jacobian."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import fd_oracle as orc
    orc.build()
    cores = usable_cores()
    scale = args.cpu_scale
    run, nnz, fcalls, desc = cpu_jacobian_runner(args.workload, args.fdtype, cores, scale)
    for _ in range(args.warmup):
        run()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        run()
    dt = time.perf_counter() - t0
    value = nnz * args.steps / dt
    # the reference itself is single-threaded Julia: also report the port run the way the reference actually runs
    run1, _, _, _ = cpu_jacobian_runner(args.workload, args.fdtype, 1, scale)
    med1, reps1 = time_cpu(run1, budget_s=6.0, max_reps=3)
    line = {
        "impl": "reference", "metric": "jacobian_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.workload, args.fdtype), "sample": desc},
        "f_evals_per_s": fcalls * args.steps / dt,
        "cpu_baseline": {"value": value, "unit": "nnz/s", "cores": cores, "kind": "port",
                         "sample": f"{desc}; OpenMP over the reference's full-length passes ({cores} threads; the reference itself is single-threaded)"},
        "e2e": {"value": value, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "single_thread": {"value": nnz / med1, "unit": "nnz/s", "cores": 1, "reps": reps1,
                          "note": "the reference's own execution model (serial broadcast loops)"},
    }
    print(json.dumps(line))


def workload_name(w, fdtype):
    return {
        "c1": f"C1: N=1000 tridiagonal f!, 3 colours, CSC J, {fdtype}",
        "c2": f"C2: N=10^7 tridiagonal f!, colorvec=((j-1) mod 3)+1, SparseMatrixCSC J, {fdtype} fdtype, 1xB200",
        "c3": f"C3: N=10^6 2-D 5-point stencil, 5 colours, BandedMatrix l=u=1000, {fdtype}",
        "c4": f"C4: N=5*10^6 random sparse f! (8 nnz/row), 64-colour colorvec, CSC J, {fdtype}, colours sharded across ranks",
        "c5": f"C5: N=10^5 dense Jacobian (no colorvec), {fdtype}",
    }[w]


# ------------------------------------------------------------------------------------------------ GPU arm
def build_gpu_problem(pkg, workload, fdtype, dev, rank, world, max_batch, use_graph=True):
    """Returns dict(J, f, x, cache, nnz, fcalls, f_launches_per_call, keep)."""
    import torch
    L = pkg._lib
    synth = L.synth()

    def native(name, ctx, mb=1):
        return pkg.NativeFn(C.cast(getattr(synth, name), C.c_void_p).value, ctx, max_batch=mb)

    if workload in ("c1", "c2"):
        n = 1000 if workload == "c1" else 10_000_000
        colptr, rowval = tridiag_pattern_torch(n, dev)
        cv = (torch.arange(n, dtype=torch.int64, device=dev) % 3) + 1
        x = torch.empty(n, dtype=torch.float64, device=dev)
        synth.fdbs_fill_x(x.data_ptr(), n, SEED + 2, None)
        J = pkg.SparseMatrixCSC(n, n, colptr, rowval, torch.full((3 * n - 2,), float("nan"), dtype=torch.float64, device=dev))
        ctx = L.TridiagCtx(n, 0)
        f = native("fdbs_tridiag_c" if fdtype == "complex" else "fdbs_tridiag", ctx, max_batch)
        cache = pkg.JacobianCache(x, fdtype, colorvec=cv, sparsity=J, max_batch=max_batch, rank=rank, world=world,
                                  use_graph=use_graph)
        return dict(J=J, f=f, x=x, cache=cache, nnz=3 * n - 2, n=n, ctx=ctx, keep=(colptr, rowval, cv))
    if workload == "c4":
        n, K, Cc = 5_000_000, 8, 64
        # generated on the device (same seed on every rank -> identical problem): row i takes K distinct colours
        # (base + j*odd_stride mod 64) and one random column of each; ELL layout [K][n]; CSC = its transpose
        g = torch.Generator(device=dev).manual_seed(11)
        base = torch.randint(0, Cc, (n,), device=dev, generator=g, dtype=torch.int64)
        stride = torch.randint(0, Cc // 2, (n,), device=dev, generator=g, dtype=torch.int64) * 2 + 1
        colors = (base[None, :] + torch.arange(K, device=dev)[:, None] * stride[None, :]) % Cc
        which = torch.randint(0, n // Cc, (K, n), device=dev, generator=g, dtype=torch.int64)
        cols64 = which * Cc + colors
        d_cols = cols64.to(torch.int32).contiguous()
        d_coef = torch.rand((K, n), device=dev, generator=g, dtype=torch.float64) * 2 - 1
        rows = torch.arange(n, device=dev, dtype=torch.int64).repeat(K)
        order = torch.argsort(cols64.reshape(-1) * n + rows)
        rowval = (rows[order] + 1).contiguous()
        colptr = torch.cat([torch.ones(1, dtype=torch.int64, device=dev),
                            1 + torch.cumsum(torch.bincount(cols64.reshape(-1), minlength=n), 0)])
        del colors, which, cols64, rows, order, base, stride
        cv = (torch.arange(n, dtype=torch.int64, device=dev) % Cc) + 1
        x = torch.empty(n, dtype=torch.float64, device=dev)
        synth.fdbs_fill_x(x.data_ptr(), n, SEED + 4, None)
        J = pkg.SparseMatrixCSC(n, n, colptr, rowval,
                                torch.full((n * K,), float("nan"), dtype=torch.float64, device=dev))
        ctx = L.EllCtx(n, K, d_cols.data_ptr(), d_coef.data_ptr(), 0)
        f = native("fdbs_ellrows", ctx, max_batch)
        cache = pkg.JacobianCache(x, fdtype, colorvec=cv, sparsity=J, max_batch=max_batch, rank=rank, world=world,
                                  partition=0, use_graph=use_graph)
        return dict(J=J, f=f, x=x, cache=cache, nnz=n * K, n=n, ctx=ctx, keep=(d_cols, d_coef, cv))
    if workload == "c3":
        g = 1000
        n = g * g
        idx = torch.arange(n, dtype=torch.int64, device=dev)
        cv = ((idx % g) + 2 * (idx // g)) % 5 + 1
        x = torch.empty(n, dtype=torch.float64, device=dev)
        synth.fdbs_fill_x(x.data_ptr(), n, SEED + 3, None)
        J = pkg.BandedMatrix(n, n, g, g, device=dev)
        ctx = L.Lap5Ctx(g, 0)
        f = native("fdbs_lap5", ctx, max_batch)
        cache = pkg.JacobianCache(x, fdtype, colorvec=cv, sparsity=J, max_batch=max_batch, use_graph=use_graph)
        return dict(J=J, f=f, x=x, cache=cache, nnz=None, n=n, ctx=ctx, keep=(cv,))
    if workload == "c5":
        n = 100_000 if world == 1 else 100_000
        mb = max(max_batch, 256)
        w = torch.rand(n, dtype=torch.float64, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
        nblk = (n + 1023) // 1024
        bs = torch.zeros(nblk * mb, dtype=torch.float64, device=dev)
        ctx = L.Rank1Ctx(n, w.data_ptr(), bs.data_ptr(), mb, 0)
        x = torch.empty(n, dtype=torch.float64, device=dev)
        synth.fdbs_fill_x(x.data_ptr(), n, SEED + 5, None)
        f = native("fdbs_rank1", ctx, mb)
        if world != 1:
            raise SystemExit("c5 is benchmarked at --gpus 1 here (column-sharded runs: tests/test_gpu_multi.py)")
        cache = pkg.JacobianCache(x, fdtype, max_batch=mb, use_graph=use_graph)
        J = pkg.zeros_colmajor(n, n, dev)
        return dict(J=J, f=f, x=x, cache=cache, nnz=n * n, n=n, ctx=ctx, keep=(w, bs))
    raise SystemExit(f"unknown workload {workload}")


def gpu_arm(args):
    import torch
    import torch.distributed as dist
    import _bootstrap
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pkg = _bootstrap.load_package()
    L = pkg._lib
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    workload, fdtype = args.workload, args.fdtype
    prob = build_gpu_problem(pkg, workload, fdtype, dev, rank, world, args.max_batch, args.graph)
    J, f, x, cache = prob["J"], prob["f"], prob["x"], prob["cache"]

    sharded = None
    colsharded = None
    if world > 1 and args.shard == "columns":
        # contiguous column blocks + slice-aware f! (few-colour problems): only the tridiagonal workloads have one here
        if workload not in ("c1", "c2") or fdtype == "complex":
            raise SystemExit("--shard columns is benchmarked on the tridiagonal workloads (c1, c2), forward / central")
        from finitediff_jl_b200 import distributed as fdist
        n_glob = prob["n"]
        keep_ctx = []

        def factory(r0, r1, x0, x1):
            c = L.TridiagRowsCtx(n_glob, r0, r1 - r0, x0, 0)
            keep_ctx.append(c)
            return pkg.NativeFn(C.cast(L.synth().fdbs_tridiag_rows, C.c_void_p).value, c, max_batch=args.max_batch)

        colsharded = fdist.ColumnShardedJacobian(J, prob["keep"][2], fdtype, dev, factory,
                                                 gather=None if args.gather == "none" else "root",
                                                 max_batch=args.max_batch, use_graph=args.graph)
    elif world > 1:
        from finitediff_jl_b200 import distributed as fdist
        if args.gather == "none":
            raise SystemExit("--gather none needs --shard columns")
        sharded = fdist.ShardedJacobian(J, cache, x.numel(), dev, gather=args.gather)

    def step():
        if colsharded is not None:
            colsharded.run(x)
        elif sharded is not None:
            sharded.run(f, x)
        else:
            pkg.finite_difference_jacobian_(J, f, x, cache)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi samples every 100 ms while a timed region can be a few ms: sample across the whole measurement phase
    # (warm-up, timed region, kernel-timing pass) — all of it is the same kernel sequence under load
    clocks = Clocks(local_rank) if rank == 0 else None
    # the first call builds the plan (index compression, colour buckets, scratch) and captures the graph: one-off cost
    barrier()
    t_first = time.perf_counter()
    step()
    barrier()
    first_call_ms = (time.perf_counter() - t_first) * 1e3
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    plan = colsharded.block.plan if colsharded is not None else cache._last_plan
    info = plan.info()
    nnz = prob["nnz"] if prob["nnz"] is not None else info["n_entries"]
    c0 = plan.counters()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    c1 = plan.counters()
    f_points = c1["f_points"] - c0["f_points"]
    # the dominant kernel, timed live with CUDA events recorded by the library around its launch (eager launches:
    # events recorded inside a captured graph cannot be timed) — same stream, right after the timed region
    tsteps = args.steps
    plan.enable_timing(True)
    plan.read_timing()
    for _ in range(tsteps):
        step()
    barrier()
    scat_ms, scat_n = plan.read_timing()
    plan.enable_timing(False)
    clk = clocks.stop() if clocks else None
    f_launch_per_point = {"c5": 2}.get(workload, 1)
    lib_launches = c1["kernel_launches"] - c0["kernel_launches"]
    f_invocations = c1["f_invocations"] - c0["f_invocations"]
    gpu_launches = lib_launches + f_invocations * f_launch_per_point
    ms_step = ms_total / args.steps
    total_nnz = nnz if workload != "c5" else prob["n"] * prob["n"]
    value = total_nnz / (ms_step * 1e-3)
    if world > 1:
        fp = torch.tensor([float(f_points)], dtype=torch.float64, device=dev)
        dist.all_reduce(fp)
        f_points_all = fp.item()
    else:
        f_points_all = f_points

    # ---- roofline of the dominant kernel (diff+scatter), this rank
    peak, peak_src = peaks()
    alg_bytes = info["alg_bytes_scatter"]
    if world > 1 and workload != "c5" and colsharded is None:
        alg_bytes = alg_bytes * info["n_local_colors"] // max(info["n_colors"], 1)
    scat_per_jac = scat_ms / tsteps if tsteps else 0.0
    launches_per_jac = scat_n / tsteps if tsteps else 0
    achieved = alg_bytes / (scat_per_jac * 1e-3) / 1e9 if scat_per_jac > 0 else None
    if info["sp_kind"] == 1:
        kname = ("diff_scatter_cols<%s,%d lanes>" % (fdtype, info["lanes"]) if info["strategy"] == 1
                 else "diff_scatter_ident<u%d,%s,FULL>" % (info["color_bits"], fdtype))
    else:
        kname = {4: "diff_scatter_band", 0: "diff_columns", 3: "diff_scatter_dest"}.get(info["sp_kind"], "diff_scatter")
    traffic = args.traffic_bytes
    roofline = {"bound": "hbm", "kernel": kname,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "traffic": traffic, "peak_source": peak_src,
                "alg_bytes_per_launch": alg_bytes / launches_per_jac if launches_per_jac else None,
                "launch_ms": scat_per_jac / launches_per_jac if launches_per_jac else None,
                "alg_bytes_per_jacobian": alg_bytes, "scatter_ms_per_jacobian": scat_per_jac,
                "scatter_launches_per_jacobian": launches_per_jac,
                "frac_moved": (traffic * launches_per_jac / (scat_per_jac * 1e-3) / 1e9 / peak) if (traffic and scat_per_jac > 0) else None,
                "note": "achieved = SURVEY.md §8(d) algorithmic bytes (Int64 indices at the ABI, fx re-read per nonzero) / "
                        "CUDA-event time of the scatter launches; the fused kernel moves fewer real bytes (int32 rows + narrow "
                        "colours, fx[r] served from cache): `traffic` = ncu dram bytes per launch, `frac_moved` = traffic-based "
                        "fraction of the same peak"}

    # ---- e2e: host buffers through the C ABI (fdb_jacobian_host), H2D x + D2H J values inside the timed region
    e2e = None
    if world == 1 and workload in ("c1", "c2", "c4") and not args.no_e2e and fdtype != "complex":
        n = prob["n"]
        hx = pkg.pinned_empty(n)
        hx[:] = x.cpu().numpy()
        hJ = pkg.pinned_empty(info["j_len"])
        fptr, cptr = C.c_void_p(f.address), f.ctx_ptr
        for _ in range(2):
            L.check(L.lib().fdb_jacobian_host(plan.handle, fptr, cptr, hx.ctypes.data, hJ.ctypes.data, None, None, 0.0, 0.0, 1.0))
        ts = []
        reps = max(3, min(args.steps, 10))
        for _ in range(reps):
            t0 = time.perf_counter()
            L.check(L.lib().fdb_jacobian_host(plan.handle, fptr, cptr, hx.ctypes.data, hJ.ctypes.data, None, None, 0.0, 0.0, 1.0))
            ts.append(time.perf_counter() - t0)
        te = statistics.median(ts)
        assert np.isfinite(hJ).all()
        e2e = {"value": nnz / te, "unit": "nnz/s", "h2d_bytes_per_step": 8 * n, "d2h_bytes_per_step": 8 * info["j_len"],
               "ms_per_step": te * 1e3, "api": "fdb_jacobian_host (C ABI, pinned host x and nzval)", "reps": reps}
    elif world > 1:
        e2e = {"value": None, "unit": "nnz/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": "host-buffer e2e is measured at --gpus 1"}

    # ---- cpu_baseline: the oracle, 1 thread (the reference is single-threaded), bounded sample, rank 0 at N=1 only
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu and workload in ("c1", "c2", "c4"):
        from oracle import fd_oracle as orc
        orc.build()
        scale = 1.0 if workload != "c4" else 0.2
        run, cnnz, cf, desc = cpu_jacobian_runner(workload, fdtype, 1, scale)
        med, reps = time_cpu(run, budget_s=14.0, max_reps=5)
        cpu = {"value": cnnz / med, "unit": "nnz/s", "cores": 1, "kind": "port",
               "sample": f"{reps} full Jacobian(s) of {desc} (median {med:.3f} s); host offers {usable_cores()} usable cores",
               "f_evals_per_s": cf / med}

    if rank == 0:
        line = {
            "metric": "jacobian_nnz_per_s", "value": value, "unit": "nnz/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(workload, fdtype), "cuda_graph": bool(args.graph), "l2": "inputs larger than L2 (no flush needed): x, the stacked "
                       "f! outputs and nzval total far more than 126 MB per step" if workload != "c1" else "C1 is L2-resident (latency config)",
                       "max_batch": args.max_batch, "gather": args.gather if world > 1 else None, "shard": args.shard if world > 1 else None, "colors_local": info["n_local_colors"], "scatter_groups": info["n_groups"],
                       "first_call_ms": first_call_ms},
            "f_evals_per_s": f_points_all / (ms_total * 1e-3),
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(gpu_launches),
            "gpu_launches_detail": {"library_kernels": int(lib_launches), "f_callback_invocations": int(f_invocations)},
            "clocks": clk,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: per workload, a few seconds in total)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=["c1", "c2", "c3", "c4", "c5"])
    ap.add_argument("--fdtype", default="forward", choices=["forward", "central", "complex"])
    ap.add_argument("--max-batch", type=int, default=1, dest="max_batch")
    ap.add_argument("--no-graph", dest="graph", action="store_false",
                    help="launch eagerly instead of replaying the captured CUDA graph of the call")
    ap.add_argument("--gather", default="root", choices=["all", "root", "all_p2p", "none"],
                    help="N>1: rank 0 ends with the full Jacobian (root: fused NVLink gather), or every rank does "
                         "(all: root gather + NCCL broadcast; all_p2p: every value stored to every peer)")
    ap.add_argument("--shard", default="colors", choices=["colors", "columns"],
                    help="N>1: what is partitioned over the GPUs — the colour set (default; c4) or contiguous column "
                         "blocks with a slice-aware f! (c2: 3 colours cannot be spread over more than 3 GPUs); with "
                         "columns, --gather root assembles J on rank 0, --gather none leaves it column-sharded")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-scale", type=float, default=None, dest="cpu_scale",
                    help="--impl reference: problem-size fraction per step (bounded sample)")
    ap.add_argument("--traffic-bytes", type=float, default=None, dest="traffic_bytes",
                    help="ncu dram bytes per launch of the scatter kernel (from profiles/), echoed into roofline.traffic")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.workload is None:
        args.workload = "c2" if max(args.gpus, world) == 1 else "c4"
    if args.fdtype == "complex" and args.workload not in ("c1", "c2"):
        raise SystemExit("--fdtype complex is benchmarked on the tridiagonal workloads (c1, c2)")
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else {"c1": 500, "c2": 200, "c3": 50, "c4": 30, "c5": 5}[args.workload]
    if args.traffic_bytes is None:
        args.traffic_bytes = known_traffic(args.workload, args.fdtype)
    if args.impl == "reference":
        if args.cpu_scale is None:
            args.cpu_scale = 1.0 if args.workload in ("c1", "c2") else 0.1
        reference_arm(args)
    else:
        gpu_arm(args)


def known_traffic(workload, fdtype):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the scatter kernel from the committed ncu capture
    (profiles/), or None when no capture exists for this workload."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d.get(f"{workload}_{fdtype}")
    return None


if __name__ == "__main__":
    main()
