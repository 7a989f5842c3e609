"""A/B of the dense-column kernel (diff_columns) on a C5-shaped column block: loads in flight per thread x row blocks per
column.  Same box, same process; scatter time from the library's own CUDA events.  Usage: python profiles/r2_ab7.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

import _bootstrap
pkg = _bootstrap.load_package()
dev = torch.device("cuda", 0)
out = []
for fdtype in ("central", "forward"):
    for depth, gx in ((1, 64), (2, 64), (4, 64), (1, 32), (1, 128), (4, 32), (4, 128), (1, 196), (4, 196), (1, 64)):
        os.environ["FDB_COLS_DEPTH"] = str(depth)
        os.environ["FDB_COLS_GX"] = str(gx)
        prob = bench.build_gpu_problem(pkg, "c5", fdtype, dev, 0, 50, 256, False)
        J, f, x, cache = prob["J"], prob["f"], prob["x"], prob["cache"]
        step = lambda: pkg.finite_difference_jacobian_(J, f, x, cache)
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        plan = cache._last_plan
        plan.enable_timing(True)
        plan.read_timing()
        for _ in range(6):
            step()
        torch.cuda.synchronize()
        ms, nl = plan.read_timing()
        par = bench.analytic_parity(pkg, "c5", fdtype, prob)
        rec = dict(fdtype=fdtype, depth=depth, gx=gx, us_per_launch=1e3 * ms / nl, launches=nl, parity=par["ok"])
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del prob, J, f, x, cache, plan
        torch.cuda.empty_cache()
