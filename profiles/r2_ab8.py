"""A/B of the step-size reduction (color_sumsq_reg): one vs two tiles of loads in flight per thread, C2 forward and
central, whole-step time (CUDA events, graph replay), alternating variants in one process on one box."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import _bootstrap
pkg = _bootstrap.load_package()
dev = torch.device("cuda", 0)
for rep in range(3):
    for fdtype in ("forward", "central"):
        for depth in (1, 2):
            os.environ["FDB_EPS_DEPTH"] = str(depth)
            prob = bench.build_gpu_problem(pkg, "c2", fdtype, dev, 0, 1, 1, True)
            J, f, x, cache = prob["J"], prob["f"], prob["x"], prob["cache"]
            step = lambda: pkg.finite_difference_jacobian_(J, f, x, cache)
            for _ in range(20):
                step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(300):
                step()
            e1.record()
            torch.cuda.synchronize()
            eps = cache._last_plan.eps()
            print(json.dumps(dict(rep=rep, fdtype=fdtype, eps_depth=depth, us_per_step=1e3 * e0.elapsed_time(e1) / 300,
                                  eps=[float(v).hex() for v in eps])), flush=True)
            del prob, J, f, x, cache
            torch.cuda.empty_cache()
