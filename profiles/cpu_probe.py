import os, time, sys
sys.path.insert(0,'.')
print("cpu_count", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max","/sys/fs/cgroup/cpu/cpu.cfs_quota_us","/sys/fs/cgroup/cpu/cpu.cfs_period_us"):
    try: print(f, open(f).read().strip())
    except Exception as e: print(f, "n/a")
print("OMP env", {k:v for k,v in os.environ.items() if k.startswith("OMP") or k.startswith("GOMP")})
import bench
for nt in (1,4,8,16,32,64,128):
    run,nnz,fc,desc=bench.cpu_jacobian_runner('c2','forward',nt,0.1)
    run()
    t0=time.perf_counter(); run(); run(); dt=(time.perf_counter()-t0)/2
    print(nt, "threads", round(dt*1e3,2), "ms", "%.3g nnz/s"%(nnz/dt), flush=True)
    if dt > 2: break
