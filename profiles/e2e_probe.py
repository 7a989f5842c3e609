"""Where does the host-buffer (e2e) Jacobian spend its time?  Raw pinned PCIe copies of the same sizes next to the
fdb_jacobian_host call (C2).  Run on the GPU box: python profiles/e2e_probe.py"""
import sys, time, statistics, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import torch
import _bootstrap
pkg = _bootstrap.load_package()

def med(fn, reps=10):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return statistics.median(ts) * 1e3

n = 1_000_000
for mb, cnt in (("x 8 MB", n), ("nzval 24 MB", 3 * n)):
    h = torch.empty(cnt, dtype=torch.float64).pin_memory(); d = torch.empty(cnt, dtype=torch.float64, device="cuda")
    h2d = med(lambda: d.copy_(h, non_blocking=True)); d2h = med(lambda: h.copy_(d, non_blocking=True))
    print(f"raw pinned {mb}: H2D {h2d:.3f} ms ({cnt*8/h2d/1e6:.1f} GB/s)  D2H {d2h:.3f} ms ({cnt*8/d2h/1e6:.1f} GB/s)")
    hp = pkg.pinned_empty(cnt); hp[:] = 1.0
    import ctypes as C
    L = pkg._lib
    f = lambda: (L.lib().fdb_memcpy_h2d(C.c_void_p(d.data_ptr()), C.c_void_p(hp.ctypes.data), cnt * 8, None), L.lib().fdb_stream_sync(None))
    g = lambda: (L.lib().fdb_memcpy_d2h(C.c_void_p(hp.ctypes.data), C.c_void_p(d.data_ptr()), cnt * 8, None), L.lib().fdb_stream_sync(None))
    print(f"fdb_host_alloc {mb}: H2D {med(f):.3f} ms  D2H {med(g):.3f} ms")

# the host-path call itself, per phase (C2 forward tridiagonal)
import subprocess, json
out = subprocess.run([sys.executable, "bench.py", "--workload", "c2", "--no-cpu", "--steps", "20"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
d = json.loads(out)
print("bench c2: device ms/step", round(d["ms_per_step"], 4), " e2e", d["e2e"])
