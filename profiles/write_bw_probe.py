"""Pure-write HBM bandwidth probe (what a store-only kernel such as the banded whole-band fill can reach):
torch fill_ and cudaMemsetAsync over 16 GB, CUDA-event timed, best of 5."""
import torch
n = 2_000_000_000
t = torch.empty(n, dtype=torch.float64, device="cuda")
for name, fn in (("fill_", lambda: t.fill_(1.0)), ("zero_", lambda: t.zero_())):
    best = 1e9
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    print(name, f"{best:.3f} ms  {n*8/best/1e6:.1f} GB/s")
src = torch.empty(n // 2, dtype=torch.float64, device="cuda"); dst = torch.empty_like(src)
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); dst.copy_(src); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("copy_", f"{best:.3f} ms  {2*src.numel()*8/best/1e6:.1f} GB/s (read+write)")

# non-constant data: broadcast a 40 MB (L2-resident) random row into the 16 GB tensor — the same traffic shape as the
# banded whole-band fill (16 GB written, sources served from L2)
del src, dst
row = torch.rand(5_000_000, dtype=torch.float64, device="cuda")
tv = t.view(400, 5_000_000)
best = 1e9
for _ in range(5):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); tv.copy_(row.view(1, -1).expand(400, -1)); b.record(); torch.cuda.synchronize()
    best = min(best, a.elapsed_time(b))
print("broadcast-copy of a 40 MB row", f"{best:.3f} ms  {n*8/best/1e6:.1f} GB/s written")

# store-only stream from our own probe kernel (16-byte streaming stores, grid-stride): constant vs distinct values,
# persistent grid (8 blocks/SM) — is the 7.5 TB/s of fill_ specific to constant data?
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import _bootstrap
synth = _bootstrap.load_package()._lib.synth()
for mode, name in ((0, "constant"), (1, "distinct values")):
    for blocks in (148 * 8, 148 * 32):
        best = 1e9
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); synth.fdbs_store_probe(t.data_ptr(), n, mode, blocks, None); b.record(); torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        print(f"store probe, {name}, {blocks} blocks", f"{best:.3f} ms  {n*8/best/1e6:.1f} GB/s written")
