#!/bin/bash
# compute-sanitizer memcheck over the late-r2 kernel forms: diff_columns with 2 / 4 loads in flight (m = 6000 and m = 10^5),
# the 16-byte rank-1 harness f!, and color_sumsq_reg with two tiles in flight (n large enough that the two-tile loop runs)
set -u
O=gpurun_out
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -x \
  -k "dense_many_row_blocks or dense_columns_bitexact or c5_column_block or dense_view_with_padding or colorvec_quirk" \
  -p no:cacheprovider > $O/r2_sanitize2_memcheck.log 2>&1
echo "memcheck exit $?" >> $O/r2_sanitize2_memcheck.log
tail -4 $O/r2_sanitize2_memcheck.log
compute-sanitizer --tool memcheck --error-exitcode 9 python - > $O/r2_sanitize2_eps.log 2>&1 <<'PY'
import torch, _bootstrap
pkg = _bootstrap.load_package()
import bench
dev = torch.device("cuda", 0)
for n in (3_000_001, 2_621_440):
    for fdtype in ("forward", "central"):
        colptr, rowval = bench.tridiag_pattern_torch(n, dev)
        cv = (torch.arange(n, dtype=torch.int64, device=dev) % 3) + 1
        x = torch.rand(n, dtype=torch.float64, device=dev) + 0.5
        J = pkg.SparseMatrixCSC(n, n, colptr, rowval, torch.full((3 * n - 2,), float("nan"), dtype=torch.float64, device=dev))
        import ctypes as C
        L = pkg._lib
        ctx = L.TridiagCtx(n, 0)
        f = pkg.NativeFn(C.cast(L.synth().fdbs_tridiag, C.c_void_p).value, ctx)
        cache = pkg.JacobianCache(x, fdtype, colorvec=cv, sparsity=J)
        pkg.finite_difference_jacobian_(J, f, x, cache)
        torch.cuda.synchronize()
        q = torch.arange(1, J.nzval.numel() + 1, device=dev)
        want = torch.where(q % 3 == 1, -2.0, 1.0).to(torch.float64)
        print(n, fdtype, "max err", float((J.nzval - want).abs().max()))
PY
echo "memcheck exit $?" >> $O/r2_sanitize2_eps.log
tail -6 $O/r2_sanitize2_eps.log
