"""On-device colouring (fdb_matrix_colors_*): the `ArrayInterface.matrix_colors(A)` step of test/coloring_tests.jl:112,117.
Checks: ArrayInterface's closed forms for structured types; for CSC patterns a VALID distance-2 colouring (verified on the
host with scipy, independently of fdb_check_coloring_csc), deterministic, with a sane colour count; and the reference's own
use of it — the Jacobian obtained through the colouring equals the uncoloured one (coloring_tests.jl:99-119)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from _util import tridiag_csc  # noqa: E402


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no GPU")
    import _bootstrap
    return _bootstrap.load_package()


def host_valid(A, cv):
    """no two columns of one colour share a row"""
    import scipy.sparse as sps
    A = sps.csc_matrix(A)
    for k in np.unique(cv):
        sub = A[:, np.nonzero(cv == k)[0]]
        if sub.shape[1] and (abs(sub).sign().sum(axis=1).max() > 1):
            return False
    return True


def test_structured_closed_forms(pkg):
    dev = torch.device("cuda:0")
    cv = pkg.matrix_colors(pkg.Tridiagonal(10, device=dev))
    assert cv.dtype == torch.int64 and cv.tolist() == [1, 2, 3, 1, 2, 3, 1, 2, 3, 1]
    cv = pkg.matrix_colors(pkg.BandedMatrix(9, 9, 2, 1, device=dev))
    assert cv.tolist() == [1, 2, 3, 4, 1, 2, 3, 4, 1]
    cv = pkg.matrix_colors(pkg.zeros_colmajor(3, 5, dev))
    assert cv.tolist() == [1, 2, 3, 4, 5]


@pytest.mark.parametrize("kind", ["tridiag", "lap5", "random", "rect", "empty_cols"])
def test_csc_coloring_valid_and_deterministic(pkg, kind):
    import scipy.sparse as sps
    rng = np.random.default_rng(3)
    if kind == "tridiag":
        n = 5001
        A = sps.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1], format="csc")
    elif kind == "lap5":
        g = 60
        n = g * g
        A = sps.kron(sps.eye(g), sps.diags([1, 1, 1], [-1, 0, 1], shape=(g, g))) + sps.kron(sps.diags([1, 1], [-1, 1], shape=(g, g)), sps.eye(g))
        A = sps.csc_matrix(A)
    elif kind == "random":
        A = sps.random(4000, 3000, density=0.002, random_state=5, format="csc")
    elif kind == "rect":
        A = sps.random(50, 2000, density=0.05, random_state=6, format="csc")
    else:
        A = sps.random(300, 400, density=0.01, random_state=7, format="csc")       # many empty columns / rows
    A = sps.csc_matrix(A)
    A.sort_indices()
    m, n = A.shape
    S = pkg.SparseMatrixCSC(m, n, torch.from_numpy(A.indptr.astype(np.int64) + 1), torch.from_numpy(A.indices.astype(np.int64) + 1), None)
    cv = pkg.matrix_colors(S)
    cvh = cv.cpu().numpy()
    assert cvh.min() >= 1 and cvh.max() == cv.n_colors
    assert host_valid(A, cvh)
    assert pkg.check_coloring(S, cv) == 0
    # deterministic: depends on the pattern alone (index arrays on the device this time)
    S2 = pkg.SparseMatrixCSC(m, n, S.colptr.cuda(), S.rowval.cuda(), None)
    assert torch.equal(pkg.matrix_colors(S2), cv)
    # colour count: at least the largest row population, at most 1 + the largest number of 2-hop neighbours
    pat = abs(A).sign()
    row_pop = int(pat.sum(axis=1).max())
    nbrs = (pat.T @ pat)
    nbrs.setdiag(0)
    nbrs.eliminate_zeros()
    max_deg = int(np.diff(sps.csc_matrix(nbrs).indptr).max()) if nbrs.nnz else 0
    assert row_pop <= cv.n_colors <= max_deg + 1
    if kind == "tridiag":
        assert cv.n_colors <= 5
    # an invalid colouring is reported
    if row_pop > 1:
        assert pkg.check_coloring(S, torch.ones(n, dtype=torch.int64)) > 0


def test_colored_jacobian_equals_uncolored(pkg, oracle):
    """coloring_tests.jl:99-119 shape: the Jacobian through matrix_colors equals the column-by-column one"""
    import scipy.sparse as sps
    dev = torch.device("cuda:0")
    L = pkg._lib
    n, K, Cc = 64 * 40, 8, 64
    rng = np.random.default_rng(11)
    colors = np.argsort(rng.random((n, Cc)), axis=1)[:, :K]
    cols = (rng.integers(0, n // Cc, size=(n, K)) * Cc + colors).astype(np.int32)
    coef = rng.uniform(-1, 1, size=(n, K))
    A = sps.csc_matrix((np.ones(n * K), (np.repeat(np.arange(n), K), cols.reshape(-1))), shape=(n, n))
    A.sort_indices()
    colptr, rowval = torch.from_numpy(A.indptr.astype(np.int64) + 1), torch.from_numpy(A.indices.astype(np.int64) + 1)
    colsT, coefT = np.ascontiguousarray(cols.T), np.ascontiguousarray(coef.T)
    d_cols, d_coef = torch.from_numpy(colsT).to(dev), torch.from_numpy(coefT).to(dev)
    x = torch.from_numpy(oracle.fill_x(n, 31)).to(dev)
    res = []
    for use_colors in (True, False):
        J = pkg.SparseMatrixCSC(n, n, colptr, rowval, torch.full((A.nnz,), float("nan"), dtype=torch.float64, device=dev))
        cv = pkg.matrix_colors(J) if use_colors else None
        ctx = L.EllCtx(n, K, d_cols.data_ptr(), d_coef.data_ptr(), 0)
        f = pkg.NativeFn(C.cast(L.synth().fdbs_ellrows, C.c_void_p).value, ctx, max_batch=64)
        cache = pkg.JacobianCache(x, "forward", colorvec=cv, sparsity=J, max_batch=64)
        pkg.finite_difference_jacobian_(J, f, x, cache)
        torch.cuda.synchronize()
        res.append((J.nzval.cpu().numpy(), ctx.calls, None if cv is None else cv.n_colors))
    (Jc, calls_c, ncol), (Jd, calls_d, _) = res
    assert calls_c == ncol + 1 and calls_d == n + 1 and ncol < 200
    np.testing.assert_allclose(Jc, Jd, rtol=0, atol=1e-5)
