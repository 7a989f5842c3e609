"""Shared helpers for the parity tests: pattern builders in the reference's (Julia) conventions —
Int64, 1-based CSC exactly as SparseMatrixCSC stores it."""
import numpy as np
import scipy.sparse as sp


def tridiag_csc(n):
    """colptr/rowval (1-based Int64) of the n x n tridiagonal pattern (sparse(second_derivative_stencil))."""
    A = sp.diags([np.ones(n - 1), np.ones(n), np.ones(n - 1)], [-1, 0, 1], format="csc")
    A.sort_indices()
    return (A.indptr.astype(np.int64) + 1), (A.indices.astype(np.int64) + 1)


def csc_from_dense_pattern(A):
    S = sp.csc_matrix(np.asarray(A) != 0)
    S.sort_indices()
    return S.indptr.astype(np.int64) + 1, S.indices.astype(np.int64) + 1


def csc_to_dense(m, n, colptr, rowval, nzval):
    J = np.zeros((m, n))
    for c in range(n):
        for p in range(colptr[c] - 1, colptr[c + 1] - 1):
            J[rowval[p] - 1, c] = nzval[p]
    return J


def band_to_dense(m, n, l, u, data):
    """BandedMatrices storage data[(l+u+1) x n] column-major flat -> dense."""
    w = l + u + 1
    J = np.zeros((m, n))
    for c in range(1, n + 1):
        for r in range(max(1, c - u), min(m, c + l) + 1):
            J[r - 1, c - 1] = data[(c - 1) * w + (u + r - c)]
    return J


def cyc_colors(n, C):
    return (np.arange(n, dtype=np.int64) % C) + 1


def f_tridiag(dx, x):
    # test/coloring_tests.jl:5-13
    n = len(x)
    dx[1:n - 1] = (x[0:n - 2] - 2 * x[1:n - 1]) + x[2:n]
    dx[0] = -2 * x[0] + x[1]
    dx[n - 1] = x[n - 2] - 2 * x[n - 1]


def tridiagonal_coo(n):
    """Structural nz of Julia's Tridiagonal(dl,d,du) as (rows, cols, slots) with slots into [dl; d; du]
    (1-based): band order sub-diagonal, diagonal, super-diagonal."""
    rows = np.concatenate([np.arange(2, n + 1), np.arange(1, n + 1), np.arange(1, n)]).astype(np.int64)
    cols = np.concatenate([np.arange(1, n), np.arange(1, n + 1), np.arange(2, n + 1)]).astype(np.int64)
    slots = np.arange(1, 3 * n - 1, dtype=np.int64)
    return rows, cols, slots


def tridiagonal_to_dense(n, buf):
    dl, d, du = buf[: n - 1], buf[n - 1: 2 * n - 1], buf[2 * n - 1:]
    return np.diag(d) + np.diag(dl, -1) + np.diag(du, 1)
