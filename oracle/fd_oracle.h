/*
 * fd_oracle.h — CPU ORACLE for the coloured finite-difference Jacobian hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and there only as the checker / reported baseline.
 *
 * What it is: a line-by-line CPU restatement (plain C, double precision, 1-based
 * Int64 indices exactly as Julia stores them) of FiniteDiff.jl v2.31.1's cached
 * in-place Jacobian driver and its decompression hooks:
 *   src/jacobians.jl:504-653   cached finite_difference_jacobian!  (driver)
 *   src/jacobians.jl:446-471   cache-less entry (prologue)
 *   src/jacobians.jl:473-488   _findstructralnz(::DenseMatrix)
 *   src/epsilons.jl:26-29,50-53,134-144   compute_epsilon / default_relstep
 *   src/iteration_utils.jl:25-32          generic COO _colorediteration!
 *   ext/FiniteDiffSparseArraysExt.jl:20-28,38-47,30,51-52   CSC hooks
 *   ext/FiniteDiffBandedMatricesExt.jl:13-27                banded hook
 *
 * Parity pinning: the reference is pure Julia and no `julia` binary exists in
 * this image, so the oracle cannot be compared with outputs of the reference
 * itself.  It IS pinned against every known-answer fixture the reference's own
 * tests hold for this path (test/coloring_tests.jl, test/cache_reuse_tests.jl,
 * test/finitedifftests.jl:398-463) — values to those tests' own tolerances,
 * f!-call counts/order and index sets exactly (tests/test_oracle_kat.py).
 * Third-party arithmetic on the path: Julia's LinearAlgebra.norm at
 * jacobians.jl:560,601 (stdlib generic_norm2 for n<32, OpenBLAS dnrm2 for n>=32;
 * Project.toml:30 julia>=1.10; neither is under /root/reference).  fdo_norm2
 * restates both; the dnrm2 half is pinned BIT FOR BIT against an OpenBLAS binary
 * (tests/golden/dnrm2_openblas.json, tests/test_oracle_norm.py); the
 * generic_norm2 half follows the published stdlib source and stays unexecuted
 * ("parity unpinned" for n<32 at bit level; the reference's KATs with n=30 pin
 * its values to their tolerances).  It only determines eps.
 */
#ifndef FD_ORACLE_H
#define FD_ORACLE_H
#include <stdint.h>
#ifndef __cplusplus
#include <complex.h>
#endif
#ifdef __cplusplus
extern "C" {
#endif

/* in-place user function f!(fx, x)  (jacobians.jl:563 `f(fx1, x1)`) */
typedef void (*fdo_fn)(void *ctx, double *fx, const double *x);

enum { FDO_FORWARD = 0, FDO_CENTRAL = 1 };

/* which decompression hook the reference would dispatch to */
enum {
  FDO_SP_NONE = 0,      /* sparsity === nothing: dense column branch, jacobians.jl:548-557,590-598 */
  FDO_SP_CSC_SAME = 1,  /* J::CSC, same pattern as sparsity: ext/Sparse..:38-47 (writes nzval[p]) */
  FDO_SP_CSC = 2,       /* sparsity::CSC, J addressed by (row,col): ext/Sparse..:20-28 */
  FDO_SP_COO = 3,       /* rows_index/cols_index lists: iteration_utils.jl:25-32 */
  FDO_SP_BANDED = 4     /* sparsity::BandedMatrix(l,u): ext/Banded..:13-27 */
};

/* how J[row,col] = v lands in memory */
enum {
  FDO_J_NZVAL = 0,      /* CSC nzval slot p (only with FDO_SP_CSC_SAME) */
  FDO_J_DENSE = 1,      /* column-major dense m x n, leading dim ldJ */
  FDO_J_BAND = 2,       /* BandedMatrices data[(l+u+1) x n], slot [u+r-c+1, c] (ext/Banded..:22) */
  FDO_J_SLOTS = 3       /* explicit 1-based slot per structural entry (structured J, e.g. Tridiagonal) */
};

typedef struct {
  int sp_kind;              /* FDO_SP_* */
  int j_kind;               /* FDO_J_*  */
  int64_t m, n;             /* size(J) */
  /* CSC sparsity (1-based, Int64, as SparseMatrixCSC stores them) */
  const int64_t *colptr;    /* n+1 */
  const int64_t *rowval;    /* nnz */
  /* COO sparsity */
  const int64_t *rows_index, *cols_index;
  int64_t nnz;
  /* explicit slots (FDO_J_SLOTS): 1-based offsets into J, one per COO entry */
  const int64_t *slots;
  /* banded */
  int64_t l, u;
  /* dense J */
  int64_t ldJ;
  /* total number of doubles in J's storage (for fill_matrix!) */
  int64_t j_len;
} fdo_problem;

/* JacobianCache fields (jacobians.jl:1-9). All caller-owned. */
typedef struct {
  double *x1;   /* n */
  double *x2;   /* n */
  double *fx;   /* m */
  double *fx1;  /* m */
} fdo_cache;

typedef struct {
  int fdtype;            /* FDO_FORWARD / FDO_CENTRAL */
  double relstep;        /* NaN -> default_relstep(fdtype, Float64); any other value as passed */
  double absstep;        /* NaN -> relstep  (jacobians.jl:510); any other value (0 included) as passed */
  double dir;            /* forward only; 1.0 default (`dir=true`) */
  const int64_t *colorvec; /* n entries, 1-based colours; NULL -> 1:n */
  const double *f_in;    /* forward: precomputed f(x) or NULL */
  const double *eps_override; /* optional: use these eps per colour instead of computing */
  double *eps_out;       /* optional: receives eps per colour (len = maxcolor) */
  int no_drift;          /* 0 = faithful ((x1+eps)-eps replayed); 1 = restore exactly */
  int nthreads;          /* <=1 single thread (the reference is single-threaded) */
  int64_t fcalls;        /* out: number of f! calls made */
} fdo_opts;

/* LinearAlgebra.norm(x::Vector{Float64}) as the reference evaluates it (see fd_oracle.c) */
double fdo_norm2(const double *v, int64_t n);
double fdo_default_relstep(int fdtype);
double fdo_compute_epsilon(int fdtype, double x, double relstep, double absstep, double dir);
int64_t fdo_max_color(const int64_t *colorvec, int64_t n);

/* jacobians.jl:473-488 — column-major scan of a dense prototype; returns nnz,
 * fills rows/cols (1-based) if non-NULL. A is column-major m x n. */
int64_t fdo_findstructralnz_dense(const double *A, int64_t m, int64_t n,
                                  int64_t *rows, int64_t *cols);

/* cached finite_difference_jacobian!(J, f, x, cache, f_in; ...) jacobians.jl:504-653.
 * x is mutable (central mode perturbs the caller's x in place and restores it).
 * Returns 0 on success, nonzero on invalid arguments. */
int fdo_finite_difference_jacobian(const fdo_problem *P, double *J, fdo_fn f, void *ctx,
                                   double *x, fdo_cache *cache, fdo_opts *opts);

/* cache-less entry jacobians.jl:446-471: evaluates fx=f(x) first (forward, no f_in),
 * allocates its own cache. */
int fdo_finite_difference_jacobian_cacheless(const fdo_problem *P, double *J, fdo_fn f, void *ctx,
                                             double *x, fdo_opts *opts);

/* Jacobian-vector product, cached in-place form: src/jvp.jl:238-274.  x1[n], fx1[m] are the JVPCache arrays. */
int fdo_finite_difference_jvp(double *jvp, fdo_fn f, void *ctx, const double *x, const double *v, int64_t m, int64_t n,
                              double *x1, double *fx1, const double *f_in, int fdtype, double relstep, double absstep,
                              double dir, const double *eps_override, double *eps_out, int64_t *fcalls);

/* complex-step variant: jacobians.jl:623-648.  f!(fx, x) works on complex128 arrays; x and J are real.
 * (C only: the complex callback type is C99 `double _Complex`.) */
#ifndef __cplusplus
typedef void (*fdo_fn_c)(void *ctx, double _Complex *fx, const double _Complex *x);
int fdo_finite_difference_jacobian_complex(const fdo_problem *P, double *J, fdo_fn_c f, void *ctx, const double *x,
                                           const int64_t *colorvec, int nthreads, int64_t *fcalls);
#endif

#ifdef __cplusplus
}
#endif
#endif
