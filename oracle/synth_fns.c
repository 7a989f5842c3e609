/*
 * synth_fns.c — CPU twins of the synthetic f!(dx, x) functions the benchmark and
 * parity tests differentiate (TEST INFRASTRUCTURE, same status as fd_oracle.c).
 * Arithmetic is written so that the CUDA twins in finitediff.jl_b200/csrc/synth_fns.cu
 * are bit-identical: explicit left-to-right evaluation, no FMA contraction
 * (-ffp-contract=off here; __dadd_rn/__dmul_rn there).
 *
 *   tridiag : test/coloring_tests.jl:5-13   dx[i] = x[i-1] - 2x[i] + x[i+1]
 *   lap5    : test/coloring_tests.jl:99-108 5-point clamped stencil on a g x g grid
 *   ellrows : random sparse rows, K entries per row, ELL layout [K][m]
 *   rank1   : dense Jacobian diag + rank-1 (bit-reproducible variant)
 */
#include <stdint.h>
#include <stddef.h>

typedef struct { int64_t n; int nthreads; } synth_tridiag_ctx;
typedef struct { int64_t g; int nthreads; } synth_lap5_ctx;
typedef struct { int64_t m; int64_t K; const int32_t *cols; const double *coef; int nthreads; } synth_ell_ctx;
typedef struct { int64_t n; const double *w; int nthreads; } synth_rank1_ctx;

#define DO_PRAGMA(x) _Pragma(#x)
#define PAR_FOR(nt) DO_PRAGMA(omp parallel for schedule(static) if ((nt) > 1) num_threads((nt) > 1 ? (nt) : 1))

/* counter-based generator: x[i] = 0.5 + u_i, u_i = top 53 bits of splitmix64(seed, i) / 2^53 */
static inline uint64_t splitmix64_at(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

void synth_fill_x(double *x, int64_t n, uint64_t seed, int nthreads) {
  PAR_FOR(nthreads)
  for (int64_t i = 0; i < n; ++i)
    x[i] = 0.5 + (double)(splitmix64_at(seed, (uint64_t)i) >> 11) * (1.0 / 9007199254740992.0);
}

void synth_tridiag(void *vctx, double *dx, const double *x) {
  const synth_tridiag_ctx *c = (const synth_tridiag_ctx *)vctx;
  const int64_t n = c->n;
  if (n == 1) { dx[0] = -2 * x[0]; return; }
  PAR_FOR(c->nthreads)
  for (int64_t i = 1; i < n - 1; ++i) dx[i] = (x[i - 1] - 2 * x[i]) + x[i + 1];
  dx[0] = -2 * x[0] + x[1];
  dx[n - 1] = x[n - 2] - 2 * x[n - 1];
}

void synth_lap5(void *vctx, double *out, const double *x) {
  const synth_lap5_ctx *c = (const synth_lap5_ctx *)vctx;
  const int64_t g = c->g;
  PAR_FOR(c->nthreads)
  for (int64_t j = 0; j < g; ++j)
    for (int64_t i = 0; i < g; ++i) {
      int64_t im = i > 0 ? i - 1 : 0, ip = i + 1 < g ? i + 1 : g - 1;
      int64_t jm = j > 0 ? j - 1 : 0, jp = j + 1 < g ? j + 1 : g - 1;
      double s = x[i + j * g];
      s = s + x[im + j * g];
      s = s + x[ip + j * g];
      s = s + x[i + jm * g];
      s = s + x[i + jp * g];
      out[i + j * g] = s;
    }
}

/* dx[i] = sum_p coef[p,i]*x[cols[p,i]]  (left to right)  + 0.1*x[cols[0,i]]^2 ; cols 0-based, ELL layout [K][m]
 * (entry p of row i at index p*m + i) */
void synth_ellrows(void *vctx, double *dx, const double *x) {
  const synth_ell_ctx *c = (const synth_ell_ctx *)vctx;
  const int64_t K = c->K, m = c->m;
  PAR_FOR(c->nthreads)
  for (int64_t i = 0; i < m; ++i) {
    double x0 = x[c->cols[i]];
    double s = c->coef[i] * x0;
    for (int64_t p = 1; p < K; ++p) s = s + c->coef[p * m + i] * x[c->cols[p * m + i]];
    s = s + 0.1 * (x0 * x0);
    dx[i] = s;
  }
}

/* f_i(x) = x_i*x_i + w_i * S,  S = (sum_j x_j)/n with the fixed-order blocked sum below
 * (blocks of 1024 summed sequentially, block sums summed sequentially) so CPU and GPU agree bitwise.
 * Jacobian = diag(2 x_i) + w * 1^T / n  (dense). */
double synth_blocked_sum(const double *x, int64_t n) {
  /* warp-shaped order (so the CUDA twin can read coalesced): per block of 1024, 32 lane partials
   * p[l] = sum_j x[b + l + 32 j] (j ascending; elements past n count as 0), then the xor butterfly
   * p[l] += p[l^o], o = 16,8,4,2,1; block sums are added in block order. */
  double total = 0.0;
  for (int64_t b = 0; b < n; b += 1024) {
    double p[32], q[32];
    for (int l = 0; l < 32; ++l) {
      double s = 0.0;
      for (int j = 0; j < 32; ++j) {
        int64_t idx = b + l + 32 * (int64_t)j;
        s = s + (idx < n ? x[idx] : 0.0);
      }
      p[l] = s;
    }
    for (int o = 16; o > 0; o >>= 1) {
      for (int l = 0; l < 32; ++l) q[l] = p[l] + p[l ^ o];
      for (int l = 0; l < 32; ++l) p[l] = q[l];
    }
    total = total + p[0];
  }
  return total;
}

void synth_rank1(void *vctx, double *dx, const double *x) {
  const synth_rank1_ctx *c = (const synth_rank1_ctx *)vctx;
  const int64_t n = c->n;
  double S = synth_blocked_sum(x, n) / (double)n;
  PAR_FOR(c->nthreads)
  for (int64_t i = 0; i < n; ++i) dx[i] = x[i] * x[i] + c->w[i] * S;
}

/* complex twin of synth_tridiag for the complex-step path: the stencil applied to real and imaginary parts
 * separately (exactly what complex addition / real scaling do), same evaluation order as the real one */
#include <complex.h>
void synth_tridiag_c(void *vctx, double _Complex *dx, const double _Complex *x) {
  const synth_tridiag_ctx *c = (const synth_tridiag_ctx *)vctx;
  const int64_t n = c->n;
  const double *xr = (const double *)x;   /* interleaved (re, im) */
  double *dr = (double *)dx;
  for (int part = 0; part < 2; ++part) {
    if (n == 1) { dr[part] = -2 * xr[part]; continue; }
    PAR_FOR(c->nthreads)
    for (int64_t i = 1; i < n - 1; ++i) dr[2 * i + part] = (xr[2 * (i - 1) + part] - 2 * xr[2 * i + part]) + xr[2 * (i + 1) + part];
    dr[part] = -2 * xr[part] + xr[2 + part];
    dr[2 * (n - 1) + part] = xr[2 * (n - 2) + part] - 2 * xr[2 * (n - 1) + part];
  }
}
