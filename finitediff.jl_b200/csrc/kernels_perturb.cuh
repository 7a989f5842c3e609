// kernels_perturb.cuh — K3: build the perturbed point(s) of one or several colours.
//
// Reference, forward (jacobians.jl:562,584):  x1 .+= eps_k*(color==k);  f!(fx1,x1);  x1 .-= eps_k*(color==k)
// central (:603-604,:619-620) additionally  x .-= eps_k*(color==k) ... x .+= eps_k*(color==k)  on the caller's x.
// The reference never restores exactly: after colour c has been processed its components hold (x+eps_c)-eps_c
// (and (x-eps_c)+eps_c in the caller's x), and every LATER colour's f! sees those drifted values.  Because a
// component belongs to one colour only, the state of x1 at colour k has the closed form
//        x1_k[j] = c(j) <  k : (x[j]+eps_c)-eps_c      (drift replay, `drift` != 0)
//                  c(j) == k :  x[j]+eps_k
//                  else      :  x[j]
// which needs no sequential dependence between colours — any GPU can build the point of any colour from the
// pristine x and the eps table.  The caller's x is never written.
#pragma once
#include "common.cuh"

namespace fdb {

template <typename CT, bool CENTRAL>
__global__ void __launch_bounds__(kThreads)
perturb_colors(const double *__restrict__ x, const CT *__restrict__ jcolor, const double *__restrict__ eps,
               const int32_t *__restrict__ klist /* global colour id of each point, device */, int32_t kcount,
               int32_t C, int drift, int64_t n, int64_t ldx, double *__restrict__ xp, double *__restrict__ xm) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t j = blockIdx.x * (int64_t)kThreads + threadIdx.x; j < n; j += stride) {
    const double v = ld_stream(x + j);
    const uint32_t c = (uint32_t)jcolor[j];
    const bool valid = c < (uint32_t)C;
    const double e = valid ? __ldg(eps + c) : 0.0;
    const double up_done = valid && drift ? (v + e) - e : v;   // colours already processed
    const double up_now = v + e;
    double dn_done = v, dn_now = v;
    if (CENTRAL) {
      dn_done = valid && drift ? (v - e) + e : v;
      dn_now = v - e;
    }
    for (int32_t b = 0; b < kcount; ++b) {
      const uint32_t k = (uint32_t)__ldg(klist + b);
      const double p = !valid ? v : (c == k ? up_now : (c < k ? up_done : v));
      st_stream(xp + (int64_t)b * ldx + j, p);
      if (CENTRAL) {
        const double q = !valid ? v : (c == k ? dn_now : (c < k ? dn_done : v));
        st_stream(xm + (int64_t)b * ldx + j, q);
      }
    }
  }
}

// ---- dense-column branch (jacobians.jl:548-557, :590-598) ----
// The batch buffer X[b] (b < B) holds B copies of x; per batch only the one perturbed component per copy changes.

// X[b][:] = x  for b < B
__global__ void __launch_bounds__(kThreads)
replicate_x(const double *__restrict__ x, int64_t n, int64_t ldx, int32_t B, double *__restrict__ X) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t j = blockIdx.x * (int64_t)kThreads + threadIdx.x; j < n; j += stride) {
    const double v = x[j];
    for (int32_t b = 0; b < B; ++b) X[(int64_t)b * ldx + j] = v;
  }
}

// batch starting at column c0 (global index): restore the previous batch's component (exact restore, as the
// reference writes the saved value back :557,:598) and set X[b][c0+b] = x[c0+b] + sign*eps[b_local].
__global__ void __launch_bounds__(kThreads)
set_components(const double *__restrict__ x, const double *__restrict__ eps_local /* indexed by local column */,
               int64_t col0_local, int64_t c0, int64_t prev_c0, int32_t B, int32_t prevB, int64_t ldx, double sign,
               double *__restrict__ X) {
  const int b = blockIdx.x * kThreads + threadIdx.x;
  if (b < prevB) {
    const int64_t pc = prev_c0 + b;
    // restore unless this slot is about to be overwritten by the new perturbation of the same component
    X[(int64_t)b * ldx + pc] = x[pc];
  }
  __syncthreads();
  if (b < B) {
    const int64_t c = c0 + b;
    X[(int64_t)b * ldx + c] = x[c] + sign * eps_local[col0_local + b];
  }
}

}  // namespace fdb
