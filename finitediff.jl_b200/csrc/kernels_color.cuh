// kernels_color.cuh — on-device colouring: the `ArrayInterface.matrix_colors(A)` step that precedes the hot path
// (test/coloring_tests.jl:112,117 obtain `colorvec` that way; the reference package itself never colours).
//
//   structured types  closed forms, exactly what ArrayInterface returns: Tridiagonal -> 1,2,3,1,2,3,...;
//                     BandedMatrix(l,u) -> cycle 1:(l+u+1); Bidiagonal -> 1,2,1,2,...; Diagonal -> all 1
//   SparseMatrixCSC   a valid distance-2 colouring of the columns (no two columns of one colour share a row — the
//                     property the decompression `J[r,c] = vfx[r]` relies on), by a deterministic Jones–Plassmann
//                     sweep: in every round the uncoloured columns whose hashed priority beats all their uncoloured
//                     2-hop neighbours take the smallest colour none of their coloured neighbours holds.  Columns
//                     coloured in one round are never neighbours, so rounds read only settled colours: the result
//                     depends on the pattern alone (not on timing, launch geometry or the GPU count).
// Solvers that resize! their problem (jacobians.jl:655-661) can therefore recolour and re-plan without a host round trip.
#pragma once
#include "common.cuh"

namespace fdb {

// colorvec[j] = (j mod period) + 1   (Int64, 1-based, as Julia holds it)
__global__ void __launch_bounds__(kThreads)
cyclic_colors(int64_t n, int64_t period, int64_t *__restrict__ colorvec) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t j = blockIdx.x * (int64_t)kThreads + threadIdx.x; j < n; j += stride) colorvec[j] = j % period + 1;
}

// ---- CSC -> CSR (row -> columns) of the pattern: counts, then fill with a cursor per row
__global__ void __launch_bounds__(kThreads)
csr_count_rows(const int64_t *__restrict__ rowval, int64_t nnz, int64_t m, int32_t *__restrict__ row_cnt, uint32_t *__restrict__ err) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t p = blockIdx.x * (int64_t)kThreads + threadIdx.x; p < nnz; p += stride) {
    const int64_t r = rowval[p];
    if (r < 1 || r > m) { atomicOr(err, 2u); continue; }
    atomicAdd(row_cnt + (r - 1), 1);
  }
}

__global__ void __launch_bounds__(kThreads)
csr_fill(const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval, int64_t n, int64_t m,
         const int32_t *__restrict__ rowptr, int32_t *__restrict__ cursor, int32_t *__restrict__ rcols) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t c = blockIdx.x * (int64_t)kThreads + threadIdx.x; c < n; c += stride) {
    for (int64_t p = colptr[c] - 1; p < colptr[c + 1] - 1; ++p) {
      const int64_t r = rowval[p] - 1;
      if (r < 0 || r >= m) continue;
      const int32_t q = atomicAdd(cursor + r, 1);
      rcols[rowptr[r] + q] = (int32_t)c;
    }
  }
}

__device__ __forceinline__ uint32_t color_priority(uint32_t c) {
  uint32_t z = c * 0x9E3779B9u + 0x7F4A7C15u;
  z = (z ^ (z >> 16)) * 0x85EBCA6Bu;
  z = (z ^ (z >> 13)) * 0xC2B2AE35u;
  return z ^ (z >> 16);
}

// One Jones–Plassmann round.  color[c] == 0: uncoloured.  `pending` counts the columns still uncoloured after the round.
// Two-buffer scheme: reads `color_in` (settled state of the previous rounds), writes `color_out`.
__global__ void __launch_bounds__(kThreads)
jp_color_round(const int64_t *__restrict__ colptr, const int64_t *__restrict__ rowval, const int32_t *__restrict__ rowptr,
               const int32_t *__restrict__ rcols, int64_t n, int64_t m, const int32_t *__restrict__ color_in,
               int32_t *__restrict__ color_out, unsigned long long *__restrict__ pending) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  unsigned long long left = 0;
  for (int64_t c = blockIdx.x * (int64_t)kThreads + threadIdx.x; c < n; c += stride) {
    const int32_t mine = color_in[c];
    if (mine != 0) { color_out[c] = mine; continue; }
    const uint32_t pc = color_priority((uint32_t)c);
    const int64_t p0 = colptr[c] - 1, p1 = colptr[c + 1] - 1;
    bool is_max = true;
    for (int64_t p = p0; p < p1 && is_max; ++p) {
      const int64_t r = rowval[p] - 1;
      if (r < 0 || r >= m) continue;
      for (int32_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
        const int32_t o = rcols[q];
        if (o == c || color_in[o] != 0) continue;
        const uint32_t po = color_priority((uint32_t)o);
        if (po > pc || (po == pc && o > c)) { is_max = false; break; }
      }
    }
    if (!is_max) { color_out[c] = 0; ++left; continue; }
    // smallest colour >= 1 not held by a coloured 2-hop neighbour: 64-colour windows
    int32_t chosen = 0;
    for (int32_t base = 0; chosen == 0; base += 64) {
      unsigned long long used = 0;
      for (int64_t p = p0; p < p1; ++p) {
        const int64_t r = rowval[p] - 1;
        if (r < 0 || r >= m) continue;
        for (int32_t q = rowptr[r]; q < rowptr[r + 1]; ++q) {
          const int32_t k = color_in[rcols[q]];
          if (k > base && k <= base + 64) used |= 1ull << (k - base - 1);
        }
      }
      if (used != ~0ull) chosen = base + __ffsll((long long)~used);
    }
    color_out[c] = chosen;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) left += __shfl_xor_sync(0xffffffffu, left, o);
  if ((threadIdx.x & 31) == 0 && left) atomicAdd(pending, left);
}

__global__ void __launch_bounds__(kThreads)
colors_to_i64(const int32_t *__restrict__ color, int64_t n, int64_t *__restrict__ colorvec, int *__restrict__ max_color) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  int mx = 0;
  for (int64_t c = blockIdx.x * (int64_t)kThreads + threadIdx.x; c < n; c += stride) {
    const int32_t k = color[c];
    colorvec[c] = k;
    mx = k > mx ? k : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const int t = __shfl_xor_sync(0xffffffffu, mx, o); mx = t > mx ? t : mx; }
  if ((threadIdx.x & 31) == 0) atomicMax(max_color, mx);
}

// validity: number of (row, colour) collisions — pairs of entries of one row whose columns share a VALID colour
__global__ void __launch_bounds__(kThreads)
count_color_conflicts(const int32_t *__restrict__ rowptr, const int32_t *__restrict__ rcols, int64_t m,
                      const int64_t *__restrict__ colorvec, unsigned long long *__restrict__ conflicts) {
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  unsigned long long bad = 0;
  for (int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x; r < m; r += stride) {
    for (int32_t a = rowptr[r]; a < rowptr[r + 1]; ++a) {
      const int64_t ka = colorvec[rcols[a]];
      if (ka < 1) continue;
      for (int32_t b = a + 1; b < rowptr[r + 1]; ++b) bad += colorvec[rcols[b]] == ka ? 1 : 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bad += __shfl_xor_sync(0xffffffffu, bad, o);
  if ((threadIdx.x & 31) == 0 && bad) atomicAdd(conflicts, bad);
}

}  // namespace fdb
