// kernels_scatter.cuh — K5: the fused divided-difference + decompression ("diff+scatter") kernels.
//
// Reference, per colour k (jacobians.jl:565-572 / :607-614 / :636-643):
//     @. vfx1 = (vfx1 - vfx) / eps_k            (a full-length pass, in place; complex step: vfx = imag(vfx)/eps)
//     _colorediteration!(J, ..., vfx1, colorvec, k, n)   -> for every structural entry (r,c) with colorvec[c]==k:
//                                                           J[r,c] = vfx1[r]
// (CSC same-pattern: ext/FiniteDiffSparseArraysExt.jl:38-47; CSC->J[r,c]: :20-28; COO: src/iteration_utils.jl:25-32;
//  banded whole-band: ext/FiniteDiffBandedMatricesExt.jl:13-27; dense column: jacobians.jl:555,597,630.)
//
// B200 formulation: the f! outputs of the colours of a group stay resident as slabs F[slab][m]; ONE launch walks J's
// value storage in storage order, and for every structural entry e (row r_e, colour k_e) computes
//     J[dest(e)] = (F[slab(k_e)][r_e] - fx[r_e]) / eps[k_e]          (central: (Fp - Fm) / (2 eps); complex: imag(F)/eps)
// at gather time — the divided difference is never materialised, index/colour streams are read once, fully
// coalesced, and J is written once with full sectors (per-colour launches would touch every sector of nzval C times).
// IEEE subtraction and division are the same operations the reference performs, so values are bit-identical
// given the same f! outputs and eps.
#pragma once
#include "common.cuh"

namespace fdb {

// finite-difference flavour of a kernel instantiation
enum : int { kForward = 0, kCentral = 1, kComplex = 2, kCopy = 3 /* slab already holds the quotients */ };

// The quotient the reference stores.  `hi` = the colour's slab (f(x+eps e_k); complex step: complex128 interleaved,
// so the imaginary part of row r sits at 2r+1 and ldF counts doubles), `lo` = f(x) (forward) / the minus slab (central).
template <int MODE>
__device__ __forceinline__ double fd_quotient(const double *__restrict__ hi, const double *__restrict__ lo, int64_t r,
                                              double e) {
  if (MODE == kCopy) return __ldg(hi + r);                           // quotient precomputed in place (diff_slabs)
  if (MODE == kComplex) return __ldg(hi + 2 * r + 1) / e;            // jacobians.jl:636  imag(vfx) / epsilon
  const double d = __ldg(hi + r) - __ldg(lo + r);
  return d / (MODE == kCentral ? 2 * e : e);                         // :565 (vfx1-vfx)/epsilon ; :607 .../2epsilon
}

// r1 tuning on B200 (C2, ncu): 8 resident blocks/SM (32 registers) 133 us, 6 blocks (40 regs) 150 us, 5 blocks 165 us
constexpr int kScatterMinBlocks = 8;
constexpr int kSmemTable = 1024;       // colours whose (slab, eps) tables are staged in (dynamic) shared memory

struct ScatterArgs {
  const int32_t *row;        // [E] 0-based row of entry e
  const void *ecolor;        // [E] 0-based colour of the entry's column (CT)
  const int64_t *dest;       // [E] destination offset, or null => identity (nzval[e])
  const double *fx;          // forward: vfx = f(x) [m]; otherwise unused
  const double *Fp;          // slabs [G][ldF]: f(x + eps_k e_k)   (complex step: complex128 slabs, ldF in doubles)
  const double *Fm;          // central: slabs f(x - eps_k e_k)
  const double *eps;         // [C]
  const int32_t *local_of;   // [C] colour -> local index on this rank, -1 if not owned
  double *J;
  double *const *peers;      // optional peer J buffers (multi-GPU fused gather), device array
  int32_t n_peers;
  int32_t C;
  int32_t l0, G;             // this launch covers local colours [l0, l0+G)
  int32_t write_invalid_zero;// entries whose column has no valid colour get 0 (fill_matrix! semantics)
  int32_t j_aligned;         // J (and every peer) 16-byte aligned
  int32_t hi_stream;         // random patterns: the slab gathers are read-once -> evict-first loads, so that they do not
                             // push f(x) (re-read by every colour) out of L2
  int64_t ldF;
  int64_t E;
};

// shared-memory tables: slab index (or -1) and eps per colour
struct ScatterTables {
  const int32_t *slab;   // null => read local_of / eps from global memory
  const double *eps;
};

__device__ __forceinline__ ScatterTables scatter_tables(const ScatterArgs &a, unsigned char *smem) {
  ScatterTables t{nullptr, nullptr};
  if (a.C <= kSmemTable) {
    double *s_eps = reinterpret_cast<double *>(smem);
    int32_t *s_slab = reinterpret_cast<int32_t *>(smem + sizeof(double) * a.C);
    for (int i = threadIdx.x; i < a.C; i += kThreads) {
      const int32_t lo = a.local_of[i];
      int32_t sl = lo < 0 ? -1 : lo - a.l0;
      if (sl >= a.G) sl = -1;
      s_slab[i] = sl;
      s_eps[i] = a.eps[i];
    }
    __syncthreads();
    t.slab = s_slab;
    t.eps = s_eps;
  }
  return t;
}

// value of one structural entry; returns false when this launch does not own the entry
template <int MODE>
__device__ __forceinline__ bool entry_value(const ScatterArgs &a, const ScatterTables &t, int32_t r, uint32_t k, double &v) {
  v = 0.0;
  if (k >= (uint32_t)a.C) return a.write_invalid_zero != 0;
  int32_t slab;
  double e;
  if (t.slab) { slab = t.slab[k]; e = t.eps[k]; }
  else {
    const int32_t lo = __ldg(a.local_of + k);
    slab = lo < 0 ? -1 : lo - a.l0;
    if (slab >= a.G) slab = -1;
    e = __ldg(a.eps + k);
  }
  if (slab < 0) return false;
  v = fd_quotient<MODE>(a.Fp + (int64_t)slab * a.ldF, MODE == kCentral ? a.Fm + (int64_t)slab * a.ldF : a.fx, r, e);
  return true;
}

// Identity destination (CSC nzval, same pattern): THE graded kernel.  Tile/pair mapping of common.cuh: per block step
// each lane owns two entry pairs; row ids arrive as one 8-byte load per pair, colours as one narrow load per pair, the
// four values leave as two 16-byte stores — every warp-level access is a single contiguous run.
//   FULL = single group on a single rank: every valid colour is resident (slab == colour), no ownership test.
template <int MODE, bool FULL>
__device__ __forceinline__ bool ident_value(const ScatterArgs &a, const ScatterTables &t, int32_t r, uint32_t k, double &v) {
  if (FULL) {
    v = 0.0;
    if (k >= (uint32_t)a.C) return true;            // column without a valid colour: stays 0 (fill_matrix!)
    const double e = t.eps ? t.eps[k] : __ldg(a.eps + k);
    if (MODE == kForward && a.hi_stream) {
      const double d = __ldcs(a.Fp + (int64_t)k * a.ldF + r) - __ldg(a.fx + r);
      v = d / e;
      return true;
    }
    v = fd_quotient<MODE>(a.Fp + (int64_t)k * a.ldF, MODE == kCentral ? a.Fm + (int64_t)k * a.ldF : a.fx, r, e);
    return true;
  }
  return entry_value<MODE>(a, t, r, k, v);
}

template <typename CT, int MODE, bool FULL, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
diff_scatter_ident(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ScatterTables t = scatter_tables(a, smem);
  const CT *__restrict__ ecolor = reinterpret_cast<const CT *>(a.ecolor);
  constexpr int kHalf = kTile / 2;
  const int64_t nfull = a.E / kTile;
  const int tid2 = 2 * threadIdx.x;
  // ---- full tiles: no bounds checks, 32-bit offsets from the tile base
  for (int64_t tile = blockIdx.x; tile < nfull; tile += gridDim.x) {
    const int32_t *__restrict__ rt = a.row + tile * kTile;
    const CT *__restrict__ ct = ecolor + tile * kTile;
    double *__restrict__ Jt = a.J + tile * kTile;
    const int2 ra = __ldcs(reinterpret_cast<const int2 *>(rt + tid2));
    const int2 rb = __ldcs(reinterpret_cast<const int2 *>(rt + kHalf + tid2));
    uint32_t ka0, ka1, kb0, kb1;
    ld_color_pair<CT>(ct + tid2, ka0, ka1);
    ld_color_pair<CT>(ct + kHalf + tid2, kb0, kb1);
    double va0, va1, vb0, vb1;
    const bool wa0 = ident_value<MODE, FULL>(a, t, ra.x, ka0, va0);
    const bool wa1 = ident_value<MODE, FULL>(a, t, ra.y, ka1, va1);
    const bool wb0 = ident_value<MODE, FULL>(a, t, rb.x, kb0, vb0);
    const bool wb1 = ident_value<MODE, FULL>(a, t, rb.y, kb1, vb1);
    if (FULL) {
      if (a.j_aligned) {
        st_stream2(Jt + tid2, va0, va1);
        st_stream2(Jt + kHalf + tid2, vb0, vb1);
      } else {
        Jt[tid2] = va0; Jt[tid2 + 1] = va1; Jt[kHalf + tid2] = vb0; Jt[kHalf + tid2 + 1] = vb1;
      }
    } else {
      const int64_t ea = tile * kTile + tid2, eb = ea + kHalf;
      if (wa0 && wa1 && a.j_aligned) {
        st_stream2(Jt + tid2, va0, va1);
        for (int p = 0; p < a.n_peers; ++p) *reinterpret_cast<double2 *>(a.peers[p] + ea) = make_double2(va0, va1);
      } else {
        if (wa0) { Jt[tid2] = va0; for (int p = 0; p < a.n_peers; ++p) a.peers[p][ea] = va0; }
        if (wa1) { Jt[tid2 + 1] = va1; for (int p = 0; p < a.n_peers; ++p) a.peers[p][ea + 1] = va1; }
      }
      if (wb0 && wb1 && a.j_aligned) {
        st_stream2(Jt + kHalf + tid2, vb0, vb1);
        for (int p = 0; p < a.n_peers; ++p) *reinterpret_cast<double2 *>(a.peers[p] + eb) = make_double2(vb0, vb1);
      } else {
        if (wb0) { Jt[kHalf + tid2] = vb0; for (int p = 0; p < a.n_peers; ++p) a.peers[p][eb] = vb0; }
        if (wb1) { Jt[kHalf + tid2 + 1] = vb1; for (int p = 0; p < a.n_peers; ++p) a.peers[p][eb + 1] = vb1; }
      }
    }
  }
  // ---- the last, partial tile (E % kTile entries): scalar, bounds-checked; one block takes it
  const int64_t rem0 = nfull * kTile;
  if (rem0 < a.E && blockIdx.x == (unsigned)(nfull % gridDim.x)) {
    for (int64_t e = rem0 + threadIdx.x; e < a.E; e += kThreads) {
      double v;
      if (ident_value<MODE, FULL>(a, t, a.row[e], (uint32_t)ecolor[e], v)) {
        a.J[e] = v;
        for (int p = 0; p < a.n_peers; ++p) a.peers[p][e] = v;
      }
    }
  }
}

__device__ __forceinline__ void store_peers(const ScatterArgs &a, int64_t off, double v) {
  for (int p = 0; p < a.n_peers; ++p) a.peers[p][off] = v;
}

// Explicit destination per entry (CSC sparsity -> dense / other-pattern CSC J, COO -> dense J or slots).
template <typename CT, int MODE>
__global__ void __launch_bounds__(kThreads)
diff_scatter_dest(const ScatterArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const ScatterTables t = scatter_tables(a, smem);
  const CT *__restrict__ ecolor = reinterpret_cast<const CT *>(a.ecolor);
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t e = blockIdx.x * (int64_t)kThreads + threadIdx.x; e < a.E; e += stride) {
    double v;
    // no zero-writes here: J was zero-filled (fill_matrix!) before the first group
    const uint32_t k = (uint32_t)ecolor[e];
    if (k < (uint32_t)a.C && entry_value<MODE>(a, t, __ldcs(a.row + e), k, v)) {
      const int64_t d = __ldcs(a.dest + e);
      a.J[d] = v;
      store_peers(a, d, v);
    }
  }
}

// ---- colour-major entry lists (CSC) ----
// The literal shape of ext/FiniteDiffSparseArraysExt.jl:38-47 — "for every column of colour k, for every stored entry:
// nzval[p] = vfx[rowval[p]]" — with the column test and the colptr walk done ONCE, at plan time: the entries of the
// colours this rank evaluates are laid out colour by colour (local colour order), column by column inside a colour, as
// two coalesced streams  cm_row[q] (row of the entry)  and  cm_slot[q] (where its value goes in J).  A launch covers the
// contiguous range of one GROUP of colours (those whose f! outputs are resident), so it only ever streams entries it
// owns: this is the form used when colours are sharded over GPUs (the stores to the peers ride in the same kernel) and
// when there are more colours than resident slabs (the slab is gathered while it is still in L2).
// Dependent-load depth is 2 (index stream -> slab gather); r1's per-column form walked cols -> colptr -> row -> slab
// (depth 4) and ran at 0.10 of the HBM roofline, one launch per colour.
struct CmArgs {
  const int32_t *row;          // [E_local] colour-major
  const void *slot;            // [E_local] int32 (nzval slot) or int64 (explicit destination)
  const int64_t *seg_start;    // [n_local + 1] first entry of every local colour
  const int32_t *local_colors; // [n_local] global colour id of every local colour
  const double *fx, *Fp, *Fm, *eps;
  const double *fx_cm;         // forward: f(x) gathered ONCE per Jacobian into colour-major order (gather_fx_cm), or null
  double *J;
  double *const *peers;
  int32_t n_peers;
  int32_t l0, G;               // this launch: local colours [l0, l0+G); slab of local colour li is li - l0
  int32_t prefetch_next;       // several resident colours in one launch: while the tiles of colour s are processed, the
                               // grid streams slab s+1 into L2 (prefetch.global.L2, one pass, sequential) — the random
                               // 8-byte gathers of the next colour then hit L2 instead of fetching a DRAM granule each
  int32_t slab_stream;         // forward, f(x) gathered by row: the slab values are read once -> evict-first loads (ld.cs), so
                               // that they do not push f(x) — re-read by every colour — out of L2 (r2 A/B 9, C4 one launch:
                               // 1.176 -> 1.020 ms; createpolicy evict-last on f(x) + no-L1 evict-first on the slab: 1.075)
  int64_t m;                   // rows of a slab (prefetch extent)
  int64_t ldF;
};

constexpr int kCmMaxGroup = 1024;          // colours per launch whose (start, eps) tables fit the static shared arrays
constexpr int kCmPerThread = 4;
constexpr int kCmTile = kThreads * kCmPerThread;

template <int MODE, typename ST>
__global__ void __launch_bounds__(kThreads)
diff_scatter_cm(const CmArgs a) {
  __shared__ int64_t s_start[kCmMaxGroup + 1];
  __shared__ double s_eps[kCmMaxGroup];
  for (int i = threadIdx.x; i <= a.G; i += kThreads) s_start[i] = a.seg_start[a.l0 + i];
  for (int i = threadIdx.x; i < a.G; i += kThreads) s_eps[i] = a.eps[a.local_colors[a.l0 + i]];
  __syncthreads();
  const ST *__restrict__ slot = reinterpret_cast<const ST *>(a.slot);
  const int64_t q0 = s_start[0], q1 = s_start[a.G];
  for (int64_t tile = q0 + (int64_t)blockIdx.x * kCmTile; tile < q1; tile += (int64_t)gridDim.x * kCmTile) {
    int32_t r[kCmPerThread];
    ST d[kCmPerThread];
    double lo_v[kCmPerThread];
    const bool lo_stream = MODE == kForward && a.fx_cm != nullptr;
#pragma unroll
    for (int u = 0; u < kCmPerThread; ++u) {
      const int64_t q = tile + u * kThreads + threadIdx.x;
      r[u] = 0; d[u] = 0; lo_v[u] = 0.0;
      if (q < q1) {
        r[u] = __ldcs(a.row + q); d[u] = __ldcs(slot + q);
        if (lo_stream) lo_v[u] = __ldcs(a.fx_cm + q);                       // coalesced; replaces the random fx[row] gather
      }
    }
    // segment of the tile's first entry (uniform binary search), then at most a few steps forward per entry
    int seg = 0;
    {
      int lo = 0, hi = a.G;   // s_start[lo] <= tile < s_start[hi]
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (s_start[mid] <= tile) lo = mid; else hi = mid; }
      seg = lo;
    }
    if (a.prefetch_next && seg + 1 < a.G) {
      // this tile's share of the next colour's slab(s): tile j of T in the segment takes lines [j*L, (j+1)*L)
      const int64_t seg_len = s_start[seg + 1] - s_start[seg];
      const int64_t T = (seg_len + kCmTile - 1) / kCmTile;
      const int64_t j = (tile - s_start[seg]) / kCmTile;
      const int64_t lines = (a.m * 8 + 127) / 128;
      const int64_t L = (lines + T - 1) / T;
      const char *nxt = reinterpret_cast<const char *>(a.Fp + (int64_t)(seg + 1) * a.ldF);
      const char *nxm = MODE == kCentral ? reinterpret_cast<const char *>(a.Fm + (int64_t)(seg + 1) * a.ldF) : nullptr;
      for (int64_t ln = j * L + threadIdx.x; ln < (j + 1) * L && ln < lines; ln += kThreads) {
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + ln * 128));
        if (MODE == kCentral) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxm + ln * 128));
      }
    }
    double v[kCmPerThread];
#pragma unroll
    for (int u = 0; u < kCmPerThread; ++u) {
      const int64_t q = tile + u * kThreads + threadIdx.x;
      v[u] = 0.0;
      if (q < q1) {
        while (q >= s_start[seg + 1]) ++seg;
        const double *hi_p = a.Fp + (int64_t)seg * a.ldF;
        const double *lo_p = MODE == kCentral ? a.Fm + (int64_t)seg * a.ldF : a.fx;
        if (lo_stream)
          v[u] = (__ldcs(hi_p + r[u]) - lo_v[u]) / s_eps[seg];              // same IEEE subtraction and division
        else if (MODE == kForward && (a.slab_stream || a.prefetch_next))     // read-once slab values: evict-first, f(x) stays
          v[u] = (__ldcs(hi_p + r[u]) - __ldg(lo_p + r[u])) / s_eps[seg];
        else
          v[u] = fd_quotient<MODE>(hi_p, lo_p, r[u], s_eps[seg]);          // fused with ext/..SparseArraysExt.jl:44
      }
    }
#pragma unroll
    for (int u = 0; u < kCmPerThread; ++u) {
      const int64_t q = tile + u * kThreads + threadIdx.x;
      if (q < q1) {
        a.J[d[u]] = v[u];
        for (int p = 0; p < a.n_peers; ++p) a.peers[p][d[u]] = v[u];
      }
    }
  }
}

// f(x) in colour-major order, once per Jacobian (forward mode): every colour's launch then reads its f(x) values as a
// coalesced 8-byte stream instead of dragging the whole f(x) vector through DRAM again (random rows: a colour's launch
// touched >= 2/3 of f(x)'s 64-byte granules — r2 ncu: 77 MB read per colour for 625 k entries, half of it f(x)).
__global__ void __launch_bounds__(kThreads)
gather_fx_cm(const int32_t *__restrict__ cm_row, const double *__restrict__ fx, int64_t count, double *__restrict__ fx_cm) {
  const int64_t stride = (int64_t)gridDim.x * kThreads * 4;
  for (int64_t q0 = (int64_t)blockIdx.x * kThreads * 4 + threadIdx.x; q0 < count; q0 += stride) {
    int32_t r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int64_t q = q0 + u * kThreads; r[u] = q < count ? __ldcs(cm_row + q) : 0; }
    double v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldg(fx + r[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int64_t q = q0 + u * kThreads; if (q < count) fx_cm[q] = v[u]; }
  }
}

// columns without a valid colour (colorvec[col] < 1): their entries stay 0 (fill_matrix! semantics); rank 0 writes them
template <typename ST>
__global__ void __launch_bounds__(kThreads)
zero_slots(const void *slot_v, int64_t q0, int64_t q1, double *__restrict__ J, double *const *peers, int32_t n_peers) {
  const ST *__restrict__ slot = reinterpret_cast<const ST *>(slot_v);
  for (int64_t q = q0 + blockIdx.x * (int64_t)kThreads + threadIdx.x; q < q1; q += (int64_t)gridDim.x * kThreads) {
    const int64_t d = slot[q];
    J[d] = 0.0;
    for (int p = 0; p < n_peers; ++p) peers[p][d] = 0.0;
  }
}

// ---- banded: ext/FiniteDiffBandedMatricesExt.jl:13-27 ----
// Every in-band (r,c), r in [max(1,c-u), min(m,c+l)], receives vfx[r] of column c's colour: the destination is the
// contiguous band column data[:,c] (slot u+r-c, 0-based) and the source rows are contiguous too, so this is a
// pure streaming transform: coalesced gather from the (L2-resident) slab, coalesced full-sector stores.
struct BandArgs {
  const void *jcolor;        // [n] colour of column c (CT)
  const double *fx, *Fp, *Fm, *eps;
  const int32_t *local_of;
  double *J;
  int32_t C, l0, G;
  int32_t write_other;       // first launch / single group: also define slots this launch does not own (0)
  int32_t to_dense;          // 1: J is dense column-major (ldJ), only in-matrix slots are written
  int64_t ldF, ldJ;
  int64_t m, n, l, u;
  int64_t cols_per_tile;     // narrow-band tiling: whole columns per block step
};

__device__ __forceinline__ void band_color_lookup(const BandArgs &a, uint32_t k, bool tables, const int32_t *s_slab,
                                                  const double *s_eps, int32_t &slab, double &e) {
  slab = -1;
  e = 1.0;
  if (k < (uint32_t)a.C) {
    if (tables) { slab = s_slab[k]; e = s_eps[k]; }
    else {
      const int32_t lo = __ldg(a.local_of + k);
      slab = lo < 0 ? -1 : lo - a.l0;
      if (slab >= a.G) slab = -1;
      e = __ldg(a.eps + k);
    }
  }
}

// The reference's own in-place pass `@. vfx1 = (vfx1 - vfx) / epsilon` (jacobians.jl:565 / :607), for the slabs of one
// group.  Only used where a quotient is stored MANY times: the whole-band fill of an under-coloured banded sparsity
// writes every (colour, row) value into (l+u+1)/C columns (C3: 400x) — dividing once per (colour, row) instead of once
// per slot removes 2*10^9 fp64 divisions from the 16 GB store stream.  The scatter then runs in kCopy mode.
template <int MODE>
__global__ void __launch_bounds__(kThreads)
diff_slabs(double *__restrict__ Fp, const double *__restrict__ Fm_or_fx, const double *__restrict__ eps,
           const int32_t *__restrict__ slab_color /* global colour of each slab of the group */, int64_t m, int64_t ldF) {
  const int s = blockIdx.y;
  const double e = __ldg(eps + __ldg(slab_color + s));
  const double denom = MODE == kCentral ? 2 * e : e;
  double *__restrict__ hi = Fp + (int64_t)s * ldF;
  const double *__restrict__ lo = MODE == kCentral ? Fm_or_fx + (int64_t)s * ldF : Fm_or_fx;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  for (int64_t r = blockIdx.x * (int64_t)kThreads + threadIdx.x; r < m; r += stride) hi[r] = (hi[r] - lo[r]) / denom;
}

// Wide bands (l+u+1 >= 64): ONE WARP PER COLUMN.  A band column is a contiguous run of l+u+1 slots whose sources
// F[slab][c-u .. c+l] are contiguous too; the warp walks it 32 slots at a time (256-byte coalesced stores, coalesced
// L2-resident gathers).  The per-column work (colour -> slab, eps) is one uniform lookup from shared-memory tables,
// amortised over the whole column (r1: the per-1024-slot tile version spent most of its time in that dependent-load
// prologue: 8.1 ms for C3 vs 3.4 ms; see profiles/).
template <typename CT, int MODE>
__global__ void __launch_bounds__(kThreads, 6)
diff_scatter_band_wide(const BandArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *s_eps = reinterpret_cast<double *>(smem);
  int32_t *s_slab = reinterpret_cast<int32_t *>(smem + sizeof(double) * (a.C <= kSmemTable ? a.C : 0));
  const bool tables = a.C <= kSmemTable;
  if (tables) {
    for (int i = threadIdx.x; i < a.C; i += kThreads) {
      const int32_t lo = a.local_of[i];
      int32_t sl = lo < 0 ? -1 : lo - a.l0;
      if (sl >= a.G) sl = -1;
      s_slab[i] = sl;
      s_eps[i] = a.eps[i];
    }
    __syncthreads();
  }
  const CT *__restrict__ jcolor = reinterpret_cast<const CT *>(a.jcolor);
  const int lane = threadIdx.x & 31;
  const int64_t w = a.l + a.u + 1;
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  for (int64_t c = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); c < a.n; c += nwarps) {
    const uint32_t k = (uint32_t)jcolor[c];
    int32_t slab;
    double e;
    band_color_lookup(a, k, tables, s_slab, s_eps, slab, e);
    const bool owned = slab >= 0;
    const bool zero_col = k >= (uint32_t)a.C && a.write_other && !a.to_dense;   // no valid colour: stays 0 (fill_matrix!)
    if (!owned && !zero_col) continue;                                         // another group's / rank's column
    const double *__restrict__ hi = a.Fp + (int64_t)(owned ? slab : 0) * a.ldF;
    const double *__restrict__ lo = MODE == kCentral ? a.Fm + (int64_t)(owned ? slab : 0) * a.ldF : a.fx;
    const int64_t r0 = c - a.u;                        // row of slot 0
    // in-matrix slots: rows r = r0+d in [0, m)  <=>  d in [d_lo, d_hi); the corner slots outside the matrix get 0
    const int64_t d_lo64 = -r0 > 0 ? -r0 : 0;
    const int64_t d_hi64 = a.m - r0 < w ? a.m - r0 : w;
    const int wi = (int)w, d_lo = (int)d_lo64, d_hi = (int)(d_hi64 > d_lo64 ? d_hi64 : d_lo64);
    if (a.to_dense) {
      double *__restrict__ out = a.J + c * a.ldJ + r0;
      if (owned)
        for (int d = d_lo + lane; d < d_hi; d += 32) out[d] = fd_quotient<MODE>(hi, lo, r0 + d, e);
    } else {
      double *__restrict__ out = a.J + c * w;
      for (int d = lane; d < d_lo; d += 32) st_stream(out + d, 0.0);
      if (owned) {
#pragma unroll 2
        for (int d = d_lo + lane; d < d_hi; d += 32) st_stream(out + d, fd_quotient<MODE>(hi, lo, r0 + d, e));
      } else {
        for (int d = d_lo + lane; d < d_hi; d += 32) st_stream(out + d, 0.0);
      }
      for (int d = d_hi + lane; d < wi; d += 32) st_stream(out + d, 0.0);
    }
  }
}

// Wide bands, band-data target, FLAT form: the (l+u+1) x n band storage is one contiguous stream (column c+1 starts
// right after column c).  Each warp takes 16-byte-aligned chunks of that stream (CH elements, CH <= l+u+1 so a chunk
// touches at most two columns) and writes them with fully aligned 16-byte stores — the warp-per-column form above
// starts every column at an 8-byte-aligned, sector-straddling address (w odd) and reached only ~4.5 of the 7.5 TB/s
// a pure store stream gets on this part (profiles/write_bw_probe.py).  One 64-bit division per chunk locates the column.
struct BandColumn {
  const double *hi, *lo;   // slab pointers already offset so that [d] addresses row c-u+d
  double e;
  int d_lo, d_hi;          // in-matrix slot range
  bool owned, zero_col;
};

template <typename CT, int MODE>
__device__ __forceinline__ BandColumn band_column(const BandArgs &a, const CT *__restrict__ jcolor, int64_t c, bool tables,
                                                  const int32_t *s_slab, const double *s_eps, int64_t w) {
  BandColumn b{};
  if (c >= a.n) return b;
  const uint32_t k = (uint32_t)jcolor[c];
  int32_t slab;
  band_color_lookup(a, k, tables, s_slab, s_eps, slab, b.e);
  b.owned = slab >= 0;
  b.zero_col = k >= (uint32_t)a.C && a.write_other;
  const int64_t r0 = c - a.u;
  b.hi = a.Fp + (int64_t)(b.owned ? slab : 0) * a.ldF + (MODE == kComplex ? 2 * r0 : r0);
  b.lo = (MODE == kCentral ? a.Fm + (int64_t)(b.owned ? slab : 0) * a.ldF : a.fx) + r0;
  const int64_t lo64 = -r0 > 0 ? -r0 : 0;
  const int64_t hi64 = a.m - r0 < w ? a.m - r0 : w;
  b.d_lo = (int)lo64;
  b.d_hi = (int)(hi64 > lo64 ? hi64 : lo64);
  return b;
}

template <int MODE>
__device__ __forceinline__ double band_value(const BandColumn &b, int d) {
  if (!b.owned || d < b.d_lo || d >= b.d_hi) return 0.0;          // corner slots / columns without a valid colour: 0
  if (MODE == kComplex) return __ldg(b.hi + 2 * d + 1) / b.e;
  if (MODE == kCopy) return __ldg(b.hi + d);
  const double df = __ldg(b.hi + d) - __ldg(b.lo + d);
  return df / (MODE == kCentral ? 2 * b.e : b.e);
}

// Interior stretch of a band column in copy mode: out[off] = src[off] for the even-aligned pairs of [s, e).  The
// sources are only 8-byte aligned relative to the destination, so they come as scalar L2 loads; kBandBatch pairs are
// loaded before the first store so that each lane keeps 64 bytes of loads in flight (the kernel is latency-bound
// otherwise: ncu showed ~60% of DRAM write bandwidth with every warp stalled on its one outstanding load pair).
constexpr int kBandBatch = 4;
__device__ __forceinline__ void band_copy_run(double *__restrict__ out, const double *__restrict__ src, int s, int e,
                                              int lane) {
  int off = s + 2 * lane;
  for (; off + 64 * (kBandBatch - 1) < e; off += 64 * kBandBatch) {
    double v[2 * kBandBatch];
#pragma unroll
    for (int u = 0; u < kBandBatch; ++u) {
      v[2 * u] = __ldg(src + off + 64 * u);
      v[2 * u + 1] = __ldg(src + off + 64 * u + 1);
    }
#pragma unroll
    for (int u = 0; u < kBandBatch; ++u) st_stream2(out + off + 64 * u, v[2 * u], v[2 * u + 1]);
  }
  for (; off < e; off += 64) st_stream2(out + off, __ldg(src + off), __ldg(src + off + 1));
}

template <typename CT, int MODE>
__global__ void __launch_bounds__(kThreads, 4)
diff_scatter_band_flat(const BandArgs a, int32_t CH /* elements per chunk: even, <= l+u+1 */) {
  extern __shared__ __align__(16) unsigned char smem[];
  double *s_eps = reinterpret_cast<double *>(smem);
  int32_t *s_slab = reinterpret_cast<int32_t *>(smem + sizeof(double) * (a.C <= kSmemTable ? a.C : 0));
  const bool tables = a.C <= kSmemTable;
  if (tables) {
    for (int i = threadIdx.x; i < a.C; i += kThreads) {
      const int32_t lo = a.local_of[i];
      int32_t sl = lo < 0 ? -1 : lo - a.l0;
      if (sl >= a.G) sl = -1;
      s_slab[i] = sl;
      s_eps[i] = a.eps[i];
    }
    __syncthreads();
  }
  const CT *__restrict__ jcolor = reinterpret_cast<const CT *>(a.jcolor);
  const int lane = threadIdx.x & 31;
  const int64_t w = a.l + a.u + 1;
  const int wi = (int)w;
  const int64_t total = w * a.n;
  const int64_t nchunks = (total + CH - 1) / CH;
  const int64_t nwarps = (int64_t)gridDim.x * (kThreads / 32);
  for (int64_t ch = (int64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); ch < nchunks; ch += nwarps) {
    const int64_t q0 = ch * CH;
    const int64_t c0 = q0 / w;
    const int d0 = (int)(q0 - c0 * w);
    const BandColumn A = band_column<CT, MODE>(a, jcolor, c0, tables, s_slab, s_eps, w);
    const BandColumn B = band_column<CT, MODE>(a, jcolor, c0 + 1, tables, s_slab, s_eps, w);
    const bool wA = A.owned || A.zero_col, wB = B.owned || B.zero_col;   // columns this launch defines
    int64_t lim = total - q0;
    if (lim > CH) lim = CH;
    double *__restrict__ out = a.J + q0;
    const int ilim = (int)lim;
    // generic pair loop over offsets [s, e): per-element column select, range tests, ownership
    auto generic = [&](int s, int e) {
      for (int off = s + 2 * lane; off < e; off += 64) {
        const int da = d0 + off, db = d0 + off + 1;
        const bool a_in_A = da < wi, b_in_A = db < wi;
        const double v0 = a_in_A ? band_value<MODE>(A, da) : band_value<MODE>(B, da - wi);
        const bool w0 = a_in_A ? wA : wB;
        if (off + 1 < e) {
          const double v1 = b_in_A ? band_value<MODE>(A, db) : band_value<MODE>(B, db - wi);
          const bool w1 = b_in_A ? wA : wB;
          if (w0 && w1) st_stream2(out + off, v0, v1);
          else { if (w0) out[off] = v0; if (w1) out[off + 1] = v1; }
        } else if (w0) {
          out[off] = v0;
        }
      }
    };
    if (MODE == kCopy) {
      // copy mode: the chunk is [0,nA) of column A then [nA,lim) of column B; interior stretches (owned column, all
      // slots inside the matrix) are plain aligned-store copies with no per-element tests: loads batch 4 deep
      const int nA = wi - d0 < ilim ? wi - d0 : ilim;
      const int nAe = nA & ~1;                           // even part of A; an odd nA leaves one straddling pair
      const int sB = (nA & 1) ? nA + 1 : nA;
      const int eB = ilim & ~1;
      const bool fastA = A.owned && d0 >= A.d_lo && d0 + nAe <= A.d_hi;
      const bool fastB = B.owned && sB < eB && (sB - nA) >= B.d_lo && (eB - nA) <= B.d_hi;
      if (fastA) {
        const double *__restrict__ src = A.hi + d0;
        band_copy_run(out, src, 0, nAe, lane);
      } else {
        generic(0, nAe);
      }
      if (nA & 1) generic(nAe, nAe + 2 < ilim ? nAe + 2 : ilim);
      if (fastB) {
        const double *__restrict__ src = B.hi - nA;        // slot d of B sits at offset nA + d
        band_copy_run(out, src, sB, eB, lane);
        if (eB < ilim) generic(eB, ilim);
      } else if (sB < ilim) {
        generic(sB, ilim);
      }
    } else {
      generic(0, ilim);
    }
  }
}

// Narrow bands: a tile = cols_per_tile whole columns, flat index inside the tile
template <typename CT, int MODE>
__global__ void __launch_bounds__(kThreads)
diff_scatter_band(const BandArgs a) {
  const CT *__restrict__ jcolor = reinterpret_cast<const CT *>(a.jcolor);
  const int64_t w = a.l + a.u + 1;
  const int64_t ntiles = (a.n + a.cols_per_tile - 1) / a.cols_per_tile;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int64_t c0 = tile * a.cols_per_tile;
    int64_t ncol = a.n - c0;
    if (ncol > a.cols_per_tile) ncol = a.cols_per_tile;
    const int32_t total = (int32_t)(ncol * w);
    const int32_t w32 = (int32_t)w;
    for (int32_t t = threadIdx.x; t < total; t += kThreads) {
      const int32_t cc = t / w32;
      const int32_t d = t - cc * w32;
      const int64_t c = c0 + cc;
      const int64_t r = c + d - a.u;
      const bool in = r >= 0 && r < a.m;
      const uint32_t k = (uint32_t)jcolor[c];
      int32_t slab;
      double e;
      band_color_lookup(a, k, false, nullptr, nullptr, slab, e);
      double v = 0.0;
      if (in && slab >= 0)
        v = fd_quotient<MODE>(a.Fp + (int64_t)slab * a.ldF, MODE == kCentral ? a.Fm + (int64_t)slab * a.ldF : a.fx, r, e);
      if (a.to_dense) {
        if (in && slab >= 0) a.J[c * a.ldJ + r] = v;
      } else if (slab >= 0) {
        a.J[c0 * w + t] = v;                       // owned column: whole data column (corner slots get 0)
      } else if (a.write_other && k >= (uint32_t)a.C) {
        a.J[c0 * w + t] = 0.0;                     // column without a valid colour: stays zero (fill_matrix!)
      }
    }
  }
}

// ---- dense column branch: J[:, c] = (fx1 - fx)/eps_c (jacobians.jl:555), (fx1 - fx_minus)/(2 eps_c) (:597),
//      imag(fx)/eps (:630) ----
template <int MODE, int DEPTH = 1>
__global__ void __launch_bounds__(kThreads)
diff_columns(const double *__restrict__ Fp, const double *__restrict__ Fm_or_fx, const double *__restrict__ eps_local,
             int64_t col0_local, int32_t B, int64_t m, int64_t ldF, int64_t ldJ, double *__restrict__ Jcols,
             int pairs_ok /* every column of hi / lo / out is 16-byte aligned */) {
  // grid.y = column within the batch, grid.x strides over rows
  const int b = blockIdx.y;
  if (b >= B) return;
  const double e = eps_local[col0_local + b];
  const double *hi = Fp + (int64_t)b * ldF;
  const double *lo = MODE == kCentral ? Fm_or_fx + (int64_t)b * ldF : Fm_or_fx;
  double *out = Jcols + (int64_t)b * ldJ;
  const int64_t stride = (int64_t)gridDim.x * kThreads;
  if (pairs_ok && MODE != kComplex) {
    // two rows per lane: 16-byte streaming loads of the slabs (read once) and 16-byte streaming stores of the column;
    // f(x) (forward) is re-read by every column and stays cacheable
    const double denom = MODE == kCentral ? 2 * e : e;
    const int64_t m2 = m >> 1;
    int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x;
    if (DEPTH > 1) {
      // DEPTH independent 16-byte loads of each slab in flight before the first quotient is formed
      for (; i + (DEPTH - 1) * stride < m2; i += DEPTH * stride) {
        double2 h[DEPTH], l[DEPTH];
#pragma unroll
        for (int q = 0; q < DEPTH; ++q) {
          h[q] = ld_stream2(hi + 2 * (i + q * stride));
          l[q] = MODE == kCentral ? ld_stream2(lo + 2 * (i + q * stride))
                                  : __ldg(reinterpret_cast<const double2 *>(lo + 2 * (i + q * stride)));
        }
#pragma unroll
        for (int q = 0; q < DEPTH; ++q)
          st_stream2(out + 2 * (i + q * stride), (h[q].x - l[q].x) / denom, (h[q].y - l[q].y) / denom);
      }
    }
    for (; i < m2; i += stride) {
      const double2 h = ld_stream2(hi + 2 * i);
      const double2 l = MODE == kCentral ? ld_stream2(lo + 2 * i) : __ldg(reinterpret_cast<const double2 *>(lo + 2 * i));
      st_stream2(out + 2 * i, (h.x - l.x) / denom, (h.y - l.y) / denom);
    }
    if ((m & 1) && blockIdx.x == 0 && threadIdx.x == 0) out[m - 1] = fd_quotient<MODE>(hi, lo, m - 1, e);
    return;
  }
  for (int64_t i = blockIdx.x * (int64_t)kThreads + threadIdx.x; i < m; i += stride)
    st_stream(out + i, fd_quotient<MODE>(hi, lo, i, e));
}

}  // namespace fdb
