// kernels_staged.cuh — K5s: the graded diff+scatter kernel with its slab gathers staged through shared memory by TMA.
//
// Same arithmetic as diff_scatter_ident<FULL> (kernels_scatter.cuh):  nzval[p] = (F[color(p)][row(p)] - fx[row(p)]) / eps
// walked in J's storage order — but for ROW-LOCAL patterns (tridiagonal, banded, stencil CSC: the rows of 1024 consecutive
// entries span a few hundred rows) the per-entry dependent gathers `row -> F[k][row]` are replaced by
//   * one bulk copy per resident slab and tile (`cp.async.bulk` global -> shared, completion on an mbarrier), issued
//     two tiles ahead by one thread: every slab row crosses the memory system exactly once, fully coalesced, and the
//     loads do not wait for the row indices;
//   * 16-bit row offsets relative to the tile's window (row16[e] = row[e] - tile_w0[tile]) instead of int32 rows:
//     2 bytes less per entry on the index stream.
// r1's ncu capture of the gather form (C2): DRAM 74.6 %, long-scoreboard 31.7 stalls per issue, L1 hit 32.8 % — a
// latency-limited dependent gather; here the only global loads a warp waits for are its own coalesced index pairs.
// Compulsory bytes per Jacobian: E*(2 + |colour| + 8) + 8*m*(slabs + 1)  (C2 forward: 650 MB; gather form: 710 MB).
#pragma once
#include "common.cuh"
#include "kernels_scatter.cuh"

namespace fdb {

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *sdst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
               "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

constexpr int kStageMaxWin = 8;        // windows per tile: C + 1 (forward: slabs + fx) or 2C (central: plus / minus slabs)
constexpr int kStagesMax = 3;          // tiles in flight per block (2 by default: 8 blocks/SM; 3 leaves 6)
constexpr int kStageMaxSmem = 46 * 1024;

struct StagedArgs {
  const uint16_t *row16;     // [E] row - tile_w0[tile]
  const void *ecolor;        // [E] (CT)
  const int32_t *tile_w0;    // [E / kTile] even first row of every full tile's window
  const int32_t *row32;      // [E] (the partial last tile takes the gather path)
  const double *fx, *Fp, *Fm, *eps;
  double *J;
  int32_t C, W;              // colours (slab == colour), window length in rows (even)
  int32_t stages;            // 2 or 3
  int64_t ldF, src_len;      // slab stride; readable doubles behind fx / every slab (even)
  int64_t E;
  int32_t j_aligned;
  int32_t reverse;           // walk the full tiles from the end of J's storage towards its start: the rows the last f! wrote
                             // most recently (the tail of the last slab) are read while they are still in L2
};

// plan time: window start (even) of every full tile, 16-bit offsets, largest span
template <typename CT>
__global__ void __launch_bounds__(kThreads)
stage_prepare(const int32_t *__restrict__ row32, int64_t ntiles, int32_t *__restrict__ tile_w0, uint16_t *__restrict__ row16,
              unsigned int *__restrict__ max_span, const CT *__restrict__ ecolor /* non-null: pack the colour into bits 12..15 */,
              int32_t C) {
  __shared__ int32_t s_min[kThreads / 32], s_max[kThreads / 32];
  __shared__ int32_t s_w0;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int32_t *rt = row32 + tile * kTile;
    int32_t r[kTile / kThreads];
    int32_t mn = INT_MAX, mx = INT_MIN;
#pragma unroll
    for (int u = 0; u < kTile / kThreads; ++u) {
      r[u] = rt[u * kThreads + threadIdx.x];
      mn = min(mn, r[u]);
      mx = max(mx, r[u]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { s_min[threadIdx.x >> 5] = mn; s_max[threadIdx.x >> 5] = mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int32_t a = s_min[0], b = s_max[0];
      for (int w = 1; w < kThreads / 32; ++w) { a = min(a, s_min[w]); b = max(b, s_max[w]); }
      a &= ~1;
      s_w0 = a;
      tile_w0[tile] = a;
      atomicMax(max_span, (unsigned int)(b - a + 1));
    }
    __syncthreads();
    const int32_t w0 = s_w0;
#pragma unroll
    for (int u = 0; u < kTile / kThreads; ++u) {
      int32_t d = r[u] - w0;
      d = d > 65535 ? 65535 : d;
      if (ecolor) {                                      // packed form: offset < 4096 in bits 0..11, colour (15 = none) above
        uint32_t k = (uint32_t)ecolor[tile * kTile + u * kThreads + threadIdx.x];
        if (k >= (uint32_t)C) k = 15u;
        d = (d & 0xFFF) | (int32_t)(k << 12);
      }
      row16[tile * kTile + u * kThreads + threadIdx.x] = (uint16_t)d;
    }
    __syncthreads();
  }
}

// MINB: resident blocks per SM the register budget is cut for (8 -> 32 registers, 6 -> 40: no spill of the prefetched
// index registers; with TMA doing the wide loads the kernel needs fewer resident warps than the gather form)
// PACKED: few colours and short windows (C <= 14, W <= 4096): the entry's colour rides in the top 4 bits of its 16-bit row
// offset — the per-entry colour stream is not read at all (2 index bytes per entry instead of 2 + |colour|)
// EMPTYBAR: a stage is handed back to the producer through a second mbarrier (one arrival per warp) instead of a block-wide
// __syncthreads(): warps that finished a tile go straight on to the next one (A/B variant)
template <typename CT, int MODE, int MINB, bool PREFETCH, bool PACKED, bool EMPTYBAR = false>
__global__ void __launch_bounds__(kThreads, MINB)
diff_scatter_staged(const StagedArgs a) {
  static_assert(MODE == kForward || MODE == kCentral, "staged scatter: forward / central");
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int C = a.C, W = a.W;
  const int nwin = MODE == kCentral ? 2 * C : C + 1;
  const int stage_elems = nwin * W;
  const int kStages = a.stages;
  double *buf = reinterpret_cast<double *>(smem_raw);                       // [stages][nwin][W]
  uint64_t *full = reinterpret_cast<uint64_t *>(buf + kStages * stage_elems);
  uint64_t *empty = full + kStagesMax;
  double *s_eps = reinterpret_cast<double *>(full + 2 * kStagesMax);        // [C]
  const CT *__restrict__ ecolor = reinterpret_cast<const CT *>(a.ecolor);
  const int64_t nfull = a.E / kTile;
  constexpr int kHalf = kTile / 2;
  const int tid2 = 2 * threadIdx.x;
  // position in the walk -> tile of J's storage (the walk order is free: every tile is independent)
  auto phys = [&](int64_t t) -> int64_t { return a.reverse ? nfull - 1 - t : t; };

  // one thread feeds the pipeline: nwin bulk copies per tile, all completing on the stage's mbarrier
  auto issue = [&](int64_t tile, int s) {
    const int32_t w0 = __ldg(a.tile_w0 + tile);
    int64_t len = a.src_len - w0;
    if (len > W) len = W;
    const uint32_t bytes = (uint32_t)len * 8u;
    double *dst = buf + s * stage_elems;
    mbar_arrive_expect_tx(full + s, bytes * (uint32_t)nwin);
    for (int k = 0; k < C; ++k) bulk_g2s(dst + k * W, a.Fp + (int64_t)k * a.ldF + w0, bytes, full + s);
    if (MODE == kCentral) {
      for (int k = 0; k < C; ++k) bulk_g2s(dst + (C + k) * W, a.Fm + (int64_t)k * a.ldF + w0, bytes, full + s);
    } else {
      bulk_g2s(dst + C * W, a.fx + w0, bytes, full + s);
    }
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kThreads / 32); }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < C; i += kThreads) s_eps[i] = a.eps[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      const int64_t t = (int64_t)blockIdx.x + (int64_t)s * gridDim.x;
      if (t < nfull) issue(phys(t), s);
    }
  }

  auto value = [&](const double *__restrict__ win, uint32_t r, uint32_t k) -> double {
    if (k >= (uint32_t)C) return 0.0;                    // column without a valid colour: stays 0 (fill_matrix!)
    const double e = s_eps[k];
    const double hi = win[k * W + r];
    const double lo = MODE == kCentral ? win[(C + k) * W + r] : win[C * W + r];
    const double d = hi - lo;                            // jacobians.jl:565 / :607 — same IEEE operations
    return d / (MODE == kCentral ? 2 * e : e);
  };

  // index pairs of a tile: coalesced, independent of the staged data — loaded ONE TILE AHEAD (software pipelining), so
  // their latency overlaps the previous tile's wait + arithmetic instead of following the block barrier
  struct Idx { ushort2 ra, rb; uint32_t ka0, ka1, kb0, kb1; };
  auto load_idx = [&](int64_t tile) {
    Idx x;
    const uint16_t *__restrict__ rt = a.row16 + tile * kTile;
    const CT *__restrict__ ct = ecolor + tile * kTile;
    x.ra = __ldcs(reinterpret_cast<const ushort2 *>(rt + tid2));
    x.rb = __ldcs(reinterpret_cast<const ushort2 *>(rt + kHalf + tid2));
    if (PACKED) {
      x.ka0 = x.ra.x >> 12; x.ka1 = x.ra.y >> 12; x.kb0 = x.rb.x >> 12; x.kb1 = x.rb.y >> 12;
      x.ra.x &= 0xFFF; x.ra.y &= 0xFFF; x.rb.x &= 0xFFF; x.rb.y &= 0xFFF;
    } else {
      ld_color_pair<CT>(ct + tid2, x.ka0, x.ka1);
      ld_color_pair<CT>(ct + kHalf + tid2, x.kb0, x.kb1);
    }
    return x;
  };
  uint32_t it = 0;
  int s = 0;
  uint32_t parity = 0;
  Idx nx{};
  if (PREFETCH && (int64_t)blockIdx.x < nfull) nx = load_idx(phys(blockIdx.x));
  for (int64_t pos = blockIdx.x; pos < nfull; pos += gridDim.x, ++it) {
    const int64_t tile = phys(pos);
    double *__restrict__ Jt = a.J + tile * kTile;
    const Idx cur = PREFETCH ? nx : load_idx(tile);
    if (PREFETCH && pos + gridDim.x < nfull) nx = load_idx(phys(pos + gridDim.x));
    const ushort2 ra = cur.ra, rb = cur.rb;
    const uint32_t ka0 = cur.ka0, ka1 = cur.ka1, kb0 = cur.kb0, kb1 = cur.kb1;
    while (!mbar_try_wait(full + s, parity)) {}
    const double *__restrict__ win = buf + s * stage_elems;
    const double va0 = value(win, ra.x, ka0), va1 = value(win, ra.y, ka1);
    const double vb0 = value(win, rb.x, kb0), vb1 = value(win, rb.y, kb1);
    if (a.j_aligned) {
      st_stream2(Jt + tid2, va0, va1);
      st_stream2(Jt + kHalf + tid2, vb0, vb1);
    } else {
      Jt[tid2] = va0; Jt[tid2 + 1] = va1; Jt[kHalf + tid2] = vb0; Jt[kHalf + tid2 + 1] = vb1;
    }
    if (EMPTYBAR) {
      __syncwarp();
      if ((threadIdx.x & 31) == 0) mbar_arrive(empty + s);              // this warp is done with stage s
      if (threadIdx.x == 0) {
        const int64_t nxt = pos + (int64_t)kStages * gridDim.x;
        if (nxt < nfull) {
          while (!mbar_try_wait(empty + s, parity)) {}                   // ... and so are the other seven: refill it
          issue(phys(nxt), s);
        }
      }
    } else {
      __syncthreads();                                    // every lane is done with stage s: refill it
      if (threadIdx.x == 0) {
        const int64_t nxt = pos + (int64_t)kStages * gridDim.x;
        if (nxt < nfull) issue(phys(nxt), s);
      }
    }
    if (++s == kStages) { s = 0; parity ^= 1u; }
  }
  // the last, partial tile (E % kTile entries): gather form, one block
  const int64_t rem0 = nfull * kTile;
  if (rem0 < a.E && blockIdx.x == (unsigned)(nfull % gridDim.x)) {
    for (int64_t e = rem0 + threadIdx.x; e < a.E; e += kThreads) {
      const uint32_t k = (uint32_t)ecolor[e];
      double v = 0.0;
      if (k < (uint32_t)C)
        v = fd_quotient<MODE>(a.Fp + (int64_t)k * a.ldF, MODE == kCentral ? a.Fm + (int64_t)k * a.ldF : a.fx, a.row32[e], s_eps[k]);
      a.J[e] = v;
    }
  }
}

}  // namespace fdb
