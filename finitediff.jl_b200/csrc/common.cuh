// common.cuh — shared device helpers for libfdjac_b200 (sm_100a only; no other arch is built).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fdb {

// Colour ids are stored 0-based and compressed to the narrowest type that holds maximum(colorvec):
// uint8 (C <= 255), uint16 (C <= 65535) or int32.  The all-ones pattern marks "no valid colour"
// (colorvec[j] < 1 in the caller's array: such a column is never perturbed and its entries stay 0,
// exactly what the reference's `colorvec[col] == color_i` test yields).
template <typename CT> struct ColorTraits;
template <> struct ColorTraits<uint8_t>  { static constexpr uint32_t invalid = 0xFFu; };
template <> struct ColorTraits<uint16_t> { static constexpr uint32_t invalid = 0xFFFFu; };
template <> struct ColorTraits<int32_t>  { static constexpr uint32_t invalid = 0x7FFFFFFFu; };

constexpr int kThreads = 256;

// Streaming kernels walk their data in tiles of kTile consecutive elements per block step.  Lane t of the block owns
// the element PAIRS (tile + 2t, tile + 2t + 1) and (tile + kTile/2 + 2t, ...): every warp-level access is one
// contiguous run (512 B for a double2, 64..256 B for the index/colour pairs) — full sectors in both directions — and
// each thread has two independent pairs in flight.
constexpr int kPairsPerThread = 2;
constexpr int kTile = kThreads * 2 * kPairsPerThread;   // 1024 elements per block step

// streaming (read-once) loads / stores: keep L1/L2 for the gathered vectors
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }
__device__ __forceinline__ double2 ld_stream2(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
__device__ __forceinline__ void st_stream2(double *p, double a, double b) {
  __stcs(reinterpret_cast<double2 *>(p), make_double2(a, b));
}

// two consecutive colour ids with ONE load (16 / 32 / 64 bit)
template <typename CT> __device__ __forceinline__ void ld_color_pair(const CT *p, uint32_t &a, uint32_t &b);
template <> __device__ __forceinline__ void ld_color_pair<uint8_t>(const uint8_t *p, uint32_t &a, uint32_t &b) {
  const uint16_t v = __ldcs(reinterpret_cast<const unsigned short *>(p));
  a = v & 0xFFu; b = v >> 8;
}
template <> __device__ __forceinline__ void ld_color_pair<uint16_t>(const uint16_t *p, uint32_t &a, uint32_t &b) {
  const uint32_t v = __ldcs(reinterpret_cast<const unsigned int *>(p));
  a = v & 0xFFFFu; b = v >> 16;
}
template <> __device__ __forceinline__ void ld_color_pair<int32_t>(const int32_t *p, uint32_t &a, uint32_t &b) {
  const int2 v = __ldcs(reinterpret_cast<const int2 *>(p));
  a = (uint32_t)v.x; b = (uint32_t)v.y;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace fdb
