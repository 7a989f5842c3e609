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

// streaming (read-once) loads / stores: keep L1 for the gathered vectors
__device__ __forceinline__ double ld_stream(const double *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(double *p, double v) { __stcs(p, v); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace fdb
