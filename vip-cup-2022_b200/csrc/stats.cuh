// Order-independent accumulators for the statistics that cross kernels (LayerNorm row statistics, pooled sums).
//
// Partial sums are produced by many threads / CTAs whose order of arrival is not fixed, and which thread sums which
// elements depends on the tile configuration a launch picks (i.e. on the batch size).  fp32 atomics would make the
// result depend on both (and with it every logit downstream).  So values are converted to 36.28 fixed point EARLY --
// pooled sums element by element, row statistics per aligned group of 8 columns -- and everything after that is 64-bit
// INTEGER addition (registers, then atomics): associative, hence bit-reproducible from run to run and independent of the
// tile schedule, the tile shape and the batch an image is in.
//
// Row statistics record (3 x int64 per row): { sum (v - p), sum (v - p)^2, bits of the pivot p }.  The pivot (the row's
// previous value of column 0, i.e. something close to the row mean) keeps the one-pass variance
//     var = s2 / C - (s1 / C)^2,   mean = p + s1 / C
// free of the cancellation the plain sum / sum-of-squares form suffers when |mean| >> sigma.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vip {

constexpr float kFxScale = 268435456.0f;          // 2^28
constexpr float kFxInv = 1.0f / 268435456.0f;

__device__ __forceinline__ long long to_fx(float v) { return __float2ll_rn(v * kFxScale); }
__device__ __forceinline__ float from_fx(long long v) { return __ll2float_rn(v) * kFxInv; }
__device__ __forceinline__ void fx_atomic_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(to_fx(v)));
}

__device__ __forceinline__ void fx_atomic_add_raw(long long* p, long long v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}

// Low plane of the two-plane residual stream.  Only epilogues touch it, one thread per row, 8 columns (16 bytes) at a
// time, so it is stored in the order the epilogue warps access it: blocks of 32 rows x 8 columns, 512 contiguous bytes
// each (lane = row).  A warp's 16-byte access is then one fully coalesced 512-byte request instead of 32 requests a row
// pitch apart.  Element (row, col) of an [rows (padded to 32), ncols] plane, in bf16 elements:
__host__ __device__ __forceinline__ size_t lo_plane_index(long long row, int col, int ncols) {
  return ((size_t)(row >> 5) * (size_t)(ncols >> 3) + (size_t)(col >> 3)) * 256 + (size_t)(row & 31) * 8 + (size_t)(col & 7);
}

struct RowMoments {
  float mean, rstd;
};
// (mean, 1 / sigma) of a row from its record
__device__ __forceinline__ RowMoments row_moments(long long s1, long long s2, long long pivot_bits, float inv_cols, float eps) {
  const float d = from_fx(s1) * inv_cols;
  const float var = fmaxf(from_fx(s2) * inv_cols - d * d, 0.0f);
  RowMoments m;
  m.mean = __int_as_float((int)pivot_bits) + d;
  m.rstd = 1.0f / sqrtf(var + eps);
  return m;
}

}  // namespace vip
