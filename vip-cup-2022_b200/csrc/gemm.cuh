// tcgen05 / TMEM / TMA GEMM core for sm_100a (hand-written PTX, no CUTLASS):
//   D[M,N] = epilogue( A[M,K] (bf16, K-major)  x  B[N,K]^T (bf16, K-major) ),  fp32 accumulation in TMEM.
// Used for every dense contraction of the backbones: 1x1 convolutions and Dense layers directly on NHWC
// activations, 3x3 convolutions as an implicit GEMM whose A tiles are gathered by im2col-mode TMA
// (models/resnet_rs/resnet_rs_model.py:64-84, models/gcvit/layers/attention.py:25,33,
// models/gcvit/layers/feature.py:20-22,98).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace vip {

enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_SWISH = 4 };

// Epilogue, applied per output element (m, n) in this order:
//   v = acc
//   v = rstd[m] * (v - mean[m] * ln_colsum[n])      if ln_stats   (LayerNorm of the A rows folded into the contraction:
//                                                    B must hold gamma-scaled weights, bias must hold beta @ W + b)
//   v += bias[n];  v = act(v);  v *= colscale[n];  v += residual[m, n];  store (bf16 or f32)
//   row_stats[m] += (sum_n (out - p), sum_n (out - p)^2)   if row_stats (vip_row_stats_finalize turns the records into the
//                                                    ln_stats of the next contraction;
//                                                    p = a per-row pivot, see stats.cuh; 64-bit fixed-point atomics:
//                                                    the result does not depend on the order of the additions)
//   gap[m / gap_rows, n] += out                      if gap        (GlobalAveragePooling partial sums, SE squeeze; fixed point)
// Two-plane residual stream (residual_lo / out_lo): the running sum of a pre-LN transformer block is carried as
// hi + lo (two bf16 planes, 16 mantissa bits); v = acc + bias + hi + lo, out = bf16(v), out_lo = bf16(v - out).  The hi
// plane is the operand of the next contraction, the lo plane only ever meets the epilogues.
struct GemmEpilogue {
  const float* bias = nullptr;              // [N] fp32 or null
  int act = ACT_NONE;
  const float* colscale = nullptr;          // [N] fp32 or null
  const __nv_bfloat16* residual = nullptr;  // [M, ldr] bf16 or null
  int ldr = 0;
  __nv_bfloat16* out_bf16 = nullptr;        // [M, ldc] (exactly one of out_bf16 / out_f32)
  float* out_f32 = nullptr;
  int ldc = 0;
  const float* ln_stats = nullptr;          // [M, 2] (mean, 1 / sigma) of the A rows (row_stats_finalize / layernorm / mlp_fused)
  const float* ln_colsum = nullptr;         // [N] column sums of the (gamma-scaled, bf16-rounded) weights
  long long* row_stats = nullptr;           // [M, 3] records accumulated with integer atomics; the caller zeroes it
  long long* gap = nullptr;                 // [ceil(M / gap_rows), N] fixed point, integer atomics; the caller zeroes it
  int gap_rows = 0;
  // SE bottleneck tail: v = relu((acc + bias[n]) * row_gate[m / gate_rows, n] + residual[m, n])
  const float* row_gate = nullptr;          // [ceil(M / gate_rows), N]
  int gate_rows = 0;
  const float* row_pivot = nullptr;            // [M, 2]: .x = pivot of the row statistics (the previous LayerNorm's mean of
                                               // the residual row: (mean, 1 / sigma) as in ln_stats); null = 0
  const __nv_bfloat16* residual_lo = nullptr;  // low plane of the residual (blocked layout, stats.cuh; null = zeros)
  __nv_bfloat16* out_lo = nullptr;             // [M, ldc] low plane of the output (needs residual, no act / LN / gate)
};

struct ConvGeom {
  int Nimg, H, W, C;        // input NHWC
  int ksize, stride, pad;   // square kernel, symmetric explicit zero padding
  int Ho, Wo;
  int ldx = 0;              // channels between consecutive pixels in memory (0 = C): a channel slice [c0, c0 + C) of a wider
                            // NHWC tensor is convolved by pointing x at channel c0 (grouped convolutions)
};

// A: [M, K] row-major bf16 (lda elements between rows), B: [N, K] row-major bf16 (ldb).  K, N, lda, ldb, ldc, ldr
// multiples of 8, 16-byte aligned bases.  Enqueues on `stream`; returns VIP_OK or a negative error code.
// rows_per_group > 0 (a multiple of 128): B holds one [N, K] matrix per group of rows_per_group rows of A, stacked.
int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
              const GemmEpilogue& epi, cudaStream_t stream, int rows_per_group = 0);

// Implicit-GEMM convolution: x bf16 NHWC, weights bf16 [Cout, ksize*ksize*C] (K order r, s, c = Keras (kh,kw,Cin,Cout)
// flattened over its first three axes, ldw elements between rows), output rows = N*Ho*Wo pixels.  C % 8 == 0.
int conv2d_bf16(const __nv_bfloat16* x, const ConvGeom& geom, const __nv_bfloat16* w, int ldw, int Cout,
                const GemmEpilogue& epi, cudaStream_t stream);

}  // namespace vip
