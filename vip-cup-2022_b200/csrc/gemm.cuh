// tcgen05 / TMEM / TMA GEMM core for sm_100a (hand-written PTX, no CUTLASS):
//   D[M,N] = epilogue( A[M,K] (bf16, K-major)  x  B[N,K]^T (bf16, K-major) ),  fp32 accumulation in TMEM.
// Used for every dense contraction of the backbones: 1x1 convolutions and Dense layers directly on NHWC
// activations, 3x3 convolutions through an im2col view (models/resnet_rs/resnet_rs_model.py:64-84,
// models/gcvit/layers/attention.py:25,33, models/gcvit/layers/feature.py:20-22).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

namespace vip {

enum GemmAct : int { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_SIGMOID = 3 };

// Epilogue, applied per output element (m, n) in this order:
//   v = acc + bias[n];  v = act(v);  v = v * colscale[n] (layer-scale gamma);  v += residual[m, n];  store
struct GemmEpilogue {
  const float* bias = nullptr;           // [N] fp32 or null
  int act = ACT_NONE;
  const float* colscale = nullptr;       // [N] fp32 or null
  const __nv_bfloat16* residual = nullptr;  // [M, ldr] bf16 or null
  int ldr = 0;
  __nv_bfloat16* out_bf16 = nullptr;     // [M, ldc] (exactly one of out_bf16 / out_f32)
  float* out_f32 = nullptr;
  int ldc = 0;
};

// A: [M, K] row-major bf16 (lda elements between rows), B: [N, K] row-major bf16 (ldb).  K % 8 == 0, lda/ldb % 8 == 0,
// 16-byte aligned bases.  N % 16 == 0.  Enqueues on `stream`; returns VIP_OK or a negative error code.
int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
              const GemmEpilogue& epi, cudaStream_t stream);

}  // namespace vip
