// Window attention of GCViT (models/gcvit/layers/attention.py:52-83) with window_partition / window_reverse
// (models/gcvit/layers/window.py:3-14) folded into the indexing: tokens are read from and written to the image-order
// [B, H, W, C] layout directly.  out = softmax(q * hd^-0.5 @ k^T + rel_bias[h]) @ v  per (window, head), head_dim 32.
//   local block : q, k, v = qkv[..., 0:C], [C:2C], [2C:3C]
//   global block: k, v = qkv[..., 0:C], [C:2C];  q = q_global[b] (shared by all windows of image b, attention.py:62-66)
// v1: one CTA per (window, head), K/V staged in shared memory, fp32 online softmax on the CUDA cores.
#include <cuda_bf16.h>

#include "common.cuh"

namespace vip {
namespace {

using bf16 = __nv_bfloat16;
constexpr int HD = 32;

__global__ void __launch_bounds__(128) window_attention_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ qg,
                                                               const float* __restrict__ rel_bias, bf16* __restrict__ out,
                                                               int H, int W, int C, int ws, int heads, int global_q,
                                                               float scale) {
  extern __shared__ float sm[];
  const int N = ws * ws;
  float* sK = sm;            // [N][HD + 1]
  float* sV = sK + N * (HD + 1);
  const int nWw = W / ws, nWh = H / ws;
  const int win = blockIdx.x, h = blockIdx.y;
  const int b = win / (nWh * nWw);
  const int wrem = win - b * nWh * nWw;
  const int wy = wrem / nWw, wx = wrem - wy * nWw;
  const int ldq = (global_q ? 2 : 3) * C;
  const int koff = (global_q ? 0 : C) + h * HD, voff = koff + C;

  for (int i = threadIdx.x; i < N * HD; i += blockDim.x) {
    const int t = i / HD, d = i - t * HD;
    const int y = wy * ws + t / ws, x = wx * ws + t % ws;
    const long long row = ((long long)b * H + y) * W + x;
    sK[t * (HD + 1) + d] = __bfloat162float(qkv[row * ldq + koff + d]);
    sV[t * (HD + 1) + d] = __bfloat162float(qkv[row * ldq + voff + d]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const int y = wy * ws + i / ws, x = wx * ws + i % ws;
    const long long row = ((long long)b * H + y) * W + x;
    float q[HD];
    const bf16* qp = global_q ? qg + ((long long)b * N + i) * C + h * HD : qkv + row * ldq + h * HD;
#pragma unroll
    for (int d = 0; d < HD; ++d) q[d] = __bfloat162float(qp[d]) * scale;
    const float* bias = rel_bias + ((long long)h * N + i) * N;
    float m = -3.0e38f, l = 0.0f, acc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) acc[d] = 0.0f;
    for (int j = 0; j < N; ++j) {
      float s = 0.0f;
#pragma unroll
      for (int d = 0; d < HD; ++d) s += q[d] * sK[j * (HD + 1) + d];
      s += __ldg(bias + j);
      const float mn = fmaxf(m, s);
      const float corr = __expf(m - mn), p = __expf(s - mn);
      l = l * corr + p;
#pragma unroll
      for (int d = 0; d < HD; ++d) acc[d] = acc[d] * corr + p * sV[j * (HD + 1) + d];
      m = mn;
    }
    const float inv = 1.0f / l;
    bf16* op = out + row * C + h * HD;
#pragma unroll
    for (int d = 0; d < HD; d += 2)
      *reinterpret_cast<__nv_bfloat162*>(op + d) = __floats2bfloat162_rn(acc[d] * inv, acc[d + 1] * inv);
  }
}

}  // namespace
}  // namespace vip

extern "C" int vip_window_attention_bf16(const void* qkv, const void* q_global, const float* rel_bias, void* out, int B,
                                         int H, int W, int C, int ws, int heads, void* stream) {
  using namespace vip;
  VIP_REQUIRE(qkv && rel_bias && out, VIP_ERR_INVALID, "vip_window_attention_bf16: null pointer");
  VIP_REQUIRE(C == heads * HD, VIP_ERR_UNSUPPORTED, "vip_window_attention_bf16: head_dim must be 32 (C=%d heads=%d)", C,
              heads);
  VIP_REQUIRE(H % ws == 0 && W % ws == 0, VIP_ERR_INVALID, "vip_window_attention_bf16: H, W must be multiples of ws");
  const int N = ws * ws;
  const int smem = 2 * N * (HD + 1) * 4;
  static bool configured = false;
  if (!configured) {
    VIP_CUDA(cudaFuncSetAttribute(window_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    configured = true;
  }
  VIP_REQUIRE(smem <= 64 * 1024, VIP_ERR_UNSUPPORTED, "vip_window_attention_bf16: window too large");
  dim3 grid(B * (H / ws) * (W / ws), heads);
  window_attention_kernel<<<grid, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)qkv, (const bf16*)q_global, rel_bias, (bf16*)out, H, W, C, ws, heads, q_global != nullptr ? 1 : 0,
      1.0f / sqrtf((float)HD));
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}
