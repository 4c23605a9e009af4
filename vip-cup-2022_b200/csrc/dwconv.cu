// Depthwise K x K convolution on NHWC bf16 activations for the second-tier backbones of ckpts/ckpts.json:
//   ConvNeXt block       ZeroPadding2D(3) + DepthwiseConv2D(7) + bias            models/tfimm/architectures/convnext.py:192-198
//   EfficientNet MBConv  DepthwiseConv2D(3 | 5, stride 1 | 2) + BN + swish      keras_cv_attention_models/efficientnet/efficientnet_v2.py:80-96
//                        (BN folded into the depthwise weights / bias by the caller; explicit, possibly asymmetric padding)
// A thread produces TX adjacent output columns x 8 channels of one output row: per filter row it loads the TX * S + K - S
// input pixels it needs once (16-byte loads, bf16 -> packed fp32 pairs) and reuses them across the K horizontal taps
// (FFMA2 on channel pairs), so an input element is fetched ~K / TX * ... times instead of K times per output.  Weights
// f32 [K, K, C] are read through the read-only path (all pixels of a warp's channel group share them).
// Optional fused epilogue: bias, activation (swish / gelu / relu), and the per-image channel sums of the ROUNDED output
// (SE squeeze) as exact fixed-point integer atomics (stats.cuh).
#include <cuda_bf16.h>

#include "common.cuh"
#include "stats.cuh"

namespace vip {
namespace {

using bf16 = __nv_bfloat16;

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmov.b64 rc, {%6,%7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 bf16pair(uint32_t w) { return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u)); }

enum DwAct : int { DW_NONE = 0, DW_SWISH = 1, DW_GELU = 2, DW_RELU = 3 };

__device__ __forceinline__ float dw_act(float v, int act) {
  if (act == DW_SWISH) return v / (1.0f + __expf(-v));
  if (act == DW_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  if (act == DW_RELU) return fmaxf(v, 0.0f);
  return v;
}

struct DwArgs {
  const bf16* x;
  const float* w;     // [K, K, C]
  const float* bias;  // [C] or null
  bf16* out;
  long long* gap;     // [N, C] fixed point or null
  int N, H, W, C, Ho, Wo, pad_t, pad_l, act;
};

template <int K, int S, int TX>
__global__ void __launch_bounds__(128) dwconv_kxk_kernel(const DwArgs a) {
  pdl_trigger();
  pdl_wait();
  constexpr int WIN = TX * S + K - S;   // input columns one thread touches per filter row
  const int c8n = a.C >> 3;
  const int xt = (a.Wo + TX - 1) / TX;
  const long long total = (long long)a.N * a.Ho * xt * c8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  long long t = idx / c8n;
  const int tx = (int)(t % xt);
  t /= xt;
  const int oy = (int)(t % a.Ho), n = (int)(t / a.Ho);
  const int ox0 = tx * TX;
  const int ix0 = ox0 * S - a.pad_l, iy0 = oy * S - a.pad_t;
  float2 acc[TX][4];
#pragma unroll
  for (int i = 0; i < TX; ++i)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[i][q] = make_float2(0.0f, 0.0f);
  const bf16* img = a.x + (long long)n * a.H * a.W * a.C + c8 * 8;
  for (int ky = 0; ky < K; ++ky) {
    const int iy = iy0 + ky;
    if (iy < 0 || iy >= a.H) continue;
    float2 in[WIN][4];
#pragma unroll
    for (int p = 0; p < WIN; ++p) {
      const int ix = ix0 + p;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (ix >= 0 && ix < a.W) v = __ldg(reinterpret_cast<const uint4*>(img + ((long long)iy * a.W + ix) * a.C));
      in[p][0] = bf16pair(v.x);
      in[p][1] = bf16pair(v.y);
      in[p][2] = bf16pair(v.z);
      in[p][3] = bf16pair(v.w);
    }
#pragma unroll
    for (int kx = 0; kx < K; ++kx) {
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(a.w + (long long)(ky * K + kx) * a.C + c8 * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(a.w + (long long)(ky * K + kx) * a.C + c8 * 8) + 1);
      const float2 wq[4] = {make_float2(w0.x, w0.y), make_float2(w0.z, w0.w), make_float2(w1.x, w1.y), make_float2(w1.z, w1.w)};
#pragma unroll
      for (int i = 0; i < TX; ++i)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[i][q] = ffma2(in[i * S + kx][q], wq[q], acc[i][q]);
    }
  }
  float bs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (a.bias != nullptr) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.bias + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(a.bias + c8 * 8) + 1);
    bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
  }
  long long gs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  bf16* orow = a.out + (((long long)n * a.Ho + oy) * a.Wo) * a.C + c8 * 8;
#pragma unroll
  for (int i = 0; i < TX; ++i) {
    if (ox0 + i >= a.Wo) break;
    uint32_t pk[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float v0 = dw_act(acc[i][q].x + bs[2 * q], a.act), v1 = dw_act(acc[i][q].y + bs[2 * q + 1], a.act);
      const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
      pk[q] = *reinterpret_cast<const uint32_t*>(&h);
      if (a.gap != nullptr) {   // exact: every rounded element goes to fixed point on its own
        gs[2 * q] += to_fx(__uint_as_float(pk[q] << 16));
        gs[2 * q + 1] += to_fx(__uint_as_float(pk[q] & 0xffff0000u));
      }
    }
    *reinterpret_cast<uint4*>(orow + (long long)(ox0 + i) * a.C) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
  if (a.gap != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) fx_atomic_add_raw(a.gap + (long long)n * a.C + c8 * 8 + k, gs[k]);
  }
}

template <int K, int S>
int launch_dw(const DwArgs& a, cudaStream_t st) {
  constexpr int TX = S == 1 ? 4 : 2;
  const long long threads = (long long)a.N * a.Ho * ((a.Wo + TX - 1) / TX) * (a.C / 8);
  VIP_LAUNCH((dwconv_kxk_kernel<K, S, TX>), (unsigned)((threads + 127) / 128), 128, 0, st, a);
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

}  // namespace
}  // namespace vip

extern "C" int vip_dwconv_bf16(const void* x, const float* w, const float* bias, void* out, int64_t* gap, int N, int H, int W,
                               int C, int ksize, int stride, int pad_top, int pad_left, int Ho, int Wo, int act, void* stream) {
  using namespace vip;
  VIP_REQUIRE(x && w && out && N > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && Ho > 0 && Wo > 0, VIP_ERR_INVALID,
              "vip_dwconv_bf16: bad argument (C %% 8 == 0)");
  VIP_REQUIRE(act >= 0 && act <= 3 && pad_top >= 0 && pad_left >= 0, VIP_ERR_INVALID, "vip_dwconv_bf16: bad act / padding");
  VIP_REQUIRE((Ho - 1) * stride - pad_top < H && (Wo - 1) * stride - pad_left < W, VIP_ERR_INVALID,
              "vip_dwconv_bf16: output geometry reaches past the input");
  DwArgs a{reinterpret_cast<const bf16*>(x), w, bias, reinterpret_cast<bf16*>(out), reinterpret_cast<long long*>(gap),
           N, H, W, C, Ho, Wo, pad_top, pad_left, act};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  switch (ksize * 10 + stride) {
    case 31: return launch_dw<3, 1>(a, st);
    case 32: return launch_dw<3, 2>(a, st);
    case 51: return launch_dw<5, 1>(a, st);
    case 52: return launch_dw<5, 2>(a, st);
    case 71: return launch_dw<7, 1>(a, st);
    default: break;
  }
  set_error("vip_dwconv_bf16: kernel %d stride %d is not built (3 | 5 with stride 1 | 2, 7 with stride 1)", ksize, stride);
  return VIP_ERR_UNSUPPORTED;
}
