// Non-GEMM layer kernels of the backbones (bf16 NHWC activations, fp32 math), all HBM-bound and written as
// 16-byte vectorised, coalesced grid-stride kernels.  Each entry point cites the reference layer it replaces.
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "stats.cuh"

namespace vip {
namespace {

using bf16 = __nv_bfloat16;

struct alignas(16) bf16x8 {
  __nv_bfloat162 v[4];
};

__device__ __forceinline__ void unpack8(const bf16x8& p, float (&f)[8]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p.v[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ bf16x8 pack8(const float (&f)[8]) {
  bf16x8 p;
#pragma unroll
  for (int i = 0; i < 4; ++i) p.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return p;
}
// x * Phi(x) with Phi through a fitted tanh form, |error| <= 3e-4 |x| (see gelu_fast in gemm.cu for the derivation)
__device__ __forceinline__ float gelu_erf(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(-0.00035307545f, x2, 0.037015257f);
  p = fmaf(p, x2, 0.79749725f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

inline int grid_for(long long work, int block) {
  long long g = (work + block - 1) / block;
  const long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// ---- im2col: Conv2D with explicit zero padding as a [M, Kp] bf16 matrix (resnet_rs_model.py:64-84; embedding.py:15;
// feature.py:98).  K order = (r, s, c), matching Keras kernels (kh, kw, Cin, Cout) flattened over the first 3 axes.
__global__ void im2col_vec8_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int C, int ks,
                                   int stride, int pad, int Ho, int Wo, int Kp) {
  pdl_trigger();
  pdl_wait();
  const int c8n = C >> 3;
  const long long per_row = (long long)ks * ks * c8n;
  const long long total = (long long)N * Ho * Wo * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / per_row;
    const int rem = (int)(i - m * per_row);
    const int tap = rem / c8n, c8 = rem - tap * c8n;
    const int r = tap / ks, s = tap - r * ks;
    const int ox = (int)(m % Wo);
    const long long t = m / Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    const int iy = oy * stride - pad + r, ix = ox * stride - pad + s;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + iy) * W + ix) * C) + c8);
    *reinterpret_cast<uint4*>(out + m * Kp + (long long)tap * C + c8 * 8) = v;
  }
}
// generic (any C, e.g. the 3-channel network input); also zero-fills the K..Kp tail.  One thread builds one whole row
// of the matrix (Kp <= 64) in registers and writes it with 16-byte stores.
template <int KP, int KS, int CI>
__global__ void im2col_row_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int stride,
                                  int pad, int Ho, int Wo) {
  pdl_trigger();
  pdl_wait();
  static_assert(KS * KS * CI <= KP && KP % 8 == 0, "row does not fit");
  const long long total = (long long)N * Ho * Wo;
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < total; m += (long long)gridDim.x * blockDim.x) {
    const int ox = (int)(m % Wo);
    const long long t = m / Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    unsigned short v[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) v[k] = 0;
#pragma unroll
    for (int r = 0; r < KS; ++r) {
      const int iy = oy * stride - pad + r;
#pragma unroll
      for (int sx = 0; sx < KS; ++sx) {
        const int ix = ox * stride - pad + sx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          const unsigned short* px = reinterpret_cast<const unsigned short*>(x) + (((long long)n * H + iy) * W + ix) * CI;
#pragma unroll
          for (int c = 0; c < CI; ++c) v[(r * KS + sx) * CI + c] = __ldg(px + c);
        }
      }
    }
    uint4* op = reinterpret_cast<uint4*>(out + m * KP);
#pragma unroll
    for (int q = 0; q < KP / 8; ++q) {
      uint4 u;
      u.x = v[q * 8 + 0] | ((unsigned)v[q * 8 + 1] << 16);
      u.y = v[q * 8 + 2] | ((unsigned)v[q * 8 + 3] << 16);
      u.z = v[q * 8 + 4] | ((unsigned)v[q * 8 + 5] << 16);
      u.w = v[q * 8 + 6] | ((unsigned)v[q * 8 + 7] << 16);
      op[q] = u;
    }
  }
}
__global__ void im2col_scalar_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int C, int ks,
                                     int stride, int pad, int Ho, int Wo, int Kp) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * Ho * Wo * Kp;
  const int K = ks * ks * C;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long m = i / Kp;
    const int k = (int)(i - m * Kp);
    bf16 v = __float2bfloat16(0.0f);
    if (k < K) {
      const int tap = k / C, c = k - tap * C;
      const int r = tap / ks, s = tap - r * ks;
      const int ox = (int)(m % Wo);
      const long long t = m / Wo;
      const int oy = (int)(t % Ho), n = (int)(t / Ho);
      const int iy = oy * stride - pad + r, ix = ox * stride - pad + s;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[(((long long)n * H + iy) * W + ix) * C + c];
    }
    out[i] = v;
  }
}

// ---- AveragePooling2D(2, 2, 'same') (resnet_rs_model.py:207-212): divisor = number of valid inputs
__global__ void avgpool2_same_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int C, int Ho,
                                     int Wo) {
  pdl_trigger();
  pdl_wait();
  const int c8n = C >> 3;
  const long long total = (long long)N * Ho * Wo * c8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    long long t = i / c8n;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    int cnt = 0;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int iy = 2 * oy + dy, ix = 2 * ox + dx;
        if (iy < H && ix < W) {
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(x + (((long long)n * H + iy) * W + ix) * C + c8 * 8), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] += f[k];
          ++cnt;
        }
      }
    const float inv = 1.0f / (float)cnt;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= inv;
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(acc);
  }
}

// ---- GlobalAveragePooling2D over [N, HW, C] -> [N, C] (bf16 and/or f32 out).  Block = (image, 64-channel slab).
__global__ void __launch_bounds__(256) global_avgpool_kernel(const bf16* __restrict__ x, bf16* __restrict__ out_bf16,
                                                             float* __restrict__ out_f32, int HW, int C) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32][65];
  const int n = blockIdx.x, c0 = blockIdx.y * 64;
  const int cl = threadIdx.x & 7, rl = threadIdx.x >> 3;  // 8 channel-octets x 32 row lanes
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (c0 + cl * 8 < C) {
    const bf16* base = x + (long long)n * HW * C + c0 + cl * 8;
    for (int p = rl; p < HW; p += 32) {
      float f[8];
      unpack8(*reinterpret_cast<const bf16x8*>(base + (long long)p * C), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[rl][cl * 8 + k] = acc[k];
  __syncthreads();
  if (threadIdx.x < 64 && c0 + threadIdx.x < C) {
    float s = 0.0f;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) s += red[r][threadIdx.x];
    s *= 1.0f / (float)HW;
    if (out_bf16) out_bf16[(long long)n * C + c0 + threadIdx.x] = __float2bfloat16_rn(s);
    if (out_f32) out_f32[(long long)n * C + c0 + threadIdx.x] = s;
  }
}

// ---- out = act(y * gate[n, c] + shortcut)   (SE excite + Add + ReLU, resnet_rs_model.py:183,278-280;
//      GCViT SE x*inputs feature.py:66, residual feature.py:109,150)
__global__ void scale_add_act_kernel(const bf16* __restrict__ y, const float* __restrict__ gate,
                                     const bf16* __restrict__ shortcut, bf16* __restrict__ out, long long total8, int HW,
                                     int C, int act) {
  pdl_trigger();
  pdl_wait();
  const int c8n = C >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    const long long pix = i / c8n;
    const int n = (int)(pix / HW);
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(y + i * 8), f);
    if (gate != nullptr) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate + (long long)n * C + c8 * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gate + (long long)n * C + c8 * 8) + 1);
      f[0] *= g0.x; f[1] *= g0.y; f[2] *= g0.z; f[3] *= g0.w;
      f[4] *= g1.x; f[5] *= g1.y; f[6] *= g1.z; f[7] *= g1.w;
    }
    if (shortcut != nullptr) {
      float s[8];
      unpack8(*reinterpret_cast<const bf16x8*>(shortcut + i * 8), s);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += s[k];
    }
    if (act == 1) {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaxf(f[k], 0.0f);
    }
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(f);
  }
}

// ---- LayerNormalization(axis=-1, eps) over [M, C] (block.py:28,39; feature.py:100-101; gcvit.py:79).  A row is handled
// by LPR lanes (8, 16 or 32, so that narrow rows do not idle most of a warp), J 16-byte chunks per lane.
// U row groups per warp iteration are loaded before any of them is reduced.
template <int LPR, int J, int U>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ out,
                                                        float* __restrict__ ln_next, float next_eps, long long M, int C, float eps) {
  pdl_trigger();
  pdl_wait();
  constexpr int RPW = 32 / LPR;  // rows per warp and group
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int c8n = C >> 3;
  const float invC = 1.0f / (float)C;
  for (long long m0 = warp0 * RPW * U; m0 < M; m0 += nwarps * RPW * U) {
    float f[U][J][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * RPW + rsel;
      const bf16* row = x + (m < M ? m : 0) * C;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c8 = sub + LPR * j;
        if (c8 < c8n) unpack8(*reinterpret_cast<const bf16x8*>(row + c8 * 8), f[u][j]);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long m = m0 + u * RPW + rsel;
      const bool live = m < M;
      float s = 0.0f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c8 = sub + LPR * j;
        if (c8 < c8n) {
#pragma unroll
          for (int k = 0; k < 8; ++k) s += f[u][j][k];
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      const float mean = s * invC;
      float v = 0.0f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c8 = sub + LPR * j;
        if (c8 < c8n) {
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float d = f[u][j][k] - mean;
            v = fmaf(d, d, v);
          }
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      const float rstd = 1.0f / sqrtf(v * invC + eps);
      // moments of the ROUNDED output row (for a LayerNorm folded into the next GEMM), sums shifted by the row's first
      // output element (stats.cuh)
      float os = 0.0f, oq = 0.0f, pivot = 0.0f;
      if (ln_next != nullptr) {
        const float g0 = __ldg(gamma), b0 = __ldg(beta);
        const float first = __bfloat162float(__float2bfloat16_rn(fmaf((f[u][0][0] - mean) * rstd, g0, b0)));  // valid in lane sub == 0
        pivot = __shfl_sync(0xffffffffu, first, rsel * LPR);
      }
#pragma unroll
      for (int j = 0; j < J; ++j) {
        const int c8 = sub + LPR * j;
        if (c8 < c8n) {
          const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8)), g1 = __ldg(reinterpret_cast<const float4*>(gamma + c8 * 8) + 1);
          const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(beta + c8 * 8) + 1);
          const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
          const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
          float o8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) o8[k] = fmaf((f[u][j][k] - mean) * rstd, gg[k], bb[k]);
          const bf16x8 pk = pack8(o8);
          if (live) *reinterpret_cast<bf16x8*>(out + m * C + c8 * 8) = pk;
          if (ln_next != nullptr) {
            float r8[8];
            unpack8(pk, r8);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              const float d = r8[k] - pivot;
              os += d;
              oq = fmaf(d, d, oq);
            }
          }
        }
      }
      if (ln_next != nullptr) {
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) {
          os += __shfl_xor_sync(0xffffffffu, os, o);
          oq += __shfl_xor_sync(0xffffffffu, oq, o);
        }
        if (sub == 0 && live) {
          const float d = os * invC;
          const float var2 = fmaxf(oq * invC - d * d, 0.0f);
          *reinterpret_cast<float2*>(ln_next + 2 * m) = make_float2(pivot + d, 1.0f / sqrtf(var2 + next_eps));
        }
      }
    }
  }
}

// ---- ZeroPadding2D(1) + DepthwiseConv2D(3, valid, no bias) (+ GELU) (feature.py:92-94,132-134).  A thread owns one
// (image, column, 2P-channel group) and walks down the rows with the 3x3 window and the weights in registers, so every
// input element is fetched three times (once per neighbouring column, L1 hits) instead of nine.  The first version was
// bound by its instruction count (334 per output row of 8 channels, ncu: issue 51 % busy at 12 warps per SM), so:
// channel PAIRS in packed fp32 (FFMA2 / FMUL2 of sm_100: half the arithmetic issue slots), the three window rows rotate by
// name (rows unrolled by three, no register copies), the loads of row y + 2 are issued before row y is computed.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmov.b64 rc, {%6,%7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nadd.rn.f32x2 rd, ra, rb;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
// gelu_erf on a pair
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
  float2 x2 = fmul2(x, x);
  x2.x = fminf(x2.x, 64.0f);
  x2.y = fminf(x2.y, 64.0f);
  float2 p = ffma2(make_float2(-0.00035307545f, -0.00035307545f), x2, make_float2(0.037015257f, 0.037015257f));
  p = ffma2(p, x2, make_float2(0.79749725f, 0.79749725f));
  const float2 a = fmul2(p, x);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
  const float2 hx = fmul2(x, make_float2(0.5f, 0.5f));
  return ffma2(hx, t, hx);
}
__device__ __forceinline__ float2 bf16pair(uint32_t w) {  // two bf16 -> two fp32
  return make_float2(__uint_as_float(w << 16), __uint_as_float(w & 0xffff0000u));
}

// What bounds the walk is bytes in flight: a thread has one row of loads outstanding however many channels it owns
// (registers per thread and threads per SM trade one for one: ~12 KB per SM, 22 % of HBM peak).  So the loads run D
// rows ahead as cp.async into thread-private shared-memory slots (no block barrier, no register cost).
template <int P, int D, int MINB>  // channel pairs per thread: 4 (16-byte accesses) or 2 (8-byte); prefetch depth in rows
__global__ void __launch_bounds__(128, MINB) dwconv3x3_kernel(const bf16* __restrict__ x, const float* __restrict__ w /*[3][3][C]*/,
                                                        bf16* __restrict__ out, long long* __restrict__ gap /*[N][C] fixed point or null*/,
                                                        int N, int H, int W, int C, int gelu) {
  pdl_trigger();
  pdl_wait();
  constexpr int V = 2 * P, BYTES = 4 * P, RING = D + 1;
  __shared__ __align__(16) uint8_t ring[RING * 3 * 128 * BYTES];
  const int cvn = C / V;
  const long long total = (long long)N * W * cvn;
  long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = idx < total;
  if (!active) idx = total - 1;   // keeps the cp.async group bookkeeping uniform; results are dropped
  const int cv = (int)(idx % cvn);
  const long long t = idx / cvn;
  const int ox = (int)(t % W), n = (int)(t / W);
  float2 wt[9][P];
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int q = 0; q < P; ++q) wt[k][q] = __ldg(reinterpret_cast<const float2*>(w + k * C + cv * V) + q);
  const bf16* img = x + (long long)n * H * W * C + cv * V;
  const bool has_l = ox > 0, has_r = ox + 1 < W;
  auto slot = [&](int y, int dx) -> uint8_t* { return ring + (((y % RING) * 3 + dx) * 128 + threadIdx.x) * BYTES; };
  auto request = [&](int y) {  // the three neighbours of input row y -> this thread's slots (zero fill outside the image)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const bool ok = y < H && (dx == 1 || (dx == 0 ? has_l : has_r));
      const bf16* src = ok ? img + ((long long)y * W + ox - 1 + dx) * C : img;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(slot(y, dx));
      const int sz = ok ? BYTES : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(src), "n"(BYTES), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  float2 win[3][3][P];  // three window rows (rotating), [col x-1, x, x+1], channel pairs
  auto take = [&](int y, float2 (&dst)[3][P]) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      if (P == 4) {
        const uint4 v = *reinterpret_cast<const uint4*>(slot(y, dx));
        dst[dx][0] = bf16pair(v.x); dst[dx][1] = bf16pair(v.y); dst[dx][P - 2] = bf16pair(v.z); dst[dx][P - 1] = bf16pair(v.w);
      } else {
        const uint2 v = *reinterpret_cast<const uint2*>(slot(y, dx));
        dst[dx][0] = bf16pair(v.x); dst[dx][1] = bf16pair(v.y);
      }
    }
  };
#pragma unroll
  for (int d = 0; d <= D; ++d) request(d);          // rows 0 .. D
#pragma unroll
  for (int dx = 0; dx < 3; ++dx)
#pragma unroll
    for (int q = 0; q < P; ++q) win[0][dx][q] = make_float2(0.0f, 0.0f);
  asm volatile("cp.async.wait_group %0;" ::"n"(D) : "memory");
  take(0, win[1]);
  bf16* op = out + (long long)n * H * W * C + (long long)ox * C + cv * V;
  float2 colsum[P];  // of the rounded outputs of this column (SE squeeze, feature.py:55)
#pragma unroll
  for (int q = 0; q < P; ++q) colsum[q] = make_float2(0.0f, 0.0f);
  // output row y from window rows (a, b, c) = rows (y - 1, y, y + 1)
  auto step = [&](int y, float2 (&ra)[3][P], float2 (&rb)[3][P], float2 (&rc)[3][P]) {
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");   // row y + 1 has landed
    take(y + 1, rc);
    request(y + 1 + D);                                                 // into the slot row y read one step ago
    float2 acc[P];
#pragma unroll
    for (int q = 0; q < P; ++q) acc[q] = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int sx = 0; sx < 3; ++sx)
#pragma unroll
      for (int q = 0; q < P; ++q) {
        acc[q] = ffma2(ra[sx][q], wt[sx][q], acc[q]);
        acc[q] = ffma2(rb[sx][q], wt[3 + sx][q], acc[q]);
        acc[q] = ffma2(rc[sx][q], wt[6 + sx][q], acc[q]);
      }
    uint32_t pk[P];
#pragma unroll
    for (int q = 0; q < P; ++q) {
      if (gelu) acc[q] = gelu_erf2(acc[q]);
      const __nv_bfloat162 h = __floats2bfloat162_rn(acc[q].x, acc[q].y);
      pk[q] = *reinterpret_cast<const uint32_t*>(&h);
      colsum[q] = fadd2(colsum[q], bf16pair(pk[q]));
    }
    if (active) {
      if (P == 4) *reinterpret_cast<uint4*>(op + (long long)y * W * C) = make_uint4(pk[0], pk[1], pk[P - 2], pk[P - 1]);
      else *reinterpret_cast<uint2*>(op + (long long)y * W * C) = make_uint2(pk[0], pk[1]);
    }
  };
  int y = 0;
  for (; y + 3 <= H; y += 3) {
    step(y, win[0], win[1], win[2]);
    step(y + 1, win[1], win[2], win[0]);
    step(y + 2, win[2], win[0], win[1]);
  }
  if (y < H) {
    step(y, win[0], win[1], win[2]);
    if (y + 1 < H) step(y + 1, win[1], win[2], win[0]);
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (gap != nullptr && active) {
#pragma unroll
    for (int q = 0; q < P; ++q) {
      fx_atomic_add(gap + (long long)n * C + cv * V + 2 * q, colsum[q].x);
      fx_atomic_add(gap + (long long)n * C + cv * V + 2 * q + 1, colsum[q].y);
    }
  }
}

// ---- ZeroPadding2D(1) + MaxPool2D(3, 2, 'valid') (feature.py:139,151-152): the padded zeros take part in the max
__global__ void maxpool3s2_zeropad_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int C,
                                          int Ho, int Wo) {
  pdl_trigger();
  pdl_wait();
  const int c8n = C >> 3;
  const long long total = (long long)N * Ho * Wo * c8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    long long t = i / c8n;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    float m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) m[k] = -3.0e38f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int iy = 2 * oy - 1 + r, ix = 2 * ox - 1 + s;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
          float f[8];
          unpack8(*reinterpret_cast<const bf16x8*>(x + (((long long)n * H + iy) * W + ix) * C + c8 * 8), f);
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
        } else {
#pragma unroll
          for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], 0.0f);
        }
      }
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(m);
  }
}

// ---- classifier head: Dense(k) + softmax | sigmoid on pooled features (gcvit.py:88, resnet_rs_model.py:474-476),
//      multi->binary (main.py:113-114) and TTA / ensemble accumulation (main.py:111,121,142) fused:
//      p_syn = k > 1 ? 1 - softmax(z)[0] : sigmoid(z);  probs[n, :] = activation;  acc[n] += weight * p_syn
__global__ void head_kernel(const float* __restrict__ feat, const float* __restrict__ w /*[C][k]*/,
                            const float* __restrict__ b, float* __restrict__ probs, double* __restrict__ acc,
                            double acc_weight, int N, int C, int k, int sigmoid_head) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= N) return;
  float z[8];
  for (int j = 0; j < k; ++j) {
    float s = 0.0f;
    for (int c = lane; c < C; c += 32) s += feat[(long long)warp * C + c] * __ldg(w + (long long)c * k + j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    z[j] = s + b[j];
  }
  if (lane == 0) {
    float p_syn;
    if (sigmoid_head) {
      for (int j = 0; j < k; ++j) {
        z[j] = 1.0f / (1.0f + expf(-z[j]));
        probs[(long long)warp * k + j] = z[j];
      }
      p_syn = z[0];
    } else {
      float mx = z[0];
      for (int j = 1; j < k; ++j) mx = fmaxf(mx, z[j]);
      float sum = 0.0f;
      for (int j = 0; j < k; ++j) {
        z[j] = expf(z[j] - mx);
        sum += z[j];
      }
      for (int j = 0; j < k; ++j) probs[(long long)warp * k + j] = z[j] / sum;
      p_syn = 1.0f - z[0] / sum;
    }
    if (k > 1 && sigmoid_head) p_syn = 1.0f - z[0];
    if (acc != nullptr) acc[warp] += acc_weight * (double)p_syn;
  }
}

// ---- f32 -> bf16 cast (network input when the preprocessing output is kept in fp32)
__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ out, long long n) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i]);
}

// ---- out = bf16(x * scale): pooled sums -> means as a GEMM operand (SE squeeze, resnet_rs_model.py:149)
__global__ void scale_cast_fx_bf16_kernel(const long long* __restrict__ x, float scale, bf16* __restrict__ out, long long n) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(from_fx(x[i]) * scale);
}

// ---- ResNeSt split attention, radix 2 (keras_cv_attention_models/resnest/resnest.py:16-24, 57-62): r-softmax of the
// attention logits over the two radix splits, then out[n, p, c] = a0[n, c] x[n, p, c] + a1[n, c] x[n, p, F + c]
__global__ void split_attention2_kernel(const bf16* __restrict__ x, const float* __restrict__ logits, bf16* __restrict__ out,
                                        long long total8, int HW, int F) {
  pdl_trigger();
  pdl_wait();
  const int f8n = F >> 3;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    const int f8 = (int)(i % f8n);
    const long long pix = i / f8n;
    const int n = (int)(pix / HW);
    const float* lg = logits + (long long)n * 2 * F + f8 * 8;
    float x0[8], x1[8], o[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + pix * 2 * F + f8 * 8), x0);
    unpack8(*reinterpret_cast<const bf16x8*>(x + pix * 2 * F + F + f8 * 8), x1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float z0 = __ldg(lg + k), z1 = __ldg(lg + F + k);
      const float a0 = 1.0f / (1.0f + __expf(z1 - z0));     // softmax over the two radix logits
      o[k] = fmaf(a0, x0[k], (1.0f - a0) * x1[k]);
    }
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(o);
  }
}

// ---- ZeroPadding2D(1) + AveragePooling2D(3, strides=2) (resnest.py:63-65): the padded zeros count, divisor 9
__global__ void avgpool3s2_zeropad_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int C, int Ho,
                                          int Wo) {
  pdl_trigger();
  pdl_wait();
  const int c8n = C >> 3;
  const long long total = (long long)N * Ho * Wo * c8n;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    long long t = i / c8n;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int dy = 0; dy < 3; ++dy) {
      const int iy = oy * 2 - 1 + dy;
      if (iy < 0 || iy >= H) continue;
      for (int dx = 0; dx < 3; ++dx) {
        const int ix = ox * 2 - 1 + dx;
        if (ix < 0 || ix >= W) continue;
        float f[8];
        unpack8(*reinterpret_cast<const bf16x8*>(x + (((long long)n * H + iy) * W + ix) * C + c8 * 8), f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] += f[k];
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] *= (1.0f / 9.0f);
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(acc);
  }
}

// ---- out = act(x) * scale (NFNet pre-activation: swish(x) * beta, models/keras_cv_attention_models/nfnets/nfnets.py:138)
__global__ void act_scale_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long total8, int act, float scale) {
  pdl_trigger();
  pdl_wait();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total8; i += (long long)gridDim.x * blockDim.x) {
    float f[8];
    unpack8(*reinterpret_cast<const bf16x8*>(x + i * 8), f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float v = f[k];
      if (act == 1) v = fmaxf(v, 0.0f);
      else if (act == 4) v = v / (1.0f + __expf(-v));
      f[k] = v * scale;
    }
    *reinterpret_cast<bf16x8*>(out + i * 8) = pack8(f);
  }
}

// ---- Efficient Channel Attention gate (common_layers.py:335-353): pooled means (fixed-point sums * inv_hw) -> zero pad ->
// Conv1D(k, no bias) along the channel axis -> sigmoid -> * out_scale (the block's attention gain and alpha, nfnets.py:162-168)
__global__ void eca_gate_kernel(const long long* __restrict__ gap, const float* __restrict__ w, float* __restrict__ gate, int N,
                                int C, int ksize, float inv_hw, float out_scale) {
  pdl_trigger();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * C) return;
  const int n = i / C, c = i - n * C, pad = ksize / 2;
  float s = 0.0f;
  for (int k = 0; k < ksize; ++k) {
    const int cc = c + k - pad;
    if (cc >= 0 && cc < C) s = fmaf(from_fx(gap[(long long)n * C + cc]) * inv_hw, __ldg(w + k), s);
  }
  gate[i] = out_scale / (1.0f + __expf(-s));
}

// ---- LayerNormalization of pooled f32 feature vectors [M, C] (ConvNeXt head: GlobalAveragePooling -> LayerNorm -> Dense,
// models/tfimm/architectures/convnext.py:432-436): one warp per row, two-pass variance, f32 in and out
__global__ void layernorm_f32_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float* __restrict__ out, int M, int C, float eps) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= M) return;
  const float* xr = x + (long long)row * C;
  float s = 0.0f;
  for (int c = lane; c < C; c += 32) s += xr[c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float v = 0.0f;
  for (int c = lane; c < C; c += 32) {
    const float d = xr[c] - mean;
    v = fmaf(d, d, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = 1.0f / sqrtf(v / (float)C + eps);
  for (int c = lane; c < C; c += 32) out[(long long)row * C + c] = fmaf((xr[c] - mean) * rstd, __ldg(gamma + c), __ldg(beta + c));
}

// ---- row statistics records (stats.cuh; accumulated by the GEMM epilogues with integer atomics) -> (mean, 1 / sigma) per
// row, the form the folded-LayerNorm consumers load (one coalesced 8-byte load per row and tile)
__global__ void row_stats_finalize_kernel(const long long* __restrict__ rec, float2* __restrict__ out, long long M, float inv_cols,
                                          float eps) {
  pdl_trigger();
  pdl_wait();
  for (long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (long long)gridDim.x * blockDim.x) {
    const RowMoments mo = row_moments(__ldg(rec + 3 * m), __ldg(rec + 3 * m + 1), __ldg(rec + 3 * m + 2), inv_cols, eps);
    out[m] = make_float2(mo.mean, mo.rstd);
  }
}

// ---- FitWindow (models/gcvit/layers/feature.py:234-256) and the crop after the blocks (layers/level.py:61): a window of an
// NHWC tensor copied into another geometry, zero outside the source.  out[n, y, x, :] = in[n, y - top, x - left, :].
// Pixels are moved as 8-byte units, so the same kernel shifts activations (C % 4 == 0) and row statistics records.
__global__ void pad_crop_kernel(const uint2* __restrict__ x, uint2* __restrict__ out, int N, int H, int W, int U, int Ho, int Wo,
                                int top, int left) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * Ho * Wo * U;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(i % U);
    long long t = i / U;
    const int ox = (int)(t % Wo);
    t /= Wo;
    const int oy = (int)(t % Ho), n = (int)(t / Ho);
    const int sy = oy - top, sx = ox - left;
    uint2 v = make_uint2(0u, 0u);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = __ldg(x + (((long long)n * H + sy) * W + sx) * U + u);
    out[i] = v;
  }
}

// ---- out[g][n][k] = bf16(w[n][k] * gate[g][k]): the SE gate of an MBConv block (feature.py:49-66,144-150) scales the INPUT
// channels of the 1x1 convolution that follows, y .* gate_g -> W; per image that is the same as contracting y with
// W diag(gate_g), so the gate is folded into per-image copies of the (tiny) weight matrix instead of a pass over y.
__global__ void scale_weights_kernel(const bf16* __restrict__ w, int ldw, const float* __restrict__ gate, bf16* __restrict__ out,
                                     int G, int N, int K) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)G * N * K;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % K);
    const long long t = i / K;
    const int n = (int)(t % N), gi = (int)(t / N);
    out[i] = __float2bfloat16_rn(__bfloat162float(w[(long long)n * ldw + k]) * gate[(long long)gi * K + k]);
  }
}

}  // namespace
}  // namespace vip

using namespace vip;

#define ST(s) reinterpret_cast<cudaStream_t>(s)
#define LAUNCH_CHECK()               \
  do {                               \
    VIP_CUDA(cudaGetLastError());    \
    count_launch();                  \
    return VIP_OK;                   \
  } while (0)

extern "C" int vip_im2col_bf16(const void* x, int N, int H, int W, int C, int ksize, int stride, int pad, int Ho, int Wo,
                               void* out, int Kp, void* stream) {
  VIP_REQUIRE(x && out && N > 0 && C > 0 && ksize > 0 && stride > 0, VIP_ERR_INVALID, "vip_im2col_bf16: bad argument");
  VIP_REQUIRE(Kp >= ksize * ksize * C && Kp % 8 == 0, VIP_ERR_INVALID, "vip_im2col_bf16: Kp must be >= k*k*C and % 8");
  const bool vec = (C % 8 == 0) && Kp == ksize * ksize * C;
  if (vec) {
    const long long work = (long long)N * Ho * Wo * ksize * ksize * (C / 8);
    VIP_LAUNCH((im2col_vec8_kernel), grid_for(work, 256), 256, 0, ST(stream), (const bf16*)x, (bf16*)out, N, H, W, C, ksize, stride, pad, Ho, Wo, Kp);
  } else if (Kp == 32 && ksize == 3 && C == 3) {
    VIP_LAUNCH((im2col_row_kernel<32, 3, 3>), grid_for((long long)N * Ho * Wo, 128), 128, 0, ST(stream), (const bf16*)x, (bf16*)out, N, H, W, stride, pad, Ho, Wo);
  } else {
    const long long work = (long long)N * Ho * Wo * Kp;
    VIP_LAUNCH((im2col_scalar_kernel), grid_for(work, 256), 256, 0, ST(stream), (const bf16*)x, (bf16*)out, N, H, W, C, ksize, stride, pad, Ho, Wo, Kp);
  }
  LAUNCH_CHECK();
}

extern "C" int vip_avgpool2_same_bf16(const void* x, int N, int H, int W, int C, void* out, void* stream) {
  VIP_REQUIRE(x && out && C % 8 == 0, VIP_ERR_INVALID, "vip_avgpool2_same_bf16: bad argument (C %% 8)");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  VIP_LAUNCH((avgpool2_same_kernel), grid_for((long long)N * Ho * Wo * (C / 8), 256), 256, 0, ST(stream), (const bf16*)x, (bf16*)out, N, H, W, C, Ho, Wo);
  LAUNCH_CHECK();
}

extern "C" int vip_global_avgpool_bf16(const void* x, int N, int HW, int C, void* out_bf16, float* out_f32, void* stream) {
  VIP_REQUIRE(x && (out_bf16 || out_f32) && C % 8 == 0, VIP_ERR_INVALID, "vip_global_avgpool_bf16: bad argument");
  dim3 grid(N, (C + 63) / 64);
  VIP_LAUNCH((global_avgpool_kernel), grid, 256, 0, ST(stream), (const bf16*)x, (bf16*)out_bf16, out_f32, HW, C);
  LAUNCH_CHECK();
}

extern "C" int vip_scale_add_act_bf16(const void* y, const float* gate, const void* shortcut, void* out, int N, int HW,
                                      int C, int act, void* stream) {
  VIP_REQUIRE(y && out && C % 8 == 0, VIP_ERR_INVALID, "vip_scale_add_act_bf16: bad argument");
  const long long total8 = (long long)N * HW * (C / 8);
  VIP_LAUNCH((scale_add_act_kernel), grid_for(total8, 256), 256, 0, ST(stream), (const bf16*)y, gate, (const bf16*)shortcut, (bf16*)out, total8, HW, C, act);
  LAUNCH_CHECK();
}

extern "C" int vip_split_attention2_bf16(const void* x, const float* logits, void* out, int N, int HW, int F, void* stream) {
  VIP_REQUIRE(x && logits && out && N > 0 && HW > 0 && F > 0 && F % 8 == 0, VIP_ERR_INVALID,
              "vip_split_attention2_bf16: bad argument (F %% 8 == 0)");
  const long long total8 = (long long)N * HW * (F / 8);
  VIP_LAUNCH((split_attention2_kernel), grid_for(total8, 256), 256, 0, ST(stream), (const bf16*)x, logits, (bf16*)out, total8, HW, F);
  LAUNCH_CHECK();
}

extern "C" int vip_avgpool3s2_bf16(const void* x, void* out, int N, int H, int W, int C, void* stream) {
  VIP_REQUIRE(x && out && C % 8 == 0, VIP_ERR_INVALID, "vip_avgpool3s2_bf16: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  VIP_LAUNCH((avgpool3s2_zeropad_kernel), grid_for((long long)N * Ho * Wo * (C / 8), 256), 256, 0, ST(stream), (const bf16*)x,
             (bf16*)out, N, H, W, C, Ho, Wo);
  LAUNCH_CHECK();
}

extern "C" int vip_act_scale_bf16(const void* x, void* out, long long n, int act, float scale, void* stream) {
  VIP_REQUIRE(x && out && n > 0 && n % 8 == 0 && (act == 0 || act == 1 || act == 4), VIP_ERR_INVALID,
              "vip_act_scale_bf16: bad argument (n %% 8 == 0; act 0 none, 1 relu, 4 swish)");
  VIP_LAUNCH((act_scale_kernel), grid_for(n / 8, 256), 256, 0, ST(stream), (const bf16*)x, (bf16*)out, n / 8, act, scale);
  LAUNCH_CHECK();
}

extern "C" int vip_eca_gate_f32(const int64_t* gap, const float* w, float* gate, int N, int C, int ksize, float inv_hw,
                                float out_scale, void* stream) {
  VIP_REQUIRE(gap && w && gate && N > 0 && C > 0 && ksize > 0 && ksize % 2 == 1, VIP_ERR_INVALID, "vip_eca_gate_f32: bad argument");
  VIP_LAUNCH((eca_gate_kernel), (N * C + 255) / 256, 256, 0, ST(stream), reinterpret_cast<const long long*>(gap), w, gate, N, C,
             ksize, inv_hw, out_scale);
  LAUNCH_CHECK();
}

extern "C" int vip_layernorm_f32(const float* x, const float* gamma, const float* beta, float* out, int M, int C, float eps,
                                 void* stream) {
  VIP_REQUIRE(x && gamma && beta && out && M > 0 && C > 0, VIP_ERR_INVALID, "vip_layernorm_f32: bad argument");
  VIP_LAUNCH((layernorm_f32_kernel), (M * 32 + 127) / 128, 128, 0, ST(stream), x, gamma, beta, out, M, C, eps);
  LAUNCH_CHECK();
}

extern "C" int vip_row_stats_finalize(const int64_t* records, long long M, int cols, float eps, float* out, void* stream) {
  VIP_REQUIRE(records && out && M > 0 && cols > 0, VIP_ERR_INVALID, "vip_row_stats_finalize: bad argument");
  VIP_LAUNCH((row_stats_finalize_kernel), grid_for(M, 256), 256, 0, ST(stream), reinterpret_cast<const long long*>(records),
             reinterpret_cast<float2*>(out), M, 1.0f / (float)cols, eps);
  LAUNCH_CHECK();
}

extern "C" int vip_layernorm_bf16(const void* x, const float* gamma, const float* beta, void* out, float* row_stats,
                                  float next_eps, long long M, int C, float eps, void* stream) {
  VIP_REQUIRE(x && out && gamma && beta && C % 8 == 0 && C <= 1024, VIP_ERR_INVALID,
              "vip_layernorm_bf16: bad argument (C %% 8 == 0, C <= 1024)");
  const bf16* xp = (const bf16*)x;
  bf16* op = (bf16*)out;
  const int c8n = C / 8;
  cudaStream_t st = ST(stream);
  // narrow rows (C <= 128): 4 lanes per row (64 contiguous bytes per row and instruction), two row groups in flight.
  // Measured at [12.8 M, 96] on B200: 16 lanes per row 1.51 ms, one thread per row 1.42, 4 lanes 1.19, 4 lanes x 2 groups 1.12.
  if (c8n <= 8) VIP_LAUNCH((layernorm_kernel<4, 2, 2>), grid_for(M * 2, 256), 256, 0, st, xp, gamma, beta, op, row_stats, next_eps, M, C, eps);
  else if (c8n <= 12) VIP_LAUNCH((layernorm_kernel<4, 3, 2>), grid_for(M * 2, 256), 256, 0, st, xp, gamma, beta, op, row_stats, next_eps, M, C, eps);
  else if (c8n <= 16) VIP_LAUNCH((layernorm_kernel<4, 4, 2>), grid_for(M * 2, 256), 256, 0, st, xp, gamma, beta, op, row_stats, next_eps, M, C, eps);
  else if (c8n <= 32) VIP_LAUNCH((layernorm_kernel<32, 1, 1>), grid_for(M * 32, 256), 256, 0, st, xp, gamma, beta, op, row_stats, next_eps, M, C, eps);
  else if (c8n <= 64) VIP_LAUNCH((layernorm_kernel<32, 2, 1>), grid_for(M * 32, 256), 256, 0, st, xp, gamma, beta, op, row_stats, next_eps, M, C, eps);
  else VIP_LAUNCH((layernorm_kernel<32, 4, 1>), grid_for(M * 32, 256), 256, 0, st, xp, gamma, beta, op, row_stats, next_eps, M, C, eps);
  LAUNCH_CHECK();
}

extern "C" int vip_dwconv3x3_bf16(const void* x, const float* w, void* out, int64_t* gap_, int N, int H, int W, int C, int gelu,
                                  void* stream) {
  long long* gap = reinterpret_cast<long long*>(gap_);
  VIP_REQUIRE(x && out && w && C % 8 == 0, VIP_ERR_INVALID, "vip_dwconv3x3_bf16: bad argument");
  // measured at [1024, 100, 100, 96] on B200: 8 channels per thread, 4 rows ahead, 3 blocks per SM (168 registers) 1.59 ms;
  // 2 blocks (202 registers) 1.61; depth 2 / 6 the same; 4 channels per thread 1.73; the register-window version 2.21
  const long long threads = (long long)N * W * (C / 8);
  VIP_LAUNCH((dwconv3x3_kernel<4, 4, 3>), (unsigned)((threads + 127) / 128), 128, 0, ST(stream), (const bf16*)x, w, (bf16*)out, gap, N, H, W, C, gelu);
  LAUNCH_CHECK();
}

extern "C" int vip_maxpool3s2_bf16(const void* x, void* out, int N, int H, int W, int C, void* stream) {
  VIP_REQUIRE(x && out && C % 8 == 0, VIP_ERR_INVALID, "vip_maxpool3s2_bf16: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  VIP_LAUNCH((maxpool3s2_zeropad_kernel), grid_for((long long)N * Ho * Wo * (C / 8), 256), 256, 0, ST(stream), (const bf16*)x, (bf16*)out, N, H, W, C, Ho, Wo);
  LAUNCH_CHECK();
}

extern "C" int vip_head_f32(const float* feat, const float* w, const float* b, float* probs, double* acc,
                            double acc_weight, int N, int C, int k, int sigmoid_head, void* stream) {
  VIP_REQUIRE(feat && w && b && probs && k >= 1 && k <= 8, VIP_ERR_INVALID, "vip_head_f32: bad argument (1 <= k <= 8)");
  VIP_LAUNCH((head_kernel), (N * 32 + 127) / 128, 128, 0, ST(stream), feat, w, b, probs, acc, acc_weight, N, C, k, sigmoid_head);
  LAUNCH_CHECK();
}

extern "C" int vip_scale_cast_fx_bf16(const int64_t* x, float scale, void* out, long long n, void* stream) {
  VIP_REQUIRE(x && out, VIP_ERR_INVALID, "vip_scale_cast_fx_bf16: null pointer");
  VIP_LAUNCH((scale_cast_fx_bf16_kernel), grid_for(n, 256), 256, 0, ST(stream), reinterpret_cast<const long long*>(x), scale, (bf16*)out, n);
  LAUNCH_CHECK();
}

extern "C" int vip_pad_crop(const void* x, int N, int H, int W, int bytes_per_pixel, void* out, int Ho, int Wo, int top, int left,
                            void* stream) {
  VIP_REQUIRE(x && out && N > 0 && H > 0 && W > 0 && Ho > 0 && Wo > 0 && bytes_per_pixel > 0 && bytes_per_pixel % 8 == 0,
              VIP_ERR_INVALID, "vip_pad_crop: bad argument (bytes_per_pixel must be a multiple of 8)");
  VIP_REQUIRE((((uintptr_t)x | (uintptr_t)out) & 7) == 0, VIP_ERR_INVALID, "vip_pad_crop: unaligned pointer");
  const int U = bytes_per_pixel / 8;
  VIP_LAUNCH((pad_crop_kernel), grid_for((long long)N * Ho * Wo * U, 256), 256, 0, ST(stream), (const uint2*)x, (uint2*)out, N, H, W,
             U, Ho, Wo, top, left);
  LAUNCH_CHECK();
}

extern "C" int vip_scale_weights_bf16(const void* w, int ldw, const float* gate, int G, int N, int K, void* out, void* stream) {
  VIP_REQUIRE(w && gate && out && G > 0 && N > 0 && K > 0 && ldw >= K, VIP_ERR_INVALID, "vip_scale_weights_bf16: bad argument");
  VIP_LAUNCH((scale_weights_kernel), grid_for((long long)G * N * K, 256), 256, 0, ST(stream), (const bf16*)w, ldw, gate, (bf16*)out, G,
             N, K);
  LAUNCH_CHECK();
}

extern "C" int vip_cast_f32_bf16(const float* x, void* out, long long n, void* stream) {
  VIP_REQUIRE(x && out, VIP_ERR_INVALID, "vip_cast_f32_bf16: null pointer");
  VIP_LAUNCH((cast_f32_bf16_kernel), grid_for(n, 256), 256, 0, ST(stream), x, (bf16*)out, n);
  LAUNCH_CHECK();
}
