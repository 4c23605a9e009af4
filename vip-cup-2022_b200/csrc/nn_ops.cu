// Non-GEMM layer kernels of the backbones (bf16 NHWC activations, fp32 math), all HBM-bound and written as
// 16-byte vectorised, coalesced grid-stride kernels.  Each entry point cites the reference layer it replaces.
#include <cuda_bf16.h>

#include "common.cuh"

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
template <int LPR, int J>
__global__ void __launch_bounds__(256) layernorm_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, bf16* __restrict__ out,
                                                        float* __restrict__ row_stats, long long M, int C, float eps) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int lane = threadIdx.x & 31, sub = lane % LPR, rsel = lane / LPR;
  const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
  const int c8n = C >> 3;
  const float invC = 1.0f / (float)C;
  for (long long m0 = warp0 * RPW; m0 < M; m0 += nwarps * RPW) {
    const long long m = m0 + rsel;
    const bool live = m < M;
    const bf16* row = x + (live ? m : 0) * C;
    float f[J][8];
    float s = 0.0f;
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int c8 = sub + LPR * j;
      if (c8 < c8n) {
        unpack8(*reinterpret_cast<const bf16x8*>(row + c8 * 8), f[j]);
#pragma unroll
        for (int k = 0; k < 8; ++k) s += f[j][k];
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
          const float d = f[j][k] - mean;
          v = fmaf(d, d, v);
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const float rstd = rsqrtf(v * invC + eps);
    float os = 0.0f, oq = 0.0f;  // statistics of the rounded output row (for a LayerNorm folded into the next GEMM)
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
        for (int k = 0; k < 8; ++k) o8[k] = fmaf((f[j][k] - mean) * rstd, gg[k], bb[k]);
        const bf16x8 pk = pack8(o8);
        if (live) *reinterpret_cast<bf16x8*>(out + m * C + c8 * 8) = pk;
        if (row_stats != nullptr) {
          float r8[8];
          unpack8(pk, r8);
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            os += r8[k];
            oq = fmaf(r8[k], r8[k], oq);
          }
        }
      }
    }
    if (row_stats != nullptr) {
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) {
        os += __shfl_xor_sync(0xffffffffu, os, o);
        oq += __shfl_xor_sync(0xffffffffu, oq, o);
      }
      if (sub == 0 && live) *reinterpret_cast<float2*>(row_stats + 2 * m) = make_float2(os, oq);
    }
  }
}

// ---- ZeroPadding2D(1) + DepthwiseConv2D(3, valid, no bias) (+ GELU) (feature.py:92-94,132-134).  A thread owns one
// (image, column, 8-channel group) and walks down the rows with the 3x3 window and the 72 weights in registers, so every
// input element is fetched three times (once per neighbouring column) instead of nine.
__global__ void __launch_bounds__(128) dwconv3x3_kernel(const bf16* __restrict__ x, const float* __restrict__ w /*[3][3][C]*/,
                                                        bf16* __restrict__ out, float* __restrict__ gap /*[N][C] or null*/,
                                                        int N, int H, int W, int C, int gelu) {
  const int c8n = C >> 3;
  const long long total = (long long)N * W * c8n;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int c8 = (int)(idx % c8n);
  const long long t = idx / c8n;
  const int ox = (int)(t % W), n = (int)(t / W);
  float wt[9][8];
#pragma unroll
  for (int k = 0; k < 9; ++k) {
    const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + k * C + c8 * 8));
    const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + k * C + c8 * 8) + 1);
    wt[k][0] = w0.x; wt[k][1] = w0.y; wt[k][2] = w0.z; wt[k][3] = w0.w;
    wt[k][4] = w1.x; wt[k][5] = w1.y; wt[k][6] = w1.z; wt[k][7] = w1.w;
  }
  const bf16* img = x + (long long)n * H * W * C + c8 * 8;
  const bool has_l = ox > 0, has_r = ox + 1 < W;
  float win[3][3][8];  // [row y-1, y, y+1][col x-1, x, x+1]
  auto load_row = [&](int y, float (&dst)[3][8]) {
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const bool ok = y >= 0 && y < H && (dx == 1 || (dx == 0 ? has_l : has_r));
      if (ok) {
        unpack8(*reinterpret_cast<const bf16x8*>(img + ((long long)y * W + ox - 1 + dx) * C), dst[dx]);
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[dx][k] = 0.0f;
      }
    }
  };
#pragma unroll
  for (int dx = 0; dx < 3; ++dx)
#pragma unroll
    for (int k = 0; k < 8; ++k) win[0][dx][k] = 0.0f;
  load_row(0, win[1]);
  bf16* op = out + (long long)n * H * W * C + (long long)ox * C + c8 * 8;
  float colsum[8] = {0, 0, 0, 0, 0, 0, 0, 0};  // of the rounded outputs of this column (SE squeeze, feature.py:55)
  for (int y = 0; y < H; ++y) {
    load_row(y + 1, win[2]);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.0f;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int sx = 0; sx < 3; ++sx)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = fmaf(win[r][sx][k], wt[r * 3 + sx][k], acc[k]);
    if (gelu) {
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = gelu_erf(acc[k]);
    }
    const bf16x8 pk = pack8(acc);
    *reinterpret_cast<bf16x8*>(op + (long long)y * W * C) = pk;
    if (gap != nullptr) {
      float r8[8];
      unpack8(pk, r8);
#pragma unroll
      for (int k = 0; k < 8; ++k) colsum[k] += r8[k];
    }
#pragma unroll
    for (int dx = 0; dx < 3; ++dx)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        win[0][dx][k] = win[1][dx][k];
        win[1][dx][k] = win[2][dx][k];
      }
  }
  if (gap != nullptr) {
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(gap + (long long)n * C + c8 * 8 + k, colsum[k]);
  }
}

// ---- ZeroPadding2D(1) + MaxPool2D(3, 2, 'valid') (feature.py:139,151-152): the padded zeros take part in the max
__global__ void maxpool3s2_zeropad_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, int N, int H, int W, int C,
                                          int Ho, int Wo) {
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
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i]);
}

// ---- out = bf16(x * scale): pooled sums -> means as a GEMM operand (SE squeeze, resnet_rs_model.py:149)
__global__ void scale_cast_f32_bf16_kernel(const float* __restrict__ x, float scale, bf16* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16_rn(x[i] * scale);
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
    im2col_vec8_kernel<<<grid_for(work, 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)out, N, H, W, C, ksize, stride,
                                                                    pad, Ho, Wo, Kp);
  } else if (Kp == 32 && ksize == 3 && C == 3) {
    im2col_row_kernel<32, 3, 3><<<grid_for((long long)N * Ho * Wo, 128), 128, 0, ST(stream)>>>((const bf16*)x, (bf16*)out, N,
                                                                                             H, W, stride, pad, Ho, Wo);
  } else {
    const long long work = (long long)N * Ho * Wo * Kp;
    im2col_scalar_kernel<<<grid_for(work, 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)out, N, H, W, C, ksize,
                                                                      stride, pad, Ho, Wo, Kp);
  }
  LAUNCH_CHECK();
}

extern "C" int vip_avgpool2_same_bf16(const void* x, int N, int H, int W, int C, void* out, void* stream) {
  VIP_REQUIRE(x && out && C % 8 == 0, VIP_ERR_INVALID, "vip_avgpool2_same_bf16: bad argument (C %% 8)");
  const int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
  avgpool2_same_kernel<<<grid_for((long long)N * Ho * Wo * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)out,
                                                                                               N, H, W, C, Ho, Wo);
  LAUNCH_CHECK();
}

extern "C" int vip_global_avgpool_bf16(const void* x, int N, int HW, int C, void* out_bf16, float* out_f32, void* stream) {
  VIP_REQUIRE(x && (out_bf16 || out_f32) && C % 8 == 0, VIP_ERR_INVALID, "vip_global_avgpool_bf16: bad argument");
  dim3 grid(N, (C + 63) / 64);
  global_avgpool_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)out_bf16, out_f32, HW, C);
  LAUNCH_CHECK();
}

extern "C" int vip_scale_add_act_bf16(const void* y, const float* gate, const void* shortcut, void* out, int N, int HW,
                                      int C, int act, void* stream) {
  VIP_REQUIRE(y && out && C % 8 == 0, VIP_ERR_INVALID, "vip_scale_add_act_bf16: bad argument");
  const long long total8 = (long long)N * HW * (C / 8);
  scale_add_act_kernel<<<grid_for(total8, 256), 256, 0, ST(stream)>>>((const bf16*)y, gate, (const bf16*)shortcut,
                                                                      (bf16*)out, total8, HW, C, act);
  LAUNCH_CHECK();
}

extern "C" int vip_layernorm_bf16(const void* x, const float* gamma, const float* beta, void* out, float* row_stats,
                                  long long M, int C, float eps, void* stream) {
  VIP_REQUIRE(x && out && gamma && beta && C % 8 == 0 && C <= 1024, VIP_ERR_INVALID,
              "vip_layernorm_bf16: bad argument (C %% 8 == 0, C <= 1024)");
  const bf16* xp = (const bf16*)x;
  bf16* op = (bf16*)out;
  const int c8n = C / 8;
  cudaStream_t st = ST(stream);
  if (c8n <= 8) layernorm_kernel<8, 1><<<grid_for(M * 8, 256), 256, 0, st>>>(xp, gamma, beta, op, row_stats, M, C, eps);
  else if (c8n <= 16) layernorm_kernel<16, 1><<<grid_for(M * 16, 256), 256, 0, st>>>(xp, gamma, beta, op, row_stats, M, C, eps);
  else if (c8n <= 32) layernorm_kernel<32, 1><<<grid_for(M * 32, 256), 256, 0, st>>>(xp, gamma, beta, op, row_stats, M, C, eps);
  else if (c8n <= 64) layernorm_kernel<32, 2><<<grid_for(M * 32, 256), 256, 0, st>>>(xp, gamma, beta, op, row_stats, M, C, eps);
  else layernorm_kernel<32, 4><<<grid_for(M * 32, 256), 256, 0, st>>>(xp, gamma, beta, op, row_stats, M, C, eps);
  LAUNCH_CHECK();
}

extern "C" int vip_dwconv3x3_bf16(const void* x, const float* w, void* out, float* gap, int N, int H, int W, int C, int gelu,
                                  void* stream) {
  VIP_REQUIRE(x && out && w && C % 8 == 0, VIP_ERR_INVALID, "vip_dwconv3x3_bf16: bad argument");
  const long long threads = (long long)N * W * (C / 8);
  dwconv3x3_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ST(stream)>>>((const bf16*)x, w, (bf16*)out, gap, N, H, W, C,
                                                                             gelu);
  LAUNCH_CHECK();
}

extern "C" int vip_maxpool3s2_bf16(const void* x, void* out, int N, int H, int W, int C, void* stream) {
  VIP_REQUIRE(x && out && C % 8 == 0, VIP_ERR_INVALID, "vip_maxpool3s2_bf16: bad argument");
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  maxpool3s2_zeropad_kernel<<<grid_for((long long)N * Ho * Wo * (C / 8), 256), 256, 0, ST(stream)>>>(
      (const bf16*)x, (bf16*)out, N, H, W, C, Ho, Wo);
  LAUNCH_CHECK();
}

extern "C" int vip_head_f32(const float* feat, const float* w, const float* b, float* probs, double* acc,
                            double acc_weight, int N, int C, int k, int sigmoid_head, void* stream) {
  VIP_REQUIRE(feat && w && b && probs && k >= 1 && k <= 8, VIP_ERR_INVALID, "vip_head_f32: bad argument (1 <= k <= 8)");
  head_kernel<<<(N * 32 + 127) / 128, 128, 0, ST(stream)>>>(feat, w, b, probs, acc, acc_weight, N, C, k, sigmoid_head);
  LAUNCH_CHECK();
}

extern "C" int vip_scale_cast_f32_bf16(const float* x, float scale, void* out, long long n, void* stream) {
  VIP_REQUIRE(x && out, VIP_ERR_INVALID, "vip_scale_cast_f32_bf16: null pointer");
  scale_cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, scale, (bf16*)out, n);
  LAUNCH_CHECK();
}

extern "C" int vip_cast_f32_bf16(const float* x, void* out, long long n, void* stream) {
  VIP_REQUIRE(x && out, VIP_ERR_INVALID, "vip_cast_f32_bf16: null pointer");
  cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(x, (bf16*)out, n);
  LAUNCH_CHECK();
}
