// libjpeg "islow" integer DCT pieces shared by the JPEG-quality emulation (preprocess.cu) and the device JPEG decoder
// (jpeg_decode.cu): the constants of jfdctint.c / jidctint.c, the inverse DCT passes, and the 8-lane block transpose.
// Restated from the published algorithm (Loeffler-Ligtenberg-Moschytz, CONST_BITS 13, PASS1_BITS 2); bit-exact against
// libjpeg-turbo's output through Pillow (tests/golden/jpeg_pillow.npz, tests/test_jpeg_decode_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vip {

constexpr int kTrStride = 72;        // words per block in the transpose scratch (64 + 8: conflict-free columns)

// ---- libjpeg islow DCTs (jfdctint.c / jidctint.c), one 8-vector per call -----------------------------
#define C0298 2446
#define C0390 3196
#define C0541 4433
#define C0765 6270
#define C0899 7373
#define C1175 9633
#define C1501 12299
#define C1847 15137
#define C1961 16069
#define C2053 16819
#define C2562 20995
#define C3072 25172

// kPass1: columns (descale 11).  Pass 2: rows, descale 18, +128 folded into the even part.
template <bool kPass1>
__device__ __forceinline__ void idct8(int (&d)[8]) {
  constexpr int n = kPass1 ? 11 : 18;
  constexpr int bias = (1 << (n - 1)) + (kPass1 ? 0 : (128 << 18));
  const int z1e = (d[2] + d[6]) * C0541;
  const int t2e = z1e - d[6] * C1847;
  const int t3e = z1e + d[2] * C0765;
  const int t0e = ((d[0] + d[4]) << 13) + bias;
  const int t1e = ((d[0] - d[4]) << 13) + bias;
  const int t10 = t0e + t3e, t13 = t0e - t3e;
  const int t11 = t1e + t2e, t12 = t1e - t2e;
  int t0 = d[7], t1 = d[5], t2 = d[3], t3 = d[1];
  const int z1 = t0 + t3, z2 = t1 + t2, z3 = t0 + t2, z4 = t1 + t3;
  const int z5 = (z3 + z4) * C1175;
  const int z3m = z5 - z3 * C1961;
  const int z4m = z5 - z4 * C0390;
  const int z1m = -z1 * C0899;
  const int z2m = -z2 * C2562;
  t0 = t0 * C0298 + z1m + z3m;
  t1 = t1 * C2053 + z2m + z4m;
  t2 = t2 * C3072 + z2m + z3m;
  t3 = t3 * C1501 + z1m + z4m;
  d[0] = (t10 + t3) >> n;
  d[7] = (t10 - t3) >> n;
  d[1] = (t11 + t2) >> n;
  d[6] = (t11 - t2) >> n;
  d[2] = (t12 + t1) >> n;
  d[5] = (t12 - t1) >> n;
  d[3] = (t13 + t0) >> n;
  d[4] = (t13 - t0) >> n;
}

// 8x8 transpose across the 8 lanes of a block group with warp shuffles: three butterfly stages (lane ^ 4, ^ 2, ^ 1), each
// exchanging the half of the registers whose index differs from the partner's in that bit -- 12 SHFL + 24 SEL, no shared
// memory, no barrier.  `r` = lane & 7 (row held on entry, column held on exit).
__device__ __forceinline__ void transpose8_shfl(int (&d)[8], int r) {
#pragma unroll
  for (int m = 4; m >= 1; m >>= 1) {
    const bool up = (r & m) != 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i & m) continue;
      const int send = up ? d[i] : d[i | m];
      const int recv = __shfl_xor_sync(0xffffffffu, send, m);
      if (up) d[i] = recv;
      else d[i | m] = recv;
    }
  }
}

// 8x8 transpose across the 8 lanes of a block group through a per-warp smem scratch.
__device__ __forceinline__ void transpose8(int (&d)[8], int* scr, int b, int r) {
  int4* wp = reinterpret_cast<int4*>(scr + b * kTrStride + r * 8);
  wp[0] = make_int4(d[0], d[1], d[2], d[3]);
  wp[1] = make_int4(d[4], d[5], d[6], d[7]);
  __syncwarp();
  const int* rp = scr + b * kTrStride + r;
#pragma unroll
  for (int k = 0; k < 8; ++k) d[k] = rp[k * 8];
  __syncwarp();
}


}  // namespace vip
