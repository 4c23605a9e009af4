// Fused preprocessing kernel for sm_100a: crop -> TF bicubic resize -> /255 -> libjpeg 4:2:0 round trip at
// quality q -> flips -> gray, one CTA per image, Y/Cb/Cr planes resident in shared memory.
//
// Reference semantics (restated in oracle/preprocess.py, which is the checker, never linked here):
//   dataset/dataset.py:31-37      cast -> tf.image.resize(bicubic) -> / 255.0
//   dataset/augment.py:110-113    tf.image.random_jpeg_quality (libjpeg-turbo baseline encode+decode)
//   dataset/augment.py:115-120    flips;  dataset/augment.py:142-146 gray
//
// Exactness rules: every fp32 product / sum of the resize is a separately rounded __fmul_rn/__fadd_rn (TF's CPU
// kernel is not FMA-contracted); x/255 uses a 3-instruction sequence proven equal to IEEE division (checked
// exhaustively by vip_selftest_div255); the JPEG part is pure 32-bit integer arithmetic.
//
// Phases per image (CTA of kThreads threads):
//   0  tap tables (4 indices + 4 weights per output row / column), quantisation tables
//   A  stripes of kStripe output rows: vertical taps -> fp32 stripe in smem -> horizontal taps, /255,
//      quantise to u8, RGB->YCbCr, 2x2 chroma average -> u8 planes in smem        (no-JPEG images store here)
//   B  8x8 blocks, 8 lanes per block: FDCT rows -> 8-lane transpose -> FDCT cols -> quantise/dequantise ->
//      IDCT cols -> 8-lane transpose -> IDCT rows -> clamp, in place (transposes: shared-memory scratch or warp shuffles)
//   C  fancy chroma upsampling, YCbCr->RGB, *1/255, flips / gray, vectorised stores
#include <cuda_bf16.h>

#include <algorithm>
#include <stdlib.h>

#include "common.cuh"
#include "jpeg_math.cuh"
#include "resize_math.cuh"

namespace vip {
namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kStripe = 8;           // output rows per phase-A stripe (even)

struct PreArgs {
  const uint8_t* src;
  const int32_t* crop;
  const int32_t* jq;
  const uint8_t* flags;
  void* dst;
  int N, Hs, Ws, Ho, Wo;
  int hc, wc;              // real chroma size ceil(Ho/2), ceil(Wo/2)
  int HpY, WpY, HpC, WpC;  // padded plane sizes (multiples of 8)
  int v_pitch;             // floats per stripe row (= 3 * Ws rounded up to 4 pixels)
  int src_pitch;           // bytes per staged source row
  int word_path;           // 1: every source row is 4-byte aligned -> cp.async word staging
  int off_wy, off_iy, off_wx, off_ix, off_qt, off_Y, off_Cb, off_Cr, off_src, off_v;  // smem byte offsets
};

__constant__ uint8_t c_luma_base[64] = {16, 11, 10, 16, 24,  40,  51,  61,  12, 12, 14, 19, 26,  58,  60,  55,
                                        14, 13, 16, 24, 40,  57,  69,  56,  14, 17, 22, 29, 51,  87,  80,  62,
                                        18, 22, 37, 56, 68,  109, 103, 77,  24, 35, 55, 64, 81,  104, 113, 92,
                                        49, 64, 78, 87, 103, 121, 120, 101, 72, 92, 95, 98, 112, 100, 103, 99};
__constant__ uint8_t c_chroma_base[64] = {17, 18, 24, 47, 99, 99, 99, 99, 18, 21, 26, 66, 99, 99, 99, 99,
                                          24, 26, 56, 99, 99, 99, 99, 99, 47, 66, 99, 99, 99, 99, 99, 99,
                                          99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99,
                                          99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99, 99};

// kPass1: rows of level-shifted samples; dc_bias is added to output 0 before scaling (folds the -128 shift).
template <bool kPass1>
__device__ __forceinline__ void fdct8(int (&d)[8]) {
  const int t0 = d[0] + d[7], t7 = d[0] - d[7];
  const int t1 = d[1] + d[6], t6 = d[1] - d[6];
  const int t2 = d[2] + d[5], t5 = d[2] - d[5];
  const int t3 = d[3] + d[4], t4 = d[3] - d[4];
  const int t10 = t0 + t3, t13 = t0 - t3;
  const int t11 = t1 + t2, t12 = t1 - t2;
  constexpr int n = kPass1 ? 11 : 15;
  constexpr int rnd = 1 << (n - 1);
  if (kPass1) {
    d[0] = (t10 + t11 - 8 * 128) << 2;   // level shift of all 8 samples only reaches the DC term
    d[4] = (t10 - t11) << 2;
  } else {
    d[0] = (t10 + t11 + 2) >> 2;
    d[4] = (t10 - t11 + 2) >> 2;
  }
  const int z1e = (t12 + t13) * C0541 + rnd;
  d[2] = (z1e + t13 * C0765) >> n;
  d[6] = (z1e - t12 * C1847) >> n;
  const int z1 = t4 + t7, z2 = t5 + t6, z3 = t4 + t6, z4 = t5 + t7;
  const int z5 = (z3 + z4) * C1175 + rnd;
  const int z3m = z5 - z3 * C1961;
  const int z4m = z5 - z4 * C0390;
  const int z1m = -z1 * C0899;
  const int z2m = -z2 * C2562;
  d[7] = (t4 * C0298 + z1m + z3m) >> n;
  d[5] = (t5 * C2053 + z2m + z4m) >> n;
  d[3] = (t6 * C3072 + z2m + z3m) >> n;
  d[1] = (t7 * C1501 + z1m + z4m) >> n;
}

// ---- output -----------------------------------------------------------------------------------------
// Stores up to two horizontally adjacent pixels (already in output order: px[0..2] is the left one).
template <bool kBf16>
__device__ __forceinline__ void store_px(void* img, size_t img_elem0, int Wo, int oy, int ox_left, const float* px,
                                         int count) {
  // img is the batch base (8-byte aligned, checked on the host); e is the element index inside the batch
  const size_t e = img_elem0 + ((size_t)oy * Wo + ox_left) * 3;
  if (kBf16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(img) + e;
    if (count == 2 && (e & 1) == 0) {
      __nv_bfloat162* p2 = reinterpret_cast<__nv_bfloat162*>(p);
      p2[0] = __floats2bfloat162_rn(px[0], px[1]);
      p2[1] = __floats2bfloat162_rn(px[2], px[3]);
      p2[2] = __floats2bfloat162_rn(px[4], px[5]);
    } else {
      for (int k = 0; k < 3 * count; ++k) p[k] = __float2bfloat16_rn(px[k]);
    }
  } else {
    float* p = reinterpret_cast<float*>(img) + e;
    if (count == 2 && (e & 1) == 0) {
      float2* p2 = reinterpret_cast<float2*>(p);
      __stcs(p2 + 0, make_float2(px[0], px[1]));
      __stcs(p2 + 1, make_float2(px[2], px[3]));
      __stcs(p2 + 2, make_float2(px[4], px[5]));
    } else {
      for (int k = 0; k < 3 * count; ++k) __stcs(p + k, px[k]);
    }
  }
}

// Emits the 2x2 quad whose top-left real pixel is (2cy, 2cx); v[dy][dx][c] are final float values.
template <bool kBf16>
__device__ __forceinline__ void emit_quad(void* img, size_t e0, int Ho, int Wo, int cy, int cx, unsigned flags,
                                          float (&v)[2][2][3]) {
  const int nx = (2 * cx + 1 < Wo) ? 2 : 1;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy) {
    const int oy = 2 * cy + dy;
    if (oy >= Ho) break;
    if (flags & VIP_FLAG_GRAY) {
      gray3(v[dy][0]);
      gray3(v[dy][1]);
    }
    const int oyo = (flags & VIP_FLAG_VFLIP) ? (Ho - 1 - oy) : oy;
    float px[6];
    if (flags & VIP_FLAG_HFLIP) {
      // mirrored: the right pixel of the pair becomes the left one
      if (nx == 2) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          px[c] = v[dy][1][c];
          px[3 + c] = v[dy][0][c];
        }
        store_px<kBf16>(img, e0, Wo, oyo, Wo - 2 - 2 * cx, px, 2);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) px[c] = v[dy][0][c];
        store_px<kBf16>(img, e0, Wo, oyo, Wo - 1 - 2 * cx, px, 1);
      }
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        px[c] = v[dy][0][c];
        px[3 + c] = v[dy][1][c];
      }
      store_px<kBf16>(img, e0, Wo, oyo, 2 * cx, px, nx);
    }
  }
}

// ---- the kernel -------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// floor(i / d) for 0 <= i, i * d < 2^32: multiply by ceil(2^32 / d) and keep the high word (d == 1 passes through)
struct FastDiv {
  unsigned m;
  int d;
  __device__ __forceinline__ explicit FastDiv(int d_) : m(0xffffffffu / (unsigned)d_ + 1u), d(d_) {}
  __device__ __forceinline__ int div(int i) const { return d == 1 ? i : (int)__umulhi((unsigned)i, m); }
};

// kShfl: the two 8x8 transposes of a block's DCT round trip go through warp shuffles (VIP_PRE_SHFL=1) instead of the
// per-warp shared-memory scratch (default: measured 5 % faster, see vip_preprocess).
template <bool kBf16, bool kShfl>
__global__ void __launch_bounds__(kThreads, 2) preprocess_kernel(const PreArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  float4* s_wy = reinterpret_cast<float4*>(smem + a.off_wy);
  short4* s_iy = reinterpret_cast<short4*>(smem + a.off_iy);
  float4* s_wx = reinterpret_cast<float4*>(smem + a.off_wx);
  short4* s_ix = reinterpret_cast<short4*>(smem + a.off_ix);
  int* s_qt = reinterpret_cast<int*>(smem + a.off_qt);  // [0..63] luma t, [64..127] chroma t, [128..255] magics
  uint8_t* s_Y = smem + a.off_Y;
  uint8_t* s_Cb = smem + a.off_Cb;
  uint8_t* s_Cr = smem + a.off_Cr;
  uint8_t* s_src = smem + a.off_src;
  float* s_v = reinterpret_cast<float*>(smem + a.off_v);

  const int tid = threadIdx.x;
  const int n = blockIdx.x;
  const int Ho = a.Ho, Wo = a.Wo;

  int y0 = 0, x0 = 0, h = a.Hs, w = a.Ws;
  if (a.crop != nullptr) {
    y0 = a.crop[4 * n + 0];
    x0 = a.crop[4 * n + 1];
    h = a.crop[4 * n + 2];
    w = a.crop[4 * n + 3];
  }
  const int q = (a.jq != nullptr) ? a.jq[n] : -1;
  const bool jpeg = q >= 0;
  const unsigned flags = (a.flags != nullptr) ? a.flags[n] : 0u;
  const int row_pitch = a.Ws * 3;
  // first byte of the crop; with word_path every source row starts 4-byte aligned, so only x0 misaligns it
  const uint8_t* src = a.src + (size_t)n * a.Hs * row_pitch + (size_t)y0 * row_pitch + (size_t)x0 * 3;
  const int mis = a.word_path ? ((x0 * 3) & 3) : 0;
  const int w3 = w * 3;
  void* dst = a.dst;
  const size_t e0 = (size_t)n * Ho * Wo * 3;

  // ---- phase 0: taps and quantisation tables
  for (int t = tid; t < Ho + Wo; t += kThreads) {
    if (t < Ho) compute_tap(t, h, Ho, &s_wy[t], &s_iy[t]);
    else        compute_tap(t - Ho, w, Wo, &s_wx[t - Ho], &s_ix[t - Ho]);
  }
  if (jpeg && tid < 128) {
    int qq = q <= 0 ? 1 : (q > 100 ? 100 : q);
    const int s = qq < 50 ? 5000 / qq : 200 - 2 * qq;
    const int base = tid < 64 ? c_luma_base[tid] : c_chroma_base[tid - 64];
    const int t = min(max((base * s + 50) / 100, 1), 255);
    s_qt[tid] = t;
    s_qt[128 + tid] = (int)((unsigned)(0x100000000ull / (unsigned)(8 * t)) + 1u);  // floor(n/8t) == umulhi(n, M)
  }
  __syncthreads();

  const int hc = a.hc, wc = a.wc;
  const int WqC = jpeg ? a.WpC : wc;  // quad columns (JPEG planes need the right-edge padding)
  const FastDiv div_q(WqC);
  const int G = (w + 3) >> 2;                                   // 4-pixel groups per source row
  const FastDiv div_g(G);
  const unsigned fsel = 0x3210u + 0x1111u * (unsigned)mis;       // byte funnel: drop the `mis` leading bytes

  // stages the source rows [lo, hi] of the crop that the stripe starting at output row oy0 reads
  auto stage_rows = [&](int oy0) {
    const int rows = min(kStripe, Ho - oy0);
    const int lo = s_iy[oy0].x, hi = s_iy[oy0 + rows - 1].w;
    const int nrows = hi - lo + 1;
    if (a.word_path) {
      const int nwords = (mis + w3 + 3) >> 2;
      const FastDiv div_w(nwords);
      const uint8_t* g0 = src - mis + (size_t)lo * row_pitch;
      for (int i = tid; i < nrows * nwords; i += kThreads) {
        const int r = div_w.div(i);
        const int j = i - r * nwords;
        cp_async4(s_src + r * a.src_pitch + 4 * j, g0 + (size_t)r * row_pitch + 4 * j);
      }
    } else {
      const FastDiv div_b(w3);
      const uint8_t* g0 = src + (size_t)lo * row_pitch;
      for (int i = tid; i < nrows * w3; i += kThreads) {
        const int r = div_b.div(i);
        const int j = i - r * w3;
        s_src[r * a.src_pitch + j] = __ldg(g0 + (size_t)r * row_pitch + j);
      }
    }
  };

  stage_rows(0);
  cp_async_wait_all();
  __syncthreads();

  // ---- phase A
  for (int cy0 = 0; cy0 < hc; cy0 += kStripe / 2) {
    const int oy0 = 2 * cy0;
    const int rows = min(kStripe, Ho - oy0);  // real rows in this stripe
    // A2: vertical taps; one item = (stripe row, group of 4 source pixels = 12 bytes = 3 words after the funnel)
    {
      const int lo = s_iy[oy0].x;
      for (int it = tid; it < rows * G; it += kThreads) {
        const int r = div_g.div(it);
        const int g = it - r * G;
        const float4 wy = s_wy[oy0 + r];
        const short4 iy = s_iy[oy0 + r];
        const uint8_t* base = s_src + 12 * g;
        const unsigned* t0 = reinterpret_cast<const unsigned*>(base + (iy.x - lo) * a.src_pitch);
        const unsigned* t1 = reinterpret_cast<const unsigned*>(base + (iy.y - lo) * a.src_pitch);
        const unsigned* t2 = reinterpret_cast<const unsigned*>(base + (iy.z - lo) * a.src_pitch);
        const unsigned* t3 = reinterpret_cast<const unsigned*>(base + (iy.w - lo) * a.src_pitch);
        unsigned A[4], B[4], C[4], D[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { A[k] = t0[k]; B[k] = t1[k]; C[k] = t2[k]; D[k] = t3[k]; }
        float o[12];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          const unsigned wa = __byte_perm(A[j], A[j + 1], fsel), wb = __byte_perm(B[j], B[j + 1], fsel);
          const unsigned wc_ = __byte_perm(C[j], C[j + 1], fsel), wd = __byte_perm(D[j], D[j + 1], fsel);
#pragma unroll
          for (int k = 0; k < 4; ++k) o[4 * j + k] = tap4(u8f(wa, k), u8f(wb, k), u8f(wc_, k), u8f(wd, k), wy);
        }
        float4* vp = reinterpret_cast<float4*>(s_v + r * a.v_pitch + 12 * g);
        vp[0] = make_float4(o[0], o[1], o[2], o[3]);
        vp[1] = make_float4(o[4], o[5], o[6], o[7]);
        vp[2] = make_float4(o[8], o[9], o[10], o[11]);
      }
    }
    __syncthreads();
    if (cy0 + kStripe / 2 < hc) stage_rows(oy0 + kStripe);  // async: lands while A3 runs
    // A3: horizontal taps per 2x2 quad, /255, quantise, colour convert
    const int nq = (kStripe / 2) * WqC;
    for (int qi = tid; qi < nq; qi += kThreads) {
      const int qr = div_q.div(qi);
      const int cx = qi - qr * WqC;
      const int cy = cy0 + qr;
      if (cy >= hc) continue;
      float fv[2][2][3];
      int cbs = 0, crs = 0;
      unsigned ypk[2] = {0u, 0u};
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int ox = min(2 * cx + dx, Wo - 1);
        const float4 wx = s_wx[ox];
        const short4 ix = s_ix[ox];
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
          const int rr = min(2 * qr + dy, rows - 1);
          const float* vr = s_v + rr * a.v_pitch;
          const float* p0 = vr + ix.x * 3;
          const float* p1 = vr + ix.y * 3;
          const float* p2 = vr + ix.z * 3;
          const float* p3 = vr + ix.w * 3;
          int rgb[3];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const float val = tap4(p0[c], p1[c], p2[c], p3[c], wx);
            const float f = div255(val);
            fv[dy][dx][c] = f;
            // tf.image.convert_image_dtype(float -> uint8, saturate=True): trunc(clip(f * 255.5, 0, 255))
            rgb[c] = trunc_small_float(fminf(fmaxf(__fmul_rn(f, 255.5f), 0.0f), 255.0f));
          }
          if (jpeg) {
            // jccolor.c rgb_ycc_convert
            const int Y = (19595 * rgb[0] + 38470 * rgb[1] + 7471 * rgb[2] + 32768) >> 16;
            cbs += (-11059 * rgb[0] - 21709 * rgb[1] + 32768 * rgb[2] + (128 << 16) + 32767) >> 16;
            crs += (32768 * rgb[0] - 27439 * rgb[1] - 5329 * rgb[2] + (128 << 16) + 32767) >> 16;
            ypk[dy] |= (unsigned)Y << (8 * dx);
          }
        }
      }
      if (jpeg) {
        if (2 * cx < a.WpY) {   // WpY is even: both pixels of the pair are inside or both outside
          *reinterpret_cast<unsigned short*>(s_Y + (2 * cy) * a.WpY + 2 * cx) = (unsigned short)ypk[0];
          *reinterpret_cast<unsigned short*>(s_Y + (2 * cy + 1) * a.WpY + 2 * cx) = (unsigned short)ypk[1];
        }
        const int bias = 1 + (cx & 1);  // jcsample.c h2v2_downsample: 1,2,1,2,...
        s_Cb[cy * a.WpC + cx] = (uint8_t)((cbs + bias) >> 2);
        s_Cr[cy * a.WpC + cx] = (uint8_t)((crs + bias) >> 2);
      } else if (cx < wc) {
        emit_quad<kBf16>(dst, e0, Ho, Wo, cy, cx, flags, fv);
      }
    }
    cp_async_wait_all();
    __syncthreads();
  }
  if (!jpeg) return;

  // bottom-edge replication to whole blocks (jcprepct.c expand_bottom_edge)
  {
    const int ylast = 2 * hc - 1;
    for (int i = tid; i < (a.HpY - 2 * hc) * a.WpY; i += kThreads) {
      const int r = i / a.WpY, x = i - r * a.WpY;
      s_Y[(2 * hc + r) * a.WpY + x] = s_Y[ylast * a.WpY + x];
    }
    for (int i = tid; i < (a.HpC - hc) * a.WpC; i += kThreads) {
      const int r = i / a.WpC, x = i - r * a.WpC;
      s_Cb[(hc + r) * a.WpC + x] = s_Cb[(hc - 1) * a.WpC + x];
      s_Cr[(hc + r) * a.WpC + x] = s_Cr[(hc - 1) * a.WpC + x];
    }
  }
  __syncthreads();

  // ---- phase B: DCT -> quantise -> dequantise -> IDCT, in place.  8 lanes per block, 4 blocks per warp.
  {
    const int warp = tid >> 5, lane = tid & 31;
    const int b = lane >> 3, r = lane & 7;
    int* scr = reinterpret_cast<int*>(s_v) + warp * (4 * kTrStride);
    const int nbxY = a.WpY >> 3, nbY = (a.HpY >> 3) * nbxY;
    const int nbxC = a.WpC >> 3, nbC = (a.HpC >> 3) * nbxC;
    const FastDiv div_bY(nbxY), div_bC(nbxC);
    const int itY = (nbY + 3) >> 2, itC = (2 * nbC + 3) >> 2;
    int cur = -1;
    int qt[8], qm[8];
    for (int it = warp; it < itY + itC; it += kWarps) {
      const int comp = it < itY ? 0 : 1;
      if (comp != cur) {
        cur = comp;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          qt[v] = s_qt[comp * 64 + v * 8 + r];
          qm[v] = s_qt[128 + comp * 64 + v * 8 + r];
        }
      }
      int bi = (comp ? it - itY : it) * 4 + b;
      const int nb = comp ? 2 * nbC : nbY;
      const bool live = bi < nb;
      bi = live ? bi : nb - 1;
      uint8_t* plane;
      int pitch, by;
      if (comp == 0) {
        plane = s_Y; pitch = a.WpY;
        by = div_bY.div(bi);
        bi -= by * nbxY;
      } else {
        const bool second = bi >= nbC;
        plane = second ? s_Cr : s_Cb;
        bi -= second ? nbC : 0;
        pitch = a.WpC;
        by = div_bC.div(bi);
        bi -= by * nbxC;
      }
      uint2* rowp = reinterpret_cast<uint2*>(plane + (by * 8 + r) * pitch + bi * 8);
      const uint2 raw = *rowp;
      int d[8];
      d[0] = raw.x & 255; d[1] = (raw.x >> 8) & 255; d[2] = (raw.x >> 16) & 255; d[3] = raw.x >> 24;
      d[4] = raw.y & 255; d[5] = (raw.y >> 8) & 255; d[6] = (raw.y >> 16) & 255; d[7] = raw.y >> 24;
      fdct8<true>(d);                 // row r of the block
      if (kShfl) transpose8_shfl(d, r);   // lane r now holds column u = r, d[v]
      else transpose8(d, scr, b, r);
      fdct8<false>(d);
#pragma unroll
      for (int v = 0; v < 8; ++v) {
        // jcdctmgr.c quantize(): sign(c) * ((|c| + (8t >> 1)) / 8t); dequantise: * t
        const int c = d[v];
        const int mag = __umulhi((unsigned)(abs(c) + 4 * qt[v]), (unsigned)qm[v]) * qt[v];
        d[v] = c < 0 ? -mag : mag;
      }
      idct8<true>(d);                 // column pass (over v)
      if (kShfl) transpose8_shfl(d, r);   // lane r holds row y = r
      else transpose8(d, scr, b, r);
      idct8<false>(d);
      uint2 o;
      o.x = __vimin_s32_relu(d[0], 255) | (__vimin_s32_relu(d[1], 255) << 8) | (__vimin_s32_relu(d[2], 255) << 16) |
            (__vimin_s32_relu(d[3], 255) << 24);
      o.y = __vimin_s32_relu(d[4], 255) | (__vimin_s32_relu(d[5], 255) << 8) | (__vimin_s32_relu(d[6], 255) << 16) |
            (__vimin_s32_relu(d[7], 255) << 24);
      if (live) *rowp = o;
    }
  }
  __syncthreads();

  // ---- phase C: fancy upsampling (jdsample.c h2v2_fancy_upsample), YCbCr->RGB (jdcolor.c), scale, store
  {
    const int nq = hc * wc;
    const FastDiv div_c(wc);
    for (int qi = tid; qi < nq; qi += kThreads) {
      const int cy = div_c.div(qi);
      const int cx = qi - cy * wc;
      const int o_m = max(cy - 1, 0) * a.WpC, o_c = cy * a.WpC, o_p = min(cy + 1, hc - 1) * a.WpC;
      const int cxm = max(cx - 1, 0), cxp = min(cx + 1, wc - 1);
      int ch[2][2][2];  // [comp][dy][dx] upsampled chroma (not centred)
#pragma unroll
      for (int comp = 0; comp < 2; ++comp) {
        const uint8_t* P = comp ? s_Cr : s_Cb;
        const int n_l = P[o_c + cxm], n_c = P[o_c + cx], n_r = P[o_c + cxp];
        // row 2cy: nearest = cy, further = cy-1;  row 2cy+1: further = cy+1
        const int u_l = 3 * n_l + P[o_m + cxm], u_c = 3 * n_c + P[o_m + cx], u_r = 3 * n_r + P[o_m + cxp];
        const int d_l = 3 * n_l + P[o_p + cxm], d_c = 3 * n_c + P[o_p + cx], d_r = 3 * n_r + P[o_p + cxp];
        ch[comp][0][0] = (3 * u_c + u_l + 8) >> 4;
        ch[comp][0][1] = (3 * u_c + u_r + 7) >> 4;
        ch[comp][1][0] = (3 * d_c + d_l + 8) >> 4;
        ch[comp][1][1] = (3 * d_c + d_r + 7) >> 4;
        // jdsample.c jinit_upsampler: planes of one or two samples per row are replicated (h2v2_upsample), not interpolated
        if (wc <= 2) ch[comp][0][0] = ch[comp][0][1] = ch[comp][1][0] = ch[comp][1][1] = n_c;
      }
      float fv[2][2][3];
#pragma unroll
      for (int dy = 0; dy < 2; ++dy) {
        const int yy = min(2 * cy + dy, a.HpY - 1);
        const int xx = min(2 * cx, a.WpY - 2);
        const unsigned ypair = *reinterpret_cast<const unsigned short*>(s_Y + yy * a.WpY + xx);
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          // R = Y + ((91881*cr' + 32768) >> 16) with cr' = cr - 128: fold Y and the -128 into one addend
          const int yb = (int)(((ypair >> (8 * dx)) & 255u) << 16) + 32768;
          const int cb = ch[0][dy][dx], cr = ch[1][dy][dx];
          const int R = __vimin_s32_relu((91881 * cr + (yb - 128 * 91881)) >> 16, 255);
          const int G = __vimin_s32_relu((-22554 * cb - 46802 * cr + (yb + 128 * (22554 + 46802))) >> 16, 255);
          const int B = __vimin_s32_relu((116130 * cb + (yb - 128 * 116130)) >> 16, 255);
          // convert_image_dtype(uint8 -> float32): cast * (1/255)
          fv[dy][dx][0] = __fmul_rn(small_int_to_float(R), 0.003921568859368562698f);
          fv[dy][dx][1] = __fmul_rn(small_int_to_float(G), 0.003921568859368562698f);
          fv[dy][dx][2] = __fmul_rn(small_int_to_float(B), 0.003921568859368562698f);
        }
      }
      emit_quad<kBf16>(dst, e0, Ho, Wo, cy, cx, flags, fv);
    }
  }
}

__global__ void div255_selftest_kernel(unsigned long long* mismatches) {
  pdl_trigger();
  pdl_wait();
  const unsigned long long total = 1ull << 32;
  unsigned long long bad = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const float x = __uint_as_float((unsigned)i);
    const float ax = fabsf(x);
    if (!(ax <= 1e30f)) continue;                 // inf / nan / out of contract
    if (ax < 1e-30f && (unsigned)i != 0u) continue;  // tiny / -0: out of contract
    const float got = div255(x);
    const float ref = __fdiv_rn(x, 255.0f);
    if (__float_as_uint(got) != __float_as_uint(ref)) ++bad;
  }
  if (bad) atomicAdd(mismatches, bad);
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

// preprocess_stream.cu
bool preprocess_stream_supported(const uint8_t* src, int Hs, int Ws, const int32_t* crop, const int32_t* jq, int Ho, int Wo,
                                 const void* dst, int dst_dtype);
int preprocess_stream(const uint8_t* src, int N, int Hs, int Ws, const uint8_t* flags, int Ho, int Wo, void* dst, int dst_dtype,
                      cudaStream_t st);
}  // namespace vip

extern "C" int vip_preprocess(const uint8_t* src, int N, int Hs, int Ws, const int32_t* crop_yxhw,
                              const int32_t* jpeg_q, const uint8_t* flags, int Ho, int Wo, void* dst, int dst_dtype,
                              void* cuda_stream) {
  using namespace vip;
  VIP_REQUIRE(N >= 0, VIP_ERR_INVALID, "vip_preprocess: N < 0");
  if (N == 0) return VIP_OK;
  VIP_REQUIRE(src != nullptr && dst != nullptr, VIP_ERR_INVALID, "vip_preprocess: null src/dst");
  VIP_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 7) == 0, VIP_ERR_INVALID, "vip_preprocess: dst must be 8-byte aligned");
  VIP_REQUIRE(Hs >= 1 && Ws >= 1 && Ho >= 1 && Wo >= 1, VIP_ERR_INVALID, "vip_preprocess: empty image");
  VIP_REQUIRE(dst_dtype == VIP_DTYPE_F32 || dst_dtype == VIP_DTYPE_BF16, VIP_ERR_INVALID,
              "vip_preprocess: dst_dtype must be VIP_DTYPE_F32 or VIP_DTYPE_BF16");
  VIP_REQUIRE(Hs <= 32767 && Ws <= 1024 && Ho <= 1024 && Wo <= 1024, VIP_ERR_UNSUPPORTED,
              "vip_preprocess: size out of range (Ws, Ho, Wo <= 1024)");
  const bool jpeg = jpeg_q != nullptr;
  VIP_REQUIRE(!jpeg || (Ho <= 256 && Wo <= 256), VIP_ERR_UNSUPPORTED,
              "vip_preprocess: JPEG emulation needs Ho, Wo <= 256 (planes live in shared memory)");
  // plain resize / normalise / flips (what main.py runs): the streaming kernel; VIP_PRE_STREAM=0 keeps the fused kernel
  static const bool stream_on = [] { const char* v = getenv("VIP_PRE_STREAM"); return v == nullptr || v[0] != '0'; }();
  if (stream_on && preprocess_stream_supported(src, Hs, Ws, crop_yxhw, jpeg_q, Ho, Wo, dst, dst_dtype)) {
    const int rc = preprocess_stream(src, N, Hs, Ws, flags, Ho, Wo, dst, dst_dtype, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (rc <= 0) return rc;   // rc > 0: geometry too wide for the streaming kernel's shared memory -> fused kernel
  }

  PreArgs a{};
  a.src = src; a.crop = crop_yxhw; a.jq = jpeg_q; a.flags = flags; a.dst = dst;
  a.N = N; a.Hs = Hs; a.Ws = Ws; a.Ho = Ho; a.Wo = Wo;
  a.hc = (Ho + 1) / 2; a.wc = (Wo + 1) / 2;
  a.HpY = align_up(Ho, 8); a.WpY = align_up(Wo, 8);
  a.HpC = align_up(a.hc, 8); a.WpC = align_up(a.wc, 8);
  a.v_pitch = align_up(Ws, 4) * 3;
  a.src_pitch = align_up(align_up(Ws, 4) * 3 + 8, 16);
  a.word_path = ((Ws * 3) % 4 == 0 && (reinterpret_cast<uintptr_t>(src) & 3) == 0) ? 1 : 0;
  // source rows one stripe can touch: taps are monotonic, span <= ceil((kStripe-1) * Hs/Ho) + 4 rows (+1 slack)
  const int src_rows = std::min(Hs, (int)(((long long)(kStripe - 1) * Hs + Ho - 1) / Ho) + 5);
  int off = 0;
  a.off_wy = off; off += Ho * 16;
  a.off_wx = off; off += Wo * 16;
  a.off_iy = off; off += Ho * 8;
  a.off_ix = off; off += Wo * 8;
  off = align_up(off, 16);
  a.off_qt = off; off += 256 * 4;
  if (jpeg) {
    a.off_Y = off; off += align_up(a.HpY * a.WpY, 16);
    a.off_Cb = off; off += align_up(a.HpC * a.WpC, 16);
    a.off_Cr = off; off += align_up(a.HpC * a.WpC, 16);
  }
  a.off_src = off; off += align_up(src_rows * a.src_pitch, 16);
  a.off_v = off;
  int vbytes = kStripe * a.v_pitch * 4;
  const int scr_bytes = kWarps * 4 * kTrStride * 4;
  if (jpeg && vbytes < scr_bytes) vbytes = scr_bytes;
  off += vbytes;
  VIP_REQUIRE(off <= 227 * 1024, VIP_ERR_UNSUPPORTED, "vip_preprocess: %d bytes of shared memory needed (max 232448)",
              off);

  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  // measured on B200 (benchmarks/pre_shfl_ab.py, 4096 images): shuffle transposes 2.58 ms, shared-memory scratch 2.47 ms --
  // 12 SHFL + 24 SEL per transpose cost more issue slots than 2 STS.128 + 8 LDS.32 in this issue-bound kernel, so the
  // scratch stays the default; VIP_PRE_SHFL=1 selects the shuffle version (bit-identical results)
  static const bool shfl = [] { const char* v = getenv("VIP_PRE_SHFL"); return v != nullptr && v[0] == '1'; }();
  auto kern = dst_dtype == VIP_DTYPE_BF16 ? (shfl ? preprocess_kernel<true, true> : preprocess_kernel<true, false>)
                                          : (shfl ? preprocess_kernel<false, true> : preprocess_kernel<false, false>);
  VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, off));
  VIP_LAUNCH((kern), N, kThreads, off, st, a);
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

extern "C" int vip_selftest_div255(uint64_t* mismatches, void* cuda_stream) {
  using namespace vip;
  VIP_REQUIRE(mismatches != nullptr, VIP_ERR_INVALID, "vip_selftest_div255: null output");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(cuda_stream);
  unsigned long long* d = nullptr;
  VIP_CUDA(cudaMalloc(&d, sizeof(unsigned long long)));
  VIP_CUDA(cudaMemsetAsync(d, 0, sizeof(unsigned long long), st));
  VIP_LAUNCH((div255_selftest_kernel), 148 * 8, 256, 0, st, d);
  count_launch();
  unsigned long long hres = 0;
  cudaError_t e = cudaMemcpyAsync(&hres, d, sizeof(hres), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d);
  if (e != cudaSuccess) return cuda_fail(e, "div255 selftest");
  *mismatches = hres;
  return VIP_OK;
}
