// Streaming preprocessing kernel for sm_100a: the path main.py actually runs (dataset/dataset.py:31-37 -- cast -> bicubic
// resize -> / 255.0 -- plus the flip / gray flags of dataset/augment.py:115-120,142-146; no crop, no JPEG emulation).
//
// The fused JPEG-emulation kernel (preprocess.cu) keeps one image per CTA with its colour planes in shared memory; used
// for this plain resize it ran at 0.15 of the HBM roofline (1024 CTAs over 296 slots, 4-byte staging, quad-ordered
// stores).  This kernel streams instead:
//   grid   = (stripes of kRows output rows, images): >= 12 waves of small CTAs at batch 1024
//   stage  : the source rows a stripe touches are one contiguous byte range of the image -> 16-byte loads into smem
//   warp   = one output row at a time: vertical taps of the row (4 source bytes per lane and step -> packed fp32 pairs,
//            separately rounded packed products and sums like TF's un-contracted CPU kernel, half the issue slots)
//            -> the warp's fp32 row buffer -> horizontal taps, / 255, gray, one lane per pixel -> the warp's output row
//            buffer (flips applied as addressing) -> 16-byte coalesced streaming stores
//   tables : tap weights / indices depend only on (Hs, Ws, Ho, Wo): computed once per geometry by tap_table_kernel into
//            a cached device buffer instead of once per CTA (the LUT entries need fp64 arithmetic, which is slow here)
//   identity (Hs, Ws) == (Ho, Wo), no flags: the taps are exactly (0, 1, 0, 0), so out = float(u8) / 255 -- a flat
//            8-bytes-in / 16-or-32-bytes-out streaming loop without staging.
// Bit-identical to preprocess.cu and to oracle/preprocess.py (tests/test_preprocess_gpu.py).
#include <cuda_bf16.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "resize_math.cuh"

namespace vip {
namespace {

constexpr int kRows = 16;       // output rows per CTA
constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

struct StreamArgs {
  const uint8_t* src;
  const uint8_t* flags;
  void* dst;
  const float4* wy;     // [Ho] tap weights, rows
  const short4* iy;     // [Ho] tap indices
  const float4* wx;     // [Wo]
  const short4* ix;     // [Wo]
  int N, Hs, Ws, Ho, Wo;
  int pitch;            // source bytes per row (3 Ws, a multiple of 4)
  int src_cap;          // bytes of the staging buffer
  int identity;
  int off_ix, off_src, off_v, off_o;   // smem byte offsets (wx first)
  int v_stride, o_stride;              // bytes per warp buffer
  float one;                           // 1.0f, opaque to the compiler (see tap4x2)
};

__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmul.rn.f32x2 rd, ra, rb;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmov.b64 rc, {%6,%7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
// tap4 (resize_math.cuh) on two independent values: ((p0 w0 + p1 w1) + p2 w2) + p3 w3, every product and sum rounded
// separately.  ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (also with -fmad=false), which would round
// once where TF rounds twice; so the sums are written as fma(x, one, y) with `one` = 1.0f read from the kernel arguments:
// exact (x * 1 + y rounds like x + y), packed, and not foldable at compile time.
__device__ __forceinline__ float2 tap4x2(float2 p0, float2 p1, float2 p2, float2 p3, const float4 w, const float2 one) {
  const float2 a = mul2(p0, make_float2(w.x, w.x)), b = mul2(p1, make_float2(w.y, w.y));
  const float2 c = mul2(p2, make_float2(w.z, w.z)), d = mul2(p3, make_float2(w.w, w.w));
  return fma2(fma2(fma2(a, one, b), one, c), one, d);
}
// div255 (resize_math.cuh) on a pair
__device__ __forceinline__ float2 div255x2(float2 x) {
  const float rc = 0.003921568859368562698f;
  const float2 q0 = mul2(x, make_float2(rc, rc));
  const float2 r = fma2(make_float2(-q0.x, -q0.y), make_float2(255.0f, 255.0f), x);
  return fma2(r, make_float2(rc, rc), q0);
}

__global__ void tap_table_kernel(int Hs, int Ws, int Ho, int Wo, float4* wy, short4* iy, float4* wx, short4* ix) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < Ho) compute_tap(t, Hs, Ho, &wy[t], &iy[t]);
  else if (t < Ho + Wo) compute_tap(t - Ho, Ws, Wo, &wx[t - Ho], &ix[t - Ho]);
}

template <bool kBf16>
__device__ __forceinline__ void put3(uint8_t* orow, int x, float r, float g, float b) {
  if (kBf16) {
    __nv_bfloat16* p = reinterpret_cast<__nv_bfloat16*>(orow) + 3 * x;
    p[0] = __float2bfloat16_rn(r);
    p[1] = __float2bfloat16_rn(g);
    p[2] = __float2bfloat16_rn(b);
  } else {
    float* p = reinterpret_cast<float*>(orow) + 3 * x;
    p[0] = r;
    p[1] = g;
    p[2] = b;
  }
}

template <bool kBf16>
__global__ void __launch_bounds__(kThreads, 3) resize_stream_kernel(const StreamArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = blockIdx.y, oy0 = blockIdx.x * kRows;
  const int Ho = a.Ho, Wo = a.Wo;
  const int rows = min(kRows, Ho - oy0);
  const unsigned flags = a.flags != nullptr ? a.flags[n] : 0u;
  constexpr int kEsz = kBf16 ? 2 : 4;
  const size_t img_in = (size_t)n * a.Hs * a.pitch;
  const size_t row_out_bytes = (size_t)Wo * 3 * kEsz;
  uint8_t* dst_img = reinterpret_cast<uint8_t*>(a.dst) + (size_t)n * Ho * row_out_bytes;

  if (a.identity && flags == 0u) {
    // out = float(u8) / 255 on the stripe's rows, flat: 8 source bytes -> 8 outputs per step
    const uint8_t* sp = a.src + img_in + (size_t)oy0 * a.pitch;
    uint8_t* dp = dst_img + (size_t)oy0 * row_out_bytes;
    const int groups = rows * a.pitch / 8;     // pitch % 8 == 0 on this path (checked on the host)
    for (int g = tid; g < groups; g += kThreads) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(sp) + g);
      float2 f[4];
      f[0] = div255x2(make_float2(u8f(w.x, 0), u8f(w.x, 1)));
      f[1] = div255x2(make_float2(u8f(w.x, 2), u8f(w.x, 3)));
      f[2] = div255x2(make_float2(u8f(w.y, 0), u8f(w.y, 1)));
      f[3] = div255x2(make_float2(u8f(w.y, 2), u8f(w.y, 3)));
      if (kBf16) {
        uint32_t o[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __nv_bfloat162 h = __floats2bfloat162_rn(f[k].x, f[k].y);
          o[k] = *reinterpret_cast<const uint32_t*>(&h);
        }
        __stcs(reinterpret_cast<uint4*>(dp) + g, make_uint4(o[0], o[1], o[2], o[3]));
      } else {
        __stcs(reinterpret_cast<float4*>(dp) + 2 * g, make_float4(f[0].x, f[0].y, f[1].x, f[1].y));
        __stcs(reinterpret_cast<float4*>(dp) + 2 * g + 1, make_float4(f[2].x, f[2].y, f[3].x, f[3].y));
      }
    }
    return;
  }

  float4* s_wx = reinterpret_cast<float4*>(smem);
  short4* s_ix = reinterpret_cast<short4*>(smem + a.off_ix);
  uint8_t* s_src = smem + a.off_src;
  float* s_v = reinterpret_cast<float*>(smem + a.off_v + warp * a.v_stride);
  uint8_t* s_o = smem + a.off_o + warp * a.o_stride;

  // ---- stage: tap tables of the columns, and the source rows [lo, hi] of this stripe as one contiguous byte range
  for (int t = tid; t < Wo; t += kThreads) {
    s_wx[t] = __ldg(a.wx + t);
    s_ix[t] = __ldg(a.ix + t);
  }
  const int lo = __ldg(a.iy + oy0).x, hi = __ldg(a.iy + oy0 + rows - 1).w;
  const size_t first = img_in + (size_t)lo * a.pitch;          // byte offset of row lo inside the batch
  const int delta = (int)(first & 15);                          // src is 16-byte aligned: align the range down
  {
    const uint4* g0 = reinterpret_cast<const uint4*>(a.src + (first - delta));
    const int nchunks = (delta + (hi - lo + 1) * a.pitch + 15) >> 4;   // the last chunk may run past the row: the host
    uint4* d0 = reinterpret_cast<uint4*>(s_src);                         // pads the allocation requirement (see below)
    for (int i = tid; i < nchunks; i += kThreads) d0[i] = __ldg(g0 + i);
  }
  __syncthreads();

  const int nwords = a.pitch >> 2;
  const float2 one = make_float2(a.one, a.one);
  for (int r = warp; r < rows; r += kWarps) {
    const int oy = oy0 + r;
    const float4 wy = __ldg(a.wy + oy);
    const short4 iy = __ldg(a.iy + oy);
    // vertical taps: 4 source bytes (of 4 rows) per lane and step
    const unsigned* t0 = reinterpret_cast<const unsigned*>(s_src + delta + (iy.x - lo) * a.pitch);
    const unsigned* t1 = reinterpret_cast<const unsigned*>(s_src + delta + (iy.y - lo) * a.pitch);
    const unsigned* t2 = reinterpret_cast<const unsigned*>(s_src + delta + (iy.z - lo) * a.pitch);
    const unsigned* t3 = reinterpret_cast<const unsigned*>(s_src + delta + (iy.w - lo) * a.pitch);
    for (int j = lane; j < nwords; j += 32) {
      const unsigned A = t0[j], B = t1[j], C = t2[j], D = t3[j];
      const float2 v01 = tap4x2(make_float2(u8f(A, 0), u8f(A, 1)), make_float2(u8f(B, 0), u8f(B, 1)),
                                make_float2(u8f(C, 0), u8f(C, 1)), make_float2(u8f(D, 0), u8f(D, 1)), wy, one);
      const float2 v23 = tap4x2(make_float2(u8f(A, 2), u8f(A, 3)), make_float2(u8f(B, 2), u8f(B, 3)),
                                make_float2(u8f(C, 2), u8f(C, 3)), make_float2(u8f(D, 2), u8f(D, 3)), wy, one);
      reinterpret_cast<float4*>(s_v)[j] = make_float4(v01.x, v01.y, v23.x, v23.y);
    }
    __syncwarp();
    // horizontal taps, / 255, gray: one lane per output pixel; flips are addressing
    for (int ox = lane; ox < Wo; ox += 32) {
      const float4 wx = s_wx[ox];
      const short4 ix = s_ix[ox];
      const float* p0 = s_v + 3 * ix.x;
      const float* p1 = s_v + 3 * ix.y;
      const float* p2 = s_v + 3 * ix.z;
      const float* p3 = s_v + 3 * ix.w;
      const float2 rg = div255x2(tap4x2(make_float2(p0[0], p0[1]), make_float2(p1[0], p1[1]), make_float2(p2[0], p2[1]),
                                        make_float2(p3[0], p3[1]), wx, one));
      const float bl = div255(tap4(p0[2], p1[2], p2[2], p3[2], wx));
      float px[3] = {rg.x, rg.y, bl};
      if (flags & VIP_FLAG_GRAY) gray3(px);
      put3<kBf16>(s_o, (flags & VIP_FLAG_HFLIP) ? Wo - 1 - ox : ox, px[0], px[1], px[2]);
    }
    __syncwarp();
    // the finished row -> global, 16 bytes per lane and step
    {
      const int orow = (flags & VIP_FLAG_VFLIP) ? Ho - 1 - oy : oy;
      uint4* gp = reinterpret_cast<uint4*>(dst_img + (size_t)orow * row_out_bytes);
      const uint4* sp = reinterpret_cast<const uint4*>(s_o);
      const int chunks = (int)(row_out_bytes >> 4);
      for (int k = lane; k < chunks; k += 32) __stcs(gp + k, sp[k]);
    }
    __syncwarp();
  }
}

struct TapTables {
  float4 *wy, *wx;
  short4 *iy, *ix;
};

// tap tables per (device, geometry), computed on first use (outside stream capture) and kept for the life of the process
int get_tables(int Hs, int Ws, int Ho, int Wo, cudaStream_t st, TapTables* out) {
  static std::mutex mu;
  static std::map<std::tuple<int, int, int, int, int>, TapTables> cache;
  int dev = 0;
  VIP_CUDA(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lock(mu);
  const auto key = std::make_tuple(dev, Hs, Ws, Ho, Wo);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return VIP_OK;
  }
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  VIP_CUDA(cudaStreamIsCapturing(st, &cap));
  VIP_REQUIRE(cap == cudaStreamCaptureStatusNone, VIP_ERR_STATE,
              "vip_preprocess: the first call for a geometry (%dx%d -> %dx%d) builds its tap tables and must not happen "
              "inside a CUDA graph capture; run it once eagerly first", Hs, Ws, Ho, Wo);
  uint8_t* buf = nullptr;
  const size_t bytes = (size_t)(Ho + Wo) * (sizeof(float4) + sizeof(short4));
  VIP_CUDA(cudaMalloc(&buf, bytes));
  TapTables t;
  t.wy = reinterpret_cast<float4*>(buf);
  t.wx = t.wy + Ho;
  t.iy = reinterpret_cast<short4*>(t.wx + Wo);
  t.ix = t.iy + Ho;
  tap_table_kernel<<<(Ho + Wo + 127) / 128, 128, 0, st>>>(Hs, Ws, Ho, Wo, t.wy, t.iy, t.wx, t.ix);
  VIP_CUDA(cudaGetLastError());
  count_launch();
  cache[key] = t;
  *out = t;
  return VIP_OK;
}

int align_up(int v, int a) { return (v + a - 1) / a * a; }

}  // namespace

// Whether the streaming kernel covers this call (otherwise the caller takes the fused kernel of preprocess.cu).
bool preprocess_stream_supported(const uint8_t* src, int Hs, int Ws, const int32_t* crop, const int32_t* jq, int Ho, int Wo,
                                 const void* dst, int dst_dtype) {
  if (crop != nullptr || jq != nullptr) return false;
  const int esz = dst_dtype == VIP_DTYPE_BF16 ? 2 : 4;
  if ((Ws * 3) % 4 != 0 || (Hs * Ws * 3) % 16 != 0) return false;          // word-addressed rows, 16-byte aligned images
  if ((Wo * 3 * esz) % 16 != 0) return false;                              // 16-byte output rows
  if ((reinterpret_cast<uintptr_t>(src) & 15) != 0 || (reinterpret_cast<uintptr_t>(dst) & 15) != 0) return false;
  if (Hs == Ho && Ws == Wo && (Ws * 3) % 8 != 0) return false;
  return Ws <= 1024 && Wo <= 1024 && Ho <= 1024;
}

int preprocess_stream(const uint8_t* src, int N, int Hs, int Ws, const uint8_t* flags, int Ho, int Wo, void* dst, int dst_dtype,
                      cudaStream_t st) {
  StreamArgs a{};
  a.src = src; a.flags = flags; a.dst = dst;
  a.N = N; a.Hs = Hs; a.Ws = Ws; a.Ho = Ho; a.Wo = Wo;
  a.pitch = Ws * 3;
  a.identity = (Hs == Ho && Ws == Wo) ? 1 : 0;
  a.one = 1.0f;
  TapTables t;
  int rc = get_tables(Hs, Ws, Ho, Wo, st, &t);
  if (rc != VIP_OK) return rc;
  a.wy = t.wy; a.iy = t.iy; a.wx = t.wx; a.ix = t.ix;
  const int esz = dst_dtype == VIP_DTYPE_BF16 ? 2 : 4;
  // source rows one stripe can touch: taps are monotonic, span <= ceil((kRows - 1) * Hs / Ho) + 4 rows (+1 slack)
  const int src_rows = std::min(Hs, (int)(((long long)(kRows - 1) * Hs + Ho - 1) / Ho) + 5);
  int off = Wo * 16;
  a.off_ix = off; off += Wo * 8;
  off = align_up(off, 16);
  a.off_src = off; a.src_cap = align_up(src_rows * a.pitch + 32, 16); off += a.src_cap;
  a.v_stride = align_up(a.pitch * 4, 16);
  a.off_v = off; off += kWarps * a.v_stride;
  a.o_stride = align_up(Wo * 3 * esz, 16);
  a.off_o = off; off += kWarps * a.o_stride;
  if (off > 227 * 1024) return 1;   // too wide for this kernel: the caller falls back to the fused kernel
  auto kern = dst_dtype == VIP_DTYPE_BF16 ? resize_stream_kernel<true> : resize_stream_kernel<false>;
  VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, off));
  VIP_LAUNCH((kern), dim3((Ho + kRows - 1) / kRows, N), kThreads, off, st, a);
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

}  // namespace vip
