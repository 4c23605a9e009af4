// Window attention of GCViT on the 5th-generation tensor cores (tcgen05 / TMEM).  Same contract as attention.cu
// (models/gcvit/layers/attention.py:52-83, window.py:3-14 folded into the addressing); that file's warp-level mma.sync
// kernel turned out to be bound by the legacy HMMA path of sm_100 (~145 TFLOP/s for the whole chip), so the two
// products move to tcgen05.mma:
//
//   S = Q K^T   A = Q tile [128 rows x 32] and B = K [keys x 32] in shared memory (128-byte swizzled rows, staged with
//               cp.async), D = S [128 x keys] fp32 in TMEM.  ws 14: one window per tile, 196 queries = two 128-row
//               tiles, 224 key columns (window rows padded 14 -> 16 so that a 32-column TMEM chunk is two whole window
//               rows); ws 7: TWO windows per tile (rows 0..63 / 64..127, key columns 0..63 / 64..127, window rows
//               padded 7 -> 8), the cross-window quadrants are computed and ignored.
//   softmax     one thread per query row (TMEM lane = thread): pass 1 reads the row and takes its maximum, pass 2 reads it
//               again, adds the relative-position bias (shared-memory table, compile-time offsets), exponentiates and
//               writes P as packed bf16 back INTO TMEM over the columns of S it has already consumed.
//   O = P V     A = P straight from TMEM (tcgen05.mma with a TMEM A operand), B = V^T [32 x keys] in shared memory
//               (transposed while staging), D = O [128 x 32] fp32 in TMEM -> registers -> / row sum -> bf16 -> global.
// 256 threads per CTA: two threads per query row (warps w and w + 4 own the same TMEM lane quarter and split the key
// columns), because the softmax -- not the MMAs -- is what the kernel spends its time on.  One (window group, head) item
// per CTA, 256 TMEM columns, so two CTAs share an SM and one's softmax overlaps the other's loads and MMAs.
#include <cuda_bf16.h>

#include "common.cuh"

namespace vip {
namespace {

using bf16 = __nv_bfloat16;
constexpr int HD = 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc)
      : "memory");
}
// K-major, 128-byte swizzle: rows 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int WS>
struct TcCfg {
  static constexpr int N = WS * WS;
  static constexpr int WP = WS <= 8 ? 8 : 16;            // padded window-row length of the key order
  static constexpr int WPT = WS <= 8 ? 2 : 1;            // windows per 128-row tile
  static constexpr int WKEYS = WS <= 8 ? 64 : 224;       // key columns one query row looks at
  static constexpr int KCOLS = WPT * WKEYS;              // columns of S = UMMA N of the first product (128 / 224)
  static constexpr int MT = (N + 127) / (128 / WPT) >= 1 ? (WS <= 8 ? 1 : 2) : 1;  // 128-row query tiles per item
  static constexpr int QPW = 128 / WPT;                  // query rows reserved per window in a tile (64 / 128)
  static constexpr int KSTEPS = KCOLS / 16;              // K steps of the second product (8 / 14)
  static constexpr int KBLK = (KCOLS + 63) / 64;         // 64-key blocks of V^T (2 / 4)
  static constexpr int TAB = (2 * WS - 1) * (2 * WS - 1);
  static constexpr int O_COL = 128;                      // TMEM column of O (beyond everything P occupies)
  static constexpr int kTmemCols = 256;
  static constexpr int off_q = 0;                                     // MT query tiles, 16 KB each
  static constexpr int off_k = MT * 128 * 128;
  static constexpr int off_v = off_k + ((KCOLS + 7) / 8) * 8 * 128;  // V^T blocks
  static constexpr int off_vr = off_v + KBLK * 32 * 128;             // V as loaded: [key][32 channels], 64-byte rows
  static constexpr int off_t = off_vr + KCOLS * 64;
  static constexpr int NC = WKEYS / 32;                  // 32-column chunks one row looks at (2 / 7)
  static constexpr int CA = (NC + 1) / 2;                // chunks of the first thread of a row; the second takes the rest
  static constexpr int off_x = off_t + ((TAB * 4 + 15) / 16) * 16 + 32;   // [2][128] row maxima, [2][128] row sums
  static constexpr int off_bar = off_x + 4 * 128 * 4;
  static constexpr int kSmem = off_bar + 32 + 1024;      // + alignment slack
  static_assert(KCOLS % 32 == 0 && KCOLS <= 256 && KCOLS / 2 <= O_COL, "TMEM plan");
};

template <int WS>
__global__ void __launch_bounds__(256, 2)
window_attention_tc_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ qg, const float* __restrict__ table,
                           bf16* __restrict__ out, int H, int W, int C, int heads, int num_windows, float scale_log2e) {
  using Cfg = TcCfg<WS>;
  constexpr int N = Cfg::N, WP = Cfg::WP, WPT = Cfg::WPT, WKEYS = Cfg::WKEYS, KCOLS = Cfg::KCOLS, QPW = Cfg::QPW;
  constexpr int TAB = Cfg::TAB, NC = Cfg::NC, CA = Cfg::CA;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sQ = smem + Cfg::off_q;
  uint8_t* sK = smem + Cfg::off_k;
  uint8_t* sV = smem + Cfg::off_v;     // V^T: [KBLK][32 rows (d)][64 keys], 128-byte swizzled rows
  uint8_t* sVr = smem + Cfg::off_vr;   // V as loaded
  float* sT = reinterpret_cast<float*>(smem + Cfg::off_t);
  float* sRed = sT + TAB;              // [8] per-warp maxima of the table
  float* sMax = reinterpret_cast<float*>(smem + Cfg::off_x);  // [2][128] partial row maxima
  float* sSum = sMax + 256;                                    // [2][128] partial row sums
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + Cfg::off_bar);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rowt = tid & 127, part = tid >> 7;   // query row of the tile, which share of its key columns
  const int h = blockIdx.x % heads, wgrp = blockIdx.x / heads;
  const int nWw = W / WS, nWh = H / WS, nWimg = nWh * nWw;
  const bool global_q = qg != nullptr;
  const int ldq = (global_q ? 2 : 3) * C;
  const int koff = (global_q ? 0 : C) + h * HD, voff = koff + C;

  if (tid == 0) {
    mbar_init(bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(Cfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }

  // window j of this tile -> (image, window row, window column); a missing second window is masked out
  auto win_coords = [&](int j, int& b, int& wy, int& wx) -> bool {
    const int win = wgrp * WPT + j;
    const bool ok = win < num_windows;
    const int wc = ok ? win : num_windows - 1;
    b = wc / nWimg;
    const int rem = wc - b * nWimg;
    wy = rem / nWw;
    wx = rem - wy * nWw;
    return ok;
  };
  auto token_row = [&](int b, int wy, int wx, int tok) -> long long {
    return ((long long)b * H + wy * WS + tok / WS) * W + wx * WS + tok % WS;
  };

  // ---- stage everything this item needs with cp.async: K (rows = key columns of S, padded key order, swizzled rows),
  //      V as it lies in memory (transposed below), and the Q tiles
  for (int i = tid; i < KCOLS * 4; i += 256) {
    const int kc = i >> 2, ch = i & 3;           // key column, 16-byte chunk (8 channels)
    const int j = kc / WKEYS, kl = kc - j * WKEYS;
    const int ky = kl / WP, kx = kl % WP;
    int b, wy, wx;
    const bool wok = win_coords(j, b, wy, wx);
    const bool valid = wok && ky < WS && kx < WS;
    const long long row = token_row(b, wy, wx, valid ? ky * WS + kx : 0);
    cp_async16_zfill(sK + kc * 128 + ((ch ^ (kc & 7)) << 4), qkv + row * ldq + koff + ch * 8, valid);
    cp_async16_zfill(sVr + kc * 64 + ch * 16, qkv + row * ldq + voff + ch * 8, valid);
  }
  const int jq = rowt / QPW;                      // window of the tile this thread's query row belongs to
  int qb, qwy, qwx;
  const bool qwin_ok = win_coords(jq, qb, qwy, qwx);
  if (part < Cfg::MT) {                           // thread group `part` stages query tile `part`
    const int tok = part * 128 + (rowt % QPW);
    const bool ok = qwin_ok && tok < N;
    const int tokc = tok < N ? tok : N - 1;
    const long long row = token_row(qb, qwy, qwx, tokc);
    const bf16* src = global_q ? qg + ((long long)qb * N + tokc) * C + h * HD : qkv + row * ldq + h * HD;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
      cp_async16_zfill(sQ + part * 16384 + rowt * 128 + ((ch ^ (rowt & 7)) << 4), src + ch * 8, ok);
  }
  float tmax = -3.0e38f;
  for (int i = tid; i < TAB; i += 256) {
    const float v = __ldg(table + (long long)h * TAB + i) * 1.4426950408889634f;
    sT[i] = v;
    tmax = fmaxf(tmax, v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, o));
  if (lane == 0) sRed[warp] = tmax;
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  // ---- V^T: element (d, key) at block (key / 64), row d, 16-byte chunk ((key % 64) / 8) ^ (d & 7), slot key % 8
  for (int i = tid; i < KCOLS * 4; i += 256) {
    const int kc = i >> 2, ch = i & 3;
    const uint4 v = *reinterpret_cast<const uint4*>(sVr + kc * 64 + ch * 16);
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    uint8_t* vb = sV + (kc >> 6) * 4096 + (kc & 7) * 2;
    const int kchunk = (kc & 63) >> 3;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int d = ch * 8 + e;
      const unsigned short val = (unsigned short)((e & 1) ? (w[e >> 1] >> 16) : (w[e >> 1] & 0xffffu));
      *reinterpret_cast<unsigned short*>(vb + d * 128 + ((kchunk ^ (d & 7)) << 4)) = val;
    }
  }
  float tabmax = sRed[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) tabmax = fmaxf(tabmax, sRed[i]);
  uint32_t phase = 0;
  uint32_t tmem_base = 0;

#pragma unroll 1
  for (int mt = 0; mt < Cfg::MT; ++mt) {
    const int tok = mt * 128 + (rowt % QPW);
    const bool row_ok = qwin_ok && tok < N;
    const int tokc = tok < N ? tok : N - 1;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (mt == 0) tmem_base = *tmem_slot;
    // ---- S = Q K^T (K = 32: two 16-element steps inside the first 64 bytes of each 128-byte row)
    if (tid == 0) {
      constexpr uint32_t idesc_s = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KCOLS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t a_desc = make_sw128_desc(smem_u32(sQ + mt * 16384)), b_desc = make_sw128_desc(smem_u32(sK));
      umma_ss(tmem_base, a_desc, b_desc, idesc_s, 0u);
      umma_ss(tmem_base, a_desc + 2, b_desc + 2, idesc_s, 1u);
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();

    // ---- softmax: the two threads of a row take chunks [0, CA) and [CA, NC) of its window's key columns
    constexpr int RPC = 32 / WP;                    // window rows per 32-column chunk
    const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
    const int cbase = jq * WKEYS;                   // first S column of this row's window
    const int c_lo = part == 0 ? 0 : CA, c_hi = part == 0 ? CA : NC;
    float mx = -3.0e38f;
#pragma unroll
    for (int lc = 0; lc < CA; ++lc) {
      const int c = c_lo + lc;
      if (c < c_hi) {
        uint32_t r[32];
        tmem_ld32(lane_addr + (uint32_t)(cbase + c * 32), r);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i % WP < WS && (WS * WP % 32 == 0 || c * RPC + i / WP < WS)) mx = fmaxf(mx, __uint_as_float(r[i]));
      }
    }
    sMax[part * 128 + rowt] = mx;
    __syncthreads();
    mx = fmaxf(sMax[rowt], sMax[128 + rowt]);
    // upper bound of the row maximum of (scaled score + bias): exact softmax after normalisation, no overflow
    const float mrow = fmaf(fmaxf(mx, -1.0e30f), scale_log2e, tabmax);
    const float* pb = sT + (tokc / WS + WS - 1) * (2 * WS - 1) + tokc % WS + WS - 1;
    float lsum = 0.0f;
    uint32_t pk[CA][16];                            // this thread's share of the row of P, packed bf16 pairs
#pragma unroll
    for (int lc = 0; lc < CA; ++lc) {
      const int c = c_lo + lc;
      if (c < c_hi) {
        uint32_t r[32];
        tmem_ld32(lane_addr + (uint32_t)(cbase + c * 32), r);
        const float* pbc = pb - c * RPC * (2 * WS - 1);
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          float pv[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int kyl = (i + e) / WP, kx = (i + e) % WP;   // compile-time
            const bool real = kx < WS && (WS * WP % 32 == 0 || c * RPC + kyl < WS);
            if (real) {
              const float sc = fmaf(__uint_as_float(r[i + e]), scale_log2e, pbc[-(kyl * (2 * WS - 1) + kx)]);
              pv[e] = fast_exp2(sc - mrow);
            } else {
              pv[e] = 0.0f;
            }
            lsum += pv[e];
          }
          pk[lc][i >> 1] = pack_bf16(pv[0], pv[1]);
        }
      }
    }
    sSum[part * 128 + rowt] = lsum;
    // P overwrites S columns that the OTHER thread of the row may still be reading: everyone finishes reading S first
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
#pragma unroll
    for (int lc = 0; lc < CA; ++lc) {
      const int c = c_lo + lc;
      if (c < c_hi) tmem_st16(lane_addr + (uint32_t)((cbase >> 1) + c * 16), pk[lc]);
    }
    if (WPT == 2) {
      // the other window's key columns of this row: zero probabilities (this thread's share of them)
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
      const int obase = (1 - jq) * WKEYS;
#pragma unroll
      for (int lc = 0; lc < CA; ++lc) {
        const int c = c_lo + lc;
        if (c < c_hi) tmem_st16(lane_addr + (uint32_t)((obase >> 1) + c * 16), z);
      }
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // ---- O = P V: A = P in TMEM (8 columns per 16-key step), B = V^T blocks in shared memory
    if (tid == 0) {
      constexpr uint32_t idesc_o = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
#pragma unroll
      for (int k = 0; k < Cfg::KSTEPS; ++k) {
        const uint64_t b_desc = make_sw128_desc(smem_u32(sV + (k >> 2) * 4096)) + 2 * (k & 3);
        umma_ts(tmem_base + Cfg::O_COL, tmem_base + 8 * k, b_desc, idesc_o, k > 0 ? 1u : 0u);
      }
      umma_commit(bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    {
      // each thread of the row normalises and stores 16 of the 32 output channels
      uint32_t r[16];
      tmem_ld16(lane_addr + (uint32_t)(Cfg::O_COL + part * 16), r);
      if (row_ok) {
        const float inv = 1.0f / (sSum[rowt] + sSum[128 + rowt]);
        uint4* op = reinterpret_cast<uint4*>(out + token_row(qb, qwy, qwx, tok) * C + h * HD + part * 16);
#pragma unroll
        for (int q4 = 0; q4 < 2; ++q4) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            w[t] = pack_bf16(__uint_as_float(r[q4 * 8 + 2 * t]) * inv, __uint_as_float(r[q4 * 8 + 2 * t + 1]) * inv);
          op[q4] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tc_fence_before();
    __syncthreads();   // TMEM and the exchange buffers are reused by the next query tile
    tc_fence_after();
  }
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::kTmemCols) : "memory");
  }
}

template <int WS>
int launch_tc(const bf16* qkv, const bf16* qg, const float* table, bf16* out, int B, int H, int W, int C, int heads,
              cudaStream_t st) {
  using Cfg = TcCfg<WS>;
  auto kern = window_attention_tc_kernel<WS>;
  static bool configured = false;
  if (!configured) {
    VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  const int num_windows = B * (H / WS) * (W / WS);
  const int groups = (num_windows + Cfg::WPT - 1) / Cfg::WPT;
  kern<<<groups * heads, 256, Cfg::kSmem, st>>>(qkv, qg, table, out, H, W, C, heads, num_windows,
                                                1.4426950408889634f / sqrtf((float)HD));
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

}  // namespace

int window_attention_tc(const void* qkv, const void* qg, const float* table, void* out, int B, int H, int W, int C, int ws,
                        int heads, cudaStream_t st) {
  if (ws == 7) return launch_tc<7>((const bf16*)qkv, (const bf16*)qg, table, (bf16*)out, B, H, W, C, heads, st);
  if (ws == 14) return launch_tc<14>((const bf16*)qkv, (const bf16*)qg, table, (bf16*)out, B, H, W, C, heads, st);
  return VIP_ERR_UNSUPPORTED;
}

}  // namespace vip
