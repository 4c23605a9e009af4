// Fused pre-LN MLP of a GCViT block (models/gcvit/layers/block.py:39-56,77-81; layers/feature.py:8-43):
//     y = x + fc2(gelu(fc1(LayerNorm(x))))        (layer-scale gamma2 folded into fc2 by the caller)
// for the wide, shallow levels (C <= 128, hidden <= 256) where the two separate contractions are bound by writing and
// re-reading the [tokens, hidden] tensor (level 0 of GCViT-small: 3.2 M tokens x 192 = 1.2 GB each way per block).
// Here the hidden activations never leave the SM:
//
//   producer warp   TMA: both weight matrices once per CTA (they stay resident in shared memory: 84 KB for 96/192),
//                   then the 128-token x tiles through a ring of stages
//   MMA issuers     H = x W1^T           A = x tile, B = W1 (shared memory, K-major), D = H fp32 in TMEM
//   (one per        Y = gelu(...) W2^T   A = the activated hidden tile read back FROM TMEM (packed bf16 written by the
//    group)                              epilogue over the columns of H it has consumed), B = W2 (shared memory), D = Y in
//                                        TMEM over the upper columns of H
//   two epilogue    thread = token = TMEM lane.  Pass 1: folded LayerNorm (rstd (acc - mean colsum) + bias), GELU, bf16 ->
//   groups          TMEM.  Pass 2: + bias + residual (the x tile still in shared memory) -> bf16 -> global, and the
//                   (sum, sum of squares) of the output row for the LayerNorm folded into the next block's qkv.
// The groups alternate tiles, so one group's MMAs and TMEM round trips run underneath the other's arithmetic.
// HBM traffic: x in, y out.  TMEM: one 256-column region per group.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdio.h>

#include "common.cuh"
#include "stats.cuh"

namespace vip {
namespace {

using bf16 = __nv_bfloat16;

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
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
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
// K-major operand tile, 128-byte swizzled rows 128 B apart, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32_nowait(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&h);
}
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
// gelu_fast on a pair (packed fp32: half the issue slots)
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
  float2 x2 = fmul2(x, x);
  x2.x = fminf(x2.x, 64.0f);
  x2.y = fminf(x2.y, 64.0f);
  float2 p = ffma2(make_float2(-0.00035307545f, -0.00035307545f), x2, make_float2(0.037015257f, 0.037015257f));
  p = ffma2(p, x2, make_float2(0.79749725f, 0.79749725f));
  const float2 u = fmul2(p, x);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hx = fmul2(x, make_float2(0.5f, 0.5f));
  return ffma2(hx, t, hx);
}
// same fitted tanh form as the GEMM epilogue (gemm.cu: gelu_fast)
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);
  float p = fmaf(-0.00035307545f, x2, 0.037015257f);
  p = fmaf(p, x2, 0.79749725f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

#ifdef VIP_MLP_TRACE
// bring-up aid (VIP_NVCC_EXTRA=-DVIP_MLP_TRACE): cycles of epilogue group 0, thread 0 of CTA 0, summed over its tiles:
// [0] tiles, [1] wait for H, [2] pass 1, [3] wait for Y, [4] pass 2
__device__ long long g_mlp_trace[8];
#endif

template <int C, int HD>
struct MlpCfg {
  static constexpr int KB1 = (C + 63) / 64;        // 64-column k-blocks of x and W1
  static constexpr int KS1 = C / 16;               // K steps of the first product
  static constexpr int KB2 = (HD + 63) / 64;       // k-blocks of W2
  static constexpr int KS2 = HD / 16;              // K steps of the second product
  static constexpr int kXBytes = KB1 * 128 * 128;  // one x tile
  static constexpr int kW1Bytes = KB1 * HD * 128;
  static constexpr int kW2Bytes = KB2 * C * 128;
  static constexpr int kStages = 3;
  static constexpr int Y_COL = HD - C;             // Y over the top columns of H; the packed hidden tile is [0, HD / 2)
  static constexpr int kThreads = 11 * 32;
  static constexpr int kSmem = 1024 + kW1Bytes + kW2Bytes + kStages * kXBytes + HD * 8 + C * 4 + 256;
  static_assert(C % 32 == 0 && HD % 32 == 0 && HD <= 256 && C <= 128 && 2 * C <= HD, "shape");
  static_assert((HD * 128) % 1024 == 0 && (C * 128) % 1024 == 0, "weight tiles must keep the swizzle atoms aligned");
};

struct MlpArgs {
  const float* ln_stats;     // [M, 2] (mean, 1 / sigma) of the rows of x
  const bf16* x_lo;          // [M, C] low plane of x (two-plane residual stream) or null
  bf16* out_lo;              // [M, C] low plane of the output or null
  const float* colsum1;      // [HD]
  const float* bias1;        // [HD]
  const float* bias2;        // [C]
  bf16* out;                 // [M, C]
  float* ln_next;            // [M, 2] (mean, 1 / sigma) of the output rows for the next folded LayerNorm, or null
  long long M;
  float next_eps;            // epsilon of that next LayerNorm
};

template <int C, int HD>
__global__ void __launch_bounds__(MlpCfg<C, HD>::kThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const MlpArgs a) {
  using Cfg = MlpCfg<C, HD>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sW1 = smem;
  uint8_t* sW2 = sW1 + Cfg::kW1Bytes;
  uint8_t* sX = sW2 + Cfg::kW2Bytes;
  float2* sP1 = reinterpret_cast<float2*>(sX + kStages * Cfg::kXBytes);   // [HD] (colsum1, bias1)
  float* sB2 = reinterpret_cast<float*>(sP1 + HD);                         // [C]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB2 + C);
  uint64_t* full = bars;               // [kStages] x tile landed
  uint64_t* empty = full + kStages;    // [kStages] x tile no longer needed (its group has read the residual)
  uint64_t* w_full = empty + kStages;  // weights landed
  uint64_t* h_full = w_full + 1;       // [2] H of group g is in TMEM
  uint64_t* h2_full = h_full + 2;      // [2] group g has written the activated hidden tile
  uint64_t* y_full = h2_full + 2;      // [2] Y of group g is in TMEM
  uint64_t* y_free = y_full + 2;       // [2] group g has read Y
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(y_free + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int total_tiles = (int)((a.M + 127) / 128);
  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 128);
    }
    mbar_init(w_full, 1);
    for (int g = 0; g < 2; ++g) {
      mbar_init(h_full + g, 1);
      mbar_init(h2_full + g, 128);
      mbar_init(y_full + g, 1);
      mbar_init(y_free + g, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < HD; i += Cfg::kThreads) sP1[i] = make_float2(__ldg(a.colsum1 + i), __ldg(a.bias1 + i));
  for (int i = tid; i < C; i += Cfg::kThreads) sB2[i] = __ldg(a.bias2 + i);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 8) {
    // ---------------- producer ----------------
    if (lane == 0) {
      mbar_expect_tx(w_full, (uint32_t)(Cfg::kW1Bytes + Cfg::kW2Bytes));
#pragma unroll
      for (int kb = 0; kb < Cfg::KB1; ++kb) tma_load_2d(sW1 + kb * HD * 128, &tmW1, w_full, kb * 64, 0);
#pragma unroll
      for (int kb = 0; kb < Cfg::KB2; ++kb) tma_load_2d(sW2 + kb * C * 128, &tmW2, w_full, kb * 64, 0);
      int n = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++n) {
        const int s = n % kStages;
        const uint32_t ph = (uint32_t)(n / kStages) & 1u;
        mbar_wait(empty + s, ph ^ 1u);
        mbar_expect_tx(full + s, (uint32_t)Cfg::kXBytes);
#pragma unroll
        for (int kb = 0; kb < Cfg::KB1; ++kb) tma_load_2d(sX + s * Cfg::kXBytes + kb * 16384, &tmX, full + s, kb * 64, tile * 128);
      }
    }
  } else if (warp >= 9) {
    // ---------------- MMA issuer of group g ----------------
    const int g = warp - 9;
    if (lane == 0) {
      constexpr uint32_t idesc1 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t tg = tmem_base + (uint32_t)(g * 256);
      mbar_wait(w_full, 0u);
      uint32_t job = 0;
      int n = g;
      for (int tile = blockIdx.x + g * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, n += 2, ++job) {
        const int s = n % kStages;
        const uint32_t ph = (uint32_t)(n / kStages) & 1u;
        mbar_wait(full + s, ph);
        mbar_wait(y_free + g, (job & 1u) ^ 1u);   // the group has read the previous Y (which lies inside this H)
        tc_fence_after();
        const uint32_t xs = smem_u32(sX + s * Cfg::kXBytes), w1s = smem_u32(sW1), w2s = smem_u32(sW2);
#pragma unroll
        for (int k = 0; k < Cfg::KS1; ++k)
          umma_ss(tg, make_sw128_desc(xs + (k >> 2) * 16384) + 2 * (k & 3), make_sw128_desc(w1s + (k >> 2) * HD * 128) + 2 * (k & 3),
                  idesc1, k > 0 ? 1u : 0u);
        umma_commit(h_full + g);
        mbar_wait(h2_full + g, job & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < Cfg::KS2; ++k)
          umma_ts(tg + Cfg::Y_COL, tg + 8 * k, make_sw128_desc(w2s + (k >> 2) * C * 128) + 2 * (k & 3), idesc2, k > 0 ? 1u : 0u);
        umma_commit(y_full + g);
      }
    }
  } else {
    // ---------------- epilogue groups ----------------
    const int g = warp >> 2;
    const int row = tid & 127;
    const uint32_t tl = tmem_base + (uint32_t)(g * 256) + ((uint32_t)((warp & 3) * 32) << 16);
    const float invC = 1.0f / (float)C;
    uint32_t job = 0;
    int n = g;
    auto load_stats = [&](int tile_) -> float2 {
      const long long m_ = (long long)tile_ * 128 + row;
      if (tile_ >= total_tiles || m_ >= a.M) return make_float2(0.0f, 1.0f);
      return __ldg(reinterpret_cast<const float2*>(a.ln_stats) + m_);
    };
    float2 st_next = load_stats(blockIdx.x + g * gridDim.x);
    for (int tile = blockIdx.x + g * gridDim.x; tile < total_tiles; tile += 2 * gridDim.x, n += 2, ++job) {
      const int s = n % kStages;
      const long long m = (long long)tile * 128 + row;
      const bool row_ok = m < a.M;
      const float2 st = st_next;
      st_next = load_stats(tile + 2 * gridDim.x);   // in flight during this tile
      const float rstd = st.y;
      const float nmr = -st.x * st.y;
#ifdef VIP_MLP_TRACE
      const long long t0 = clock64();
#endif
      mbar_wait(h_full + g, job & 1u);
      tc_fence_after();
#ifdef VIP_MLP_TRACE
      const long long t1 = clock64();
#endif
      // ---- pass 1: hidden = gelu(rstd (acc - mean colsum) + bias) -> packed bf16 over the consumed columns of H
      {
        uint32_t r[2][32];
        tmem_ld32_nowait(tl, r[0]);
#pragma unroll
        for (int c = 0; c < HD / 32; ++c) {
          tmem_ld_wait();
          if (c + 1 < HD / 32) tmem_ld32_nowait(tl + (c + 1) * 32, r[(c + 1) & 1]);
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 p0 = sP1[c * 32 + i], p1 = sP1[c * 32 + i + 1];
            const float2 sh = ffma2(make_float2(nmr, nmr), make_float2(p0.x, p1.x), make_float2(p0.y, p1.y));
            const float2 v = gelu_fast2(ffma2(make_float2(__uint_as_float(r[c & 1][i]), __uint_as_float(r[c & 1][i + 1])),
                                              make_float2(rstd, rstd), sh));
            pk[i >> 1] = pack_bf16(v.x, v.y);
          }
          tmem_st16(tl + c * 16, pk);
        }
        tmem_st_wait();
      }
      tc_fence_before();
      mbar_arrive(h2_full + g);
#ifdef VIP_MLP_TRACE
      const long long t2 = clock64();
#endif
      // ---- pass 2: y = acc + bias + x -> bf16 -> global; statistics of the output row
      mbar_wait(y_full + g, job & 1u);
      tc_fence_after();
#ifdef VIP_MLP_TRACE
      const long long t3 = clock64();
#endif
      float rs_sum = 0.0f, rs_sq = 0.0f;
      const uint8_t* xrow = sX + s * Cfg::kXBytes + row * 128;
      bf16* orow = a.out + m * C;
      // low planes: 32-row x 8-column blocks (stats.cuh), this row's slot of column block 0
      const bf16* lrow = (a.x_lo != nullptr && row_ok) ? a.x_lo + lo_plane_index(m, 0, C) : nullptr;
      bf16* olrow = (a.out_lo != nullptr && row_ok) ? a.out_lo + lo_plane_index(m, 0, C) : nullptr;
      const float pivot = st.x;   // pivot of the output row's moments: the mean of the input row (stats.cuh)
      uint32_t ry[C / 32][32];
#pragma unroll
      for (int c = 0; c < C / 32; ++c) tmem_ld32_nowait(tl + Cfg::Y_COL + c * 32, ry[c]);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(y_free + g);      // Y is in registers: the next H may overwrite it
#pragma unroll
      for (int c = 0; c < C / 32; ++c) {
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          const int col = c * 32 + q8 * 8;            // first of 8 channels; k-block col / 64, 16-byte chunk (col % 64) / 8
          const uint4 xr = *reinterpret_cast<const uint4*>(xrow + (col >> 6) * 16384 + ((((col & 63) >> 3) ^ (row & 7)) << 4));
          const uint32_t xw[4] = {xr.x, xr.y, xr.z, xr.w};
          const float4 ba = *reinterpret_cast<const float4*>(sB2 + col), bb = *reinterpret_cast<const float4*>(sB2 + col + 4);
          const float bs[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
          uint32_t w[4], wl[4] = {0u, 0u, 0u, 0u};
          if (lrow != nullptr) {
            const uint4 ul = __ldg(reinterpret_cast<const uint4*>(lrow + (size_t)(col >> 3) * 256));
            wl[0] = ul.x; wl[1] = ul.y; wl[2] = ul.z; wl[3] = ul.w;
          }
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float v0 = __uint_as_float(ry[c][q8 * 8 + 2 * t]) + bs[2 * t] + (__uint_as_float(xw[t] << 16) + __uint_as_float(wl[t] << 16));
            const float v1 = __uint_as_float(ry[c][q8 * 8 + 2 * t + 1]) + bs[2 * t + 1] +
                             (__uint_as_float(xw[t] & 0xffff0000u) + __uint_as_float(wl[t] & 0xffff0000u));
            const float d0 = v0 - pivot, d1 = v1 - pivot;
            rs_sum += d0 + d1;
            rs_sq = fmaf(d0, d0, rs_sq);
            rs_sq = fmaf(d1, d1, rs_sq);
            w[t] = pack_bf16(v0, v1);
            wl[t] = pack_bf16(v0 - __uint_as_float(w[t] << 16), v1 - __uint_as_float(w[t] & 0xffff0000u));
          }
          if (row_ok) *reinterpret_cast<uint4*>(orow + col) = make_uint4(w[0], w[1], w[2], w[3]);
          if (olrow != nullptr) *reinterpret_cast<uint4*>(olrow + (size_t)(col >> 3) * 256) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
        }
      }
      mbar_arrive(empty + s);   // the residual has been read: the stage may be refilled
      if (a.ln_next != nullptr && row_ok) {
        // this thread owns the whole row: the moments of the next LayerNorm directly (pivoted one-pass form, stats.cuh)
        const float d = rs_sum * invC;
        const float var = fmaxf(rs_sq * invC - d * d, 0.0f);
        *reinterpret_cast<float2*>(a.ln_next + 2 * m) = make_float2(pivot + d, 1.0f / sqrtf(var + a.next_eps));
      }
#ifdef VIP_MLP_TRACE
      if (blockIdx.x == 0 && tid == 0) {
        const long long t4 = clock64();
        g_mlp_trace[0] += 1;
        g_mlp_trace[1] += t1 - t0;
        g_mlp_trace[2] += t2 - t1;
        g_mlp_trace[3] += t3 - t2;
        g_mlp_trace[4] += t4 - t3;
      }
#endif
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// [rows, cols] bf16 row-major (ld elements between rows), box = [box_rows, 64 cols], 128-byte swizzle, zero fill outside
int make_tmap(CUtensorMap* tm, const void* base, long long rows, int cols, int ld, int box_rows) {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA, "cuTensorMapEncodeTiled (fused MLP) failed with CUresult %d (rows=%lld cols=%d ld=%d)",
              (int)r, rows, cols, ld);
  return VIP_OK;
}

template <int C, int HD>
int launch_mlp(const bf16* x, long long M, const bf16* w1, int ldw1, const bf16* w2, int ldw2, const MlpArgs& a, cudaStream_t st) {
  using Cfg = MlpCfg<C, HD>;
  auto kern = mlp_fused_kernel<C, HD>;
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    VIP_CUDA(cudaGetDevice(&dev));
    VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    VIP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  CUtensorMap tmX, tmW1, tmW2;
  int rc = make_tmap(&tmX, x, M, C, C, 128);
  if (rc != VIP_OK) return rc;
  rc = make_tmap(&tmW1, w1, HD, C, ldw1, HD);
  if (rc != VIP_OK) return rc;
  rc = make_tmap(&tmW2, w2, C, HD, ldw2, C);
  if (rc != VIP_OK) return rc;
  const int tiles = (int)((M + 127) / 128);
  const int grid = tiles < sms ? tiles : sms;
  VIP_LAUNCH(kern, grid, Cfg::kThreads, Cfg::kSmem, st, tmX, tmW1, tmW2, a);
  VIP_CUDA(cudaGetLastError());
  count_launch();
#ifdef VIP_MLP_TRACE
  {
    long long t[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(t, g_mlp_trace, sizeof(t));
    if (t[0] > 0)
      fprintf(stderr, "[mlp trace %d/%d] tiles %lld  wait H %lld  pass 1 %lld  wait Y %lld  pass 2 %lld cycles/tile\n", C, HD, t[0],
              t[1] / t[0], t[2] / t[0], t[3] / t[0], t[4] / t[0]);
    long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_mlp_trace, z, sizeof(z));
  }
#endif
  return VIP_OK;
}

}  // namespace
}  // namespace vip

extern "C" int vip_mlp_fused_bf16(const void* x, const void* x_lo, long long M, int C, int hidden, const float* ln_stats,
                                  float next_eps, const void* w1, int ldw1, const float* colsum1, const float* bias1,
                                  const void* w2, int ldw2, const float* bias2, void* out, void* out_lo, float* ln_next,
                                  void* stream) {
  using namespace vip;
  VIP_REQUIRE(x && ln_stats && w1 && colsum1 && bias1 && w2 && bias2 && out, VIP_ERR_INVALID, "vip_mlp_fused_bf16: null pointer");
  VIP_REQUIRE(M > 0 && M < (1LL << 37), VIP_ERR_INVALID, "vip_mlp_fused_bf16: bad M");
  VIP_REQUIRE(ldw1 % 8 == 0 && ldw2 % 8 == 0 && ldw1 >= C && ldw2 >= hidden, VIP_ERR_INVALID,
              "vip_mlp_fused_bf16: weight leading dimensions (ldw1=%d ldw2=%d)", ldw1, ldw2);
  MlpArgs a;
  a.ln_stats = ln_stats;
  a.x_lo = (const bf16*)x_lo;
  a.out_lo = (bf16*)out_lo;
  a.colsum1 = colsum1;
  a.bias1 = bias1;
  a.bias2 = bias2;
  a.out = (bf16*)out;
  a.ln_next = ln_next;
  a.M = M;
  a.next_eps = next_eps;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (C == 96 && hidden == 192) return launch_mlp<96, 192>((const bf16*)x, M, (const bf16*)w1, ldw1, (const bf16*)w2, ldw2, a, st);
  if (C == 64 && hidden == 192) return launch_mlp<64, 192>((const bf16*)x, M, (const bf16*)w1, ldw1, (const bf16*)w2, ldw2, a, st);
  set_error("vip_mlp_fused_bf16: (C=%d, hidden=%d) is not built; use two contractions", C, hidden);
  return VIP_ERR_UNSUPPORTED;
}
