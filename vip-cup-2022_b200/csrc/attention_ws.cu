// Window attention of GCViT (models/gcvit/layers/attention.py:52-83, window.py:3-14 folded into the addressing) as a
// persistent, warp-specialised tcgen05 kernel -- the default implementation.  attention.cu (warp-level mma.sync) is
// bound by the legacy HMMA rate of sm_100, a first one-item-per-CTA tcgen05 version (removed) by latency; here
// one CTA per SM walks a list of items and overlaps the phases of different items:
//
//   item      one 128-row query tile group x one PAIR of heads (64 channels = one 128-byte swizzled row):
//               ws 7 : two windows (rows 0..48 and 64..112 of the tile, key columns 0..48 and 64..112)
//               ws 14: one window, 196 queries = two 128-row tiles, 196 keys (208 columns)
//   warp 8    producer: TMA box loads {64 channels, ws, ws} of q, k, v straight from the image-order [B, H, W, 3C] tensor
//             (a window IS a box of that tensor) into a ring of shared-memory stages; tokens land densely in key order
//   warp 9/10 MMA issuers, one per softmax group: S = Q K^T (both operands K-major, the head of the pair picked by a
//             64-byte offset inside the swizzle atom), then O = P V with P read from TMEM and V used as it lies in
//             memory (MN-major B operand: no transposition)
//   warps 0-7 two softmax groups of 128 threads (thread = query row = TMEM lane), group g owns head 2 * hp + g of the
//             item: row maximum, exp2, P written back to TMEM as packed bf16, later O / row sum -> bf16 -> global.
//             While one group exponentiates, the other group's MMAs, TMEM loads and stores are in flight.
// TMEM: 256 columns per group (S, P, O); 512 allocated, one CTA per SM.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>

#include "common.cuh"

#ifndef VIP_ATTN_WS7_STAGES
#define VIP_ATTN_WS7_STAGES 4   // measured at 56x56, batch 1024: 3 stages 317 us, 4 stages 303 us
#endif
#ifndef VIP_ATTN_POLY
#define VIP_ATTN_POLY 1   // measured on B200 (ws 14, batch 1024): 0 -> 218 us, 1 -> 192, 2 -> 199, 3 -> 220
#endif

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
// 128-byte swizzled rows 128 B apart, 8-row groups 1024 B apart.  K-major operands: rows = M/N index, the K slice is
// chosen by a 32-byte step of the start address; MN-major operand (V): rows = K index (keys), the 32 channels of a head
// by a 64-byte step.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
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
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

#ifdef VIP_ATTN_TRACE
// bring-up aid (build with VIP_NVCC_EXTRA=-DVIP_ATTN_TRACE): cycle counts of the MMA issuer of group 0 in CTA 0, summed over
// its jobs: [0] jobs, [1] QK issue -> S complete, [2] S complete -> P arrived, [3] P arrived -> O complete
__device__ long long g_attn_trace[8];
#endif

// packed fp32 pairs (FFMA2 / FADD2 of sm_100): half the issue slots of the softmax arithmetic
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rc, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nmov.b64 rc, {%6,%7};\n"
      "fma.rn.f32x2 rd, ra, rb, rc;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) {
  float2 d;
  asm("{\n.reg .b64 ra, rb, rd;\nmov.b64 ra, {%2,%3};\nmov.b64 rb, {%4,%5};\nadd.rn.f32x2 rd, ra, rb;\nmov.b64 {%0,%1}, rd;\n}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}

// exp2 of a pair of non-positive arguments on the FMA / ALU pipes (no MUFU): x = n + f with n = round(x), 2^f by a
// degree-3 polynomial on [-0.5, 0.5] (max relative error 7.5e-5, well under the bf16 rounding of P), 2^n by adding n to
// the exponent field.  The softmax is bound by MUFU.EX2 issue (one warp instruction per 8 cycles and scheduler), so a
// share of the exponentials is moved here.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);   // 1.5 * 2^23: the low mantissa bits of x + magic hold n
  const float2 t = fadd2(x, magic);
  const float2 nf = fadd2(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = fadd2(x, make_float2(-nf.x, -nf.y));
  float2 p = ffma2(make_float2(0.05517144873738289f, 0.05517144873738289f), f, make_float2(0.2426108419895172f, 0.2426108419895172f));
  p = ffma2(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = ffma2(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

template <int WS>
struct WsCfg {
  static constexpr int N = WS * WS;                       // tokens of a window (49 / 196)
  static constexpr int WPT = WS <= 8 ? 2 : 1;             // windows per 128-row tile
  static constexpr int MT = WS <= 8 ? 1 : 2;              // 128-row query tiles per item
  static constexpr int QPW = 128 / WPT;                   // tile rows reserved per window
  static constexpr int WKEYS = WS <= 8 ? 64 : 208;        // S columns reserved per window
  static constexpr int KEYS = WPT * WKEYS;                // UMMA N of S (128 / 208)
  static constexpr int NCH = (N + 31) / 32;               // 32-column chunks of S one row reads (2 / 7)
  static constexpr int KSTEPS = KEYS / 16;                // K steps of P V (8 / 13)
  static constexpr int TAB = (2 * WS - 1) * (2 * WS - 1);
  static constexpr int kQBytes = MT * 128 * 128, kKBytes = KEYS * 128;
  static constexpr int kStageBytes = kQBytes + 2 * kKBytes;   // q, k, v (multiples of 1024)
  static constexpr int kStages = WS <= 8 ? VIP_ATTN_WS7_STAGES : 2;
  static constexpr int kBoxBytes = N * 128;
  // TMEM columns of one softmax group
  static constexpr int S_COL = 0;
  static constexpr int P_COL = WS <= 8 ? 128 : 0;         // ws 14: P overwrites the columns of S already consumed
  static constexpr int O_COL = WS <= 8 ? 192 : 208;       // ws 7: two O buffers (192, 224), read out one job late
  // P reaches the MMA issuer in parts (ws 14: after chunks 1, 3, 5 and 6 = K steps 0-3, 4-7, 8-11, 12), so that most of
  // P V runs underneath the exponentials of the later chunks
  static constexpr int NPART = WS <= 8 ? 1 : 4;
  static constexpr int kPolyOf4 = VIP_ATTN_POLY;          // of every 4 pairs of exponentials, how many avoid the MUFU pipe
  static constexpr int kThreads = 11 * 32;
  static_assert(kStageBytes % 1024 == 0 && O_COL + (WS <= 8 ? 64 : 32) <= 256, "layout");
  static int smem_bytes(int heads) {
    return 1024 + kStages * kStageBytes + ((heads * TAB * 4 + 15) / 16) * 16 + ((heads * 4 + 15) / 16) * 16 + 256;
  }
};

template <int WS>
__global__ void __launch_bounds__(WsCfg<WS>::kThreads, 1)
window_attention_ws_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmQG,
                           const float* __restrict__ table, bf16* __restrict__ out, int H, int W, int C, int heads,
                           int num_windows, int global_q, float scale_log2e) {
  using Cfg = WsCfg<WS>;
  constexpr int N = Cfg::N, WPT = Cfg::WPT, MT = Cfg::MT, QPW = Cfg::QPW, WKEYS = Cfg::WKEYS, KEYS = Cfg::KEYS;
  constexpr int TAB = Cfg::TAB, kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // (pointer arithmetic on the array, not an integer round trip: the compiler keeps the shared address space -> LDS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* stage0 = smem;
  float* sT = reinterpret_cast<float*>(smem + kStages * Cfg::kStageBytes);      // [heads][TAB], times log2 e
  float* sTmax = sT + ((heads * TAB + 3) / 4) * 4;                               // [heads]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(sTmax) + ((heads * 4 + 15) / 16) * 16);
  uint64_t* full = bars;                 // [kStages] TMA -> MMA warps
  uint64_t* empty = full + kStages;      // [kStages] both MMA warps -> producer
  uint64_t* s_full = empty + kStages;    // [2] S of group g is in TMEM
  uint64_t* p_full = s_full + 2;         // [2][4] group g has written part k of P (the last part: and is done with S)
  uint64_t* o_full = p_full + 8;         // [2][2] O buffer b of group g is in TMEM (ws 14 uses buffer 0 only)
  uint64_t* o_free = o_full + 4;         // [2][2] group g has read O buffer b
  uint64_t* s_free = o_free + 4;         // [2] group g holds its S row in registers (ws 7): the next Q K^T may overwrite S
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);  // (barrier block: 2 * kStages + 22 words)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nWw = W / WS, nWimg = (H / WS) * nWw;
  const int num_hp = (heads + 1) >> 1;
  const int num_tiles = (num_windows + WPT - 1) / WPT;
  const int total_items = num_tiles * num_hp;
  const int ldq = (global_q ? 2 : 3) * C;

  pdl_trigger();
  if (tid == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 2);
    }
    for (int g = 0; g < 2; ++g) {
      mbar_init(s_full + g, 1);
      for (int k = 0; k < 4; ++k) mbar_init(p_full + g * 4 + k, 128);
      for (int b = 0; b < 2; ++b) {
        mbar_init(o_full + g * 2 + b, 1);
        mbar_init(o_free + g * 2 + b, 128);
      }
      mbar_init(s_free + g, 128);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 8) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // the padding rows of the stages are never written by the TMA boxes: zero them once (0 x stale NaN would poison P V)
  for (int i = tid; i < kStages * Cfg::kStageBytes / 16; i += Cfg::kThreads)
    reinterpret_cast<uint4*>(stage0)[i] = make_uint4(0, 0, 0, 0);
  for (int i = tid; i < heads * TAB; i += Cfg::kThreads) sT[i] = __ldg(table + i) * 1.4426950408889634f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid < heads) {
    float m = -3.0e38f;
    for (int i = 0; i < TAB; ++i) m = fmaxf(m, sT[tid * TAB + i]);
    sTmax[tid] = m;
  }
  __syncthreads();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // set-up (barriers, TMEM, zeroed stages, bias tables = weights) overlapped the preceding kernel's tail

  if (warp == 8) {
    // ---------------- producer ----------------
    if (lane == 0) {
      int n = 0;
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const int s = n % kStages;
        const uint32_t ph = (uint32_t)(n / kStages) & 1u;
        mbar_wait(empty + s, ph ^ 1u);
        const int tile = item / num_hp, hp = item - tile * num_hp;
        uint8_t* sQ = stage0 + s * Cfg::kStageBytes;
        uint8_t* sK = sQ + Cfg::kQBytes;
        uint8_t* sV = sK + Cfg::kKBytes;
        const int nwin = min(WPT, num_windows - tile * WPT);
        mbar_expect_tx(full + s, (uint32_t)(nwin * 3 * Cfg::kBoxBytes));
        for (int j = 0; j < nwin; ++j) {
          const int win = tile * WPT + j;
          const int b = win / nWimg, rem = win - b * nWimg;
          const int wy = rem / nWw, wx = rem - wy * nWw;
          const int x0 = wx * WS, y0 = b * H + wy * WS;
          if (global_q) tma_load_3d(sQ + j * QPW * 128, &tmQG, full + s, hp * 64, 0, b * WS);
          else tma_load_3d(sQ + j * QPW * 128, &tmQKV, full + s, hp * 64, x0, y0);
          const int koff = (global_q ? 0 : C) + hp * 64;
          tma_load_3d(sK + j * WKEYS * 128, &tmQKV, full + s, koff, x0, y0);
          tma_load_3d(sV + j * WKEYS * 128, &tmQKV, full + s, koff + C, x0, y0);
        }
      }
    }
  } else if (warp >= 9) {
    // ---------------- MMA issuer of softmax group g ----------------
    const int g = warp - 9;
    if (lane == 0) {
      constexpr uint32_t idesc_s =
          (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(KEYS >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      constexpr uint32_t idesc_o =  // B (= V) MN-major
          (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(HD >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t tg = tmem_base + (uint32_t)(g * 256);
      int n = 0;
      uint32_t job = 0;
      if constexpr (WS <= 8) {
        // ws 7: the group copies its 49 scores into registers at once and releases S, so Q K^T of job j + 1 is issued
        // BEFORE P V of job j and runs underneath the softmax of job j; O alternates between two TMEM buffers and is
        // read out one job late.  Neither MMA round trip is on the group's critical path any more.
        bool have_prev = false;
        int prev_s = 0;
        uint64_t prev_vdesc = 0;
        auto issue_pv = [&](uint32_t j) {      // P V of job j into O buffer j & 1
          const uint32_t b = j & 1u, u = j >> 1;
          mbar_wait(p_full + g * 4, j & 1u);
          mbar_wait(o_free + g * 2 + b, (u & 1u) ^ 1u);   // the previous use of this buffer has been read out
          tc_fence_after();
#pragma unroll
          for (int k = 0; k < Cfg::KSTEPS; ++k)
            umma_ts(tg + Cfg::O_COL + 32 * b, tg + Cfg::P_COL + 8 * k, prev_vdesc + (uint64_t)(k * 128), idesc_o, k > 0 ? 1u : 0u);
          umma_commit(o_full + g * 2 + b);
          umma_commit(empty + prev_s);
        };
        for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
          const int s = n % kStages;
          const uint32_t ph = (uint32_t)(n / kStages) & 1u;
          const int tile = item / num_hp, hp = item - tile * num_hp;
          mbar_wait(full + s, ph);
          if (hp * 2 + g >= heads) {
            mbar_arrive(empty + s);
            continue;
          }
          mbar_wait(s_free + g, (job & 1u) ^ 1u);         // S of the previous job is in the group's registers
          tc_fence_after();
          const uint32_t sQ = smem_u32(stage0 + s * Cfg::kStageBytes);
          const uint32_t sK = sQ + Cfg::kQBytes, sV = sK + Cfg::kKBytes;
          const uint64_t qdesc = make_sw128_desc(sQ) + 4 * g, kdesc = make_sw128_desc(sK) + 4 * g;
          umma_ss(tg + Cfg::S_COL, qdesc, kdesc, idesc_s, 0u);
          umma_ss(tg + Cfg::S_COL, qdesc + 2, kdesc + 2, idesc_s, 1u);
          umma_commit(s_full + g);
          if (have_prev) issue_pv(job - 1u);
          have_prev = true;
          prev_s = s;
          prev_vdesc = make_sw128_desc(sV) + 4 * g;
          ++job;
        }
        if (have_prev) issue_pv(job - 1u);
      } else
      for (int item = blockIdx.x; item < total_items; item += gridDim.x, ++n) {
        const int s = n % kStages;
        const uint32_t ph = (uint32_t)(n / kStages) & 1u;
        const int tile = item / num_hp, hp = item - tile * num_hp;
        const bool active = hp * 2 + g < heads;
        if (!active) {  // odd number of heads: nothing to do for this half of the last pair
          mbar_wait(full + s, ph);
          mbar_arrive(empty + s);
          continue;
        }
        mbar_wait(full + s, ph);
        tc_fence_after();
        const uint32_t sQ = smem_u32(stage0 + s * Cfg::kStageBytes);
        const uint32_t sK = sQ + Cfg::kQBytes, sV = sK + Cfg::kKBytes;
        const uint64_t kdesc = make_sw128_desc(sK) + 4 * g, vdesc = make_sw128_desc(sV) + 4 * g;
#pragma unroll 1
        for (int mt = 0; mt < MT; ++mt, ++job) {
          const uint64_t qdesc = make_sw128_desc(sQ + mt * 16384) + 4 * g;
          umma_ss(tg + Cfg::S_COL, qdesc, kdesc, idesc_s, 0u);
          umma_ss(tg + Cfg::S_COL, qdesc + 2, kdesc + 2, idesc_s, 1u);
          umma_commit(s_full + g);
#ifdef VIP_ATTN_TRACE
          const long long t0 = clock64();
          mbar_wait(s_full + g, job & 1u);
          const long long t1 = clock64();
#endif
          mbar_wait(o_free + g * 2, (job & 1u) ^ 1u);   // O of the previous job has been read
#pragma unroll
          for (int part = 0; part < Cfg::NPART; ++part) {
            mbar_wait(p_full + g * 4 + part, job & 1u);
            tc_fence_after();
            constexpr int SPP = Cfg::NPART == 1 ? Cfg::KSTEPS : 4;   // K steps per part
#pragma unroll
            for (int k = part * SPP; k < (part + 1) * SPP && k < Cfg::KSTEPS; ++k)
              umma_ts(tg + Cfg::O_COL, tg + Cfg::P_COL + 8 * k, vdesc + (uint64_t)(k * 128), idesc_o, k > 0 ? 1u : 0u);
          }
#ifdef VIP_ATTN_TRACE
          const long long t2 = clock64();
#endif
          umma_commit(o_full + g * 2);
          if (mt == MT - 1) umma_commit(empty + s);
#ifdef VIP_ATTN_TRACE
          mbar_wait(o_full + g * 2, job & 1u);
          const long long t3 = clock64();
          if (blockIdx.x == 0 && g == 0) {
            g_attn_trace[0] += 1;
            g_attn_trace[1] += t1 - t0;
            g_attn_trace[2] += t2 - t1;
            g_attn_trace[3] += t3 - t2;
          }
#endif
        }
      }
    }
  } else {
    // ---------------- softmax groups ----------------
    const int g = warp >> 2;
    const int row = tid & 127;
    const int jq = row / QPW;                       // window of the tile this row belongs to
    const uint32_t tl = tmem_base + (uint32_t)(g * 256) + ((uint32_t)((warp & 3) * 32) << 16);
    if (WPT == 2) {  // the P columns of the OTHER window of the tile stay zero for the whole kernel
      uint32_t z[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) z[i] = 0u;
      tmem_st16(tl + Cfg::P_COL + (1 - jq) * 32, z);
      tmem_st16(tl + Cfg::P_COL + (1 - jq) * 32 + 16, z);
      tmem_st_wait();
    }
    uint32_t job = 0;
    // ws 7: O is read out one job late from one of two TMEM buffers (see the MMA issuer)
    bool pend = false;
    float pend_inv = 0.0f;
    uint4* pend_dst = nullptr;               // null: the row does not exist
    auto wait_prev_pv = [&]() {              // P V of job - 1 has completed: P may be overwritten, O buffer is valid
      const uint32_t j = job - 1u;
      mbar_wait(o_full + g * 2 + (j & 1u), (j >> 1) & 1u);
      tc_fence_after();
    };
    auto read_out_prev = [&]() {             // O of job - 1 -> registers -> / row sum -> bf16 -> global
      const uint32_t j = job - 1u;
      uint32_t ro[32];
      tmem_ld32_nowait(tl + Cfg::O_COL + 32 * (j & 1u), ro);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free + g * 2 + (j & 1u));
      if (pend_dst != nullptr) {
#pragma unroll
        for (int q4 = 0; q4 < 4; ++q4) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t)
            w[t] = pack_bf16(__uint_as_float(ro[q4 * 8 + 2 * t]) * pend_inv, __uint_as_float(ro[q4 * 8 + 2 * t + 1]) * pend_inv);
          pend_dst[q4] = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    };
    for (int item = blockIdx.x; item < total_items; item += gridDim.x) {
      const int tile = item / num_hp, hp = item - tile * num_hp;
      const int h = hp * 2 + g;
      if (h >= heads) continue;
      const int win = tile * WPT + jq;
      const bool win_ok = win < num_windows;
      const int wc = win_ok ? win : num_windows - 1;
      const int b = wc / nWimg, rem = wc - b * nWimg;
      const int wy = rem / nWw, wx = rem - wy * nWw;
      const float* tab = sT + h * TAB;
      const float tabmax = sTmax[h];
#pragma unroll 1
      for (int mt = 0; mt < MT; ++mt, ++job) {
        const int tok = mt * 128 + (row % QPW);
        const bool row_ok = win_ok && tok < N;
        const int tokc = tok < N ? tok : N - 1;
        // a warp none of whose rows exist only keeps the barrier protocol going
        const bool warp_live = mt * 128 + ((row & ~31) % QPW) < N;
        mbar_wait(s_full + g, job & 1u);
        tc_fence_after();
        float lsum = 1.0f;
        if (warp_live) {
          const uint32_t srow = tl + Cfg::S_COL + jq * WKEYS;
          const float* pb = tab + (tokc / WS + WS - 1) * (2 * WS - 1) + tokc % WS + WS - 1;
          const uint32_t prow = tl + Cfg::P_COL + jq * (WKEYS / 2);
          uint32_t r[2][32];
          float mx = -3.0e38f, mrow;
          float2 ls[2] = {make_float2(0.0f, 0.0f), make_float2(0.0f, 0.0f)};
          // P chunk c = exp2(score * scale + bias - mrow) of the 32 raw scores in `rc`, packed bf16 pairs back into TMEM
          auto emit = [&](const int c, const uint32_t (&rc)[32]) {
            uint32_t pk[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              const int j = c * 32 + i;              // keys j, j + 1 = (j / WS, j % WS), compile-time after unrolling
              float2 pv = make_float2(0.0f, 0.0f);
              if (j + 1 < N) {
                const float2 bias = make_float2(pb[-((j / WS) * (2 * WS - 1) + j % WS)],
                                                pb[-(((j + 1) / WS) * (2 * WS - 1) + (j + 1) % WS)]);
                float2 sc = ffma2(make_float2(__uint_as_float(rc[i]), __uint_as_float(rc[i + 1])),
                                  make_float2(scale_log2e, scale_log2e), bias);
                sc = fadd2(sc, make_float2(-mrow, -mrow));
                if (((i >> 1) & 3) < Cfg::kPolyOf4) pv = exp2_poly2(sc);
                else pv = make_float2(fast_exp2(sc.x), fast_exp2(sc.y));
                ls[(i >> 1) & 1] = fadd2(ls[(i >> 1) & 1], pv);   // four independent partial sums
              } else if (j < N) {
                const float sc = fmaf(__uint_as_float(rc[i]), scale_log2e, pb[-((j / WS) * (2 * WS - 1) + j % WS)]);
                pv.x = fast_exp2(sc - mrow);
                ls[0].x += pv.x;
              }
              pk[i >> 1] = pack_bf16(pv.x, pv.y);
            }
            if (Cfg::NPART > 1 && c > 0 && (c & 1) == 0) {
              // the stores of the chunks before this one have long landed: hand that part of P to the MMA issuer
              tmem_st_wait();
              tc_fence_before();
              mbar_arrive(p_full + g * 4 + (c >> 1) - 1);
            }
            // only the columns that exist in P (ws 14: 104, the last chunk is half a chunk)
            if (c * 32 + 32 <= KEYS || WPT == 2) {
              tmem_st16(prow + c * 16, pk);
            } else {
              asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(prow + c * 16),
                           "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                           : "memory");
            }
          };
          if (Cfg::NCH == 2) {
            // the whole row fits in registers: one trip to TMEM serves the maximum and the exponentials
            tmem_ld32_nowait(srow, r[0]);
            tmem_ld32_nowait(srow + 32, r[1]);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(s_free + g);           // the next Q K^T may overwrite S
#pragma unroll
            for (int i = 0; i < 64; ++i)
              if (i < N) mx = fmaxf(mx, __uint_as_float(r[i >> 5][i & 31]));
            // upper bound of the maximum of (scaled score + bias): softmax is exact after normalisation, no overflow
            mrow = fmaf(mx, scale_log2e, tabmax);
            if (pend) wait_prev_pv();          // (long done) before P is overwritten
            emit(0, r[0]);
            emit(1, r[1]);
          } else {
            // two passes over the row, the TMEM load of chunk c + 1 in flight while chunk c is worked on
            tmem_ld32_nowait(srow, r[0]);
#pragma unroll
            for (int c = 0; c < Cfg::NCH; ++c) {
              tmem_ld_wait();
              if (c + 1 < Cfg::NCH) tmem_ld32_nowait(srow + (c + 1) * 32, r[(c + 1) & 1]);
              else tmem_ld32_nowait(srow, r[(c + 1) & 1]);   // first chunk of the second pass
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (c * 32 + i < N) mx = fmaxf(mx, __uint_as_float(r[c & 1][i]));
            }
            mrow = fmaf(mx, scale_log2e, tabmax);
            constexpr int B0 = Cfg::NCH & 1;   // buffer that holds chunk 0 of the second pass
#pragma unroll
            for (int c = 0; c < Cfg::NCH; ++c) {
              tmem_ld_wait();
              if (c + 1 < Cfg::NCH) tmem_ld32_nowait(srow + (c + 1) * 32, r[(B0 + c + 1) & 1]);
              emit(c, r[(B0 + c) & 1]);
            }
          }
          lsum = (ls[0].x + ls[0].y) + (ls[1].x + ls[1].y);
          tmem_st_wait();
        } else {
#pragma unroll
          for (int k = 0; k + 1 < Cfg::NPART; ++k) mbar_arrive(p_full + g * 4 + k);
        }
        tc_fence_before();
        mbar_arrive(p_full + g * 4 + Cfg::NPART - 1);
        if constexpr (WS <= 8) {
          if (pend) read_out_prev();
          pend = true;
          pend_inv = 1.0f / lsum;
          pend_dst = row_ok ? reinterpret_cast<uint4*>(out + (((long long)b * H + wy * WS + tok / WS) * W + wx * WS + tok % WS) * C + h * HD)
                            : nullptr;
          continue;
        }
        // ---- O / row sum -> bf16 -> global
        mbar_wait(o_full + g * 2, job & 1u);
        tc_fence_after();
        uint32_t r[32];
        if (warp_live) {
          tmem_ld32_nowait(tl + Cfg::O_COL, r);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(o_free + g * 2);
        if (row_ok) {
          // 64 contiguous bytes per thread (two full sectors).  Staging the rows through shared memory for wider
          // coalescing was measured slower: the address arithmetic per staged row costs more than the stores save.
          const float inv = 1.0f / lsum;
          const long long grow = ((long long)b * H + wy * WS + tok / WS) * W + wx * WS + tok % WS;
          uint4* op = reinterpret_cast<uint4*>(out + grow * C + h * HD);
#pragma unroll
          for (int q4 = 0; q4 < 4; ++q4) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t)
              w[t] = pack_bf16(__uint_as_float(r[q4 * 8 + 2 * t]) * inv, __uint_as_float(r[q4 * 8 + 2 * t + 1]) * inv);
            op[q4] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
    }
    if (pend) {   // ws 7: drain the last job
      wait_prev_pv();
      read_out_prev();
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

// [rows_y, rows_x, ch] bf16 tensor, box = {64 channels, ws, ws}, 128-byte swizzle, zero fill outside
int make_window_tmap(CUtensorMap* tm, const void* base, int ch, int nx, long long ny, int ws) {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return reinterpret_cast<EncodeTiledFn>(p);
  }();
  VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[3] = {(cuuint64_t)ch, (cuuint64_t)nx, (cuuint64_t)ny};
  const cuuint64_t gstride[2] = {(cuuint64_t)ch * 2, (cuuint64_t)nx * ch * 2};
  const cuuint32_t box[3] = {64, (cuuint32_t)ws, (cuuint32_t)ws};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA, "cuTensorMapEncodeTiled (window box) failed with CUresult %d (ch=%d nx=%d ny=%lld ws=%d)",
              (int)r, ch, nx, ny, ws);
  return VIP_OK;
}

template <int WS>
int launch_ws(const bf16* qkv, const bf16* qg, const float* table, bf16* out, int B, int H, int W, int C, int heads,
              cudaStream_t st) {
  using Cfg = WsCfg<WS>;
  auto kern = window_attention_ws_kernel<WS>;
  static int sms = 0, smem_max = 0;
  if (sms == 0) {
    int dev = 0;
    VIP_CUDA(cudaGetDevice(&dev));
    VIP_CUDA(cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    VIP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  }
  const int smem = Cfg::smem_bytes(heads);
  if (smem > smem_max) {   // the caller falls back to the mma.sync kernel; the reason stays readable in vip_last_error()
    set_error("window_attention_ws: %d heads x %d bias entries do not fit in shared memory (%d > %d bytes): mma.sync kernel used",
              heads, Cfg::TAB, smem, smem_max);
    return VIP_ERR_UNSUPPORTED;
  }
  const int ldq = (qg ? 2 : 3) * C;
  CUtensorMap tmQKV, tmQG;
  int rc = make_window_tmap(&tmQKV, qkv, ldq, W, (long long)B * H, WS);
  if (rc != VIP_OK) return rc;
  rc = qg ? make_window_tmap(&tmQG, qg, C, WS, (long long)B * WS, WS) : make_window_tmap(&tmQG, qkv, ldq, W, (long long)B * H, WS);
  if (rc != VIP_OK) return rc;
  const int num_windows = B * (H / WS) * (W / WS);
  const int items = ((num_windows + Cfg::WPT - 1) / Cfg::WPT) * ((heads + 1) / 2);
  const int grid = items < sms ? items : sms;
  VIP_LAUNCH(kern, grid, Cfg::kThreads, smem, st, tmQKV, tmQG, table, out, H, W, C, heads, num_windows, qg ? 1 : 0,
             1.4426950408889634f / sqrtf((float)HD));
  VIP_CUDA(cudaGetLastError());
  count_launch();
#ifdef VIP_ATTN_TRACE
  {
    long long t[8];
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(t, g_attn_trace, sizeof(t));
    if (t[0] > 0)
      fprintf(stderr, "[attn trace ws%d] jobs %lld  QK->S %lld  S->P %lld  P->O %lld cycles/job\n", WS, t[0], t[1] / t[0],
              t[2] / t[0], t[3] / t[0]);
    long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    cudaMemcpyToSymbol(g_attn_trace, z, sizeof(z));
  }
#endif
  return VIP_OK;
}

}  // namespace

// VIP_ERR_UNSUPPORTED when the shape does not fit (the bias tables of all heads must fit in shared memory)
int window_attention_ws(const void* qkv, const void* qg, const float* table, void* out, int B, int H, int W, int C, int ws,
                        int heads, cudaStream_t st) {
  if (ws == 7) return launch_ws<7>((const bf16*)qkv, (const bf16*)qg, table, (bf16*)out, B, H, W, C, heads, st);
  if (ws == 14) return launch_ws<14>((const bf16*)qkv, (const bf16*)qg, table, (bf16*)out, B, H, W, C, heads, st);
  return VIP_ERR_UNSUPPORTED;
}

}  // namespace vip
