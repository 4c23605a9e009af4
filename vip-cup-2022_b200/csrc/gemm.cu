// tcgen05 / TMEM / TMA GEMM + implicit-GEMM convolution for sm_100a.  See gemm.cuh for the contract.
//
// Persistent kernel, one CTA per SM, static round-robin over 128 x BN output tiles.  Warp roles (320 threads):
//   warp 0      TMA producer.  A k-blocks [128 x 64] come either from a 2-D tiled tensor map over A[M,K] or, for a
//               convolution, from an im2col-mode tensor map over the NHWC activation (one (filter tap, 64-channel
//               block) per k-block; padding and stride are resolved by the TMA unit, nothing is materialised).  B
//               k-blocks [BN x 64] come from a 2-D map over the weights.  128-byte swizzle, kStages-deep smem ring,
//               completion on "full" mbarriers.  The residual tile (if any) is also brought in by TMA, into the
//               output staging buffers.
//   warp 1      TMEM allocation (2 x BN fp32 columns = two accumulator stages) + MMA issuer: one lane issues
//               tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN, K=16) x4 per k-block; tcgen05.commit releases the
//               ring slot and, after the last k-block of a tile, hands the accumulator stage to the epilogue.  The
//               MMAs of tile i+1 run while the epilogue drains tile i.
//   warps 2..9  epilogue (256 threads): warp w owns TMEM lanes 32*(w%4).., the two warps of a lane quarter split each
//               64-column chunk in halves.  tcgen05.ld -> registers -> folded LayerNorm / bias / activation /
//               layer-scale / residual -> bf16 -> swizzled smem staging -> one TMA store per 64-column chunk (rows
//               beyond M and columns beyond N are clipped by the TMA unit).  Optional row statistics (for the
//               LayerNorm folded into the NEXT contraction) and global-average-pool partial sums are accumulated here.
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "gemm.cuh"
#include "stats.cuh"

namespace vip {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int kABytes = BM * BK * 2;
constexpr int kCBufBytes = BM * 128;  // 128 rows x 64 bf16
constexpr int kMaxCols = 8192;        // widest N the neutral-parameter vectors cover

// Neutral per-column parameter vectors: an absent bias / ln_colsum points at zeros, an absent colscale at ones, so the
// epilogue inner loop carries no per-feature branches.
__device__ float g_zeros[kMaxCols];
__device__ float g_ones[kMaxCols];
__global__ void init_neutral_kernel() {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < kMaxCols; i += gridDim.x * blockDim.x) {
    g_zeros[i] = 0.0f;
    g_ones[i] = 1.0f;
  }
}

struct GemmArgs {
  int M, N, K;
  int num_kb, m_tiles, n_tiles;
  int conv;  // 0: plain GEMM; 1: A from the im2col-mode map (linear pixel tiles); 2: A from a tiled 4-D map (patch tiles)
  int cblocks, C, ks, stride, pad, Wo, HoWo;
  // patch tiles (conv == 2): a tile is bni images x bh rows x bw columns of the output (<= 128 pixels)
  int bw, bh, bni, tiles_w, tiles_h, prow, Ho, Nimg;
  int mode;  // EpiMode
  // grouped B (plain GEMM only): rows [i * bgroup_mtiles * 128, ...) of A are contracted with rows
  // [i * bgroup_rows, (i + 1) * bgroup_rows) of B -- one weight matrix per image (0 = one B for all rows)
  int bgroup_mtiles, bgroup_rows;
  int dbg;            // experiments (VIP_GEMM_DEBUG): 4 = MMA warp does not wait for operands, 8 = MMA warp issues no MMAs,
                      // 16 = folded-LN consumer does not load the statistics records, 32 = producer does not write its
                      // records, 64 = low planes are not touched
  long long* trace;  // VIP_GEMM_TRACE=1: per-tile clock64() stamps of CTA 0 ([tile][8]); null otherwise
  GemmEpilogue epi;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#ifdef VIP_MBAR_DEBUG
// debug build: a wait that does not complete within ~2^27 polls reports where it is stuck and traps
__device__ __noinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, int line) {
  for (unsigned long long n = 0;; ++n) {
    uint32_t ok;
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (n > (1ull << 25)) {
      printf("mbar stuck: line %d block %d thread %d parity %u\n", line, (int)blockIdx.x, (int)threadIdx.x, parity);
      __trap();
    }
  }
}
#define mbar_wait(bar, parity) mbar_wait_dbg(bar, parity, __LINE__)
#else
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
#endif
// polling wait (mbarrier.test_wait): for barriers whose arrivals come from the OTHER CTA of a pair; a thread parked in
// try_wait is not woken promptly by remote arrivals
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SPIN_%=:\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra SDONE_%=;\n"
      "bra SPIN_%=;\n"
      "SDONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// im2col-mode load of [pixels x channels]: coordinates (c, w, h, n) of the first pixel's filter-window origin in the
// input, offsets (s, r) of the filter tap
__device__ __forceinline__ void tma_load_im2col_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c, int w,
                                                   int h, int n, uint16_t off_w, uint16_t off_h) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.im2col.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2], {%7, %8};" ::"r"(smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c), "r"(w), "r"(h), "r"(n), "h"(off_w), "h"(off_h)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* tmap, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(tmap),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// pair-mode TMA load: data lands in this CTA's shared memory, the bytes are credited to the LEADER's barrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint32_t leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(leader_bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, 128-byte swizzle, dense [rows][64 bf16] tile: rows are 128 B apart, 8-row groups 1024 B apart.
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address
  d |= (uint64_t)1 << 16;                         // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;               // stride byte offset
  d |= (uint64_t)1 << 46;                         // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                         // SWIZZLE_128B
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

// GELU.  Keras' 'gelu' is the exact erf form (models/gcvit/layers/feature.py:21).  Evaluated as
//   x * Phi(x),  Phi(x) = 0.5 (1 + tanh(x (a0 + a1 x^2 + a2 x^4)))
// with (a0, a1, a2) fitted to atanh(erf(x / sqrt 2)) (max |formula error| 2.6e-5 over all x; the textbook two-term tanh
// form is 20x worse) and the hardware tanh.approx.f32 (relative error 2^-11): |error| <= 3e-4 |x|, 0.07 ulp of the bf16
// value the result is stored as.  One MUFU + 7 FMA-pipe instructions: the erf / exp forms need two MUFU per element,
// which makes the fc1 epilogues MUFU-bound (16 results / clock / SM).
__device__ __forceinline__ float gelu_fast(float x) {
  const float x2 = fminf(x * x, 64.0f);  // the fitted polynomial is monotone up to |x| = 8, where tanh has saturated
  float p = fmaf(-0.00035307545f, x2, 0.037015257f);
  p = fmaf(p, x2, 0.79749725f);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(p * x));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// kDeep: one CTA per SM with a deep operand ring (long K loops).  !kDeep: two CTAs per SM with a short ring, for the
// output-heavy contractions (K <= 256) whose time goes into the epilogue; BN <= 128 so that both CTAs get TMEM.
// kPair (implies kDeep): a cluster of two CTAs on the two SMs of a TPC computes one 256 x BN tile with
// tcgen05.mma.cta_group::2: each CTA stages its own 128 rows of A and HALF of the B k-block, the leader's MMA reads both
// halves.  Operand traffic into each SM's shared memory drops by a third (BN = 256), which is what bounds the
// single-CTA kernel (A + B fill plus the MMA's own reads exceed the 128 B/clk of shared memory).
template <int BN, bool kDeep, bool kPair = false>
struct Cfg {
  static constexpr int kBBytes = (kPair ? BN / 2 : BN) * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = kPair ? (BN == 256 ? 5 : 6) : kDeep ? (BN == 256 ? 3 : BN == 128 ? 5 : 6) : (BN == 128 ? 2 : 3);
  static constexpr int kCBufs = kDeep ? 4 : 2;   // ring of output staging buffers (one 64-column chunk each)
  static constexpr int kMinBlocks = kDeep ? 1 : 2;
  static constexpr int kChunks = BN / 64;   // 64-column chunks per tile
  // epilogue groups of 4 warps (each covers all 128 rows): one per chunk for the 256-wide deep tiles, otherwise two
  // Two epilogue groups in every configuration: at BN = 256 each group takes two chunks.  Four groups (16 epilogue warps,
  // 576 threads) cap the kernel at 96 registers with spills and measured 3-11 % slower (profiles/r2_gemm_groups_ab.txt).
  static constexpr int kGroups = 2;
  static constexpr int kThreads = 64 + 128 * kGroups;
  static constexpr int kTmemCols = 2 * BN;  // 128, 256 or 512: a power of two
  static constexpr int kSmem = kStages * kStageBytes + kCBufs * kCBufBytes + 1024 + 256;
  static_assert(kDeep || BN <= 128, "two CTAs per SM need at most 256 TMEM columns each");
  static_assert(!kPair || (kDeep && BN >= 128), "pair mode is a deep-ring configuration");
};

// One group of 8 output columns of one row: folded LayerNorm + bias (2 FMAs), activation, column scale.
template <int ACT>
__device__ __forceinline__ void epi_group(float (&v)[8], float rstd, float nmr, const float* __restrict__ colsum,
                                          const float* __restrict__ bias, const float* __restrict__ colscale, int n) {
  const float4 s0 = __ldg(reinterpret_cast<const float4*>(colsum + n)), s1 = __ldg(reinterpret_cast<const float4*>(colsum + n) + 1);
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(bias + n) + 1);
  const float4 c0 = __ldg(reinterpret_cast<const float4*>(colscale + n)), c1 = __ldg(reinterpret_cast<const float4*>(colscale + n) + 1);
  const float cs[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
  const float bs[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  const float sc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float t = fmaf(v[i], rstd, fmaf(nmr, cs[i], bs[i]));  // rstd * (acc - mean * colsum) + bias
    if (ACT == ACT_RELU) t = fmaxf(t, 0.0f);
    if (ACT == ACT_GELU) t = gelu_fast(t);
    if (ACT == ACT_SIGMOID) t = 1.0f / (1.0f + __expf(-t));
    if (ACT == ACT_SWISH) t = t / (1.0f + __expf(-t));   // x * sigmoid(x) (EfficientNet / NFNet activations)
    v[i] = t * sc[i];
  }
}

__device__ __forceinline__ void tma_store_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// Specialised epilogues (compile-time feature sets) for the combinations the backbones use; everything else takes the
// generic path with neutral parameter vectors.
enum EpiMode : int { EPI_GENERIC = 0, EPI_NONE, EPI_RELU, EPI_GELU, EPI_LN, EPI_LN_GELU, EPI_RES, EPI_SE };

// 32 columns of one output row: TMEM registers -> (folded LN) + bias -> activation -> (+ residual, row statistics) ->
// bf16 -> swizzled staging row.  `nb` = first column, `cbase16` = 16-byte chunk index of that column inside the row.
template <int MODE>
__device__ __forceinline__ void epi_row32(const uint32_t (&r)[32], uint8_t* crow, uint32_t swz, uint32_t cbase16, int nb,
                                          int N, float rstd, float nmr, const float* __restrict__ colsum,
                                          const float* __restrict__ bias, float& rs_sum, float& rs_sq,
                                          const float* __restrict__ gate_row = nullptr, float pivot = 0.0f,
                                          const __nv_bfloat16* __restrict__ lo_in = nullptr,
                                          __nv_bfloat16* __restrict__ lo_out = nullptr, const uint4* lo_regs = nullptr) {
  constexpr bool kLN = MODE == EPI_LN || MODE == EPI_LN_GELU;
  constexpr bool kGelu = MODE == EPI_GELU || MODE == EPI_LN_GELU;
#pragma unroll
  for (int q8 = 0; q8 < 4; ++q8) {
    const int n = nb + q8 * 8;
    if (n >= N) break;  // N % 8 == 0: a group of 8 columns is entirely inside or outside
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(bias + n) + 1);
    float sh[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    if (kLN) {
      const float4 s0 = __ldg(reinterpret_cast<const float4*>(colsum + n)), s1 = __ldg(reinterpret_cast<const float4*>(colsum + n) + 1);
      const float cs[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) sh[i] = fmaf(nmr, cs[i], sh[i]);
    }
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float a = __uint_as_float(r[q8 * 8 + i]);
      float t = kLN ? fmaf(a, rstd, sh[i]) : a + sh[i];
      if (MODE == EPI_RELU) t = fmaxf(t, 0.0f);
      if (kGelu) t = gelu_fast(t);
      v[i] = t;
    }
    uint4* cp = reinterpret_cast<uint4*>(crow + ((cbase16 + q8) ^ swz) * 16);
    if (MODE == EPI_SE) {
      // excite, add the shortcut, ReLU (resnet_rs_model.py:183,278-280)
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gate_row + n)), g1 = __ldg(reinterpret_cast<const float4*>(gate_row + n) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const uint4 u = *cp;
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        v[2 * t] = fmaxf(fmaf(v[2 * t], gg[2 * t], __uint_as_float(w[t] << 16)), 0.0f);
        v[2 * t + 1] = fmaxf(fmaf(v[2 * t + 1], gg[2 * t + 1], __uint_as_float(w[t] & 0xffff0000u)), 0.0f);
      }
    }
    if (MODE == EPI_RES) {
      const uint4 u = *cp;
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        v[2 * t] += __uint_as_float(w[t] << 16);
        v[2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
      }
      if (lo_in != nullptr) {   // low plane of the two-plane residual stream (null for rows beyond M)
        // lo_regs: fetched before the wait for the tile's MMAs (deep configuration), else one L2 round trip per group here
        const uint4 ul = lo_regs != nullptr ? lo_regs[q8] : __ldg(reinterpret_cast<const uint4*>(lo_in + (size_t)(n >> 3) * 256));
        const uint32_t wl[4] = {ul.x, ul.y, ul.z, ul.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          v[2 * t] += __uint_as_float(wl[t] << 16);
          v[2 * t + 1] += __uint_as_float(wl[t] & 0xffff0000u);
        }
      }
      // row statistics: fp32 partial sums over this 32-column piece, in column order (the caller converts them to fixed
      // point once per piece)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[i] - pivot;
        rs_sum += d;
        rs_sq = fmaf(d, d, rs_sq);
      }
    }
    uint32_t w[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
      w[t] = *reinterpret_cast<const uint32_t*>(&h2);
    }
    *cp = make_uint4(w[0], w[1], w[2], w[3]);
    if (MODE == EPI_RES && lo_out != nullptr) {   // what the bf16 rounding of the high plane dropped
      uint32_t wl[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * t] - __uint_as_float(w[t] << 16),
                                                        v[2 * t + 1] - __uint_as_float(w[t] & 0xffff0000u));
        wl[t] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      *reinterpret_cast<uint4*>(lo_out + (size_t)(n >> 3) * 256) = make_uint4(wl[0], wl[1], wl[2], wl[3]);
    }
  }
}

// Every other feature combination: neutral parameter vectors, run-time activation, bf16 or f32 output.
__device__ __forceinline__ void epi_generic32(const uint32_t (&r)[32], uint8_t* crow, uint32_t swz, uint32_t cbase16, int nb,
                                              int N, int row, int M, const GemmEpilogue& e, bool has_res, float rstd,
                                              float nmr, const float* __restrict__ p_colsum,
                                              const float* __restrict__ p_bias, const float* __restrict__ p_colscale,
                                              float& rs_sum, float& rs_sq, float pivot) {
#pragma unroll
  for (int q8 = 0; q8 < 4; ++q8) {
    const int n = nb + q8 * 8;
    if (n >= N) break;  // N % 8 == 0: a group of 8 columns is entirely inside or outside
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[q8 * 8 + i]);
    switch (e.act) {
      case ACT_RELU: epi_group<ACT_RELU>(v, rstd, nmr, p_colsum, p_bias, p_colscale, n); break;
      case ACT_GELU: epi_group<ACT_GELU>(v, rstd, nmr, p_colsum, p_bias, p_colscale, n); break;
      case ACT_SIGMOID: epi_group<ACT_SIGMOID>(v, rstd, nmr, p_colsum, p_bias, p_colscale, n); break;
      case ACT_SWISH: epi_group<ACT_SWISH>(v, rstd, nmr, p_colsum, p_bias, p_colscale, n); break;
      default: epi_group<ACT_NONE>(v, rstd, nmr, p_colsum, p_bias, p_colscale, n); break;
    }
    uint4* cp = reinterpret_cast<uint4*>(crow + ((cbase16 + q8) ^ swz) * 16);
    if (has_res) {
      const uint4 u = *cp;
      const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        v[2 * t] += __uint_as_float(w[t] << 16);
        v[2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
      }
    }
    if (e.out_bf16 != nullptr) {
      uint32_t w[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * t], v[2 * t + 1]);
        w[t] = *reinterpret_cast<const uint32_t*>(&h2);
        const float lo = __uint_as_float(w[t] << 16) - pivot, hi = __uint_as_float(w[t] & 0xffff0000u) - pivot;
        rs_sum += lo + hi;
        rs_sq = fmaf(lo, lo, fmaf(hi, hi, rs_sq));
      }
      *cp = make_uint4(w[0], w[1], w[2], w[3]);
    } else if (row < M) {
      float4* op = reinterpret_cast<float4*>(e.out_f32 + (size_t)row * e.ldc + n);
      op[0] = make_float4(v[0], v[1], v[2], v[3]);
      op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
  }
}

template <int BN, bool kDeep, bool kPair = false>
__global__ void __launch_bounds__(Cfg<BN, kDeep, kPair>::kThreads, Cfg<BN, kDeep, kPair>::kMinBlocks)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmR, const GemmArgs g) {
  using C = Cfg<BN, kDeep, kPair>;
  constexpr int kCtas = kPair ? 2 : 1;          // CTAs that share one tile
  const uint32_t cta_rank = kPair ? cluster_ctarank() : 0u;
  const int cta_tile0 = blockIdx.x / kCtas, cta_tile_step = gridDim.x / kCtas;
  constexpr int kStages = C::kStages;
  constexpr int kCBufs = C::kCBufs;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* cbuf = smem + kStages * C::kStageBytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(cbuf + kCBufs * kCBufBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tfull_bar = empty_bar + kStages;   // [2] accumulator stage filled by the MMAs
  uint64_t* tempty_bar = tfull_bar + 2;        // [2] accumulator stage drained by the epilogue
  uint64_t* rfull_bar = tempty_bar + 2;        // [kCBufs] residual chunk landed in a staging buffer
  uint64_t* cfree_bar = rfull_bar + kCBufs;    // [kCBufs] staging buffer free for the next residual chunk
  uint64_t* pfull_bar = cfree_bar + kCBufs;    // [kStages] pair mode, leader only: the PEER's ring slot has landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pfull_bar + kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles;
  const GemmEpilogue& e = g.epi;
  const bool has_res = e.residual != nullptr;
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    if (e.out_bf16 != nullptr) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    if (has_res) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmR) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
      mbar_init(&pfull_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull_bar[s], 1);
      mbar_init(&tempty_bar[s], 4 * C::kGroups * kCtas);  // one arrival per epilogue warp (of both CTAs in pair mode)
    }
    for (int b = 0; b < kCBufs; ++b) {
      mbar_init(&rfull_bar[b], 1);
      mbar_init(&cfree_bar[b], 1);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (kPair) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                   "n"(C::kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // both CTAs' barriers are initialised before anything arrives on them remotely
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above (descriptor prefetch, barriers, TMEM) overlapped the tail of the preceding kernel; from here on the
  // kernel reads and writes activations
  pdl_wait();

  if (warp == 0) {
    // ================= TMA producer =================
    if (lane == 0) {
      uint32_t it = 0, tcountp = 0;
      for (int tile = cta_tile0; tile < total_tiles; tile += cta_tile_step) {
        const int m0 = ((tile / g.n_tiles) * kCtas + (int)cta_rank) * BM, n0 = (tile % g.n_tiles) * BN;
        int cw = 0, ch = 0, cn = 0;
        if (g.conv == 1) {
          cn = m0 / g.HoWo;
          const int rem = m0 - cn * g.HoWo;
          const int p0 = rem / g.Wo, q0 = rem - p0 * g.Wo;
          cw = q0 * g.stride - g.pad;
          ch = p0 * g.stride - g.pad;
        } else if (g.conv == 2) {
          const int mt = tile / g.n_tiles;
          const int pw = mt % g.tiles_w, t2 = mt / g.tiles_w;
          cw = pw * g.bw * g.stride - g.pad;
          ch = (t2 % g.tiles_h) * g.bh * g.stride - g.pad;
          cn = (t2 / g.tiles_h) * g.bni;
        }
        const uint32_t stage_tx = (g.conv == 2 ? (uint32_t)g.prow * 128u : (uint32_t)kABytes) + (uint32_t)C::kBBytes;
        if (g.trace != nullptr && blockIdx.x == 0) g.trace[(tile / cta_tile_step) * 16 + 0] = clock64();
        for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          if (kPair) mbar_wait_spin(&empty_bar[s], ph ^ 1);
          else mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* a_dst = smem + s * C::kStageBytes;
          if (kPair) {
            // this CTA stages its 128 rows of A and its half of the B k-block; completion stays on the LOCAL barrier
            // (crediting the peer's bytes to the leader's barrier directly costs one cross-SM message per 128-byte row
            // and throttles the ring to ~20 B/clk; the peer forwards ONE arrival per slot instead, see the MMA warp)
            mbar_expect_tx(&full_bar[s], stage_tx);
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
            tma_load_2d(a_dst + kABytes, &tmB, &full_bar[s], kb * BK, n0 + (int)cta_rank * (BN / 2));
            continue;
          }
          mbar_expect_tx(&full_bar[s], stage_tx);
          if (g.conv) {
            const int tap = kb / g.cblocks, cb = kb - tap * g.cblocks;
            const int r = tap / g.ks, sx = tap - r * g.ks;
            if (g.conv == 1) tma_load_im2col_4d(a_dst, &tmA, &full_bar[s], cb * BK, cw, ch, cn, (uint16_t)sx, (uint16_t)r);
            else tma_load_4d(a_dst, &tmA, &full_bar[s], cb * BK, cw + sx, ch + r, cn);
            tma_load_2d(a_dst + kABytes, &tmB, &full_bar[s], tap * g.C + cb * BK, n0);
          } else {
            tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
            const int brow = g.bgroup_mtiles > 0 ? ((tile / g.n_tiles) / g.bgroup_mtiles) * g.bgroup_rows : 0;
            tma_load_2d(a_dst + kABytes, &tmB, &full_bar[s], kb * BK, n0 + brow);
          }
        }
        if (g.trace != nullptr && blockIdx.x == 0) g.trace[(tile / cta_tile_step) * 16 + 1] = clock64();
        if (has_res) {
          // after this tile's operand loads are in flight: bring the residual chunks into the staging buffers of this
          // tile's set, each as soon as the store that last used the buffer has finished reading it.  Every buffer of
          // the set takes part in the handshake, also the ones a ragged last N tile does not fill.
          constexpr int kSetsP = kCBufs / C::kChunks;
          const uint32_t set = tcountp % kSetsP, u = tcountp / kSetsP;
          for (int j = 0; j < C::kChunks; ++j) {
            const uint32_t b = set * C::kChunks + j;
            mbar_wait(&cfree_bar[b], (u & 1) ^ 1);
            if (n0 + j * 64 < g.N) {
              mbar_expect_tx(&rfull_bar[b], kCBufBytes);
              tma_load_2d(cbuf + b * kCBufBytes, &tmR, &rfull_bar[b], n0 + j * 64, m0);
            } else {
              mbar_arrive(&rfull_bar[b]);
            }
          }
        }
        ++tcountp;
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      constexpr uint32_t idesc_pair = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      uint32_t it = 0, tcount = 0;
      if (kPair && cta_rank != 0) {
        // peer CTA: no MMAs to issue; tell the leader when each of this CTA's ring slots has landed
        for (int tile = cta_tile0; tile < total_tiles; tile += cta_tile_step) {
          for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
            const int s = it % kStages;
            mbar_wait(&full_bar[s], (it / kStages) & 1);
            mbar_arrive_cluster(mapa_u32(smem_u32(&pfull_bar[s]), 0));
          }
        }
      }
      it = 0;
      for (int tile = cta_tile0; tile < total_tiles && cta_rank == 0; tile += cta_tile_step, ++tcount) {
        const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
        if (kPair) mbar_wait_spin(&tempty_bar[as], aph ^ 1);
        else mbar_wait(&tempty_bar[as], aph ^ 1);
        tcgen05_fence_after();
        if (g.trace != nullptr && blockIdx.x == 0) g.trace[(tile / cta_tile_step) * 16 + 2] = clock64();
        const uint32_t d_tmem = tmem_base + as * BN;
        for (int kb = 0; kb < g.num_kb; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          if (!(g.dbg & 4)) {
            mbar_wait(&full_bar[s], ph);
            if (kPair) mbar_wait_spin(&pfull_bar[s], ph);
          }
          tcgen05_fence_after();
          if (g.trace != nullptr && blockIdx.x == 0 && kb == 0) g.trace[(tile / cta_tile_step) * 16 + 3] = clock64();
          const uint32_t a_addr = smem_u32(smem + s * C::kStageBytes);
          const uint64_t a_desc = make_sw128_desc(a_addr);
          const uint64_t b_desc = make_sw128_desc(a_addr + kABytes);
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            if (g.dbg & 8) break;
            // advance 16 elements = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
            if (kPair) umma_bf16_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc_pair, (kb | k) != 0 ? 1u : 0u);
            else umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // implies tcgen05.fence::before_thread_sync; pair mode: releases the ring slot in BOTH CTAs
          if (kPair) umma_commit_pair(&empty_bar[s]);
          else umma_commit(&empty_bar[s]);
        }
        if (kPair) umma_commit_pair(&tfull_bar[as]);
        else umma_commit(&tfull_bar[as]);
        if (g.trace != nullptr && blockIdx.x == 0) g.trace[(tile / cta_tile_step) * 16 + 4] = clock64();
      }
    }
  } else {
    // ================= epilogue =================
    // Two groups of 4 warps, each covering all 128 rows of the tile.  A group owns BN/2 columns = kCPG 64-column chunk
    // buffers per tile, fills them, syncs once among its 128 threads and issues its TMA stores as one bulk group; the two
    // groups never wait for each other (BN = 64: they fill the two halves of ONE chunk buffer and share a barrier).
    const int te = threadIdx.x - 64;           // 0 .. 128 * kGroups - 1
    const int quarter = warp & 3;              // TMEM lanes 32*quarter .. +31
    const int group = (warp - 2) >> 2;
    const int tg = te & 127;
    const int rt = quarter * 32 + lane;        // row inside the tile
    const uint32_t swz = (uint32_t)(rt & 7);
    constexpr bool kShared = BN == 64;
    constexpr int kCPG = kShared ? 1 : C::kChunks / C::kGroups;
    constexpr int kSets = kCBufs / C::kChunks;  // consecutive tiles whose staging buffers are disjoint
    const bool storer = kShared ? te == 0 : tg == 0;
    const float* p_colsum = e.ln_colsum != nullptr ? e.ln_colsum : g_zeros;
    const float* p_bias = e.bias != nullptr ? e.bias : g_zeros;
    const float* p_colscale = e.colscale != nullptr ? e.colscale : g_ones;
    auto group_sync = [&]() {
      if (kShared) asm volatile("bar.sync 1, 256;" ::: "memory");
      else asm volatile("bar.sync %0, 128;" ::"r"(2 + group) : "memory");
    };
    uint32_t tcount = 0;  // tiles processed by this CTA; tile t stages into buffer set t % kSets
    // row statistics of the folded LayerNorm: fetched one tile ahead so that the load never sits on the critical path
    auto load_stats = [&](int tile_) -> float2 {   // (mean, 1 / sigma) of this thread's row of that tile
      const long long row_ = ((long long)(tile_ / g.n_tiles) * kCtas + cta_rank) * BM + rt;
      if (e.ln_stats == nullptr || tile_ >= total_tiles || row_ >= g.M || (g.dbg & 16)) return make_float2(0.0f, 1.0f);
      return __ldg(reinterpret_cast<const float2*>(e.ln_stats) + row_);
    };
    float2 st_next = load_stats(cta_tile0);
    auto load_pivot = [&](int tile_) -> float {
      const long long row_ = ((long long)(tile_ / g.n_tiles) * kCtas + cta_rank) * BM + rt;
      if (e.row_pivot == nullptr || e.row_stats == nullptr || tile_ >= total_tiles || row_ >= g.M) return 0.0f;
      return __ldg(e.row_pivot + 2 * row_);
    };
    float pv_next = load_pivot(cta_tile0);
    for (int tile = cta_tile0; tile < total_tiles; tile += cta_tile_step, ++tcount) {
      const int m0 = ((tile / g.n_tiles) * kCtas + (int)cta_rank) * BM, n0 = (tile % g.n_tiles) * BN;
      const uint32_t as = tcount & 1, aph = (tcount >> 1) & 1;
      const int row = m0 + rt;
      const int nvalid = min(C::kChunks, (g.N - n0 + 63) >> 6);   // chunk buffers this tile fills
      const uint32_t bset = (tcount % kSets) * C::kChunks;         // first buffer of this tile's set
      int pq0 = 0, pp0 = 0, pn0 = 0;  // patch tile origin (output column, row, image)
      if (g.conv == 2) {
        const int mt = tile / g.n_tiles;
        const int t2 = mt / g.tiles_w;
        pq0 = (mt % g.tiles_w) * g.bw;
        pp0 = (t2 % g.tiles_h) * g.bh;
        pn0 = (t2 / g.tiles_h) * g.bni;
      }
      float rstd = 1.0f, nmr = 0.0f;  // 1/sigma and -mean/sigma of this row (identity without a folded LayerNorm)
      if (e.ln_stats != nullptr) {
        const float2 st = st_next;
        st_next = load_stats(tile + cta_tile_step);
        rstd = st.y;
        nmr = -st.x * st.y;
      }
      // two-plane residual stream: this row's low planes; pivot of the row statistics = the row's previous value of
      // column 0 (every thread of the row fetches the same two numbers, whichever N tile it works on)
      const bool row_ok = row < g.M;
      // (low planes are stored in 32-row x 8-column blocks, stats.cuh: lo_* point at this row's slot of column block 0,
      //  column block k lies 256 elements further)
      const __nv_bfloat16* lo_in = (e.residual_lo != nullptr && row_ok && !(g.dbg & 64)) ? e.residual_lo + lo_plane_index(row, 0, g.N) : nullptr;
      __nv_bfloat16* lo_out = (e.out_lo != nullptr && row_ok && !(g.dbg & 64)) ? e.out_lo + lo_plane_index(row, 0, g.N) : nullptr;
      // Low-plane values of this thread's 64 columns, fetched NOW (the tile's MMAs are still running): in the body each
      // group of 8 columns would expose its own L2 / DRAM round trip, four in a row per 32-column piece (timeline: 2.6-4.3 k
      // cycles of "math" per piece against ~1 k without the planes).  Registers allow it in the one-CTA-per-SM BN = 128
      // configuration (the level-2 / level-3 proj and fc2 of GCViT); elsewhere the blocks are only pulled towards L2.
      constexpr bool kLoRegs = kDeep && !kPair && BN == 128;
      uint4 lo_pre[2][4];
      bool lo_pre_valid = false;
      if (kLoRegs && lo_in != nullptr && g.mode == EPI_RES) {
        lo_pre_valid = true;
        const int jf = group * kCPG;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
#pragma unroll
          for (int q8 = 0; q8 < 4; ++q8) {
            const int n = n0 + jf * 64 + hh * 32 + q8 * 8;
            lo_pre[hh][q8] = n < g.N ? __ldg(reinterpret_cast<const uint4*>(lo_in + (size_t)(n >> 3) * 256)) : make_uint4(0u, 0u, 0u, 0u);
          }
      } else if (lo_in != nullptr) {
        // the low-plane blocks this thread will read, towards L2 now (no registers or shared memory to spare here)
        const int jf = kShared ? 0 : group * kCPG;
        for (int j = jf; j < jf + kCPG && n0 + j * 64 < g.N; ++j) {
#pragma unroll
          for (int q = (kShared ? group * 4 : 0); q < (kShared ? group * 4 + 4 : 8); ++q)
            if (n0 + j * 64 + q * 8 < g.N)
              asm volatile("prefetch.global.L2 [%0];" ::"l"(lo_in + (size_t)((n0 + j * 64 + q * 8) >> 3) * 256));
        }
      }
      // pivot of the row statistics: the mean the previous LayerNorm saw for this row (fetched one tile ahead)
      const float pivot = pv_next;
      pv_next = load_pivot(tile + cta_tile_step);
      const float* gate_row = e.row_gate != nullptr ? e.row_gate + (size_t)(min(row, g.M - 1) / e.gate_rows) * g.N : nullptr;
      {
        // SE gate of the row's image for this group's chunk: towards L1 while the MMAs run (measured 1-3 % on the conv_3
        // shapes of ResNet-RS; the same prefetch of the bias / column-sum lines measured +1.5 % on the LN epilogues: dropped)
        const int c0 = n0 + (kShared ? group * 32 : group * kCPG * 64);
        if (c0 < g.N) {
          if (gate_row != nullptr) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(gate_row + c0));
            if (!kShared) asm volatile("prefetch.global.L1 [%0];" ::"l"(gate_row + min(c0 + 32, g.N - 8)));
          }
        }
      }
      if (kPair) mbar_wait_spin(&tfull_bar[as], aph);
      else mbar_wait(&tfull_bar[as], aph);
      tcgen05_fence_after();
      if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 5] = clock64();

      // ---- staging buffers of this tile must have been read out by the stores that last used them
      if (storer) {
        if (kSets >= 2 && has_res) {
          // every store issued so far is done: the previous tile's buffers are free for the tile AFTER this one, whose
          // residual the producer can now fetch while this epilogue runs
          tma_store_wait_read_all();
          if (tcount > 0) {
            const uint32_t pset = ((tcount - 1) % kSets) * C::kChunks;
            for (int j = 0; j < C::kChunks; ++j)
              if (kShared || j / kCPG == group) mbar_arrive(&cfree_bar[pset + j]);
          }
        } else if (kSets >= 2) {
          asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kSets - 1) : "memory");
        } else {
          tma_store_wait_read_all();
        }
      }
      if (!has_res) group_sync();   // (with a residual the wait on its arrival below orders the buffer reuse)

      // Row statistics: an fp32 partial per aligned 32-column piece (fixed column order), converted to fixed point and
      // from then on added as integers.  The pieces are the same whatever tile width / epilogue grouping the launch
      // picked, so the totals depend neither on the tile schedule nor on the batch size (stats.cuh).
      long long rs_sum = 0, rs_sq = 0;
      // chunks of this group; the warp's last TMEM read of the tile hands the accumulator stage back to the MMA warp
      const int j_first = kShared ? 0 : group * kCPG;
      const int j_last = min(kShared ? 0 : group * kCPG + kCPG - 1, nvalid - 1);
      auto release_tmem = [&]() {
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (kPair) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[as]), 0));  // the leader's MMA warp waits for both
          else mbar_arrive(&tempty_bar[as]);
        }
      };
      if (j_last < j_first) release_tmem();   // nothing to read for this group (narrow N)
      if (has_res) {
        // chunks of this group that a ragged last N tile leaves empty still take part in the residual handshake: wait
        // for the producer's (data-less) arrival, otherwise this group could free the buffer a second time before the
        // producer has consumed the first release and the barrier phases would alias
        for (int j = max(j_last + 1, j_first); j <= (kShared ? 0 : group * kCPG + kCPG - 1); ++j)
          mbar_wait(&rfull_bar[bset + (uint32_t)j], (tcount / kSets) & 1);
      }
#pragma unroll 1
      for (int j = j_first; j <= j_last; ++j) {
        const uint32_t b = bset + (uint32_t)j;
        uint8_t* cbase = cbuf + b * kCBufBytes;
        uint8_t* crow = cbase + rt * 128;
        bool res_ready = !has_res;
#pragma unroll(kLoRegs ? 2 : 1)
        for (int hh = (kShared ? group : 0); hh < (kShared ? group + 1 : 2); ++hh) {
          uint32_t r[32];
          if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 7 + 3 * (hh & 1)] = clock64();
          tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(as * BN + j * 64 + hh * 32), r);
          if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 8 + 3 * (hh & 1)] = clock64();
          if (j == j_last && hh == (kShared ? group : 1)) release_tmem();
          if (!res_ready) {
            mbar_wait(&rfull_bar[b], (tcount / kSets) & 1);
            res_ready = true;
          }
          const int nb = n0 + j * 64 + hh * 32;
          const uint32_t c16 = (uint32_t)(hh * 4);
          float ps = 0.0f, pq = 0.0f;
          switch (g.mode) {
            case EPI_NONE: epi_row32<EPI_NONE>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq); break;
            case EPI_RELU: epi_row32<EPI_RELU>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq); break;
            case EPI_GELU: epi_row32<EPI_GELU>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq); break;
            case EPI_LN: epi_row32<EPI_LN>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq); break;
            case EPI_LN_GELU: epi_row32<EPI_LN_GELU>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq); break;
            case EPI_SE: epi_row32<EPI_SE>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq, gate_row); break;
            case EPI_RES:
              epi_row32<EPI_RES>(r, crow, swz, c16, nb, g.N, rstd, nmr, p_colsum, p_bias, ps, pq, nullptr, pivot, lo_in, lo_out,
                                 (kLoRegs && lo_pre_valid) ? lo_pre[hh & 1] : nullptr);
              break;
            default:
              epi_generic32(r, crow, swz, c16, nb, g.N, row, g.M, e, has_res, rstd, nmr, p_colsum, p_bias, p_colscale, ps, pq, pivot);
              break;
          }
          if (e.row_stats != nullptr) {
            rs_sum += to_fx(ps);
            rs_sq += to_fx(pq);
          }
        }
      }
      // ---- the group's chunks are staged: one fence, one barrier, its stores as one bulk group
      if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 12] = clock64();
      fence_async_smem();
      if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 13] = clock64();
      group_sync();
      if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 9] = clock64();
      if (storer && e.out_bf16 != nullptr) {
        for (int j = j_first; j <= j_last; ++j) {
          const uint8_t* cbase = cbuf + (bset + (uint32_t)j) * kCBufBytes;
          if (g.conv == 2) tma_store_4d(&tmC, cbase, n0 + j * 64, pq0, pp0, pn0);
          else tma_store_2d(&tmC, cbase, n0 + j * 64, m0);
        }
        tma_store_commit();
      }
      if (storer && has_res && kSets < 2) {
        // the next tile reuses these very buffers: its residual can only be fetched once the stores have read them
        tma_store_wait_read_all();
        for (int j = 0; j < C::kChunks; ++j)
          if (kShared || j / kCPG == group) mbar_arrive(&cfree_bar[bset + j]);
      }
      if (e.gap != nullptr) {
        // column sums of the staged bf16 chunks (split at image boundaries): thread -> 2 columns x kGapRows rows
        constexpr int kGapRows = kShared ? 16 : 32;
        const int cpair = te & 31, rq = kShared ? te >> 5 : tg >> 5;
        for (int j = j_first; j <= j_last; ++j) {
          const uint8_t* cbase = cbuf + (bset + (uint32_t)j) * kCBufBytes;
          const int n = n0 + j * 64 + cpair * 2;
          if (n >= g.N) continue;
          long long s0 = 0, s1 = 0;   // exact: every bf16 element is converted to fixed point before it is added
          if (g.conv == 2) {
            // patch tile: row rr = ((image, patch row, patch column)); rows outside the map / batch are skipped
            const int per_img = g.bw * g.bh;
            int cur = -1;
            for (int k = 0; k < kGapRows; ++k) {
              const int rr = rq * kGapRows + k;
              if (rr >= g.prow) break;
              const int ni = rr / per_img, rem = rr - ni * per_img;
              const int py = rem / g.bw, px = rem - py * g.bw;
              if (pn0 + ni >= g.Nimg || pp0 + py >= g.Ho || pq0 + px >= g.Wo) continue;
              if (ni != cur) {
                if (cur >= 0) {
                  fx_atomic_add_raw(e.gap + (size_t)(pn0 + cur) * g.N + n, s0);
                  fx_atomic_add_raw(e.gap + (size_t)(pn0 + cur) * g.N + n + 1, s1);
                }
                s0 = s1 = 0;
                cur = ni;
              }
              const uint32_t w = *reinterpret_cast<const uint32_t*>(
                  cbase + rr * 128 + ((((uint32_t)(cpair >> 2)) ^ (uint32_t)(rr & 7)) << 4) + (cpair & 3) * 4);
              s0 += to_fx(__uint_as_float(w << 16));
              s1 += to_fx(__uint_as_float(w & 0xffff0000u));
            }
            if (cur >= 0) {
              fx_atomic_add_raw(e.gap + (size_t)(pn0 + cur) * g.N + n, s0);
              fx_atomic_add_raw(e.gap + (size_t)(pn0 + cur) * g.N + n + 1, s1);
            }
          } else if (m0 + rq * kGapRows < g.M) {
            int rr = rq * kGapRows;
            int img = (m0 + rr) / e.gap_rows;
            int next = (img + 1) * e.gap_rows - m0;  // first tile row of the next image
            for (int k = 0; k < kGapRows; ++k, ++rr) {
              if (m0 + rr >= g.M) break;
              if (rr == next) {
                fx_atomic_add_raw(e.gap + (size_t)img * g.N + n, s0);
                fx_atomic_add_raw(e.gap + (size_t)img * g.N + n + 1, s1);
                s0 = s1 = 0;
                ++img;
                next += e.gap_rows;
              }
              const uint32_t w = *reinterpret_cast<const uint32_t*>(
                  cbase + rr * 128 + ((((uint32_t)(cpair >> 2)) ^ (uint32_t)(rr & 7)) << 4) + (cpair & 3) * 4);
              s0 += to_fx(__uint_as_float(w << 16));
              s1 += to_fx(__uint_as_float(w & 0xffff0000u));
            }
            fx_atomic_add_raw(e.gap + (size_t)img * g.N + n, s0);
            fx_atomic_add_raw(e.gap + (size_t)img * g.N + n + 1, s1);
          }
        }
      }
      if (g.trace != nullptr && blockIdx.x == 0 && te == 0) g.trace[(tile / cta_tile_step) * 16 + 6] = clock64();
      if (e.row_stats != nullptr && row < g.M && j_last >= j_first && !(g.dbg & 32)) {
        long long* rec = e.row_stats + 3 * (size_t)row;
        fx_atomic_add_raw(rec, rs_sum);
        fx_atomic_add_raw(rec + 1, rs_sq);
        if (n0 == 0 && group == 0) rec[2] = (long long)__float_as_int(pivot);   // the thread that holds column 0
      }
    }
    if (storer) tma_store_wait_all();

  }
  tcgen05_fence_before();
  __syncthreads();
  if (kPair) cluster_sync_all();   // no CTA of the pair leaves (or frees TMEM) while the other may still reach into it
  if (warp == 1) {
    tcgen05_fence_after();
    if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(C::kTmemCols) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
typedef CUresult (*EncodeIm2colFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const int*, const int*, cuuint32_t, cuuint32_t, const cuuint32_t*,
                                   CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                   CUtensorMapFloatOOBfill);

void* driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    return p;
  return nullptr;
}

// [rows, cols] bf16 row-major with `ld` elements between rows; box = [box_rows, 64 cols], 128B swizzle, zero OOB fill
int make_tmap_2d(CUtensorMap* tm, const void* base, long long rows, int cols, int ld, int box_rows) {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_fn("cuTensorMapEncodeTiled"));
  VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA,
              "cuTensorMapEncodeTiled failed with CUresult %d (rows=%lld cols=%d ld=%d box_rows=%d)", (int)r, rows, cols,
              ld, box_rows);
  return VIP_OK;
}

// NHWC activation as an im2col source: box = 128 output pixels x 64 channels
int make_tmap_im2col(CUtensorMap* tm, const void* base, const ConvGeom& c) {
  static EncodeIm2colFn fn = reinterpret_cast<EncodeIm2colFn>(driver_fn("cuTensorMapEncodeIm2col"));
  VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeIm2col is not available from the driver");
  const cuuint64_t ld = (cuuint64_t)(c.ldx > 0 ? c.ldx : c.C);
  const cuuint64_t gdim[4] = {(cuuint64_t)c.C, (cuuint64_t)c.W, (cuuint64_t)c.H, (cuuint64_t)c.Nimg};
  const cuuint64_t gstride[3] = {ld * 2, (cuuint64_t)c.W * ld * 2, (cuuint64_t)c.H * c.W * ld * 2};
  const int lower[2] = {-c.pad, -c.pad};                                        // (W, H)
  const int upper[2] = {c.pad - (c.ksize - 1), c.pad - (c.ksize - 1)};
  const cuuint32_t estr[4] = {1, (cuuint32_t)c.stride, (cuuint32_t)c.stride, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, lower, upper,
                        (cuuint32_t)BK, (cuuint32_t)BM, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA, "cuTensorMapEncodeIm2col failed with CUresult %d (N=%d H=%d W=%d C=%d k=%d s=%d p=%d)",
              (int)r, c.Nimg, c.H, c.W, c.C, c.ksize, c.stride, c.pad);
  return VIP_OK;
}

// NHWC tensor as a 4-D tiled map (C, W, H, N): box = 64 channels x bw x bh x bni positions taken every `stride`
int make_tmap_nhwc(CUtensorMap* tm, const void* base, int Nimg, int H, int W, int C, int bw, int bh, int bni, int stride) {
  static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_fn("cuTensorMapEncodeTiled"));
  VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)Nimg};
  const cuuint64_t gstride[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(bw * stride), (cuuint32_t)(bh * stride), (cuuint32_t)bni};
  const cuuint32_t estr[4] = {1, (cuuint32_t)stride, (cuuint32_t)stride, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA,
              "cuTensorMapEncodeTiled (NHWC) failed with CUresult %d (N=%d H=%d W=%d C=%d box %dx%dx%d stride %d)", (int)r,
              Nimg, H, W, C, bw, bh, bni, stride);
  return VIP_OK;
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

int ensure_neutral(cudaStream_t st) {
  static bool done = false;  // per process (one device per process in this library's use)
  if (!done) {
    init_neutral_kernel<<<8, 256, 0, st>>>();   // (once per process, plain launch)
    VIP_CUDA(cudaGetLastError());
    done = true;
  }
  return VIP_OK;
}

template <int BN, bool kDeep, bool kPair = false>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC, const CUtensorMap& tmR,
           const GemmArgs& g, cudaStream_t st) {
  using C = Cfg<BN, kDeep, kPair>;
  auto kern = gemm_tcgen05_kernel<BN, kDeep, kPair>;
  static bool configured = false;  // per-process; attribute is per-function
  if (!configured) {
    VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmem));
    configured = true;
  }
  const int tiles = g.m_tiles * g.n_tiles;
  const int ctas_per_tile = kPair ? 2 : 1;
  int slots = num_sms() * C::kMinBlocks / ctas_per_tile;
  const int grid = (tiles < slots ? tiles : slots) * ctas_per_tile;
  static const bool cluster_test = [] { const char* v = getenv("VIP_GEMM_CLUSTER_TEST"); return v != nullptr && v[0] == '1'; }();
  if (kPair || (cluster_test && grid % 2 == 0)) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(C::kThreads);
    cfg.dynamicSmemBytes = C::kSmem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    VIP_CUDA(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmR, g));
  } else {
    VIP_LAUNCH(kern, grid, C::kThreads, C::kSmem, st, tmA, tmB, tmC, tmR, g);
  }
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

// Tile width by a small cost model (cycles per tile x waves over the resident CTAs).  Per k-block a tile needs
// max(MMA issue: 2*bn cycles for M=128, operand delivery: bytes / ~64 B/clk/SM from L2 -- an im2col-mode A k-block takes
// ~650 cycles whatever the width); its epilogue needs ~600 cycles per 64-column chunk (halved when two CTAs share the SM).
int pick_bn(long long M, int N, bool deep, bool conv, int num_kb) {
  const int slots = num_sms() * (deep ? 1 : 2);
  const long long mt = (M + BM - 1) / BM;
  int best = 64;
  double best_cost = 1e30;
  for (int bn : {256, 128, 64}) {
    if (!deep && bn == 256) continue;
    const int nt = (N + bn - 1) / bn;
    const long long tiles = mt * nt;
    const double load = conv ? 650.0 + bn * 128 / 64.0 : (kABytes + bn * 128) / 64.0;
    const double mma = 2.0 * bn;
    const double main = num_kb * (load > mma ? load : mma);
    const int chunks = (N < bn ? N + 63 : bn) / 64;
    const double epi = chunks * (deep ? 600.0 : 300.0);
    const double per_tile = (main > epi ? main : epi) + 200.0;
    const double waves = (double)((tiles + slots - 1) / slots);
    const double cost = waves * per_tile * (deep ? 1.0 : 2.0);  // two co-resident CTAs share the SM's bandwidth
    if (cost < best_cost) { best_cost = cost; best = bn; }
  }
  return best;
}

int pick_mode(const GemmEpilogue& e) {
  if (e.row_gate != nullptr) return EPI_SE;
  if (e.out_bf16 == nullptr || e.colscale != nullptr || e.act == ACT_SIGMOID || e.act == ACT_SWISH) return EPI_GENERIC;
  const bool ln = e.ln_stats != nullptr, res = e.residual != nullptr;
  if (res) return (!ln && e.act == ACT_NONE) ? EPI_RES : EPI_GENERIC;
  if (e.row_stats != nullptr) return EPI_GENERIC;
  if (ln) return e.act == ACT_NONE ? EPI_LN : e.act == ACT_GELU ? EPI_LN_GELU : EPI_GENERIC;
  return e.act == ACT_NONE ? EPI_NONE : e.act == ACT_RELU ? EPI_RELU : EPI_GELU;
}

int check_epilogue(const GemmEpilogue& epi, int N) {
  VIP_REQUIRE((epi.out_bf16 != nullptr) != (epi.out_f32 != nullptr), VIP_ERR_INVALID,
              "gemm: exactly one of out_bf16 / out_f32 must be set");
  VIP_REQUIRE(epi.ldc % 8 == 0 && epi.ldc >= N, VIP_ERR_UNSUPPORTED, "gemm: ldc must be a multiple of 8 and >= N");
  VIP_REQUIRE(epi.residual == nullptr || (epi.ldr % 8 == 0 && ((uintptr_t)epi.residual & 15) == 0), VIP_ERR_UNSUPPORTED,
              "gemm: residual must be 16-byte aligned with ldr %% 8 == 0");
  VIP_REQUIRE((epi.ln_stats == nullptr) == (epi.ln_colsum == nullptr), VIP_ERR_INVALID,
              "gemm: ln_stats and ln_colsum come together");
  VIP_REQUIRE((epi.row_stats == nullptr && epi.gap == nullptr) || epi.out_bf16 != nullptr, VIP_ERR_UNSUPPORTED,
              "gemm: row_stats / gap need a bf16 output");
  VIP_REQUIRE(epi.gap == nullptr || epi.gap_rows > 0, VIP_ERR_INVALID, "gemm: gap_rows must be set with gap");
  VIP_REQUIRE(epi.gap == nullptr || epi.residual == nullptr, VIP_ERR_UNSUPPORTED,
              "gemm: gap with a residual is not supported (the pooled sums read the staging buffers the next residual tile lands in)");
  VIP_REQUIRE((epi.residual_lo == nullptr && epi.out_lo == nullptr) ||
                  (epi.residual != nullptr && epi.out_bf16 != nullptr && epi.act == ACT_NONE && epi.ln_stats == nullptr &&
                   epi.colscale == nullptr && epi.row_gate == nullptr),
              VIP_ERR_UNSUPPORTED, "gemm: residual_lo / out_lo need a residual, a bf16 output and a plain (bias only) epilogue");
  VIP_REQUIRE((((uintptr_t)epi.residual_lo | (uintptr_t)epi.out_lo) & 15) == 0, VIP_ERR_INVALID, "gemm: unaligned low plane");
  VIP_REQUIRE(epi.row_gate == nullptr ||
                  (epi.gate_rows > 0 && epi.residual != nullptr && epi.act == ACT_RELU && epi.out_bf16 != nullptr &&
                   epi.ln_stats == nullptr && epi.colscale == nullptr && epi.row_stats == nullptr),
              VIP_ERR_UNSUPPORTED, "gemm: row_gate needs gate_rows, a residual, act relu and a bf16 output (SE bottleneck tail)");
  const void* outp = epi.out_bf16 ? (const void*)epi.out_bf16 : (const void*)epi.out_f32;
  VIP_REQUIRE(((uintptr_t)outp & 15) == 0, VIP_ERR_INVALID, "gemm: output must be 16-byte aligned");
  return VIP_OK;
}

int run(const CUtensorMap& tmA, const __nv_bfloat16* B, int ldb, long long M, int N, int K, GemmArgs& g,
        const GemmEpilogue& epi, cudaStream_t stream, const CUtensorMap* tmC_patch = nullptr, int rows_per_group = 0) {
  VIP_REQUIRE(N <= kMaxCols, VIP_ERR_UNSUPPORTED, "gemm: N = %d exceeds %d", N, kMaxCols);
  int rc0 = ensure_neutral(stream);
  if (rc0 != VIP_OK) return rc0;
  // K > 256: the MMA loop dominates (one CTA per SM, deep ring); otherwise the epilogue does (two CTAs per SM).  With a
  // residual the deep configuration is also taken from K > 128 on: its two staging-buffer sets let the producer fetch the
  // next tile's residual a whole tile ahead, which the two-buffer shallow configuration cannot.
  static const int deep_res_kb = [] { const char* v = getenv("VIP_GEMM_DEEP_RES_KB"); return v != nullptr ? atoi(v) : 4; }();
  static const int deep_kb = [] { const char* v = getenv("VIP_GEMM_DEEP_KB"); return v != nullptr ? atoi(v) : 4; }();
  const bool deep = g.num_kb > deep_kb || (epi.residual != nullptr && g.num_kb > deep_res_kb);
  int bn = pick_bn(M, N, deep, g.conv == 1, g.num_kb);
  {
    static const int force_bn = [] { const char* v = getenv("VIP_GEMM_FORCE_BN"); return v != nullptr ? atoi(v) : 0; }();
    if (force_bn == 64 || force_bn == 128 || (force_bn == 256 && deep)) bn = force_bn;   // tile-width experiments
  }
  // pair mode (two CTAs, tcgen05.mma.cta_group::2, 256-row tiles) for the plain deep GEMMs; VIP_GEMM_PAIR=0 disables
  static const bool pair_env = [] { const char* v = getenv("VIP_GEMM_PAIR"); return v == nullptr || v[0] != '0'; }();
  // Measured (profiles/README.md): 76 % of the cuBLAS peak at 8192^3 against 71 % for one CTA per tile, but no gain below
  // K ~ 4096, where the tile prologue and the epilogue dominate: enabled for long K loops only.
  static const int pair_min_kb = [] { const char* v = getenv("VIP_GEMM_PAIR_MIN_KB"); return v != nullptr ? atoi(v) : 64; }();
  const bool pair = pair_env && deep && g.conv == 0 && M >= 2 * BM && N >= 128 && g.num_kb >= pair_min_kb && rows_per_group == 0;
  if (pair && bn < 128) bn = 128;
  long long b_rows = N;
  if (rows_per_group > 0) {   // one [N, K] weight matrix per group of rows_per_group rows of A
    g.bgroup_mtiles = rows_per_group / BM;
    g.bgroup_rows = N;
    b_rows = (long long)N * ((M + rows_per_group - 1) / rows_per_group);
  }
  CUtensorMap tmB, tmC, tmR;
  int rc = make_tmap_2d(&tmB, B, b_rows, K, ldb, pair ? bn / 2 : bn);
  if (rc != VIP_OK) return rc;
  tmC = tmB;
  tmR = tmB;
  if (tmC_patch != nullptr) {
    tmC = *tmC_patch;
  } else if (epi.out_bf16 != nullptr) {
    rc = make_tmap_2d(&tmC, epi.out_bf16, M, N, epi.ldc, BM);
    if (rc != VIP_OK) return rc;
  }
  if (epi.residual != nullptr) {
    rc = make_tmap_2d(&tmR, epi.residual, M, N, epi.ldr, BM);
    if (rc != VIP_OK) return rc;
  }
  g.mode = pick_mode(epi);
  g.trace = nullptr;
  {
    static const int dbg_env = [] { const char* v = getenv("VIP_GEMM_DEBUG"); return v != nullptr ? atoi(v) : 0; }();
    g.dbg = dbg_env;
  }
  static const bool trace_env = [] { const char* v = getenv("VIP_GEMM_TRACE"); return v != nullptr && v[0] == '1'; }();
  long long* trace_dev = nullptr;
  const int trace_tiles = 2048;
  if (trace_env) {
    if (cudaMalloc(&trace_dev, trace_tiles * 16 * sizeof(long long)) == cudaSuccess) {
      cudaMemsetAsync(trace_dev, 0, trace_tiles * 16 * sizeof(long long), stream);
      g.trace = trace_dev;
    }
  }
  g.M = (int)M;
  g.N = N;
  g.K = K;
  g.m_tiles = pair ? (int)((M + 2 * BM - 1) / (2 * BM)) : (int)((M + BM - 1) / BM);
  g.n_tiles = (N + bn - 1) / bn;
  g.epi = epi;
  int lrc;
  if (pair) {
    lrc = bn == 256 ? launch<256, true, true>(tmA, tmB, tmC, tmR, g, stream) : launch<128, true, true>(tmA, tmB, tmC, tmR, g, stream);
  } else if (deep) {
    switch (bn) {
      case 256: lrc = launch<256, true>(tmA, tmB, tmC, tmR, g, stream); break;
      case 128: lrc = launch<128, true>(tmA, tmB, tmC, tmR, g, stream); break;
      default: lrc = launch<64, true>(tmA, tmB, tmC, tmR, g, stream); break;
    }
  } else if (bn == 128) {
    lrc = launch<128, false>(tmA, tmB, tmC, tmR, g, stream);
  } else {
    lrc = launch<64, false>(tmA, tmB, tmC, tmR, g, stream);
  }
  if (trace_dev != nullptr) {
    // experiment aid: timeline of CTA 0 (cycles relative to its first stamp)
    static int printed = 0;
    cudaStreamSynchronize(stream);
    if (printed < 2) {
      ++printed;
      static long long h[2048 * 16];
      cudaMemcpy(h, trace_dev, sizeof(h), cudaMemcpyDeviceToHost);
      const long long t0 = h[0];
      fprintf(stderr, "gemm trace M=%lld N=%d K=%d bn=%d deep=%d pair=%d mode=%d conv=%d: tile | load-issue begin end | mma tmem-free first-full commit | epi begin end\n",
              M, N, K, bn, (int)deep, (int)pair, g.mode, g.conv);
      for (int t = 0; t < 24 && h[t * 16] != 0; ++t) {
        const long long* q = h + t * 16;
        // epilogue thread 0: wait before its first TMEM load, the two loads (issue -> data) and the math after each, the
        // fence, the group barrier, the store issue + statistics tail
        fprintf(stderr, "  %3d | %7lld %7lld | %7lld %7lld %7lld | %7lld %7lld | pre %lld ld0 %lld math0 %lld ld1 %lld math1 %lld fence %lld bar %lld tail %lld\n", t,
                q[0] - t0, q[1] - t0, q[2] - t0, q[3] - t0, q[4] - t0, q[5] - t0, q[6] - t0, q[7] - q[5], q[8] - q[7],
                q[10] - q[8], q[11] - q[10], q[12] - q[11], q[13] - q[12], q[9] - q[13], q[6] - q[9]);
      }
    }
    cudaFree(trace_dev);
  }
  return lrc;
}

}  // namespace

int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
              const GemmEpilogue& epi, cudaStream_t stream, int rows_per_group) {
  VIP_REQUIRE(M > 0 && N > 0 && K > 0, VIP_ERR_INVALID, "gemm_bf16: empty problem %dx%dx%d", M, N, K);
  VIP_REQUIRE(rows_per_group >= 0 && rows_per_group % BM == 0, VIP_ERR_UNSUPPORTED,
              "gemm_bf16: rows_per_group = %d must be a multiple of %d (tiles may not straddle weight groups)", rows_per_group, BM);
  VIP_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, VIP_ERR_UNSUPPORTED,
              "gemm_bf16: K, lda, ldb must be multiples of 8 (16-byte TMA rows): K=%d lda=%d ldb=%d", K, lda, ldb);
  VIP_REQUIRE(N % 8 == 0, VIP_ERR_UNSUPPORTED, "gemm_bf16: N must be a multiple of 8 (N=%d)", N);
  VIP_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, VIP_ERR_INVALID, "gemm_bf16: unaligned operand");
  int rc = check_epilogue(epi, N);
  if (rc != VIP_OK) return rc;
  CUtensorMap tmA;
  rc = make_tmap_2d(&tmA, A, M, K, lda, BM);
  if (rc != VIP_OK) return rc;
  GemmArgs g{};
  g.num_kb = (K + BK - 1) / BK;
  return run(tmA, B, ldb, M, N, K, g, epi, stream, nullptr, rows_per_group);
}

int conv2d_bf16(const __nv_bfloat16* x, const ConvGeom& c, const __nv_bfloat16* w, int ldw, int Cout,
                const GemmEpilogue& epi, cudaStream_t stream) {
  VIP_REQUIRE(c.Nimg > 0 && c.H > 0 && c.W > 0 && c.C > 0 && Cout > 0, VIP_ERR_INVALID, "conv2d_bf16: empty problem");
  VIP_REQUIRE(c.C % 8 == 0 && Cout % 8 == 0 && ldw % 8 == 0, VIP_ERR_UNSUPPORTED,
              "conv2d_bf16: C, Cout, ldw must be multiples of 8 (C=%d Cout=%d ldw=%d)", c.C, Cout, ldw);
  VIP_REQUIRE(c.ksize >= 1 && c.ksize <= 7 && c.stride >= 1 && c.stride <= 8 && c.pad >= 0 && c.pad < c.ksize,
              VIP_ERR_UNSUPPORTED, "conv2d_bf16: kernel %d stride %d pad %d not supported", c.ksize, c.stride, c.pad);
  VIP_REQUIRE(c.Ho == (c.H + 2 * c.pad - c.ksize) / c.stride + 1 && c.Wo == (c.W + 2 * c.pad - c.ksize) / c.stride + 1,
              VIP_ERR_INVALID, "conv2d_bf16: Ho/Wo do not match the geometry");
  VIP_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w & 15) == 0, VIP_ERR_INVALID, "conv2d_bf16: unaligned operand");
  VIP_REQUIRE(c.ldx == 0 || (c.ldx >= c.C && c.ldx % 8 == 0), VIP_ERR_INVALID, "conv2d_bf16: ldx must be a multiple of 8 and >= C");
  int rc = check_epilogue(epi, Cout);
  if (rc != VIP_OK) return rc;
  GemmArgs g{};
  g.cblocks = (c.C + BK - 1) / BK;
  g.C = c.C;
  g.ks = c.ksize;
  g.stride = c.stride;
  g.pad = c.pad;
  g.Wo = c.Wo;
  g.Ho = c.Ho;
  g.Nimg = c.Nimg;
  g.HoWo = c.Ho * c.Wo;
  g.num_kb = c.ksize * c.ksize * g.cblocks;
  const long long M = (long long)c.Nimg * c.Ho * c.Wo;
  VIP_REQUIRE(M < (1LL << 31), VIP_ERR_UNSUPPORTED, "conv2d_bf16: too many output pixels");
  const int K = c.ksize * c.ksize * c.C;
  const bool deep = g.num_kb > 4;

  // Two ways to feed the A operand.  im2col-mode TMA packs 128 consecutive output pixels per tile (no padding rows) but
  // delivers a k-block in ~650 cycles; a tiled 4-D box (a patch of bni x bh x bw output pixels, shifted per filter tap,
  // zero-filled outside the image) arrives ~4x faster but wastes the rows the patch shape cannot fill.  Pick by a
  // simple cost model: tiles x max(load cycles, MMA cycles) per k-block.
  int bw = 0, bh = 0, bni = 0;
  long long patches = 0;
  {
    double best = 0.0;
    for (int w_ = c.Wo < 128 ? c.Wo : 128; w_ >= 1; --w_) {
      if (w_ * c.stride > 256) continue;
      int h_ = 128 / w_;
      if (h_ > c.Ho) h_ = c.Ho;
      if (h_ * c.stride > 256) h_ = 256 / c.stride;
      int n_ = (h_ == c.Ho && w_ == c.Wo) ? 128 / (w_ * h_) : 1;
      if (n_ > c.Nimg) n_ = c.Nimg;
      if (n_ < 1) n_ = 1;
      const long long t = (long long)((c.Wo + w_ - 1) / w_) * ((c.Ho + h_ - 1) / h_) * ((c.Nimg + n_ - 1) / n_);
      const double eff = (double)M / ((double)t * BM);
      if (eff > best + 1e-9) { best = eff; bw = w_; bh = h_; bni = n_; patches = t; }
    }
  }
  const bool patch_ok = epi.out_bf16 != nullptr && epi.residual == nullptr && epi.row_stats == nullptr &&
                        epi.ln_stats == nullptr && epi.row_gate == nullptr && patches > 0 && patches * BM < (1LL << 31);
  // Measured on B200: both feeds run at the same ~650 cycles per k-block (the 9x re-read of the input through L2 is the
  // limit, not the TMA mode), so the patch tiles' padding rows are pure loss; the path stays selectable for experiments
  // (VIP_CONV_PATCH=1) and as the base of a halo-reuse variant.
  static const bool patch_env = [] { const char* v = getenv("VIP_CONV_PATCH"); return v != nullptr && v[0] == '1'; }();
  const bool use_patch = patch_ok && patch_env && c.ldx == 0;
  CUtensorMap tmA;
  if (!use_patch) {
    rc = make_tmap_im2col(&tmA, x, c);
    if (rc != VIP_OK) return rc;
    g.conv = 1;
    return run(tmA, w, ldw, M, Cout, K, g, epi, stream);
  }
  rc = make_tmap_nhwc(&tmA, x, c.Nimg, c.H, c.W, c.C, bw, bh, bni, c.stride);
  if (rc != VIP_OK) return rc;
  CUtensorMap tmC;
  {
    // output [Nimg, Ho, Wo, ldc]: same patch box, unit stride; only the first Cout channels of a row are addressed
    static EncodeTiledFn fn = reinterpret_cast<EncodeTiledFn>(driver_fn("cuTensorMapEncodeTiled"));
    VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t gdim[4] = {(cuuint64_t)Cout, (cuuint64_t)c.Wo, (cuuint64_t)c.Ho, (cuuint64_t)c.Nimg};
    const cuuint64_t gstride[3] = {(cuuint64_t)epi.ldc * 2, (cuuint64_t)c.Wo * epi.ldc * 2, (cuuint64_t)c.Ho * c.Wo * epi.ldc * 2};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bni};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    const CUresult r = fn(&tmC, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, epi.out_bf16, gdim, gstride, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA, "cuTensorMapEncodeTiled (conv output) failed with CUresult %d", (int)r);
  }
  g.conv = 2;
  g.bw = bw;
  g.bh = bh;
  g.bni = bni;
  g.tiles_w = (c.Wo + bw - 1) / bw;
  g.tiles_h = (c.Ho + bh - 1) / bh;
  g.prow = bw * bh * bni;
  return run(tmA, w, ldw, patches * BM, Cout, K, g, epi, stream, &tmC);
}

}  // namespace vip

// ---- C ABI -----------------------------------------------------------------------------------------------------
namespace {
vip::GemmEpilogue to_epilogue(const vip_epilogue_t* p) {
  vip::GemmEpilogue e;
  e.bias = p->bias;
  e.act = p->act;
  e.colscale = p->colscale;
  e.residual = reinterpret_cast<const __nv_bfloat16*>(p->residual);
  e.ldr = p->ldr;
  e.ldc = p->ldc;
  if (p->out_dtype == VIP_DTYPE_BF16) e.out_bf16 = reinterpret_cast<__nv_bfloat16*>(p->out);
  else e.out_f32 = reinterpret_cast<float*>(p->out);
  e.ln_stats = p->ln_stats;
  e.ln_colsum = p->ln_colsum;
  e.row_stats = reinterpret_cast<long long*>(p->row_stats);
  e.gap = reinterpret_cast<long long*>(p->gap);
  e.row_pivot = p->row_pivot;
  e.residual_lo = reinterpret_cast<const __nv_bfloat16*>(p->residual_lo);
  e.out_lo = reinterpret_cast<__nv_bfloat16*>(p->out_lo);
  e.gap_rows = p->gap_rows;
  e.row_gate = p->row_gate;
  e.gate_rows = p->gate_rows;
  return e;
}
}  // namespace

extern "C" int vip_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                             int act, const float* colscale, const void* residual, int ldr, void* out, int ldc,
                             int out_dtype, void* cuda_stream) {
  vip_epilogue_t p{};
  p.bias = bias;
  p.act = act;
  p.colscale = colscale;
  p.residual = residual;
  p.ldr = ldr;
  p.out = out;
  p.ldc = ldc;
  p.out_dtype = out_dtype;
  return vip_gemm_bf16_ex(A, lda, B, ldb, M, N, K, &p, cuda_stream);
}

extern "C" int vip_gemm_bf16_ex(const void* A, int lda, const void* B, int ldb, int M, int N, int K,
                                const vip_epilogue_t* epi, void* cuda_stream) {
  VIP_REQUIRE(A && B && epi && epi->out, VIP_ERR_INVALID, "vip_gemm_bf16_ex: null pointer");
  return vip::gemm_bf16(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb,
                        M, N, K, to_epilogue(epi), reinterpret_cast<cudaStream_t>(cuda_stream));
}

extern "C" int vip_gemm_grouped_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, int rows_per_group,
                                     const vip_epilogue_t* epi, void* cuda_stream) {
  VIP_REQUIRE(A && B && epi && epi->out, VIP_ERR_INVALID, "vip_gemm_grouped_bf16: null pointer");
  VIP_REQUIRE(rows_per_group > 0, VIP_ERR_INVALID, "vip_gemm_grouped_bf16: rows_per_group must be positive");
  return vip::gemm_bf16(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb,
                        M, N, K, to_epilogue(epi), reinterpret_cast<cudaStream_t>(cuda_stream), rows_per_group);
}

extern "C" int vip_conv2d_slice_bf16(const void* x, int N, int H, int W, int C, int ldx, const void* w, int ldw, int Cout,
                                     int ksize, int stride, int pad, const vip_epilogue_t* epi, void* cuda_stream) {
  VIP_REQUIRE(x && w && epi && epi->out, VIP_ERR_INVALID, "vip_conv2d_slice_bf16: null pointer");
  VIP_REQUIRE(stride > 0 && ksize > 0, VIP_ERR_INVALID, "vip_conv2d_slice_bf16: bad kernel / stride");
  vip::ConvGeom c{N, H, W, C, ksize, stride, pad, 0, 0, ldx};
  c.Ho = (H + 2 * pad - ksize) / stride + 1;
  c.Wo = (W + 2 * pad - ksize) / stride + 1;
  return vip::conv2d_bf16(reinterpret_cast<const __nv_bfloat16*>(x), c, reinterpret_cast<const __nv_bfloat16*>(w), ldw,
                          Cout, to_epilogue(epi), reinterpret_cast<cudaStream_t>(cuda_stream));
}

extern "C" int vip_conv2d_bf16(const void* x, int N, int H, int W, int C, const void* w, int ldw, int Cout, int ksize,
                               int stride, int pad, const vip_epilogue_t* epi, void* cuda_stream) {
  VIP_REQUIRE(x && w && epi && epi->out, VIP_ERR_INVALID, "vip_conv2d_bf16: null pointer");
  vip::ConvGeom c{N, H, W, C, ksize, stride, pad, 0, 0};
  VIP_REQUIRE(stride > 0 && ksize > 0, VIP_ERR_INVALID, "vip_conv2d_bf16: bad kernel / stride");
  c.Ho = (H + 2 * pad - ksize) / stride + 1;
  c.Wo = (W + 2 * pad - ksize) / stride + 1;
  return vip::conv2d_bf16(reinterpret_cast<const __nv_bfloat16*>(x), c, reinterpret_cast<const __nv_bfloat16*>(w), ldw,
                          Cout, to_epilogue(epi), reinterpret_cast<cudaStream_t>(cuda_stream));
}
