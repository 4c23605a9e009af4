// Window attention of GCViT (models/gcvit/layers/attention.py:52-83) with window_partition / window_reverse
// (models/gcvit/layers/window.py:3-14) folded into the indexing: tokens are read from and written to the image-order
// [B, H, W, C] layout directly.  out = softmax(q * hd^-0.5 @ k^T + rel_bias[h]) @ v  per (window, head), head_dim 32.
//   local block : q, k, v = qkv[..., 0:C], [C:2C], [2C:3C]
//   global block: k, v = qkv[..., 0:C], [C:2C];  q = q_global[b] (shared by all windows of image b, attention.py:62-66)
// The relative position bias is gathered on the fly from the compact table [(2ws-1)^2] of the head
// (attention.py:39-50,71-75): index = (yi - yj + ws - 1) * (2ws - 1) + (xi - xj + ws - 1).
//
// v2: tensor-core kernel.  The problems are tiny (49x49x32 or 196x196x32 per (window, head)), so the 128-row tcgen05
// atom would waste more than half of its rows and P would have to round-trip shared memory; the kernel is bounded by
// HBM traffic (read 3C + write C per token) and by the exp throughput of the softmax, not by the tensor pipe.  It uses
// warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate) with S, P and O kept in registers (P is re-used as the A
// fragment of the second product without leaving the register file):
//   stage   q / k / v rows of the (window, head) pair -> shared memory with 16-byte cp.async (zero-filled padding rows)
//   S       16 query rows x all keys per warp step, fp32
//   softmax scale, + bias (table lookup), row max / sum across the 4 lanes that share a row, exp2
//   O       P (bf16 A fragments built from the S accumulators) x V (ldmatrix.trans), / row sum
//   store   through the (consumed) q rows in shared memory -> coalesced 16-byte stores
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"

namespace vip {
namespace {

using bf16 = __nv_bfloat16;
constexpr int HD = 32;
constexpr int PITCH = 40;  // bf16 per staged row (80 B): conflict-free fragment loads and ldmatrix rows

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
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
struct AttnCfg {
  static constexpr int N = WS * WS;                 // tokens per window
  static constexpr int WP = WS <= 8 ? 8 : 16;       // key rows padded to a whole number of 8-key tiles: the window row of
                                                    // a key tile is then a compile-time constant (cheap bias addressing)
  static constexpr int KEYS = WS * WP;              // padded keys (56 / 224)
  static constexpr int KT = (KEYS + 15) / 16;       // 16-key steps of P @ V
  static constexpr int NT = 2 * KT;                 // 8-key tiles of S
  static constexpr int KROWS = KT * 16;             // staged key / value rows (zero padded)
  static constexpr int QT = (N + 15) / 16;          // 16-row query tiles
  static constexpr int QROWS = QT * 16;
  static constexpr int NB = WS <= 8 ? 1 : 2;        // key blocks of the online softmax (bounds the S registers)
  static constexpr int NTB = NT / NB, KTB = KT / NB;
  static constexpr int TAB = (2 * WS - 1) * (2 * WS - 1);
  static constexpr int TPAD = 4;                    // padded keys index up to 2 entries before the table
  static constexpr int kWarpsPerPair = WS <= 8 ? 1 : 4;
  static constexpr int kPairsPerCta = 4 / kWarpsPerPair;
  static constexpr int kPairBytes = (QROWS + 2 * KROWS) * PITCH * 2 + (((TAB + TPAD) * 4 + 15) / 16) * 16;
  static constexpr int kSmem = kPairsPerCta * kPairBytes;
  static_assert(NT % NB == 0 && KT % NB == 0, "key blocks must split evenly");
};

template <int WS>
__global__ void __launch_bounds__(128, WS <= 8 ? 3 : 4)
window_attention_mma_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ qg, const float* __restrict__ table /*[heads][TAB]*/,
                            bf16* __restrict__ out, int H, int W, int C, int heads, int num_pairs, float scale_log2e) {
  using Cfg = AttnCfg<WS>;
  constexpr int N = Cfg::N, WP = Cfg::WP, KT = Cfg::KT, KROWS = Cfg::KROWS, QT = Cfg::QT, QROWS = Cfg::QROWS;
  constexpr int NB = Cfg::NB, NTB = Cfg::NTB, KTB = Cfg::KTB, TAB = Cfg::TAB, TPAD = Cfg::TPAD;
  constexpr int WPP = Cfg::kWarpsPerPair;
  extern __shared__ __align__(16) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair_in_cta = warp / WPP, w_in_pair = warp % WPP;
  const int ltid = w_in_pair * 32 + lane;  // thread index inside the pair's group
  constexpr int GT = WPP * 32;
  const int pair = blockIdx.x * Cfg::kPairsPerCta + pair_in_cta;
  const bool live = pair < num_pairs;      // dead groups still take part in the barriers
  const int p = live ? pair : num_pairs - 1;
  const int h = p % heads, win = p / heads;
  const int nWw = W / WS, nWh = H / WS;
  const int b = win / (nWh * nWw);
  const int wrem = win - b * nWh * nWw;
  const int wy = wrem / nWw, wx = wrem - wy * nWw;
  const bool global_q = qg != nullptr;
  const int ldq = (global_q ? 2 : 3) * C;
  const int koff = (global_q ? 0 : C) + h * HD, voff = koff + C;

  bf16* sQ = reinterpret_cast<bf16*>(smem + pair_in_cta * Cfg::kPairBytes);
  bf16* sK = sQ + QROWS * PITCH;
  bf16* sV = sK + KROWS * PITCH;
  float* sT = reinterpret_cast<float*>(sV + KROWS * PITCH) + TPAD;

  // ---- stage q (token order), k / v (padded key order: row = window row * WP + window column; padding rows are zero)
  //      and the bias table (pre-multiplied by log2 e); 16-byte chunks, 4 per 32-channel head slice
  for (int i = ltid; i < (QROWS + 2 * KROWS) * 4; i += GT) {
    const int r = i >> 2, ch = i & 3;
    bool valid;
    int tok;
    const bf16* src;
    if (r < QROWS) {
      valid = r < N;
      tok = valid ? r : 0;
    } else {
      const int pr = (r - QROWS) % KROWS;
      const int ky = pr / WP, kx = pr % WP;
      valid = ky < WS && kx < WS;
      tok = valid ? ky * WS + kx : 0;
    }
    const int y = wy * WS + tok / WS, x = wx * WS + tok % WS;
    const long long row = ((long long)b * H + y) * W + x;
    if (r < QROWS) src = global_q ? qg + ((long long)b * N + tok) * C + h * HD : qkv + row * ldq + h * HD;
    else src = qkv + row * ldq + (r < QROWS + KROWS ? koff : voff);
    cp_async16_zfill(sQ + r * PITCH + ch * 8, src + ch * 8, valid);
  }
  for (int i = ltid; i < TAB; i += GT) sT[i] = __ldg(table + (long long)h * TAB + i) * 1.4426950408889634f;
  asm volatile("cp.async.wait_all;" ::: "memory");
  if (WPP == 1) __syncwarp();
  else __syncthreads();

  const int g = lane >> 2, t = lane & 3;
  for (int mt = w_in_pair; mt < QT; mt += WPP) {
    uint32_t qa[2][4];
    {
      const uint32_t* q0 = reinterpret_cast<const uint32_t*>(sQ + (mt * 16 + g) * PITCH);
      const uint32_t* q1 = reinterpret_cast<const uint32_t*>(sQ + (mt * 16 + g + 8) * PITCH);
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        qa[ks][0] = q0[ks * 8 + t];
        qa[ks][1] = q1[ks * 8 + t];
        qa[ks][2] = q0[ks * 8 + 4 + t];
        qa[ks][3] = q1[ks * 8 + 4 + t];
      }
    }
    // bias rows of the two query rows this lane holds (r0 = mt*16+g, r1 = r0+8): sT[qo - ky*(2WS-1) - kx]
    const int r0 = min(mt * 16 + g, N - 1), r1 = min(mt * 16 + g + 8, N - 1);
    const float* pb0 = sT + (r0 / WS + WS - 1) * (2 * WS - 1) + r0 % WS + WS - 1 - 2 * t;
    const float* pb1 = sT + (r1 / WS + WS - 1) * (2 * WS - 1) + r1 % WS + WS - 1 - 2 * t;
    float m0 = -3.0e38f, m1 = -3.0e38f, l0 = 0.0f, l1 = 0.0f;
    float o[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.0f;
#pragma unroll
    for (int blk = 0; blk < NB; ++blk) {
      // ---- S = Q K^T for 16 query rows x NTB key tiles
      float s[NTB][4];
#pragma unroll
      for (int i = 0; i < NTB; ++i) {
        const int nt = blk * NTB + i;
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.0f;
        if ((nt * 8) / WP < WS) {  // a tile of padding keys only needs no product
          const uint32_t* kr = reinterpret_cast<const uint32_t*>(sK + (nt * 8 + g) * PITCH);
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) mma16816(s[i], qa[ks], kr[ks * 8 + t], kr[ks * 8 + 4 + t]);
        }
      }
      // ---- scale + relative position bias, mask the padding keys, block maximum
      float mb0 = -3.0e38f, mb1 = -3.0e38f;
#pragma unroll
      for (int i = 0; i < NTB; ++i) {
        const int nt = blk * NTB + i;
        const int ky = (nt * 8) / WP, kx0 = (nt * 8) % WP;  // compile-time
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          // kx = kx0 + 2t + e is a real window column iff < WS
          const bool real = ky < WS && (kx0 + 7 < WS || kx0 + 2 * t + e < WS);
          if (real) {
            const int off = -(ky * (2 * WS - 1) + kx0 + e);
            s[i][e] = fmaf(s[i][e], scale_log2e, pb0[off]);
            s[i][2 + e] = fmaf(s[i][2 + e], scale_log2e, pb1[off]);
          } else {
            s[i][e] = -3.0e38f;
            s[i][2 + e] = -3.0e38f;
          }
          mb0 = fmaxf(mb0, s[i][e]);
          mb1 = fmaxf(mb1, s[i][2 + e]);
        }
      }
      mb0 = fmaxf(mb0, __shfl_xor_sync(0xffffffffu, mb0, 1));
      mb0 = fmaxf(mb0, __shfl_xor_sync(0xffffffffu, mb0, 2));
      mb1 = fmaxf(mb1, __shfl_xor_sync(0xffffffffu, mb1, 1));
      mb1 = fmaxf(mb1, __shfl_xor_sync(0xffffffffu, mb1, 2));
      const float mn0 = fmaxf(m0, mb0), mn1 = fmaxf(m1, mb1);
      if (NB > 1) {  // online softmax: rescale what the earlier key blocks contributed
        const float c0 = fast_exp2(m0 - mn0), c1 = fast_exp2(m1 - mn1);
        l0 *= c0;
        l1 *= c1;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          o[nt][0] *= c0;
          o[nt][1] *= c0;
          o[nt][2] *= c1;
          o[nt][3] *= c1;
        }
      }
      m0 = mn0;
      m1 = mn1;
#pragma unroll
      for (int i = 0; i < NTB; ++i) {
        s[i][0] = fast_exp2(s[i][0] - m0);
        s[i][1] = fast_exp2(s[i][1] - m0);
        s[i][2] = fast_exp2(s[i][2] - m1);
        s[i][3] = fast_exp2(s[i][3] - m1);
        l0 += s[i][0] + s[i][1];
        l1 += s[i][2] + s[i][3];
      }
      // ---- O += P V
#pragma unroll
      for (int kk = 0; kk < KTB; ++kk) {
        uint32_t pa[4];
        pa[0] = pack_bf16(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack_bf16(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack_bf16(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack_bf16(s[2 * kk + 1][2], s[2 * kk + 1][3]);
        const int mj = lane >> 3;  // which 8x8 matrix this lane addresses
        const bf16* vrow = sV + ((blk * KTB + kk) * 16 + (mj & 1) * 8 + (lane & 7)) * PITCH + (mj >> 1) * 8;
#pragma unroll
        for (int dp = 0; dp < 2; ++dp) {
          uint32_t vb[4];
          ldmatrix_x4_trans(vb, vrow + dp * 16);
          mma16816(o[2 * dp], pa, vb[0], vb[1]);
          mma16816(o[2 * dp + 1], pa, vb[2], vb[3]);
        }
      }
    }
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    // ---- normalise, stage through the consumed q rows, coalesced store
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __syncwarp();
    {
      uint32_t* o0 = reinterpret_cast<uint32_t*>(sQ + (mt * 16 + g) * PITCH);
      uint32_t* o1 = reinterpret_cast<uint32_t*>(sQ + (mt * 16 + g + 8) * PITCH);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        o0[nt * 4 + t] = pack_bf16(o[nt][0] * i0, o[nt][1] * i0);
        o1[nt * 4 + t] = pack_bf16(o[nt][2] * i1, o[nt][3] * i1);
      }
    }
    __syncwarp();
    if (live) {
#pragma unroll
      for (int c = lane; c < 64; c += 32) {
        const int rl = c >> 2, ch = c & 3;
        const int tok = mt * 16 + rl;
        if (tok < N) {
          const int y = wy * WS + tok / WS, x = wx * WS + tok % WS;
          const long long row = ((long long)b * H + y) * W + x;
          *reinterpret_cast<uint4*>(out + row * C + h * HD + ch * 8) =
              *reinterpret_cast<const uint4*>(sQ + (mt * 16 + rl) * PITCH + ch * 8);
        }
      }
    }
  }
}

template <int WS>
int launch_attention(const bf16* qkv, const bf16* qg, const float* table, bf16* out, int B, int H, int W, int C, int heads,
                     cudaStream_t st) {
  using Cfg = AttnCfg<WS>;
  auto kern = window_attention_mma_kernel<WS>;
  static bool configured = false;
  if (!configured) {
    VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
    configured = true;
  }
  const int num_pairs = B * (H / WS) * (W / WS) * heads;
  const int grid = (num_pairs + Cfg::kPairsPerCta - 1) / Cfg::kPairsPerCta;
  kern<<<grid, 128, Cfg::kSmem, st>>>(qkv, qg, table, out, H, W, C, heads, num_pairs,
                                      1.4426950408889634f / sqrtf((float)HD));
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

}  // namespace
}  // namespace vip

namespace vip {
int window_attention_ws(const void* qkv, const void* qg, const float* table, void* out, int B, int H, int W, int C, int ws,
                        int heads, cudaStream_t st);
}

extern "C" int vip_window_attention_bf16(const void* qkv, const void* q_global, const float* rel_table, void* out, int B,
                                         int H, int W, int C, int ws, int heads, void* stream) {
  using namespace vip;
  VIP_REQUIRE(qkv && rel_table && out, VIP_ERR_INVALID, "vip_window_attention_bf16: null pointer");
  VIP_REQUIRE(B > 0 && heads > 0, VIP_ERR_INVALID, "vip_window_attention_bf16: empty problem");
  VIP_REQUIRE(C == heads * HD, VIP_ERR_UNSUPPORTED, "vip_window_attention_bf16: head_dim must be 32 (C=%d heads=%d)", C,
              heads);
  VIP_REQUIRE(H % ws == 0 && W % ws == 0, VIP_ERR_INVALID, "vip_window_attention_bf16: H, W must be multiples of ws");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // Two implementations, same results; VIP_ATTN_IMPL = ws (default) | mma.  The persistent warp-specialised tcgen05 kernel
  // (attention_ws.cu) is the product path; the warp-level mma.sync kernel of this file runs at the legacy-HMMA roofline of
  // sm_100 (~145 TFLOP/s chip-wide) and serves the shapes whose bias tables do not fit next to the operand ring in shared
  // memory (ws 14 with more than ~40 heads; tests/test_attention_gpu.py drives that trigger).
  static const int impl = [] {
    const char* v = getenv("VIP_ATTN_IMPL");
    return (v != nullptr && v[0] == 'm') ? 0 : 2;
  }();
  if (impl == 2 && (ws == 7 || ws == 14)) {
    const int rc = window_attention_ws(qkv, q_global, rel_table, out, B, H, W, C, ws, heads, st);
    if (rc != VIP_ERR_UNSUPPORTED) return rc;   // else: bias tables too large for shared memory -> mma.sync kernel
  }
  if (ws == 7)
    return launch_attention<7>((const bf16*)qkv, (const bf16*)q_global, rel_table, (bf16*)out, B, H, W, C, heads, st);
  if (ws == 14)
    return launch_attention<14>((const bf16*)qkv, (const bf16*)q_global, rel_table, (bf16*)out, B, H, W, C, heads, st);
  VIP_REQUIRE(false, VIP_ERR_UNSUPPORTED, "vip_window_attention_bf16: window size %d not built (7 and 14 are)", ws);
  return VIP_ERR_UNSUPPORTED;
}
