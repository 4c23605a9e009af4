// tcgen05 / TMEM / TMA GEMM for sm_100a.  See gemm.cuh for the contract.
//
// One CTA computes one 128 x BN output tile.  Warp roles (192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor.2d (128B swizzle) of A[128 x 64] and B[BN x 64] k-blocks into a
//               kStages-deep shared-memory ring, completion on "full" mbarriers
//   warp 1      TMEM allocation + MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//               (M=128, N=BN, K=16) x4 per k-block, tcgen05.commit releases the ring slot ("empty" mbarrier) and,
//               after the last k-block, signals the epilogue ("tmem_full")
//   warps 2..5  epilogue: tcgen05.ld 32 lanes x 32 columns -> registers -> bias / activation / layer-scale /
//               residual -> bf16 or fp32 global stores (each thread owns one output row)
#include <cuda.h>

#include "gemm.cuh"

namespace vip {
namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int kGemmThreads = 192;

struct GemmArgs {
  int M, N, K;
  GemmEpilogue epi;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
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

__device__ __forceinline__ float apply_act(float v, int act) {
  if (act == ACT_RELU) return fmaxf(v, 0.0f);
  if (act == ACT_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));  // Keras 'gelu' = exact erf form
  if (act == ACT_SIGMOID) return 1.0f / (1.0f + __expf(-v));
  return v;
}

template <int BN>
struct TmemCols {
  static constexpr int value = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
};

template <int BN, int kStages>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmArgs g) {
  constexpr int kABytes = BM * BK * 2;
  constexpr int kBBytes = BN * BK * 2;
  constexpr int kStageBytes = kABytes + kBBytes;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full_bar = empty_bar + kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int num_kb = (g.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "n"(TmemCols<BN>::value)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        mbar_expect_tx(&full_bar[s], kStageBytes);
        uint8_t* a_dst = smem + s * kStageBytes;
        tma_load_2d(a_dst, &tmA, &full_bar[s], kb * BK, m0);
        tma_load_2d(a_dst + kABytes, &tmB, &full_bar[s], kb * BK, n0);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // instruction descriptor: D fp32, A/B bf16, both K-major, N = BN, M = 128
      constexpr uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tcgen05_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * kStageBytes);
        const uint64_t a_desc = make_sw128_desc(a_addr);
        const uint64_t b_desc = make_sw128_desc(a_addr + kABytes);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k) {
          // advance 16 elements = 32 bytes inside the swizzle atom: +2 in the (addr >> 4) field
          umma_bf16(tmem_base, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // implies tcgen05.fence::before_thread_sync
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // epilogue: warp w may touch TMEM lanes 32*(w%4) .. +31
    const int q = warp & 3;
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    const int row = m0 + q * 32 + lane;
    const GemmEpilogue& e = g.epi;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, r);
      if (row < g.M) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const int n = n0 + c;
        if (e.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += __ldg(e.bias + n + j);
        }
        if (e.act != ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], e.act);
        }
        if (e.colscale != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= __ldg(e.colscale + n + j);
        }
        if (e.residual != nullptr) {
          const uint4* rp = reinterpret_cast<const uint4*>(e.residual + (size_t)row * e.ldr + n);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const uint4 u = __ldg(rp + j4);
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              v[j4 * 8 + 2 * t] += __uint_as_float(w[t] << 16);
              v[j4 * 8 + 2 * t + 1] += __uint_as_float(w[t] & 0xffff0000u);
            }
          }
        }
        if (e.out_bf16 != nullptr) {
          uint4* op = reinterpret_cast<uint4*>(e.out_bf16 + (size_t)row * e.ldc + n);
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            uint32_t w[4];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const __nv_bfloat162 h = __floats2bfloat162_rn(v[j4 * 8 + 2 * t], v[j4 * 8 + 2 * t + 1]);
              w[t] = *reinterpret_cast<const uint32_t*>(&h);
            }
            op[j4] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        } else {
          float4* op = reinterpret_cast<float4*>(e.out_f32 + (size_t)row * e.ldc + n);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) op[j4] = make_float4(v[4 * j4], v[4 * j4 + 1], v[4 * j4 + 2], v[4 * j4 + 3]);
        }
      }
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TmemCols<BN>::value) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows, cols] bf16 row-major with `ld` elements between rows; box = [box_rows, 64 cols], 128B swizzle, zero OOB fill
int make_tmap_2d(CUtensorMap* tm, const void* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  VIP_REQUIRE(fn != nullptr, VIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VIP_REQUIRE(r == CUDA_SUCCESS, VIP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rows=%d cols=%d ld=%d)",
              (int)r, rows, cols, ld);
  return VIP_OK;
}

template <int BN, int kStages>
int launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& g, cudaStream_t st) {
  constexpr int smem = kStages * (BM * BK * 2 + BN * BK * 2) + (2 * kStages + 1) * 8 + 16 + 1024;
  auto kern = gemm_tcgen05_kernel<BN, kStages>;
  static bool configured = false;  // per-process; attribute is per-function
  if (!configured) {
    VIP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  dim3 grid((g.M + BM - 1) / BM, g.N / BN);
  kern<<<grid, kGemmThreads, smem, st>>>(tmA, tmB, g);
  VIP_CUDA(cudaGetLastError());
  count_launch();
  return VIP_OK;
}

}  // namespace

int gemm_bf16(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, int M, int N, int K,
              const GemmEpilogue& epi, cudaStream_t stream) {
  VIP_REQUIRE(M > 0 && N > 0 && K > 0, VIP_ERR_INVALID, "gemm_bf16: empty problem %dx%dx%d", M, N, K);
  VIP_REQUIRE(K % 8 == 0 && lda % 8 == 0 && ldb % 8 == 0, VIP_ERR_UNSUPPORTED,
              "gemm_bf16: K, lda, ldb must be multiples of 8 (16-byte TMA rows): K=%d lda=%d ldb=%d", K, lda, ldb);
  VIP_REQUIRE(N % 32 == 0, VIP_ERR_UNSUPPORTED, "gemm_bf16: N must be a multiple of 32 (N=%d)", N);
  VIP_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, VIP_ERR_INVALID, "gemm_bf16: unaligned operand");
  VIP_REQUIRE((epi.out_bf16 != nullptr) != (epi.out_f32 != nullptr), VIP_ERR_INVALID,
              "gemm_bf16: exactly one of out_bf16 / out_f32 must be set");
  VIP_REQUIRE(epi.ldc % 8 == 0 && (epi.residual == nullptr || epi.ldr % 8 == 0), VIP_ERR_UNSUPPORTED,
              "gemm_bf16: ldc / ldr must be multiples of 8");
  // tile width: widest of 256 / 128 / 64 / 32 that divides N and still fills the machine reasonably
  const int mt = (M + BM - 1) / BM;
  int bn = 32;
  for (int cand : {256, 128, 64, 32}) {
    if (N % cand == 0 && (cand <= 64 || (long long)mt * (N / cand) >= 120 || cand == 32)) { bn = cand; break; }
  }
  if (N % bn != 0) bn = 32;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, A, M, K, lda, BM);
  if (rc != VIP_OK) return rc;
  rc = make_tmap_2d(&tmB, B, N, K, ldb, bn);
  if (rc != VIP_OK) return rc;
  GemmArgs g{M, N, K, epi};
  switch (bn) {
    case 256: return launch<256, 4>(tmA, tmB, g, stream);
    case 128: return launch<128, 6>(tmA, tmB, g, stream);
    case 64: return launch<64, 8>(tmA, tmB, g, stream);
    default: return launch<32, 8>(tmA, tmB, g, stream);
  }
}

}  // namespace vip

// C-ABI test / utility entry point (also usable by a host that wants the raw contraction).
extern "C" int vip_gemm_bf16(const void* A, int lda, const void* B, int ldb, int M, int N, int K, const float* bias,
                             int act, const float* colscale, const void* residual, int ldr, void* out, int ldc,
                             int out_dtype, void* cuda_stream) {
  vip::GemmEpilogue e;
  e.bias = bias;
  e.act = act;
  e.colscale = colscale;
  e.residual = reinterpret_cast<const __nv_bfloat16*>(residual);
  e.ldr = ldr;
  e.ldc = ldc;
  if (out_dtype == VIP_DTYPE_BF16) e.out_bf16 = reinterpret_cast<__nv_bfloat16*>(out);
  else e.out_f32 = reinterpret_cast<float*>(out);
  return vip::gemm_bf16(reinterpret_cast<const __nv_bfloat16*>(A), lda, reinterpret_cast<const __nv_bfloat16*>(B), ldb,
                        M, N, K, e, reinterpret_cast<cudaStream_t>(cuda_stream));
}
