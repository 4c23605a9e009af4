// Host-buffer entry points: chunked H2D -> kernel -> D2H pipelines over two internal streams so that PCIe copies
// in both directions overlap the preprocessing kernel.  This is what stands where dataset.build_dataset
// (dataset/dataset.py:64-102) hands batches to the caller.
#include <algorithm>

#include "common.cuh"

namespace vip {
namespace {

struct HostPipe {
  int device = -1;
  cudaStream_t stream[2] = {nullptr, nullptr};
  uint8_t* d_src[2] = {nullptr, nullptr};
  uint8_t* d_dst[2] = {nullptr, nullptr};
  int32_t* d_crop[2] = {nullptr, nullptr};
  int32_t* d_q[2] = {nullptr, nullptr};
  uint8_t* d_flags[2] = {nullptr, nullptr};
  size_t src_cap = 0, dst_cap = 0;
  int aux_cap = 0;
};

thread_local HostPipe g_pipe;

int ensure_pipe(size_t src_bytes, size_t dst_bytes, int chunk) {
  int dev = 0;
  VIP_CUDA(cudaGetDevice(&dev));
  HostPipe& p = g_pipe;
  if (p.device != dev) {
    p = HostPipe();  // buffers of another device are left to that device's context teardown
    p.device = dev;
    for (int i = 0; i < 2; ++i) VIP_CUDA(cudaStreamCreateWithFlags(&p.stream[i], cudaStreamNonBlocking));
  }
  for (int i = 0; i < 2; ++i) {
    if (p.src_cap < src_bytes) {
      if (p.d_src[i]) VIP_CUDA(cudaFree(p.d_src[i]));
      VIP_CUDA(cudaMalloc(&p.d_src[i], src_bytes));
    }
    if (p.dst_cap < dst_bytes) {
      if (p.d_dst[i]) VIP_CUDA(cudaFree(p.d_dst[i]));
      VIP_CUDA(cudaMalloc(&p.d_dst[i], dst_bytes));
    }
    if (p.aux_cap < chunk) {
      if (p.d_crop[i]) VIP_CUDA(cudaFree(p.d_crop[i]));
      if (p.d_q[i]) VIP_CUDA(cudaFree(p.d_q[i]));
      if (p.d_flags[i]) VIP_CUDA(cudaFree(p.d_flags[i]));
      VIP_CUDA(cudaMalloc(&p.d_crop[i], sizeof(int32_t) * 4 * chunk));
      VIP_CUDA(cudaMalloc(&p.d_q[i], sizeof(int32_t) * chunk));
      VIP_CUDA(cudaMalloc(&p.d_flags[i], chunk));
    }
  }
  p.src_cap = std::max(p.src_cap, src_bytes);
  p.dst_cap = std::max(p.dst_cap, dst_bytes);
  p.aux_cap = std::max(p.aux_cap, chunk);
  return VIP_OK;
}

}  // namespace
}  // namespace vip

extern "C" int vip_preprocess_host(const uint8_t* src, int N, int Hs, int Ws, const int32_t* crop_yxhw,
                                   const int32_t* jpeg_q, const uint8_t* flags, int Ho, int Wo, void* dst,
                                   int dst_dtype) {
  using namespace vip;
  VIP_REQUIRE(N >= 0, VIP_ERR_INVALID, "vip_preprocess_host: N < 0");
  if (N == 0) return VIP_OK;
  VIP_REQUIRE(src != nullptr && dst != nullptr, VIP_ERR_INVALID, "vip_preprocess_host: null src/dst");
  VIP_REQUIRE(dst_dtype == VIP_DTYPE_F32 || dst_dtype == VIP_DTYPE_BF16, VIP_ERR_INVALID,
              "vip_preprocess_host: bad dst_dtype");
  VIP_REQUIRE(Hs >= 1 && Ws >= 1 && Ho >= 1 && Wo >= 1, VIP_ERR_INVALID, "vip_preprocess_host: empty image");
  const size_t src_img = (size_t)Hs * Ws * 3;
  size_t dst_img = (size_t)Ho * Wo * 3 * (dst_dtype == VIP_DTYPE_BF16 ? 2 : 4);
  // chunk: about 64 MiB of output per stage, image count kept even so every chunk base stays 8-byte aligned
  int chunk = (int)std::max<size_t>(2, ((size_t)64 << 20) / dst_img);
  chunk = std::min(chunk & ~1, std::max(2, (N + 1) & ~1));
  int rc = ensure_pipe(src_img * chunk, dst_img * chunk, chunk);
  if (rc != VIP_OK) return rc;
  HostPipe& p = g_pipe;
  int slot = 0;
  for (int i0 = 0; i0 < N; i0 += chunk, slot ^= 1) {
    const int n = std::min(chunk, N - i0);
    cudaStream_t st = p.stream[slot];
    VIP_CUDA(cudaMemcpyAsync(p.d_src[slot], src + (size_t)i0 * src_img, src_img * n, cudaMemcpyHostToDevice, st));
    if (crop_yxhw)
      VIP_CUDA(cudaMemcpyAsync(p.d_crop[slot], crop_yxhw + 4 * (size_t)i0, sizeof(int32_t) * 4 * n,
                               cudaMemcpyHostToDevice, st));
    if (jpeg_q)
      VIP_CUDA(cudaMemcpyAsync(p.d_q[slot], jpeg_q + i0, sizeof(int32_t) * n, cudaMemcpyHostToDevice, st));
    if (flags) VIP_CUDA(cudaMemcpyAsync(p.d_flags[slot], flags + i0, n, cudaMemcpyHostToDevice, st));
    rc = vip_preprocess(p.d_src[slot], n, Hs, Ws, crop_yxhw ? p.d_crop[slot] : nullptr,
                        jpeg_q ? p.d_q[slot] : nullptr, flags ? p.d_flags[slot] : nullptr, Ho, Wo, p.d_dst[slot],
                        dst_dtype, st);
    if (rc != VIP_OK) return rc;
    VIP_CUDA(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(dst) + (size_t)i0 * dst_img, p.d_dst[slot], dst_img * n,
                             cudaMemcpyDeviceToHost, st));
  }
  VIP_CUDA(cudaStreamSynchronize(p.stream[0]));
  VIP_CUDA(cudaStreamSynchronize(p.stream[1]));
  return VIP_OK;
}
