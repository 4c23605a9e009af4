// Shared host/device helpers for libvipcup.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vipcup.h"

namespace vip {

// thread-local error text behind vip_last_error()
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
void count_launch(int n = 1);

#define VIP_CUDA(call)                                             \
  do {                                                             \
    cudaError_t e_ = (call);                                       \
    if (e_ != cudaSuccess) return ::vip::cuda_fail(e_, #call);     \
  } while (0)

#define VIP_REQUIRE(cond, code, ...)  \
  do {                                \
    if (!(cond)) {                    \
      ::vip::set_error(__VA_ARGS__);  \
      return (code);                  \
    }                                 \
  } while (0)

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Programmatic dependent launch (opt-in, VIP_PDL=1): kernels are launched with the programmatic-stream-serialization
// attribute, call pdl_trigger() first thing (the next kernel of the stream may be scheduled as soon as every CTA of this
// one has started) and pdl_wait() before they touch global memory that an earlier kernel may still read or write (the
// wait returns when the preceding kernel has completed and flushed), so that launch latency, CTA start-up and a kernel's
// private set-up overlap the tail of its predecessor.  Measured on the bench workload it LOSES 2.3 ms per 82 ms step
// (early-resident CTAs of the next kernel compete with the tail of the running one), so the attribute is off unless
// asked for; without it the two device calls are no-ops.
bool pdl_enabled();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define VIP_LAUNCH(kernel, grid, block, smem, st, ...) \
  VIP_CUDA(::vip::launch_pdl(kernel, dim3(grid), dim3(block), (size_t)(smem), st, __VA_ARGS__))
#endif

}  // namespace vip
