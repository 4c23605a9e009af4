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

}  // namespace vip
