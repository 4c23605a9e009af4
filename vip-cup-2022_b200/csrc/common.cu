// Error plumbing and library-level entry points of libvipcup.so.
#include <stdlib.h>
#include <stdarg.h>
#include <stdio.h>

#include "common.cuh"

namespace vip {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
  return VIP_ERR_CUDA;
}

bool pdl_enabled() {
  // off by default: measured on B200 (bench.py, 457 launches per step) 81.8 ms without and 84.1 ms with the attribute
  static const bool on = [] { const char* v = getenv("VIP_PDL"); return v != nullptr && v[0] == '1'; }();
  return on;
}

void count_launch(int n) { g_launches += n; }

}  // namespace vip

extern "C" {

const char* vip_version(void) { return "vipcup-b200 0.1 (sm_100a)"; }
const char* vip_last_error(void) { return vip::g_err; }
int64_t vip_launch_count(void) { return vip::g_launches; }
void vip_launch_count_reset(void) { vip::g_launches = 0; }

int vip_memset_async(void* ptr, int value, size_t bytes, void* cuda_stream) {
  VIP_REQUIRE(ptr != nullptr || bytes == 0, VIP_ERR_INVALID, "vip_memset_async: null pointer");
  if (bytes == 0) return VIP_OK;
  VIP_CUDA(cudaMemsetAsync(ptr, value, bytes, reinterpret_cast<cudaStream_t>(cuda_stream)));
  return VIP_OK;
}

}  // extern "C"
