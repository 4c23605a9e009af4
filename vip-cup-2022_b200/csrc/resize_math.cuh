// Exact arithmetic shared by the preprocessing kernels (preprocess.cu: fused JPEG-emulation pipeline; preprocess_stream.cu:
// streaming resize / normalise): TF's bicubic tap computation, the separately-rounded 4-tap sums, x / 255 and the small
// integer <-> float conversions.  Reference semantics: dataset/dataset.py:31-37 (tf.image.resize(bicubic), / 255.0),
// restated in oracle/preprocess.py (the checker, never linked here).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vip {

// x / 255.0f, correctly rounded, for every finite x with 1e-30 <= |x| <= 1e30 and +0 (Markstein: RN(1/255)
// multiply, exact remainder by FMA, one correction).  Exhaustively verified on the CPU and by
// vip_selftest_div255 on the device.
__device__ __forceinline__ float div255(float x) {
  const float rc = 0.003921568859368562698f;
  const float q0 = __fmul_rn(x, rc);
  const float r = __fmaf_rn(-q0, 255.0f, x);
  return __fmaf_rn(r, rc, q0);
}

// Keys cubic (a = -0.5) LUT entry exactly as TF's InitCoeffsTable: double arithmetic on a float abscissa,
// rounded to float once.  i in [0, 1024].
__device__ inline float coeff_near(int i) {
  const double x = (double)((float)i * 0.0009765625f);
  double t = __dadd_rn(__dmul_rn(1.5, x), -2.5);
  t = __dmul_rn(__dmul_rn(t, x), x);
  return __double2float_rn(__dadd_rn(t, 1.0));
}
__device__ inline float coeff_far(int i) {
  const double x = (double)((float)i * 0.0009765625f + 1.0f);
  double t = __dadd_rn(__dmul_rn(-0.5, x), 2.5);
  t = __dadd_rn(__dmul_rn(t, x), -4.0);
  t = __dmul_rn(t, x);
  return __double2float_rn(__dadd_rn(t, 2.0));
}

// TF GetWeightsAndIndices<HalfPixelScaler, use_keys_cubic=true>
__device__ inline void compute_tap(int o, int in_size, int out_size, float4* w_out, short4* i_out) {
  const float scale = __fdiv_rn((float)in_size, (float)out_size);
  const float loc = __fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), scale), 0.5f);
  const float fl = floorf(loc);
  const int il = (int)fl;
  const float delta = __fsub_rn(loc, fl);
  const int off = __float2int_rn(__fmul_rn(delta, 1024.0f));
  const int lim = in_size - 1;
  const int r0 = il - 1, r1 = il, r2 = il + 1, r3 = il + 2;
  const int i0 = min(max(r0, 0), lim), i1 = min(max(r1, 0), lim);
  const int i2 = min(max(r2, 0), lim), i3 = min(max(r3, 0), lim);
  float w0 = (i0 == r0) ? coeff_far(off) : 0.0f;
  float w1 = (i1 == r1) ? coeff_near(off) : 0.0f;
  float w2 = (i2 == r2) ? coeff_near(1024 - off) : 0.0f;
  float w3 = (i3 == r3) ? coeff_far(1024 - off) : 0.0f;
  const float sum = __fadd_rn(__fadd_rn(__fadd_rn(w0, w1), w2), w3);
  if (fabsf(sum) >= 1000.0f * 1.17549435e-38f) {
    const float inv = __fdiv_rn(1.0f, sum);
    w0 = __fmul_rn(w0, inv);
    w1 = __fmul_rn(w1, inv);
    w2 = __fmul_rn(w2, inv);
    w3 = __fmul_rn(w3, inv);
  }
  *w_out = make_float4(w0, w1, w2, w3);
  *i_out = make_short4((short)i0, (short)i1, (short)i2, (short)i3);
}

__device__ __forceinline__ float tap4(float p0, float p1, float p2, float p3, const float4 w) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(p0, w.x), __fmul_rn(p1, w.y)), __fmul_rn(p2, w.z)),
                   __fmul_rn(p3, w.w));
}

// Exact small-integer <-> float conversions on the FMA / ALU pipes (I2F / F2I issue on the quarter-rate XU pipe and
// there are ~0.8 M of them per image): 2^23 + n has n in its low mantissa bits for 0 <= n < 2^23.
__device__ __forceinline__ float u8f(unsigned word, int k) {   // (float) byte k of word
  return __fsub_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u | (unsigned)k)), 8388608.0f);
}
__device__ __forceinline__ float small_int_to_float(int n) {   // 0 <= n < 2^23
  return __fsub_rn(__uint_as_float(0x4B000000u | (unsigned)n), 8388608.0f);
}
__device__ __forceinline__ int trunc_small_float(float x) {    // (int)x for 0 <= x < 2^22 (round toward zero = floor)
  return (int)(__float_as_uint(__fadd_rz(x, 8388608.0f)) & 0x7fffffu);
}

__device__ __forceinline__ void gray3(float* p) {
  const float g = __fadd_rn(__fadd_rn(__fmul_rn(p[0], 0.2989f), __fmul_rn(p[1], 0.5870f)), __fmul_rn(p[2], 0.1140f));
  p[0] = g;
  p[1] = g;
  p[2] = g;
}


}  // namespace vip
