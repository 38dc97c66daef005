// rr_math.cuh -- device statement of the numerics contract.
//
// The reference kernel (src/Trace.cl) calls OpenCL builtins whose precision is
// implementation-defined (native_cos/sin/log/powr/sqrt, fast_normalize, and --
// under its "-cl-fast-relaxed-math" build flag, src/image.hpp:49 -- tan, sqrt,
// normalize).  To make parity a bit-level statement this path evaluates them
// with fixed algorithms built only from IEEE binary32 add/mul/div/sqrt
// (round-to-nearest) and integer bit operations.  The translation unit is
// compiled with -fmad=false, so no multiply-add is contracted and a CPU that
// evaluates the same expressions in the same order produces the same bits.
//
// Algorithms: single-precision Cephes-style kernels (octant reduction with a
// three-part pi/4 for sin/cos; frexp + degree-8 polynomial for log; floor +
// degree-6 polynomial for exp2).
#pragma once
#include <cstdint>

namespace rr {

__device__ __forceinline__ float sin_kernel(float x, float z) {
  float y = -1.9515295891e-4f * z + 8.3321608736e-3f;
  y = y * z - 1.6666654611e-1f;
  return y * z * x + x;
}
__device__ __forceinline__ float cos_kernel(float z) {
  float y = 2.443315711809948e-5f * z - 1.388731625493765e-3f;
  y = y * z + 4.166664568298827e-2f;
  return y * z * z - 0.5f * z + 1.0f;
}

// Shared octant reduction: returns reduced argument, octant in j (0..3 after
// folding) and whether the folded half-turn flips the sign.
__device__ __forceinline__ float reduce_octant(float ax, uint32_t& j, bool& halfTurn) {
  j = (uint32_t)(ax * 1.27323954473516f);
  float y = (float)j;
  if (j & 1u) {
    j += 1u;
    y += 1.0f;
  }
  j &= 7u;
  halfTurn = j > 3u;
  if (halfTurn) j -= 4u;
  return ((ax - y * 0.78515625f) - y * 2.4187564849853515625e-4f) - y * 3.77489497744594108e-8f;
}

// native_cos of the contract; |x| < 8192, otherwise 1.
__device__ __forceinline__ float cos_c(float x) {
  float ax = fabsf(x);
  if (!(ax < 8192.0f)) return 1.0f;
  uint32_t j;
  bool neg;
  float r = reduce_octant(ax, j, neg);
  if (j > 1u) neg = !neg;
  float z = r * r;
  float v = (j == 1u || j == 2u) ? sin_kernel(r, z) : cos_kernel(z);
  return neg ? -v : v;
}

// native_sin of the contract; |x| < 8192, otherwise 0.
__device__ __forceinline__ float sin_c(float x) {
  bool neg0 = x < 0.0f;
  float ax = neg0 ? -x : x;
  if (!(ax < 8192.0f)) return 0.0f;
  uint32_t j;
  bool half;
  float r = reduce_octant(ax, j, half);
  bool neg = neg0 != half;
  float z = r * r;
  float v = (j == 1u || j == 2u) ? cos_kernel(z) : sin_kernel(r, z);
  return neg ? -v : v;
}

__device__ __forceinline__ float tan_c(float x) { return sin_c(x) / cos_c(x); }

// native_log of the contract.
__device__ __forceinline__ float log_c(float x) {
  if (!(x > 0.0f)) return -1.0e30f;
  if (x > 3.0e38f) return 1.0e30f;
  int e = 0;
  if (x < 1.17549435e-38f) {
    x = x * 8388608.0f;
    e = -23;
  }
  uint32_t b = __float_as_uint(x);
  e += (int)(b >> 23) - 126;
  float m = __uint_as_float((b & 0x007fffffu) | 0x3f000000u);
  if (m < 0.707106781186547524f) {
    e -= 1;
    m = m + m - 1.0f;
  } else {
    m = m - 1.0f;
  }
  float z = m * m;
  float y = 7.0376836292e-2f * m - 1.1514610310e-1f;
  y = y * m + 1.1676998740e-1f;
  y = y * m - 1.2420140846e-1f;
  y = y * m + 1.4249322787e-1f;
  y = y * m - 1.6668057665e-1f;
  y = y * m + 2.0000714765e-1f;
  y = y * m - 2.4999993993e-1f;
  y = y * m + 3.3333331174e-1f;
  y = y * m * z;
  float fe = (float)e;
  y = y + -2.12194440e-4f * fe;
  y = y + -0.5f * z;
  float r = m + y;
  r = r + 0.693359375f * fe;
  return r;
}

// 2^t; results below 2^-125 flush to 0, t is clamped to 127.
__device__ __forceinline__ float exp2_c(float t) {
  if (!(t > -125.0f)) return 0.0f;
  if (t > 127.0f) t = 127.0f;
  float fi = floorf(t);
  float f = t - fi;
  int i = (int)fi;
  if (f > 0.5f) {
    i += 1;
    f = f - 1.0f;
  }
  float p = 1.535336188319500e-4f * f + 1.339887440266574e-3f;
  p = p * f + 9.618437357674640e-3f;
  p = p * f + 5.550332471162809e-2f;
  p = p * f + 2.402264791363012e-1f;
  p = p * f + 6.931472028550421e-1f;
  p = p * f + 1.0f;
  return p * __uint_as_float((uint32_t)(i + 127) << 23);
}

// native_powr of the contract, x >= 0.
__device__ __forceinline__ float powr_c(float x, float y) {
  if (!(x > 0.0f)) return 0.0f;
  return exp2_c(y * (log_c(x) * 1.44269504088896341f));
}

}  // namespace rr
