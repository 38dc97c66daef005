/*
 * oracle/rr_math_ref.h -- TEST INFRASTRUCTURE (oracle side). Not product code.
 *
 * The reference kernel leans on OpenCL builtins whose precision is
 * implementation-defined: native_cos / native_sin / native_log / native_powr /
 * native_sqrt / fast_normalize and (under -cl-fast-relaxed-math,
 * /root/reference/src/image.hpp:49) also tan / sqrt / normalize.  To make
 * "parity" a bit-level statement we pin ONE definition of each of them -- the
 * numerics contract in DESIGN.md section 3 -- built only from IEEE-754 binary32
 * add / mul / div / sqrt (round-to-nearest, no FMA contraction) and integer bit
 * manipulation, so that it evaluates to identical bits on any x86-64 host and
 * on the GPU.  This header is the CPU statement of that contract; the CUDA side
 * has its own independent statement (ripoff_raytracer_b200/csrc/rr_math.cuh)
 * and tests/test_gpu_parity.py::test_numerics_contract_is_bit_identical_on_device compares the two bit for bit.
 *
 * Polynomials follow the classic single-precision Cephes forms (public
 * algorithm: octant reduction with a three-part pi/4, minimax polynomials on
 * the reduced interval).  Accuracy against double precision, measured by
 * tests/test_oracle.py::test_numerics_contract_accuracy on the domains the path uses: sin / cos 1.5 ulp on
 * [-50, 50], log 0.8 ulp, exp2 1.2 ulp, tan 2.8 ulp on [-1.5, 1.5], powr(x, 1/2.2) 8.7 ulp on [0, 1] (its error
 * in the 8-bit output is 3e-5 of one level: it moves a pixel only where c*255 is within that of an integer).
 *
 * Compile with -ffp-contract=off (the Makefile does).
 */
#ifndef RR_MATH_REF_H
#define RR_MATH_REF_H

#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t rrm_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static inline float rrm_from_bits(uint32_t u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}

#define RRM_FOPI 1.27323954473516f /* 4/pi */
#define RRM_DP1 0.78515625f
#define RRM_DP2 2.4187564849853515625e-4f
#define RRM_DP3 3.77489497744594108e-8f

static inline float rrm_sin_poly(float x, float z) {
  float y = -1.9515295891e-4f * z + 8.3321608736e-3f;
  y = y * z - 1.6666654611e-1f;
  return y * z * x + x;
}
static inline float rrm_cos_poly(float z) {
  float y = 2.443315711809948e-5f * z - 1.388731625493765e-3f;
  y = y * z + 4.166664568298827e-2f;
  return y * z * z - 0.5f * z + 1.0f;
}

/* native_cos of the contract.  Domain |x| < 8192; outside it returns 1. */
static inline float rr_cosf_ref(float x) {
  x = fabsf(x);
  if (!(x < 8192.0f)) return 1.0f;
  uint32_t j = (uint32_t)(x * RRM_FOPI);
  float y = (float)j;
  if (j & 1u) {
    j += 1u;
    y += 1.0f;
  }
  j &= 7u;
  int neg = 0;
  if (j > 3u) {
    j -= 4u;
    neg = !neg;
  }
  if (j > 1u) neg = !neg;
  x = ((x - y * RRM_DP1) - y * RRM_DP2) - y * RRM_DP3;
  float z = x * x;
  float r = (j == 1u || j == 2u) ? rrm_sin_poly(x, z) : rrm_cos_poly(z);
  return neg ? -r : r;
}

/* native_sin of the contract.  Domain |x| < 8192; outside it returns 0. */
static inline float rr_sinf_ref(float x) {
  int neg = 0;
  if (x < 0.0f) {
    x = -x;
    neg = 1;
  }
  if (!(x < 8192.0f)) return 0.0f;
  uint32_t j = (uint32_t)(x * RRM_FOPI);
  float y = (float)j;
  if (j & 1u) {
    j += 1u;
    y += 1.0f;
  }
  j &= 7u;
  if (j > 3u) {
    j -= 4u;
    neg = !neg;
  }
  x = ((x - y * RRM_DP1) - y * RRM_DP2) - y * RRM_DP3;
  float z = x * x;
  float r = (j == 1u || j == 2u) ? rrm_cos_poly(z) : rrm_sin_poly(x, z);
  return neg ? -r : r;
}

/* tan of the contract (host-side only: camera fov). */
static inline float rr_tanf_ref(float x) { return rr_sinf_ref(x) / rr_cosf_ref(x); }

/* native_log of the contract.  x > 0 (normal or subnormal). x <= 0 or NaN
 * returns -inf-like large negative (-1e30f); +inf returns +1e30f. */
static inline float rr_logf_ref(float x) {
  if (!(x > 0.0f)) return -1.0e30f;
  if (x > 3.0e38f) return 1.0e30f;
  int e = 0;
  if (x < 1.17549435e-38f) { /* subnormal: renormalise */
    x = x * 8388608.0f;      /* 2^23, exact */
    e = -23;
  }
  uint32_t b = rrm_bits(x);
  e += (int)(b >> 23) - 126; /* frexp exponent: x = m * 2^e, m in [0.5,1) */
  float m = rrm_from_bits((b & 0x007fffffu) | 0x3f000000u);
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

/* 2^t for t <= 128; results below 2^-125 flush to 0. */
static inline float rr_exp2f_ref(float t) {
  if (!(t > -125.0f)) return 0.0f; /* also NaN */
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
  return p * rrm_from_bits((uint32_t)(i + 127) << 23);
}

/* native_powr of the contract: x >= 0.  powr(0,y>0) = 0. */
static inline float rr_powrf_ref(float x, float y) {
  if (!(x > 0.0f)) return 0.0f;
  return rr_exp2f_ref(y * (rr_logf_ref(x) * 1.44269504088896341f));
}

#endif /* RR_MATH_REF_H */
