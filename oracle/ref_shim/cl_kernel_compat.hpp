// oracle/ref_shim/cl_kernel_compat.hpp -- TEST INFRASTRUCTURE (Oracle A).
//
// A minimal "OpenCL C device" written in C++ so that the reference's kernel
// text (/root/reference/src/Trace.cl, read at BUILD time, never copied into
// this repo) compiles natively for the host CPU.  There is no OpenCL runtime in
// this image (no PoCL, no CL headers, no ICD), so this stands in
// for the device compiler.  Everything an OpenCL implementation is free to
// choose -- the precision of native_* / fast_* builtins and, under the
// reference's "-cl-fast-relaxed-math" build flag (src/image.hpp:49), of
// tan / sqrt / normalize / length too -- is pinned to the numerics contract in
// oracle/rr_math_ref.h + DESIGN.md section 3: IEEE binary32, round to nearest,
// no FMA contraction, left-to-right dot products.
//
// Included INSIDE `namespace clk { ... }` by ref_driver.cpp, before the
// sed-adapted kernel text.  sed rules (oracle/Makefile): "(float3)(" -> "mk3(",
// "(float2)(" -> "mk2(", "(uchar4)(" -> "mkuc4(", and the two bare `private`
// storage-class lines (Trace.cl:330,332) are dropped.

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned char uchar;

#define __kernel
#define __global
#define __private

struct float2 {
  float x, y;
  float& operator[](int i) { return i ? y : x; }
};
inline float2 mk2(float a, float b) { return float2{a, b}; }
inline float2 operator*(float2 a, float s) { return float2{a.x * s, a.y * s}; }
inline float2 operator*(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
inline float2 operator/(float2 a, float2 b) { return float2{a.x / b.x, a.y / b.y}; }
inline float2 operator-(float2 a, float2 b) { return float2{a.x - b.x, a.y - b.y}; }
inline float2 floor(float2 a) { return float2{::floorf(a.x), ::floorf(a.y)}; }

// OpenCL float3 occupies 16 bytes, 16-byte aligned.
struct alignas(16) float3 {
  float x, y, z, pad_;
  float3() = default;
  float3(float s) : x(s), y(s), z(s), pad_(0.0f) {}
  float3(float a, float b, float c) : x(a), y(b), z(c), pad_(0.0f) {}
};
inline float3 mk3(float s) { return float3(s); }
inline float3 mk3(float a, float b, float c) { return float3(a, b, c); }

struct uchar4 {
  uchar x, y, z, w;
};
inline uchar4 mkuc4(uchar a, uchar b, uchar c, uchar d) { return uchar4{a, b, c, d}; }

inline float3 operator+(float3 a, float3 b) { return float3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline float3 operator-(float3 a, float3 b) { return float3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline float3 operator*(float3 a, float3 b) { return float3(a.x * b.x, a.y * b.y, a.z * b.z); }
inline float3 operator/(float3 a, float3 b) { return float3(a.x / b.x, a.y / b.y, a.z / b.z); }
inline float3 operator*(float3 a, float s) { return float3(a.x * s, a.y * s, a.z * s); }
inline float3 operator*(float s, float3 a) { return float3(s * a.x, s * a.y, s * a.z); }
inline float3 operator/(float3 a, float s) { return float3(a.x / s, a.y / s, a.z / s); }
inline float3 operator/(float s, float3 a) { return float3(s / a.x, s / a.y, s / a.z); }
inline float3 operator-(float3 a) { return float3(-a.x, -a.y, -a.z); }
inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
inline float3& operator*=(float3& a, float s) { a = a * s; return a; }
inline float3& operator/=(float3& a, float s) { a = a / s; return a; }

// geometric builtins: left-to-right, no contraction
inline float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float3 cross(float3 a, float3 b) {
  return float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float sqrt(float x) { return ::sqrtf(x); }
inline float native_sqrt(float x) { return ::sqrtf(x); }
inline float length(float3 a) { return ::sqrtf(dot(a, a)); }
inline float3 fast_normalize(float3 a) {
  float inv = 1.0f / ::sqrtf(dot(a, a));
  return a * inv;
}
inline float3 normalize(float3 a) { return fast_normalize(a); }

// common / math builtins
inline float fabs(float x) { return ::fabsf(x); }
inline float fmin(float a, float b) { return ::fminf(a, b); }
inline float fmax(float a, float b) { return ::fmaxf(a, b); }
inline float min(float a, float b) { return ::fminf(a, b); }
inline float max(float a, float b) { return ::fmaxf(a, b); }
inline float3 fmin(float3 a, float3 b) { return float3(fmin(a.x, b.x), fmin(a.y, b.y), fmin(a.z, b.z)); }
inline float3 fmax(float3 a, float3 b) { return float3(fmax(a.x, b.x), fmax(a.y, b.y), fmax(a.z, b.z)); }
inline float floor(float x) { return ::floorf(x); }
inline float clamp(float x, float lo, float hi) { return fmin(fmax(x, lo), hi); }
inline float3 clamp(float3 v, float lo, float hi) {
  return float3(clamp(v.x, lo, hi), clamp(v.y, lo, hi), clamp(v.z, lo, hi));
}
inline float sign(float x) { return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f); }
inline float radians(float deg) { return deg * 0.017453292519943295f; }
inline bool isfinite(float x) { return std::isfinite(x); }

// implementation-defined builtins -> numerics contract
inline float native_cos(float x) { return rr_cosf_ref(x); }
inline float native_sin(float x) { return rr_sinf_ref(x); }
inline float native_log(float x) { return rr_logf_ref(x); }
inline float tan(float x) { return rr_tanf_ref(x); }
inline float3 native_powr(float3 a, float3 b) {
  return float3(rr_powrf_ref(a.x, b.x), rr_powrf_ref(a.y, b.y), rr_powrf_ref(a.z, b.z));
}

// work-item functions: the driver sets the id before calling the kernel body
static thread_local size_t g_global_id[3] = {0, 0, 0};
inline size_t get_global_id(uint dim) { return g_global_id[dim]; }
