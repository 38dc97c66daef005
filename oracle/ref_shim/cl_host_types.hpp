// oracle/ref_shim/cl_host_types.hpp -- TEST INFRASTRUCTURE (Oracle A).
//
// Stand-in for the handful of Khronos <CL/cl.h> host types that the
// reference's CL-free host headers use (src/readobj.hpp, src/math.hpp).  The
// image ships no OpenCL headers.  Layout matches cl_platform.h:
// cl_float4 is a 16-byte-aligned union whose first member is `float s[4]`,
// and cl_float3 is an alias of cl_float4.
#pragma once
#include <cfloat>
#include <cstddef>
#include <cstdint>

typedef union alignas(16) {
  float s[4];
} cl_float4;
typedef cl_float4 cl_float3;
typedef union alignas(8) {
  float s[2];
} cl_float2;
typedef uint64_t cl_ulong;
typedef uint32_t cl_uint;
typedef int32_t cl_int;
#define CL_FLT_MAX FLT_MAX
#define CL_FLT_MIN FLT_MIN
