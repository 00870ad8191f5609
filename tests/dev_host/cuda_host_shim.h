// cuda_host_shim.h -- TEST INFRASTRUCTURE: lets g++ compile the DEVICE headers of libadapt_b200 (csrc/pt_common.cuh, pt_shade.cuh,
// pt_trace.cuh, pt_volume.cuh) as ordinary host C++, so the very functions the kernels call can be driven path by path on the CPU
// and compared with the oracle (tests/dev_host/dev_host.cpp).  Only intrinsics are provided here; no algorithm lives in this file.
#pragma once
#include <cuda_runtime.h>      // vector types (float3, float4, make_float4 ...); __host__ / __device__ expand to ignored attributes

#include <math.h>
#include <stdint.h>
#include <string.h>

#include <cmath>

#ifndef __CUDACC__
using std::isnan; using std::isfinite;
#define __noinline__
template <typename T> inline T __ldg(const T* p) { return *p; }
inline float __frcp_rn(float x) { return 1.0f / x; }
inline uint32_t __float_as_uint(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline void sincosf(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
#endif
