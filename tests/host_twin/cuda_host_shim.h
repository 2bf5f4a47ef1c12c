// cuda_host_shim.h -- TEST INFRASTRUCTURE ONLY.  Lets g++ compile tvc_ai_b200/csrc/tvc_device.cuh (the device-side model of the
// env step) for the HOST, so that the CPU test-suite can run the very source the kernels are built from against the fp64
// oracle (tests/test_host_twin.py) and a kernel/oracle logic difference can be found without a GPU.  The product never
// includes this file: libtvc_b200.so is nvcc-compiled device code only, and there is no CPU fallback.
#pragma once
#include <cuda_runtime.h>   // float2/float4/uint4 + make_*; __device__ / __forceinline__ become (ignored) attributes under g++

#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>   // every libstdc++ header this build needs comes BEFORE the __noinline__ macro below (libstdc++ spells
#include <vector>   // its own attribute __attribute__((__noinline__)))

#define TVC_HOST_TWIN 1
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif
// glibc declares __logf / __sincosf (internal aliases, not exported): route the CUDA fast intrinsics to the libm functions
#define __logf(x) logf(x)
#define __expf(x) expf(x)
#define __sincosf(x, s, c) sincosf((x), (s), (c))

static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline float __double2float_rn(double d) { return (float)d; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline void sincospif(float x, float *s, float *c) { *s = (float)sin(3.141592653589793 * (double)x); *c = (float)cos(3.141592653589793 * (double)x); }
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline void __syncthreads() {}
template <typename T> static inline T __ldcg(const T *p) { return *p; }
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
