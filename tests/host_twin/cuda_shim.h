// cuda_shim.h -- TEST INFRASTRUCTURE.  Lets g++ compile the per-env device functions of gym_rotor_b200/csrc/*.cuh
// for the HOST so that they can be unit-tested against the golden vectors without a GPU (tests/test_host_twin.py).
// float64 and float32 instantiations (the headers' few inline-PTX helpers fall back to plain C outside device passes,
// QR_PTX in qr_math.cuh).  Nothing in the package loads this: the product has no CPU path.
#pragma once
#ifdef __CUDACC__
#error "host-only shim"
#endif
#include <cuda_runtime.h>   // vector types, __device__ / __forceinline__ as host no-ops
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

// round-to-nearest single operations (the TU is compiled with -ffp-contract=off, so plain operators are exact twins)
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
static inline int __float_as_int(float f) { int v; memcpy(&v, &f, 4); return v; }
static inline double __longlong_as_double(long long v) { double f; memcpy(&f, &v, 8); return f; }
static inline long long __double_as_longlong(double f) { long long v; memcpy(&v, &f, 8); return v; }
static inline float rsqrtf(float a) { return 1.0f / sqrtf(a); }
// glibc declares (internal) functions of these names: map the CUDA fast-math intrinsics with macros instead
#define __sincosf(a, s, c) sincosf(a, s, c)
#define __expf(a) expf(a)
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
static inline size_t __cvta_generic_to_shared(const void*) { return 0; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((unsigned long long)a * b) >> 32); }
template <typename T> static inline T __ldg(const T* p) { return *p; }
