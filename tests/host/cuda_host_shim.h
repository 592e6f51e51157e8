// cuda_host_shim.h -- lets g++ compile the traversal headers (csrc/bvh.cuh, csrc/device_math.cuh) for the CPU test
// harness: the device intrinsics they use, restated with <cmath> / <cfenv>.  Test infrastructure only.
#pragma once
#include <cuda_runtime.h>
#include <cfenv>
#include <cmath>
#include <cstdint>
#include <cstring>

static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fsqrt_rn(float a) { return sqrtf(a); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
// directed rounding through double: the double result of one float operation is exact (sub) or exact to 2^-53 (mul, fma
// of floats: products of two floats are exact in double; the sum is rounded once), then rounded to float in the wanted direction
static inline float round_dir(double x, int dir) {
    float r = (float)x;
    if (dir < 0 && (double)r > x) r = nextafterf(r, -INFINITY);
    if (dir > 0 && (double)r < x) r = nextafterf(r, INFINITY);
    return r;
}
static inline float __fsub_rd(float a, float b) { return round_dir((double)a - (double)b, -1); }
static inline float __fsub_ru(float a, float b) { return round_dir((double)a - (double)b, +1); }
static inline float __fmul_rd(float a, float b) { return round_dir((double)a * (double)b, -1); }
static inline float __fmul_ru(float a, float b) { return round_dir((double)a * (double)b, +1); }
static inline float __fmaf_rd(float a, float b, float c) { return round_dir(std::fma((double)a, (double)b, (double)c), -1); }
static inline float __fmaf_ru(float a, float b, float c) { return round_dir(std::fma((double)a, (double)b, (double)c), +1); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
static inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline long long __double_as_longlong(double d) { long long i; memcpy(&i, &d, 8); return i; }
static inline double __longlong_as_double(long long i) { double d; memcpy(&d, &i, 8); return d; }
// never executed by the harness (queue_push of the wavefront kernels); declared so that device_math.cuh parses
struct ShimDim3 { unsigned x, y, z; };
static const ShimDim3 threadIdx = {0, 0, 0};
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p += v; return o; }
