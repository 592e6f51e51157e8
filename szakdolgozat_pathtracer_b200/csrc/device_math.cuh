// device_math.cuh -- scalar building blocks of the sm_100a kernels.
//
// Everything that feeds a hit decision or a radiance value is written in a
// fixed operation order over IEEE single-precision +,-,*,/,sqrt and the whole
// library is compiled with -fmad=false (no FMA contraction, default
// -prec-div/-prec-sqrt), so results are bit-identical to the CPU oracle's
// (oracle/oracle_math.h, g++ -ffp-contract=off).  No CUDA libm transcendental
// is used on the radiance path: sin/cos/atan2/asin are the Cephes
// single-precision algorithms (<= 2 ulp on the ranges used), pow(x,5) is
// ((x*x)*(x*x))*x, the display transform's pow(x, 1/2.2) and pow(x, 1/2.4) are det_powf (double precision, one
// rounding).
//
// Two arithmetic modes share this source (PTB_FAST, see below):
//   exact (PTB_FAST = 0, nvcc -fmad=false): what the paragraph above describes; the parity-checked build.
//   fast  (PTB_FAST = 1, nvcc -fmad=true, fast_kernels.cu): FMA contraction everywhere, ar_div / ar_rcp / ar_sqrt /
//         ar_rsqrt are single MUFU approximations, sin/cos are MUFU.SIN/COS.  What decides a HIT ID stays exact in
//         both builds because it is written with the ex_* intrinsics (IEEE round-to-nearest operations that no
//         compiler flag contracts or approximates): camera rays (start_sample), the traversal set-up and the watertight
//         ray-triangle test (bvh.cuh).  The reference itself is built with --use_fast_math; the
//         fast build is gated by tests/test_gpu_fast_mode.py (primary-hit IDs bit-exact, image RMSE bound).
//
// Vector helpers follow the OptiX SDK's sutil/vec_math.h semantics the
// reference relies on (optixSphere.cu:10): v / s multiplies by 1.0f / s,
// normalize(v) = v * (1.0f / sqrtf(dot(v,v))), lerp(a,b,t) = a + t*(b-a),
// reflect(i,n) = i - 2.0f*n*dot(n,i), faceforward(n,i,r) = n*copysign(1,dot(i,r)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PTB_DEV __device__ __forceinline__

// The kernels are compiled twice (see views.cuh): PTB_NS = ptb with -fmad=false (exact, bit-identical to the oracle) and
// PTB_NS = ptb_fast with -fmad=true and PTB_FAST = 1 (fast_kernels.cu).
#ifndef PTB_NS
#define PTB_NS ptb
#endif
#ifndef PTB_FAST
#define PTB_FAST 0
#endif

namespace PTB_NS {

// ---- ex_*: IEEE operations immune to -fmad / -prec-div / -prec-sqrt (identical to the plain operators in the exact build)
PTB_DEV float ex_mul(float a, float b) { return __fmul_rn(a, b); }
PTB_DEV float ex_add(float a, float b) { return __fadd_rn(a, b); }
PTB_DEV float ex_sub(float a, float b) { return __fsub_rn(a, b); }
PTB_DEV float ex_div(float a, float b) { return __fdiv_rn(a, b); }
PTB_DEV float ex_sqrt(float a) { return __fsqrt_rn(a); }
// ---- ar_*: the build's division / reciprocal / square root
#if PTB_FAST
PTB_DEV float ar_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PTB_DEV float ar_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PTB_DEV float ar_rsqrt(float x) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
PTB_DEV float ar_div(float a, float b) { return a * ar_rcp(b); }
#else
PTB_DEV float ar_rcp(float x) { return 1.0f / x; }
PTB_DEV float ar_sqrt(float x) { return sqrtf(x); }
PTB_DEV float ar_rsqrt(float x) { return 1.0f / sqrtf(x); }
PTB_DEV float ar_div(float a, float b) { return a / b; }
#endif

PTB_DEV float3 mk3(float x, float y, float z) { return make_float3(x, y, z); }
PTB_DEV float3 mk3(float s) { return make_float3(s, s, s); }
PTB_DEV float3 mk3(float4 v) { return make_float3(v.x, v.y, v.z); }

PTB_DEV float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
PTB_DEV float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
PTB_DEV float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
PTB_DEV float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
PTB_DEV float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
PTB_DEV float3 operator*(float s, float3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
PTB_DEV float3 operator+(float3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
PTB_DEV float3 operator-(float3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
PTB_DEV float3 operator/(float3 a, float3 b) { return mk3(ar_div(a.x, b.x), ar_div(a.y, b.y), ar_div(a.z, b.z)); }
PTB_DEV float3 operator/(float3 a, float s) { float inv = ar_rcp(s); return a * inv; }
PTB_DEV float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
PTB_DEV float3 cross(float3 a, float3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
PTB_DEV float length(float3 a) { return ar_sqrt(dot(a, a)); }
PTB_DEV float3 normalize(float3 a) { float inv = ar_rsqrt(dot(a, a)); return a * inv; }
// the same with ex_* operations (camera rays)
PTB_DEV float3 ex_normalize(float3 a) {
    const float d = ex_add(ex_add(ex_mul(a.x, a.x), ex_mul(a.y, a.y)), ex_mul(a.z, a.z));
    const float inv = ex_div(1.0f, ex_sqrt(d));
    return mk3(ex_mul(a.x, inv), ex_mul(a.y, inv), ex_mul(a.z, inv));
}
PTB_DEV float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
PTB_DEV float4 lerp(float4 a, float4 b, float t) {
    return make_float4(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w));
}
PTB_DEV float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
PTB_DEV float3 clamp3(float3 v, float a, float b) { return mk3(clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b)); }
PTB_DEV float3 reflect(float3 i, float3 n) { return i - 2.0f * n * dot(n, i); }
PTB_DEV float3 faceforward(float3 n, float3 i, float3 nref) { return n * copysignf(1.0f, dot(i, nref)); }
PTB_DEV float comp(float3 v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

// ---- RNG (optixSphere.cu:24-35) ----------------------------------------------
// pcg_hash returns float in the reference: the 32-bit hash is rounded to float
// and converted back (cvt.rzi.u32.f32 saturates 2^32 to 0xFFFFFFFF).
PTB_DEV float pcg_hash_f(uint32_t input) {
    uint32_t state = input * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (float)((word >> 22u) ^ word);
}
PTB_DEV float myrnd(uint32_t& seed) {
    // seed = pcg_hash(seed) converts the float back to uint (saturating); (float)seed is then the SAME float again
    // (2^32 saturates to 0xFFFFFFFF, which rounds back to 2^32), and dividing by (float)UINT_MAX == 2^32 is an exact
    // scaling: one int->float conversion per draw instead of two, bit-identical to optixSphere.cu:32-35.
    const float hf = pcg_hash_f(seed);
    seed = (uint32_t)hf;
    return ex_mul(hf, 2.3283064365386962890625e-10f);
}

// ---- detmath (same algorithms as oracle/oracle_math.h) -------------------------
// written with ex_* operations: bit-identical in both builds (the depth-of-field camera ray uses it, start_sample)
PTB_DEV void det_sincosf_ex(float xx, float* s_out, float* c_out) {
    const float FOPI = 1.27323954473516f;
    const float DP1 = 0.78515625f, DP2 = 2.4187564849853515625e-4f, DP3 = 3.77489497744594108e-8f;
    float x = fabsf(xx);
    int j = (int)ex_mul(FOPI, x);
    float y = (float)j;
    if (j & 1) { j += 1; y = ex_add(y, 1.0f); }
    j &= 7;
    x = ex_sub(ex_sub(ex_sub(x, ex_mul(y, DP1)), ex_mul(y, DP2)), ex_mul(y, DP3));
    const float z = ex_mul(x, x);
    // ((-1.95e-4 z + 8.33e-3) z - 1.67e-1) z x + x
    const float ps = ex_add(ex_mul(ex_mul(ex_sub(ex_mul(ex_add(ex_mul(-1.9515295891e-4f, z), 8.3321608736e-3f), z), 1.6666654611e-1f), z), x), x);
    // ((2.44e-5 z - 1.39e-3) z + 4.17e-2) z z - 0.5 z + 1
    const float pc = ex_add(ex_sub(ex_mul(ex_mul(ex_add(ex_mul(ex_sub(ex_mul(2.443315711809948e-5f, z), 1.388731625493765e-3f), z), 4.166664568298827e-2f), z), z), ex_mul(0.5f, z)), 1.0f);
    float s, c;
    if (j == 0) { s = ps; c = pc; }
    else if (j == 2) { s = pc; c = -ps; }
    else if (j == 4) { s = -ps; c = -pc; }
    else { s = -pc; c = ps; }
    if (xx < 0.0f) s = -s;
    *s_out = s; *c_out = c;
}
PTB_DEV void det_sincosf(float xx, float* s_out, float* c_out) {
#if PTB_FAST
    __sincosf(xx, s_out, c_out);  // MUFU.SIN / MUFU.COS; arguments are in [0, 2 pi]
#else
    det_sincosf_ex(xx, s_out, c_out);
#endif
}
PTB_DEV float det_atan_pos(float t) {
    float y;
    if (t > 2.414213562373095f) { y = 1.5707963267948966f; t = -ar_rcp(t); }
    else if (t > 0.4142135623730950f) { y = 0.7853981633974483f; t = ar_div(t - 1.0f, t + 1.0f); }
    else y = 0.0f;
    float z = t * t;
    y += (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * t + t;
    return y;
}
PTB_DEV float det_atan2f(float y, float x) {
    const float PI_F = 3.14159265358979323846f, PIO2_F = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIO2_F;
        if (y < 0.0f) return -PIO2_F;
        return 0.0f;
    }
    if (y == 0.0f) return x > 0.0f ? 0.0f : PI_F;
    float a = det_atan_pos(fabsf(ar_div(y, x)));
    if (x < 0.0f) a = PI_F - a;
    return y < 0.0f ? -a : a;
}
PTB_DEV float det_asinf(float xx) {
    float a = fabsf(xx);
    if (a > 1.0f) a = 1.0f;
    float z, x;
    bool big = a > 0.5f;
    if (big) { z = 0.5f * (1.0f - a); x = ar_sqrt(z); }
    else { x = a; z = x * x; }
    float r = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z + 1.6666752422e-1f) * z * x + x;
    if (big) { r = r + r; r = 1.5707963267948966f - r; }
    return xx < 0.0f ? -r : r;
}
PTB_DEV float det_pow5(float x) { float x2 = x * x; float x4 = x2 * x2; return x4 * x; }
// pow(x, y) of the display transform (cu:425, SDK make_color): exp2(y * log2(x)) in IEEE double, fixed operation order,
// one rounding to float -- the same code as oracle/oracle_math.h det_powf, so the 8-bit frame is bit-identical too.
#if !PTB_FAST  // display transform: exact build only (double arithmetic must not be contracted)
PTB_DEV float det_powf(float xf, float yf) {
    if (yf == 0.0f || xf == 1.0f) return 1.0f;
    if (xf != xf || yf != yf) return xf + yf;
    if (xf < 0.0f) return __int_as_float(0x7fc00000);
    if (xf == 0.0f) return yf > 0.0f ? 0.0f : __int_as_float(0x7f800000);
    if (xf > 3.0e38f) return yf > 0.0f ? xf : 0.0f;
    const double x = (double)xf;
    const long long b = __double_as_longlong(x);
    int e = (int)((b >> 52) & 0x7ffll) - 1023;
    double m = __longlong_as_double((b & 0x000fffffffffffffll) | 0x3ff0000000000000ll);
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0), s2 = s * s;
    double p = 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0; p = p * s2 + 1.0 / 19.0; p = p * s2 + 1.0 / 17.0; p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0; p = p * s2 + 1.0 / 11.0; p = p * s2 + 1.0 / 9.0; p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0; p = p * s2 + 1.0 / 3.0; p = p * s2 + 1.0;
    const double log2x = (double)e + (2.0 * s * p) * 1.4426950408889634;
    const double t = (double)yf * log2x;
    if (t >= 128.0) return __int_as_float(0x7f800000);
    if (t < -160.0) return 0.0f;
    const double n = floor(t + 0.5);
    const double z = (t - n) * 0.6931471805599453;
    double q = 1.0 / 87178291200.0;
    q = q * z + 1.0 / 6227020800.0; q = q * z + 1.0 / 479001600.0; q = q * z + 1.0 / 39916800.0; q = q * z + 1.0 / 3628800.0;
    q = q * z + 1.0 / 362880.0; q = q * z + 1.0 / 40320.0; q = q * z + 1.0 / 5040.0; q = q * z + 1.0 / 720.0;
    q = q * z + 1.0 / 120.0; q = q * z + 1.0 / 24.0; q = q * z + 1.0 / 6.0; q = q * z + 0.5; q = q * z + 1.0; q = q * z + 1.0;
    const double scale = __longlong_as_double(((long long)n + 1023ll) << 52);
    return (float)(q * scale);
}
#endif

// ---- orthonormal basis (optixSphere.cu:38-61) -----------------------------------
struct Onb {
    float3 t, b, n;
    PTB_DEV explicit Onb(float3 normal) {
        n = normalize(normal);
        float3 up = fabsf(n.y) < 0.9999f ? mk3(0.0f, 1.0f, 0.0f) : mk3(1.0f, 0.0f, 0.0f);
        t = normalize(cross(up, n));
        b = normalize(cross(n, t));
    }
    PTB_DEV float3 inverse_transform(float3 p) const { return p.x * t + p.y * n + p.z * b; }
};

// ---- warp-aggregated queue append ---------------------------------------------------
// All 32 lanes of the warp must call this (converged); lanes with pred == false
// append nothing.  One atomicAdd per warp, order inside the warp is preserved.
PTB_DEV void queue_push(uint32_t* __restrict__ queue, uint32_t* __restrict__ counter, bool pred, uint32_t value) {
    const unsigned mask = __ballot_sync(0xffffffffu, pred);
    if (mask == 0u) return;
    const unsigned lane = threadIdx.x & 31u;
    const int leader = __ffs(mask) - 1;
    uint32_t base = 0;
    if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + (uint32_t)__popc(mask & ((1u << lane) - 1u))] = value;
}

}  // namespace PTB_NS
