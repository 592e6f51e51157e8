// oracle/oracle_math.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Scalar host math used by the CPU oracle (oracle_pt.cpp) and by the shim that
// compiles the reference's own optixSphere.cu on the host (ref_shim/).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may link this.
//
// Three groups of functions:
//  1. float2/3/4 helpers restating the OptiX SDK 8.0.0 sutil/vec_math.h
//     definitions the reference relies on (optixSphere.cu:10).  The SDK is NOT
//     vendored in /root/reference, so these are restated from the published
//     SDK source as remembered: PARITY UNPINNED at this boundary.
//  2. The reference's RNG (optixSphere.cu:24-35), including CUDA's saturating
//     float->uint conversion which is undefined behaviour in host C++.
//  3. "detmath": deterministic single-precision sin/cos/atan2/asin (and a double-precision pow) built only
//     from IEEE +,-,*,/,sqrt in a fixed order (Cephes single-precision
//     algorithms), so that the CPU oracle and the CUDA kernels produce
//     bit-identical results.  The reference calls CUDA libm (cosf/sinf/atan2f/
//     asinf, optixSphere.cu:249-250,288-289,499,543-544); detmath deviates from
//     a correctly rounded libm by <= 2 ulp on the argument ranges used.
//
// Everything must be compiled with -ffp-contract=off (no FMA contraction).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace orc {

struct v2 { float x, y; };
struct v3 { float x, y, z; };
struct v4 { float x, y, z, w; };

static inline v3 mk3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 mk3(float s) { v3 r = {s, s, s}; return r; }
static inline v4 mk4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }

// ---- sutil/vec_math.h restatement (float3) --------------------------------
static inline v3 operator+(v3 a, v3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 operator-(v3 a, v3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 operator-(v3 a) { return mk3(-a.x, -a.y, -a.z); }
static inline v3 operator*(v3 a, v3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 operator*(v3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline v3 operator*(float s, v3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline v3 operator+(v3 a, float s) { return mk3(a.x + s, a.y + s, a.z + s); }
static inline v3 operator-(v3 a, float s) { return mk3(a.x - s, a.y - s, a.z - s); }
// float3 / float3 : component-wise true division
static inline v3 operator/(v3 a, v3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
// float3 / float  : SDK multiplies by the reciprocal
static inline v3 operator/(v3 a, float s) { float inv = 1.0f / s; return a * inv; }
static inline float dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline float length(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 normalize(v3 a) { float inv = 1.0f / sqrtf(dot(a, a)); return a * inv; }
static inline float lerpf(float a, float b, float t) { return a + t * (b - a); }
static inline v3 lerp(v3 a, v3 b, float t) { return a + t * (b - a); }
static inline v4 lerp(v4 a, v4 b, float t) {
    return mk4(a.x + t * (b.x - a.x), a.y + t * (b.y - a.y), a.z + t * (b.z - a.z), a.w + t * (b.w - a.w));
}
static inline float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
static inline v3 clamp(v3 v, float a, float b) { return mk3(clampf(v.x, a, b), clampf(v.y, a, b), clampf(v.z, a, b)); }
static inline v3 reflect(v3 i, v3 n) { return i - 2.0f * n * dot(n, i); }
static inline v3 faceforward(v3 n, v3 i, v3 nref) { return n * copysignf(1.0f, dot(i, nref)); }

// ---- RNG: optixSphere.cu:24-35 ---------------------------------------------
// pcg_hash is declared to return float, so the 32-bit hash is rounded to the
// nearest float and then converted back to unsigned int in myrnd.
// CUDA's cvt.rzi.u32.f32 saturates (2^32 -> 0xFFFFFFFF); x86 cvttss2si wraps
// to 0.  sat_cuda selects which one (the oracle rule is CUDA's).
static inline float pcg_hash_f(uint32_t input) {
    uint32_t state = input * 747796405u + 2891336453u;
    uint32_t word = ((state >> ((state >> 28u) + 4u)) ^ state) * 277803737u;
    return (float)((word >> 22u) ^ word);
}
static inline uint32_t f2u_cuda(float f, bool sat_cuda) {
    if (f >= 4294967296.0f) return sat_cuda ? 0xFFFFFFFFu : 0u;
    return (uint32_t)f;
}
static inline float myrnd(uint32_t& seed, bool sat_cuda = true) {
    seed = f2u_cuda(pcg_hash_f(seed), sat_cuda);
    return (float)seed / 4294967296.0f;  // (float)UINT_MAX == 2^32
}

// ---- detmath ---------------------------------------------------------------
// Cephes sinf/cosf: octant reduction with a 3-term Cody-Waite split of pi/4.
// Valid for 0 <= |x| < 8192; the path only uses x in [0, 2*pi].
static inline void det_sincosf(float xx, float* s_out, float* c_out) {
    const float FOPI = 1.27323954473516f;
    const float DP1 = 0.78515625f, DP2 = 2.4187564849853515625e-4f, DP3 = 3.77489497744594108e-8f;
    float x = fabsf(xx);
    int j = (int)(FOPI * x);
    float y = (float)j;
    if (j & 1) { j += 1; y += 1.0f; }
    j &= 7;
    x = ((x - y * DP1) - y * DP2) - y * DP3;
    float z = x * x;
    float ps = ((-1.9515295891e-4f * z + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z * x + x;
    float pc = ((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z * z
               - 0.5f * z + 1.0f;
    float s, c;
    switch (j) {
        case 0: s = ps;  c = pc;  break;
        case 2: s = pc;  c = -ps; break;
        case 4: s = -ps; c = -pc; break;
        default: s = -pc; c = ps; break;  // 6
    }
    if (xx < 0.0f) s = -s;
    *s_out = s; *c_out = c;
}
static inline float det_sinf(float x) { float s, c; det_sincosf(x, &s, &c); return s; }
static inline float det_cosf(float x) { float s, c; det_sincosf(x, &s, &c); return c; }

// Cephes atanf on t >= 0.
static inline float det_atan_pos(float t) {
    float y;
    if (t > 2.414213562373095f) { y = 1.5707963267948966f; t = -(1.0f / t); }
    else if (t > 0.4142135623730950f) { y = 0.7853981633974483f; t = (t - 1.0f) / (t + 1.0f); }
    else y = 0.0f;
    float z = t * t;
    y += (((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z * t + t;
    return y;
}
static inline float det_atan2f(float y, float x) {
    const float PI_F = 3.14159265358979323846f, PIO2_F = 1.5707963267948966f;
    if (x == 0.0f) {
        if (y > 0.0f) return PIO2_F;
        if (y < 0.0f) return -PIO2_F;
        return 0.0f;
    }
    if (y == 0.0f) return x > 0.0f ? 0.0f : PI_F;
    float a = det_atan_pos(fabsf(y / x));
    if (x < 0.0f) a = PI_F - a;
    return y < 0.0f ? -a : a;
}
// Cephes asinf, |x| <= 1 (values beyond are clamped by the caller's data:
// a normalised direction component).
static inline float det_asinf(float xx) {
    float a = fabsf(xx);
    if (a > 1.0f) a = 1.0f;
    float z, x;
    bool big = a > 0.5f;
    if (big) { z = 0.5f * (1.0f - a); x = sqrtf(z); }
    else { x = a; z = x * x; }
    float r = ((((4.2163199048e-2f * z + 2.4181311049e-2f) * z + 4.5470025998e-2f) * z + 7.4953002686e-2f) * z
               + 1.6666752422e-1f) * z * x + x;
    if (big) { r = r + r; r = 1.5707963267948966f - r; }
    return xx < 0.0f ? -r : r;
}
// (1-c)^5 as used by both Schlick terms (optixSphere.cu:483,491).  The
// reference calls powf(x, 5.0f); the oracle rule is x2=x*x; x4=x2*x2; x4*x.
static inline float det_pow5(float x) { float x2 = x * x; float x4 = x2 * x2; return x4 * x; }

// pow(x, y) of the display transform (optixSphere.cu:425 `pow(rgb, 1/2.2)`; SDK make_color's powf(c, 1/2.4f)).  The
// reference calls CUDA's powf; glibc's differs from it in the last bit for some inputs, which moved 8-bit frame values
// by one LSB.  detmath evaluates exp2(y * log2(x)) in IEEE DOUBLE precision with a fixed operation order over +,-,*,/
// (atanh series for the logarithm, Taylor series for the exponential, both truncated below 1e-16) and rounds ONCE to
// float: the correctly rounded x^y unless the exact value lies within ~1e-8 ulp of a rounding boundary.  Same code
// on the device (csrc/device_math.cuh), so frame bytes are bit-identical.
static inline double det_bits_to_double(uint64_t b) { double d; memcpy(&d, &b, 8); return d; }
static inline float det_powf(float xf, float yf) {
    if (yf == 0.0f || xf == 1.0f) return 1.0f;
    if (xf != xf || yf != yf) return xf + yf;
    if (xf < 0.0f) return xf * 0.0f / 0.0f;  // NaN; never produced by the display transform (inputs are clamped to [0, 1])
    if (xf == 0.0f) return yf > 0.0f ? 0.0f : 1.0f / 0.0f;
    if (xf > 3.0e38f) return yf > 0.0f ? xf : 0.0f;
    const double x = (double)xf;  // a float subnormal is a normal double
    uint64_t b; memcpy(&b, &x, 8);
    int e = (int)((b >> 52) & 0x7ffu) - 1023;
    double m = det_bits_to_double((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);  // [1, 2)
    if (m > 1.4142135623730951) { m = m * 0.5; e += 1; }
    const double s = (m - 1.0) / (m + 1.0), s2 = s * s;  // |s| <= 0.1716
    double p = 1.0 / 23.0;
    p = p * s2 + 1.0 / 21.0; p = p * s2 + 1.0 / 19.0; p = p * s2 + 1.0 / 17.0; p = p * s2 + 1.0 / 15.0;
    p = p * s2 + 1.0 / 13.0; p = p * s2 + 1.0 / 11.0; p = p * s2 + 1.0 / 9.0; p = p * s2 + 1.0 / 7.0;
    p = p * s2 + 1.0 / 5.0; p = p * s2 + 1.0 / 3.0; p = p * s2 + 1.0;
    const double log2x = (double)e + (2.0 * s * p) * 1.4426950408889634;
    const double t = (double)yf * log2x;
    if (t >= 128.0) return 1.0f / 0.0f;
    if (t < -160.0) return 0.0f;
    const double n = floor(t + 0.5);
    const double z = (t - n) * 0.6931471805599453;  // |z| <= 0.3466
    double q = 1.0 / 87178291200.0;  // 1/14!
    q = q * z + 1.0 / 6227020800.0; q = q * z + 1.0 / 479001600.0; q = q * z + 1.0 / 39916800.0; q = q * z + 1.0 / 3628800.0;
    q = q * z + 1.0 / 362880.0; q = q * z + 1.0 / 40320.0; q = q * z + 1.0 / 5040.0; q = q * z + 1.0 / 720.0;
    q = q * z + 1.0 / 120.0; q = q * z + 1.0 / 24.0; q = q * z + 1.0 / 6.0; q = q * z + 0.5; q = q * z + 1.0; q = q * z + 1.0;
    const double scale = det_bits_to_double((uint64_t)((long long)n + 1023ll) << 52);  // 2^n, -160 <= n <= 128
    return (float)(q * scale);
}

}  // namespace orc
