// oracle/ref_shim/shim_all.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Host stand-ins that let g++ compile the reference's device file
// /root/reference/optixSphere.cu UNMODIFIED, from where it lies, into
// oracle/_ref/libref_pt.so (recipe: oracle/Makefile).  The reference as
// shipped cannot be built here (no OptiX SDK, no sutil, no RT cores); this
// shim replaces exactly the pieces that are absent:
//   * CUDA vector types and qualifiers (__device__, __global__, ...);
//   * OptiX SDK 8.0.0 sutil/vec_math.h + cuda/helpers.h, restated from the
//     published SDK source as remembered (PARITY UNPINNED at this boundary);
//   * OptiX device intrinsics (optixTraverse/optixInvoke/optixGetPayload_N/
//     ...), implemented over thread-local state; optixTraverse asks the
//     oracle's closest-hit query (oracle_isect.h) instead of RT cores;
//   * CUDA libm: cosf/sinf/atan2f/asinf and powf(x,5) are routed to the
//     oracle's detmath so that reference-on-host and oracle can be compared
//     bit for bit; powf with any other exponent (the display transform) is
//     detmath's double-precision det_powf on both sides.
//   * make_floatN(...) are macros that brace-initialise, which sequences the
//     reference's `make_float2(myrnd(seed), myrnd(seed))` (cu:328, cu:15)
//     left to right (oracle rule R2; g++ would evaluate right to left).
#pragma once
#include <climits>
#include <cmath>
#include <math.h>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../oracle_math.h"

#define __forceinline__ inline
#define __device__
#define __host__
#define __global__
#define __constant__

struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };
struct uchar4 { unsigned char x, y, z, w; };
struct uint3 { unsigned int x, y, z; };

typedef unsigned long long OptixTraversableHandle;
typedef unsigned long long CUdeviceptr;
typedef unsigned int OptixPayloadTypeID;
typedef unsigned int OptixVisibilityMask;
enum { OPTIX_PAYLOAD_TYPE_ID_0 = 1u };
enum { OPTIX_RAY_FLAG_NONE = 0u, OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1u, OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 4u };

#define M_PIf 3.14159265358979323846f

// ---- make_floatN with a defined argument evaluation order -------------------
static inline float3 shim_mk3(float s) { return float3{s, s, s}; }
static inline float3 shim_mk3(const float4& v) { return float3{v.x, v.y, v.z}; }
static inline float3 shim_mk3(const float3& v) { return v; }
static inline float4 shim_mk4(const float3& v, float w) { return float4{v.x, v.y, v.z, w}; }
#define SHIM_SEL4(_1, _2, _3, _4, NAME, ...) NAME
#define SHIM_MK2_2(a, b) float2{(float)(a), (float)(b)}
#define make_float2(...) SHIM_MK2_2(__VA_ARGS__)
#define SHIM_MK3_3(a, b, c) float3{(float)(a), (float)(b), (float)(c)}
#define SHIM_MK3_1(a) shim_mk3(a)
#define SHIM_MK3_BAD(...) static_assert(false, "make_float3 arity")
#define make_float3(...) SHIM_SEL4(__VA_ARGS__, SHIM_MK3_BAD, SHIM_MK3_3, SHIM_MK3_BAD, SHIM_MK3_1)(__VA_ARGS__)
#define SHIM_MK4_4(a, b, c, d) float4{(float)(a), (float)(b), (float)(c), (float)(d)}
#define SHIM_MK4_2(a, b) shim_mk4(a, b)
#define make_float4(...) SHIM_SEL4(__VA_ARGS__, SHIM_MK4_4, SHIM_MK3_BAD, SHIM_MK4_2, SHIM_MK3_BAD)(__VA_ARGS__)

// ---- sutil/vec_math.h (restated) --------------------------------------------
static inline float2 operator*(const float2& a, float s) { return float2{a.x * s, a.y * s}; }
static inline float2 operator*(float s, const float2& a) { return float2{s * a.x, s * a.y}; }
static inline float2 operator+(const float2& a, const float2& b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 operator-(const float2& a, float s) { return float2{a.x - s, a.y - s}; }

static inline float3 operator+(const float3& a, const float3& b) { return float3{a.x + b.x, a.y + b.y, a.z + b.z}; }
static inline float3 operator-(const float3& a, const float3& b) { return float3{a.x - b.x, a.y - b.y, a.z - b.z}; }
static inline float3 operator-(const float3& a) { return float3{-a.x, -a.y, -a.z}; }
static inline float3 operator*(const float3& a, const float3& b) { return float3{a.x * b.x, a.y * b.y, a.z * b.z}; }
static inline float3 operator*(const float3& a, float s) { return float3{a.x * s, a.y * s, a.z * s}; }
static inline float3 operator*(float s, const float3& a) { return float3{s * a.x, s * a.y, s * a.z}; }
static inline float3 operator+(const float3& a, float s) { return float3{a.x + s, a.y + s, a.z + s}; }
static inline float3 operator+(float s, const float3& a) { return float3{s + a.x, s + a.y, s + a.z}; }
static inline float3 operator-(const float3& a, float s) { return float3{a.x - s, a.y - s, a.z - s}; }
static inline float3 operator/(const float3& a, const float3& b) { return float3{a.x / b.x, a.y / b.y, a.z / b.z}; }
static inline float3 operator/(const float3& a, float s) { float inv = 1.0f / s; return a * inv; }
static inline void operator+=(float3& a, const float3& b) { a.x += b.x; a.y += b.y; a.z += b.z; }
static inline void operator*=(float3& a, const float3& b) { a.x *= b.x; a.y *= b.y; a.z *= b.z; }
static inline void operator*=(float3& a, float s) { a.x *= s; a.y *= s; a.z *= s; }
static inline void operator/=(float3& a, float s) { float inv = 1.0f / s; a *= inv; }

static inline float4 operator+(const float4& a, const float4& b) { return float4{a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w}; }
static inline float4 operator-(const float4& a, const float4& b) { return float4{a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w}; }
static inline float4 operator*(float s, const float4& a) { return float4{s * a.x, s * a.y, s * a.z, s * a.w}; }

static inline float dot(const float3& a, const float3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline float3 cross(const float3& a, const float3& b) {
    return float3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
static inline float length(const float3& v) { return sqrtf(dot(v, v)); }
static inline float3 normalize(const float3& v) { float invLen = 1.0f / sqrtf(dot(v, v)); return v * invLen; }
static inline float lerp(float a, float b, float t) { return a + t * (b - a); }
static inline float3 lerp(const float3& a, const float3& b, float t) { return a + t * (b - a); }
static inline float4 lerp(const float4& a, const float4& b, float t) { return a + t * (b - a); }
static inline float clamp(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }
static inline float3 clamp(const float3& v, float a, float b) { return float3{clamp(v.x, a, b), clamp(v.y, a, b), clamp(v.z, a, b)}; }
static inline float3 reflect(const float3& i, const float3& n) { return i - 2.0f * n * dot(n, i); }
static inline float3 faceforward(const float3& n, const float3& i, const float3& nref) { return n * copysignf(1.0f, dot(i, nref)); }
// only referenced from the unreachable glass branch (cu:846)
static inline bool refract(float3& r, const float3& i, const float3& n, float ior) {
    float3 nn = n; float negNdotV = dot(i, nn); float eta;
    if (negNdotV > 0.0f) { eta = ior; nn = -n; negNdotV = -negNdotV; } else { eta = 1.0f / ior; }
    const float k = 1.0f - eta * eta * (1.0f - negNdotV * negNdotV);
    if (k < 0.0f) { r = float3{0.0f, 0.0f, 0.0f}; return false; }
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}

// ---- cuda/helpers.h make_color (restated) ------------------------------------
static inline float3 toSRGB(const float3& c) {
    float invGamma = 1.0f / 2.4f;
    float3 powed = float3{orc::det_powf(c.x, invGamma), orc::det_powf(c.y, invGamma), orc::det_powf(c.z, invGamma)};
    return float3{c.x < 0.0031308f ? 12.92f * c.x : 1.055f * powed.x - 0.055f,
                  c.y < 0.0031308f ? 12.92f * c.y : 1.055f * powed.y - 0.055f,
                  c.z < 0.0031308f ? 12.92f * c.z : 1.055f * powed.z - 0.055f};
}
static inline unsigned char quantizeUnsigned8Bits(float x) {
    x = clamp(x, 0.0f, 1.0f);
    unsigned int q = (unsigned int)(x * 256.0f);
    return (unsigned char)(q < 255u ? q : 255u);
}
static inline uchar4 make_color(const float3& c) {
    float3 srgb = toSRGB(clamp(c, 0.0f, 1.0f));
    return uchar4{quantizeUnsigned8Bits(srgb.x), quantizeUnsigned8Bits(srgb.y), quantizeUnsigned8Bits(srgb.z), 255u};
}

// ---- CUDA intrinsics -----------------------------------------------------------
static inline unsigned int __float_as_uint(float f) { unsigned int u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned int u) { float f; memcpy(&f, &u, 4); return f; }

// ---- libm routing (see header comment) ----------------------------------------
static inline float shim_powf(float a, float b) { return b == 5.0f ? orc::det_pow5(a) : orc::det_powf(a, b); }
#define cosf(x) orc::det_cosf(x)
#define sinf(x) orc::det_sinf(x)
#define atan2f(y, x) orc::det_atan2f(y, x)
#define asinf(x) orc::det_asinf(x)
#define powf(a, b) shim_powf(a, b)

// ---- OptiX device intrinsics over thread-local state ---------------------------
namespace refshim {
struct TraceResult { int prim; float t, b1, b2; };
typedef TraceResult (*TraceFn)(const void* user, const float o[3], const float d[3], float tmin, float tmax);
struct Global {
    TraceFn trace; const void* user;
    const void* raygen_data; const void* miss_data; const void* hitgroup_base; size_t hitgroup_stride;
    const uint32_t* mat_ids;
    int32_t* primary_hit; unsigned width;
};
struct Tls {
    unsigned int p[19];
    uint3 idx, dim;
    const void* sbt;
    float3 ray_o, ray_d;
    TraceResult hit;
    unsigned long long segments, hits, misses, paths;
    unsigned segs_this_pixel;
    unsigned segs_this_path;
};
extern Global g;
extern thread_local Tls tls;
struct PathHang {};
}  // namespace refshim

extern "C" void __closesthit__radiance();
extern "C" void __miss__radiance();

static inline uint3 optixGetLaunchIndex() { return refshim::tls.idx; }
static inline uint3 optixGetLaunchDimensions() { return refshim::tls.dim; }
static inline CUdeviceptr optixGetSbtDataPointer() { return (CUdeviceptr)(uintptr_t)refshim::tls.sbt; }
static inline unsigned int optixGetPrimitiveIndex() { return (unsigned int)refshim::tls.hit.prim; }
static inline float3 optixGetWorldRayDirection() { return refshim::tls.ray_d; }
static inline float3 optixGetWorldRayOrigin() { return refshim::tls.ray_o; }
static inline float optixGetRayTmax() { return refshim::tls.hit.t; }
static inline OptixTraversableHandle optixGetGASTraversableHandle() { return 0; }
static inline unsigned int optixGetSbtGASIndex() { return 0; }
static inline float2 optixGetTriangleBarycentrics() { return float2{refshim::tls.hit.b1, refshim::tls.hit.b2}; }
static inline void optixSetPayloadTypes(unsigned int) {}
static inline void optixReorder() {}
static inline bool optixHitObjectIsHit() { return refshim::tls.hit.prim >= 0; }

#define SHIM_PAYLOAD(N) \
    static inline unsigned int optixGetPayload_##N() { return refshim::tls.p[N]; } \
    static inline void optixSetPayload_##N(unsigned int v) { refshim::tls.p[N] = v; }
SHIM_PAYLOAD(0) SHIM_PAYLOAD(1) SHIM_PAYLOAD(2) SHIM_PAYLOAD(3) SHIM_PAYLOAD(4) SHIM_PAYLOAD(5) SHIM_PAYLOAD(6)
SHIM_PAYLOAD(7) SHIM_PAYLOAD(8) SHIM_PAYLOAD(9) SHIM_PAYLOAD(10) SHIM_PAYLOAD(11) SHIM_PAYLOAD(12) SHIM_PAYLOAD(13)
SHIM_PAYLOAD(14) SHIM_PAYLOAD(15) SHIM_PAYLOAD(16) SHIM_PAYLOAD(17) SHIM_PAYLOAD(18)

static inline void shim_do_trace(const float3& o, const float3& d, float tmin, float tmax) {
    using namespace refshim;
    float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    tls.ray_o = o; tls.ray_d = d;
    tls.hit = g.trace(g.user, oo, dd, tmin, tmax);
}

// optixTraverse with a payload type id and 19 payload registers (cu:99-112)
static inline void optixTraverse(OptixPayloadTypeID, OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax,
                                 float, OptixVisibilityMask, unsigned int, unsigned int, unsigned int, unsigned int,
                                 unsigned int& u0, unsigned int& u1, unsigned int& u2, unsigned int& u3, unsigned int& u4,
                                 unsigned int& u5, unsigned int& u6, unsigned int& u7, unsigned int& u8, unsigned int& u9,
                                 unsigned int& u10, unsigned int& u11, unsigned int& u12, unsigned int& u13,
                                 unsigned int& u14, unsigned int& u15, unsigned int& u16, unsigned int& u17,
                                 unsigned int& u18) {
    using namespace refshim;
    unsigned int v[19] = {u0, u1, u2, u3, u4, u5, u6, u7, u8, u9, u10, u11, u12, u13, u14, u15, u16, u17, u18};
    memcpy(tls.p, v, sizeof(v));
    shim_do_trace(o, d, tmin, tmax);
    tls.segments++;
    if (tls.segs_this_pixel == 0 && g.primary_hit) g.primary_hit[(size_t)tls.idx.y * g.width + tls.idx.x] = tls.hit.prim;
    tls.segs_this_pixel++;
    // The reference loops forever when done && !(p > 0) (cu:382-387); bail out.
    if (++tls.segs_this_path > 100000u) throw PathHang();
}
// probe-ray overload (cu:143-154, traceOcclusion; never called by the reference)
static inline void optixTraverse(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax, float,
                                 OptixVisibilityMask, unsigned int, unsigned int, unsigned int, unsigned int) {
    shim_do_trace(o, d, tmin, tmax);
}
static inline void optixInvoke(OptixPayloadTypeID, unsigned int& u0, unsigned int& u1, unsigned int& u2, unsigned int& u3,
                               unsigned int& u4, unsigned int& u5, unsigned int& u6, unsigned int& u7, unsigned int& u8,
                               unsigned int& u9, unsigned int& u10, unsigned int& u11, unsigned int& u12,
                               unsigned int& u13, unsigned int& u14, unsigned int& u15, unsigned int& u16,
                               unsigned int& u17, unsigned int& u18) {
    using namespace refshim;
    const void* saved = tls.sbt;
    if (tls.hit.prim >= 0) {
        tls.hits++;
        tls.sbt = (const char*)g.hitgroup_base + g.hitgroup_stride * g.mat_ids[tls.hit.prim];
        __closesthit__radiance();
    } else {
        tls.misses++;
        tls.sbt = g.miss_data;
        __miss__radiance();
    }
    tls.sbt = saved;
    unsigned int* r[19] = {&u0, &u1, &u2, &u3, &u4, &u5, &u6, &u7, &u8, &u9, &u10, &u11, &u12, &u13, &u14, &u15, &u16, &u17, &u18};
    for (int i = 0; i < 19; ++i) *r[i] = tls.p[i];
}
