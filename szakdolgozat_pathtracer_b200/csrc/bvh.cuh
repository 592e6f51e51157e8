// bvh.cuh -- flattened BVH layout in HBM/L2 and the closest-hit traversal.
//
// This replaces what the reference delegates to RT cores: optixTraverse over a
// single triangle GAS with the built-in triangle intersector
// (optixSphere.cu:99-112, optixSphere.cpp:897-913, 1007-1011).
//
// Layout (all 16-byte vector loads):
//   node  = 4 x float4 = 64 B (two children per node, boxes stored in the parent):
//     n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)
//     n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//     n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)
//     n3 = (child0, child1, 0, 0) as int bits; child >= 0: node index,
//          child < 0: leaf code, ~child = (first triangle << 3) | (count - 1),
//          1 <= count <= 8, first < 2^28
//   triangle (leaf order) = 3 x float4 = 48 B: (v0.xyz, original prim id bits),
//          (v1.xyz, 0), (v2.xyz, 0)
//   node 0 is the root and always an internal node.
//
// Hit rule (identical to oracle/oracle_isect.h): Woop/Benthin/Wald watertight
// test, tmin < t < tmax, closest t wins, equal t => lowest ORIGINAL prim id
// (so the result does not depend on traversal order), barycentrics weight
// vertices 1 and 2.
#pragma once
#include "device_math.cuh"

namespace ptb {

#define PTB_BVH_STACK 64

struct HitRec { float t, b1, b2; int prim; };

struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };

PTB_DEV RayShear ray_shear(float3 d) {
    RayShear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (comp(d, kz) < 0.0f) { int t = kx; kx = ky; ky = t; }
    r.kx = kx; r.ky = ky; r.kz = kz;
    float dz = comp(d, kz);
    r.Sx = comp(d, kx) / dz;
    r.Sy = comp(d, ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// true and (t,b1,b2) when tmin < t < tmax
PTB_DEV bool ray_tri(float3 org, const RayShear& rs, float3 p0, float3 p1, float3 p2, float tmin, float tmax,
                     float* t_out, float* b1_out, float* b2_out) {
    float3 A = p0 - org, B = p1 - org, C = p2 - org;
    float Akz = comp(A, rs.kz), Bkz = comp(B, rs.kz), Ckz = comp(C, rs.kz);
    float Ax = comp(A, rs.kx) - rs.Sx * Akz, Ay = comp(A, rs.ky) - rs.Sy * Akz;
    float Bx = comp(B, rs.kx) - rs.Sx * Bkz, By = comp(B, rs.ky) - rs.Sy * Bkz;
    float Cx = comp(C, rs.kx) - rs.Sx * Ckz, Cy = comp(C, rs.ky) - rs.Sy * Ckz;
    float U = Cx * By - Cy * Bx;
    float V = Ax * Cy - Ay * Cx;
    float W = Bx * Ay - By * Ax;
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = (double)Cx * (double)By, CyBx = (double)Cy * (double)Bx;
        U = (float)(CxBy - CyBx);
        double AxCy = (double)Ax * (double)Cy, AyCx = (double)Ay * (double)Cx;
        V = (float)(AxCy - AyCx);
        double BxAy = (double)Bx * (double)Ay, ByAx = (double)By * (double)Ax;
        W = (float)(BxAy - ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = U + V + W;
    if (det == 0.0f) return false;
    float Az = rs.Sz * Akz, Bz = rs.Sz * Bkz, Cz = rs.Sz * Ckz;
    float T = U * Az + V * Bz + W * Cz;
    float rdet = 1.0f / det;
    float t = T * rdet;
    if (!(t > tmin && t < tmax)) return false;
    *t_out = t; *b1_out = V * rdet; *b2_out = W * rdet;
    return true;
}

struct TravCounters { uint32_t nodes, tris; };

// Conservative slab test of one child box against [tmin, tbest].  Boxes are
// padded at build time and the far distance is widened by 2 ulp, so a box is
// never rejected when one of its triangles would pass ray_tri.
PTB_DEV bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, float3 o, float3 id, float tmin,
                  float tbest, float* tnear) {
    float t0 = (lox - o.x) * id.x, t1 = (hix - o.x) * id.x;
    float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
    t0 = (loy - o.y) * id.y; t1 = (hiy - o.y) * id.y;
    lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
    t0 = (loz - o.z) * id.z; t1 = (hiz - o.z) * id.z;
    lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
    hi = hi * 1.0000005f;
    lo = fmaxf(lo * 0.9999995f, tmin * 0.999f);
    *tnear = lo;
    return lo <= hi && lo <= tbest;
}

template <bool COUNT>
PTB_DEV HitRec bvh_closest_hit(const float4* __restrict__ nodes, const float4* __restrict__ tris, float3 o, float3 d,
                               float tmin, float tmax, TravCounters* cnt) {
    HitRec best; best.t = tmax; best.b1 = 0.0f; best.b2 = 0.0f; best.prim = -1;
    const RayShear rs = ray_shear(d);
    const float3 id = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int stack[PTB_BVH_STACK];  // one entry per level; the builder rejects trees deeper than the stack
    int sp = 0;
    int node = 0;  // current internal node (>= 0) or leaf code (< 0)
    const int SENTINEL = 0x7fffffff;
    stack[sp++] = SENTINEL;
    while (node != SENTINEL) {
        if (node >= 0) {
            const float4 n0 = __ldg(nodes + (size_t)node * 4 + 0);
            const float4 n1 = __ldg(nodes + (size_t)node * 4 + 1);
            const float4 n2 = __ldg(nodes + (size_t)node * 4 + 2);
            const float4 n3 = __ldg(nodes + (size_t)node * 4 + 3);
            if (COUNT) cnt->nodes++;
            float tn0, tn1;
            const bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, o, id, tmin, best.t, &tn0);
            const bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, o, id, tmin, best.t, &tn1);
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                // nearer child first, the other one goes on the stack
                const bool swap = tn1 < tn0;
                stack[sp++] = swap ? c0 : c1;
                node = swap ? c1 : c0;
            } else if (h0) node = c0;
            else if (h1) node = c1;
            else node = stack[--sp];
        } else {
            const int code = ~node;
            const int first = code >> 3, count = (code & 7) + 1;
            for (int i = 0; i < count; ++i) {
                const float4 a = __ldg(tris + (size_t)(first + i) * 3 + 0);
                const float4 b = __ldg(tris + (size_t)(first + i) * 3 + 1);
                const float4 c = __ldg(tris + (size_t)(first + i) * 3 + 2);
                if (COUNT) cnt->tris++;
                float t, b1, b2;
                // test against the ray's own tmax so that ties can be resolved by prim id
                if (ray_tri(o, rs, mk3(a), mk3(b), mk3(c), tmin, tmax, &t, &b1, &b2)) {
                    const int prim = __float_as_int(a.w);
                    if (t < best.t || (t == best.t && best.prim >= 0 && prim < best.prim)) {
                        best.t = t; best.b1 = b1; best.b2 = b2; best.prim = prim;
                    }
                }
            }
            node = stack[--sp];
        }
    }
    return best;
}

}  // namespace ptb
