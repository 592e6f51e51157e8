// oracle/oracle_isect.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Closest-hit query of the CPU oracle.  The reference delegates this to OptiX
// (optixTraverse over a triangle GAS with the built-in triangle intersector,
// optixSphere.cu:99-112, optixSphere.cpp:897-913,1007-1011), which is closed
// source: PARITY UNPINNED.  The oracle rule:
//   * ray/triangle test = Woop/Benthin/Wald 2013 watertight test, single
//     precision with the double-precision fallback on zero edge functions;
//   * a hit needs tmin < t < tmax;
//   * closest t wins, equal t => lowest original primitive index;
//   * barycentrics (b1,b2) weight vertices 1 and 2 (as optixGetTriangleBarycentrics).
// Brute force is the definition; the BVH below must return the same answer
// (tests check it) and only exists so that the oracle finishes in seconds.
#pragma once
#include <algorithm>
#include <cstdint>
#include <vector>
#include "oracle_math.h"

namespace orc {

struct RayShear {
    int kx, ky, kz;
    float Sx, Sy, Sz;
};

static inline float comp(const v3& v, int k) { return k == 0 ? v.x : (k == 1 ? v.y : v.z); }

static inline RayShear ray_shear(v3 d) {
    RayShear r;
    float ax = fabsf(d.x), ay = fabsf(d.y), az = fabsf(d.z);
    int kz = 0; float m = ax;
    if (ay > m) { kz = 1; m = ay; }
    if (az > m) { kz = 2; }
    int kx = kz + 1; if (kx == 3) kx = 0;
    int ky = kx + 1; if (ky == 3) ky = 0;
    if (comp(d, kz) < 0.0f) std::swap(kx, ky);
    r.kx = kx; r.ky = ky; r.kz = kz;
    float dz = comp(d, kz);
    r.Sx = comp(d, kx) / dz;
    r.Sy = comp(d, ky) / dz;
    r.Sz = 1.0f / dz;
    return r;
}

// Returns true and fills t,b1,b2 when tmin < t < tmax.
static inline bool ray_tri(const v3& org, const RayShear& rs, const v3& p0, const v3& p1, const v3& p2,
                           float tmin, float tmax, float* t_out, float* b1_out, float* b2_out) {
    v3 A = p0 - org, B = p1 - org, C = p2 - org;
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

struct Hit { int32_t prim; float t, b1, b2; };

struct TriSoup {
    const float* verts;  // float4[3N]
    uint32_t n;
    v3 vert(uint32_t prim, int k) const {
        const float* p = verts + (size_t)(prim * 3 + k) * 4;
        return mk3(p[0], p[1], p[2]);
    }
};

static inline Hit closest_brute(const TriSoup& s, v3 org, v3 dir, float tmin, float tmax) {
    Hit h = {-1, tmax, 0.0f, 0.0f};
    RayShear rs = ray_shear(dir);
    for (uint32_t i = 0; i < s.n; ++i) {
        float t, b1, b2;
        if (ray_tri(org, rs, s.vert(i, 0), s.vert(i, 1), s.vert(i, 2), tmin, h.t, &t, &b1, &b2)) {
            h.prim = (int32_t)i; h.t = t; h.b1 = b1; h.b2 = b2;  // strict t < h.t: lowest id wins ties
        }
    }
    return h;
}

// ---- the oracle's own CPU BVH (median-of-centroid split on the widest axis,
// leaves <= 4 triangles, generously padded boxes).  Not the product's builder.
struct CpuBvh {
    struct Node { float lo[3], hi[3]; int32_t left, right; int32_t first, count; };
    std::vector<Node> nodes;
    std::vector<uint32_t> order;

    void build(const TriSoup& s) {
        nodes.clear(); order.resize(s.n);
        std::vector<float> cent((size_t)s.n * 3), blo((size_t)s.n * 3), bhi((size_t)s.n * 3);
        for (uint32_t i = 0; i < s.n; ++i) {
            order[i] = i;
            v3 a = s.vert(i, 0), b = s.vert(i, 1), c = s.vert(i, 2);
            float lo[3] = {std::min(a.x, std::min(b.x, c.x)), std::min(a.y, std::min(b.y, c.y)), std::min(a.z, std::min(b.z, c.z))};
            float hi[3] = {std::max(a.x, std::max(b.x, c.x)), std::max(a.y, std::max(b.y, c.y)), std::max(a.z, std::max(b.z, c.z))};
            float mag = 0.0f;
            for (int k = 0; k < 3; ++k) mag = std::max(mag, std::max(fabsf(lo[k]), fabsf(hi[k])));
            float pad = mag * 1e-4f + 1e-6f;
            for (int k = 0; k < 3; ++k) {
                blo[(size_t)i * 3 + k] = lo[k] - pad; bhi[(size_t)i * 3 + k] = hi[k] + pad;
                cent[(size_t)i * 3 + k] = 0.5f * (lo[k] + hi[k]);
            }
        }
        if (s.n == 0) return;
        nodes.reserve((size_t)s.n);
        build_rec(0, s.n, cent, blo, bhi);
    }
    int32_t build_rec(uint32_t first, uint32_t count, const std::vector<float>& cent,
                      const std::vector<float>& blo, const std::vector<float>& bhi) {
        Node nd; nd.left = nd.right = -1; nd.first = (int32_t)first; nd.count = (int32_t)count;
        float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
        for (int k = 0; k < 3; ++k) { nd.lo[k] = 1e30f; nd.hi[k] = -1e30f; }
        for (uint32_t i = first; i < first + count; ++i) {
            uint32_t p = order[i];
            for (int k = 0; k < 3; ++k) {
                nd.lo[k] = std::min(nd.lo[k], blo[(size_t)p * 3 + k]); nd.hi[k] = std::max(nd.hi[k], bhi[(size_t)p * 3 + k]);
                clo[k] = std::min(clo[k], cent[(size_t)p * 3 + k]); chi[k] = std::max(chi[k], cent[(size_t)p * 3 + k]);
            }
        }
        int32_t me = (int32_t)nodes.size();
        nodes.push_back(nd);
        if (count <= 4) return me;
        int ax = 0; float ext = chi[0] - clo[0];
        for (int k = 1; k < 3; ++k) if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; ax = k; }
        uint32_t mid = first + count / 2;
        if (ext > 0.0f) {
            std::nth_element(order.begin() + first, order.begin() + mid, order.begin() + first + count,
                             [&](uint32_t a, uint32_t b) { return cent[(size_t)a * 3 + ax] < cent[(size_t)b * 3 + ax]; });
        }
        int32_t l = build_rec(first, mid - first, cent, blo, bhi);
        int32_t r = build_rec(mid, first + count - mid, cent, blo, bhi);
        nodes[me].left = l; nodes[me].right = r; nodes[me].count = 0;
        return me;
    }
    static inline bool slab(const Node& n, v3 o, v3 id, float tmin, float tmax, float* tn) {
        float t0 = (n.lo[0] - o.x) * id.x, t1 = (n.hi[0] - o.x) * id.x;
        float lo = fminf(t0, t1), hi = fmaxf(t0, t1);
        t0 = (n.lo[1] - o.y) * id.y; t1 = (n.hi[1] - o.y) * id.y;
        lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
        t0 = (n.lo[2] - o.z) * id.z; t1 = (n.hi[2] - o.z) * id.z;
        lo = fmaxf(lo, fminf(t0, t1)); hi = fminf(hi, fmaxf(t0, t1));
        hi *= 1.0000005f;  // conservative
        lo = fmaxf(lo, tmin * 0.999f);
        *tn = lo;
        return lo <= hi && lo <= tmax && hi >= tmin * 0.999f;
    }
    Hit closest(const TriSoup& s, v3 org, v3 dir, float tmin, float tmax) const {
        Hit h = {-1, tmax, 0.0f, 0.0f};
        if (nodes.empty()) return h;
        RayShear rs = ray_shear(dir);
        v3 id = mk3(1.0f / dir.x, 1.0f / dir.y, 1.0f / dir.z);
        int32_t stack[128]; int sp = 0;
        stack[sp++] = 0;
        while (sp) {
            const Node& n = nodes[stack[--sp]];
            float tn;
            if (!slab(n, org, id, tmin, h.t, &tn)) continue;
            if (n.left < 0) {
                for (int32_t i = n.first; i < n.first + n.count; ++i) {
                    uint32_t p = order[i];
                    float t, b1, b2;
                    // accept t <= h.t so that ties can be resolved by primitive index
                    if (ray_tri(org, rs, s.vert(p, 0), s.vert(p, 1), s.vert(p, 2), tmin, tmax, &t, &b1, &b2)) {
                        if (t < h.t || (t == h.t && h.prim >= 0 && (int32_t)p < h.prim)) {
                            h.prim = (int32_t)p; h.t = t; h.b1 = b1; h.b2 = b2;
                        }
                    }
                }
            } else {
                stack[sp++] = n.left; stack[sp++] = n.right;
            }
        }
        return h;
    }
};

}  // namespace orc
