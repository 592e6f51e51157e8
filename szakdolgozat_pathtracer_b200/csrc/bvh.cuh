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
#include "views.cuh"

namespace PTB_NS {

#define PTB_BVH_STACK 128

struct HitRec { float t, b1, b2; int prim; };

struct RayShear { int kx, ky, kz; float Sx, Sy, Sz; };

// Everything from here to ray_tri decides which triangle a ray hits, so it is written with the ex_* operations (IEEE
// round to nearest, immune to FMA contraction and to approximate division): bit-identical in the exact and the fast build.
// id = (1/d.x, 1/d.y, 1/d.z): Sz = 1 / d[kz] is one of its components (the same IEEE division), so it is not recomputed
PTB_DEV RayShear ray_shear(float3 d, float3 id) {
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
    r.Sx = ex_div(comp(d, kx), dz);
    r.Sy = ex_div(comp(d, ky), dz);
    r.Sz = comp(id, kz);
    return r;
}

// true and (t,b1,b2) when tmin < t < tmax
PTB_DEV bool ray_tri(float3 org, const RayShear& rs, float3 p0, float3 p1, float3 p2, float tmin, float tmax,
                     float* t_out, float* b1_out, float* b2_out) {
    const float3 A = mk3(ex_sub(p0.x, org.x), ex_sub(p0.y, org.y), ex_sub(p0.z, org.z));
    const float3 B = mk3(ex_sub(p1.x, org.x), ex_sub(p1.y, org.y), ex_sub(p1.z, org.z));
    const float3 C = mk3(ex_sub(p2.x, org.x), ex_sub(p2.y, org.y), ex_sub(p2.z, org.z));
    float Akz = comp(A, rs.kz), Bkz = comp(B, rs.kz), Ckz = comp(C, rs.kz);
    float Ax = ex_sub(comp(A, rs.kx), ex_mul(rs.Sx, Akz)), Ay = ex_sub(comp(A, rs.ky), ex_mul(rs.Sy, Akz));
    float Bx = ex_sub(comp(B, rs.kx), ex_mul(rs.Sx, Bkz)), By = ex_sub(comp(B, rs.ky), ex_mul(rs.Sy, Bkz));
    float Cx = ex_sub(comp(C, rs.kx), ex_mul(rs.Sx, Ckz)), Cy = ex_sub(comp(C, rs.ky), ex_mul(rs.Sy, Ckz));
    float U = ex_sub(ex_mul(Cx, By), ex_mul(Cy, Bx));
    float V = ex_sub(ex_mul(Ax, Cy), ex_mul(Ay, Cx));
    float W = ex_sub(ex_mul(Bx, Ay), ex_mul(By, Ax));
    if (U == 0.0f || V == 0.0f || W == 0.0f) {
        double CxBy = __dmul_rn((double)Cx, (double)By), CyBx = __dmul_rn((double)Cy, (double)Bx);
        U = (float)__dsub_rn(CxBy, CyBx);
        double AxCy = __dmul_rn((double)Ax, (double)Cy), AyCx = __dmul_rn((double)Ay, (double)Cx);
        V = (float)__dsub_rn(AxCy, AyCx);
        double BxAy = __dmul_rn((double)Bx, (double)Ay), ByAx = __dmul_rn((double)By, (double)Ax);
        W = (float)__dsub_rn(BxAy, ByAx);
    }
    if ((U < 0.0f || V < 0.0f || W < 0.0f) && (U > 0.0f || V > 0.0f || W > 0.0f)) return false;
    float det = ex_add(ex_add(U, V), W);
    if (det == 0.0f) return false;
    float Az = ex_mul(rs.Sz, Akz), Bz = ex_mul(rs.Sz, Bkz), Cz = ex_mul(rs.Sz, Ckz);
    float T = ex_add(ex_add(ex_mul(U, Az), ex_mul(V, Bz)), ex_mul(W, Cz));
    float rdet = ex_div(1.0f, det);
    float t = ex_mul(T, rdet);
    if (!(t > tmin && t < tmax)) return false;
    *t_out = t; *b1_out = ex_mul(V, rdet); *b2_out = ex_mul(W, rdet);
    return true;
}

struct TravCounters { uint32_t nodes, tris; };

// Prefetch of what a stack entry points at (node or first triangle of a leaf) into L1, issued when the entry is pushed:
// by the time it is popped the fetch has been under way for a whole subtree.  Only for trees that do not fit the caches
// (4-wide traversal, PTB_WIDE_BVH_MIN_TRIS): C4 +1.5 %.
PTB_DEV void trav_prefetch4(const float4* __restrict__ nodes4, const float4* __restrict__ tris, int code) {
#ifndef PTB_HOST_SIM   // (tests/host: the traversal compiled for the CPU)
    const void* a = code >= 0 ? (const void*)(nodes4 + (size_t)code * 8) : (const void*)(tris + (size_t)((~code) >> 3) * 3);
    asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
#endif
}

// Conservative slab test of one child box against [tmin, tbest].  Leaf boxes are padded at build time
// (bvh_build.cu: k_leaf_boxes) and the far distance is widened, so a box is never rejected when one of its
// triangles would pass ray_tri.  tminp = tmin * 0.999.
PTB_DEV bool slab(float lox, float hix, float loy, float hiy, float loz, float hiz, float3 o, float3 id, float tminp,
                  float tbest, float* tnear) {
    const float x0 = ex_mul(ex_sub(lox, o.x), id.x), x1 = ex_mul(ex_sub(hix, o.x), id.x);
    const float y0 = ex_mul(ex_sub(loy, o.y), id.y), y1 = ex_mul(ex_sub(hiy, o.y), id.y);
    const float z0 = ex_mul(ex_sub(loz, o.z), id.z), z1 = ex_mul(ex_sub(hiz, o.z), id.z);
    const float lo = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fmaxf(fminf(z0, z1), tminp));
    const float hi = ex_mul(fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1)), 1.000001f);
    *tnear = lo;
    return lo <= hi && lo <= tbest;
}

#define PTB_TRAV_SENTINEL 0x7fffffff
#ifndef PTB_TRACE_QUANTUM
#define PTB_TRACE_QUANTUM 8  // traversal steps between two ray-fetch points (measured: 8 >= 16 > 4 on C2)
#endif

// Traversal state of one ray.  The stack lives in the caller's local memory (one entry per tree level; the
// builder rejects trees deeper than PTB_BVH_STACK).
struct Trav {
    float3 o, id;
    RayShear rs;
    float tmin, tminp, tmax;
    HitRec best;
    int node;  // current internal node (>= 0), leaf code (< 0) or PTB_TRAV_SENTINEL when finished
               // (8-wide traversal: child_base of the current node group)
    int sp;
    unsigned grp;  // 8-wide traversal only: hit bits (31..24) | imask (7..0) of the current node group
};

PTB_DEV void trav_begin(Trav& t, int* stack, float3 o, float3 d, float tmin, float tmax) {
    t.o = o; t.id = mk3(ex_div(1.0f, d.x), ex_div(1.0f, d.y), ex_div(1.0f, d.z));
    t.rs = ray_shear(d, t.id);
    t.tmin = tmin; t.tminp = ex_mul(tmin, 0.999f); t.tmax = tmax;
    t.best.t = tmax; t.best.b1 = 0.0f; t.best.b2 = 0.0f; t.best.prim = -1;
    stack[0] = PTB_TRAV_SENTINEL;
    t.sp = 1;
    t.node = 0;
    t.grp = 0u;
}

// ray octant of the 8-wide traversal: bit set <=> the ray travels towards + on that axis (sign of 1 / d, so that -0 counts as -)
PTB_DEV unsigned ray_octant(float3 id) { return (id.x < 0.0f ? 0u : 1u) | (id.y < 0.0f ? 0u : 2u) | (id.z < 0.0f ? 0u : 4u); }

// 8-wide traversal: the root is the only member of a pseudo group (child_base 0, imask 1, its hit bit set); empty stack
PTB_DEV void trav_begin8(Trav& t, float3 o, float3 d, float tmin, float tmax) {
    t.o = o; t.id = mk3(ex_div(1.0f, d.x), ex_div(1.0f, d.y), ex_div(1.0f, d.z));
    t.rs = ray_shear(d, t.id);
    t.tmin = tmin; t.tminp = ex_mul(tmin, 0.999f); t.tmax = tmax;
    t.best.t = tmax; t.best.b1 = 0.0f; t.best.b2 = 0.0f; t.best.prim = -1;
    t.sp = 0;
    t.node = 0;
    t.grp = (1u << (24u + ray_octant(t.id))) | 1u;
}

// tests the triangles of one leaf against the ray (closest-hit rule of the header comment)
template <bool COUNT>
PTB_DEV void trav_leaf(Trav& t, const float4* __restrict__ tris, int code, TravCounters* cnt) {
    const int first = code >> 3, count = (code & 7) + 1;
    for (int i = 0; i < count; ++i) {
        const float4* tp = tris + (size_t)(first + i) * 3;
        const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
        if (COUNT) cnt->tris++;
        float th, b1, b2;
        // test against the ray's own tmax so that ties can be resolved by prim id
        if (ray_tri(t.o, t.rs, mk3(a), mk3(b), mk3(c), t.tmin, t.tmax, &th, &b1, &b2)) {
            const int prim = __float_as_int(a.w);
            if (th < t.best.t || (th == t.best.t && t.best.prim >= 0 && prim < t.best.prim)) {
                t.best.t = th; t.best.b1 = b1; t.best.b2 = b2; t.best.prim = prim;
            }
        }
    }
}

// Runs at most `budget` steps ("while-while": descend through internal nodes until a leaf is reached, then test
// the leaf's triangles).  Returns true when the ray is finished.
template <bool COUNT>
PTB_DEV bool trav_run(Trav& t, int* stack, const float4* __restrict__ nodes, const float4* __restrict__ tris, int budget,
                      TravCounters* cnt) {
    while (budget > 0) {
        while ((unsigned)t.node < (unsigned)PTB_TRAV_SENTINEL && budget > 0) {
            const float4* np = nodes + (size_t)t.node * 4;
            const float4 n0 = __ldg(np + 0), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3);
            if (COUNT) cnt->nodes++;
            float tn0, tn1;
            const bool h0 = slab(n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, t.o, t.id, t.tminp, t.best.t, &tn0);
            const bool h1 = slab(n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, t.o, t.id, t.tminp, t.best.t, &tn1);
            const int c0 = __float_as_int(n3.x), c1 = __float_as_int(n3.y);
            if (h0 && h1) {
                const bool swap = tn1 < tn0;  // nearer child first, the other one goes on the stack
                stack[t.sp++] = swap ? c0 : c1;
                t.node = swap ? c1 : c0;
            } else if (h0) t.node = c0;
            else if (h1) t.node = c1;
            else t.node = stack[--t.sp];
            --budget;
        }
        if (t.node == PTB_TRAV_SENTINEL) return true;
        if (t.node < 0) {
            trav_leaf<COUNT>(t, tris, ~t.node, cnt);
            t.node = stack[--t.sp];
            budget -= 2;
            if (t.node == PTB_TRAV_SENTINEL) return true;
        }
    }
    return false;
}

// 4-wide traversal over the collapsed tree (bvh_build.cu: k_collapse4).  Node = 8 x float4 (128 B):
//   q0 lo.x[4]  q1 hi.x[4]  q2 lo.y[4]  q3 hi.y[4]  q4 lo.z[4]  q5 hi.z[4]  q6 child codes[4]  (q7 pad)
// Child codes as in the 2-wide tree (>= 0: node index, < 0: leaf code); unused slots carry NaN boxes.  Same hit rule, same
// conservative slab test, so the result is identical to trav_run; the children that are hit are visited nearest first
// (the others go on the stack farthest first).  One step costs about 2.3 x a 2-wide step but a ray needs half as many
// DEPENDENT node fetches.
#define PTB_SWAP_IF(c, a, b, ia, ib) do { if (c) { const float tf_ = a; a = b; b = tf_; const int ti_ = ia; ia = ib; ib = ti_; } } while (0)
template <bool COUNT>
PTB_DEV bool trav_run4(Trav& t, int* stack, const float4* __restrict__ nodes4, const float4* __restrict__ tris, int budget,
                       TravCounters* cnt) {
    while (budget > 0) {
        while ((unsigned)t.node < (unsigned)PTB_TRAV_SENTINEL && budget > 0) {
            const float4* np = nodes4 + (size_t)t.node * 8;
            const float4 lx = __ldg(np + 0), hx = __ldg(np + 1), ly = __ldg(np + 2), hy = __ldg(np + 3), lz = __ldg(np + 4), hz = __ldg(np + 5);
            const float4 cc = __ldg(np + 6);
            if (COUNT) cnt->nodes++;
            float d0, d1, d2, d3;
            const bool h0 = slab(lx.x, hx.x, ly.x, hy.x, lz.x, hz.x, t.o, t.id, t.tminp, t.best.t, &d0);
            const bool h1 = slab(lx.y, hx.y, ly.y, hy.y, lz.y, hz.y, t.o, t.id, t.tminp, t.best.t, &d1);
            const bool h2 = slab(lx.z, hx.z, ly.z, hy.z, lz.z, hz.z, t.o, t.id, t.tminp, t.best.t, &d2);
            const bool h3 = slab(lx.w, hx.w, ly.w, hy.w, lz.w, hz.w, t.o, t.id, t.tminp, t.best.t, &d3);
            const float far = 3.4e38f;
            if (!h0) d0 = far; if (!h1) d1 = far; if (!h2) d2 = far; if (!h3) d3 = far;
            int c0 = __float_as_int(cc.x), c1 = __float_as_int(cc.y), c2 = __float_as_int(cc.z), c3 = __float_as_int(cc.w);
            // sorting network, ascending entry distance (misses sink to the end)
            PTB_SWAP_IF(d1 < d0, d0, d1, c0, c1); PTB_SWAP_IF(d3 < d2, d2, d3, c2, c3);
            PTB_SWAP_IF(d2 < d0, d0, d2, c0, c2); PTB_SWAP_IF(d3 < d1, d1, d3, c1, c3);
            PTB_SWAP_IF(d2 < d1, d1, d2, c1, c2);
            const int nh = (int)h0 + (int)h1 + (int)h2 + (int)h3;
            if (nh == 0) t.node = stack[--t.sp];
            else {
                if (nh > 3) { stack[t.sp++] = c3; trav_prefetch4(nodes4, tris, c3); }
                if (nh > 2) { stack[t.sp++] = c2; trav_prefetch4(nodes4, tris, c2); }
                if (nh > 1) { stack[t.sp++] = c1; trav_prefetch4(nodes4, tris, c1); }
                t.node = c0;
            }
            --budget;
        }
        if (t.node == PTB_TRAV_SENTINEL) return true;
        if (t.node < 0) {
            trav_leaf<COUNT>(t, tris, ~t.node, cnt);
            t.node = stack[--t.sp];
            budget -= 2;
            if (t.node == PTB_TRAV_SENTINEL) return true;
        }
    }
    return false;
}

// ---- 8-wide traversal over the quantised tree (bvh8.cuh: node format, bvh_build.cu: k_collapse8_level) ------------------
// A stack entry is a NODE GROUP: (child_base, hit bits 31..24 | imask 7..0) -- the children of one node that the ray
// hit and has not entered yet.  Hit bits are stored at 24 + (slot ^ ray octant): the highest set bit is the child whose
// slot lies farthest AGAINST the ray direction, i.e. the one the ray reaches first, so children are entered in (approximate)
// front-to-back order without sorting distances.  Triangles of the leaf children that were hit come out of the same node
// test as a 24-bit mask over the node's contiguous triangle range and are tested at once.
// The box test is conservative by construction: (p - o) / d is bounded from below and above with directed rounding, the
// quantised planes are added with fma.rd / fma.ru, and the quantised boxes contain the 2-wide tree's (padded) boxes.
// Same hit rule as trav_run, so the result is identical.
// byte k of w as a float without the conversion unit (I2F runs at a quarter of the FP32 rate and a node test needs 48 of
// them): the byte is permuted into the mantissa of 2^23 and 2^23 is subtracted -- both exact
PTB_DEV float u8f(unsigned w, int k) {
#ifdef PTB_HOST_SIM
    return (float)((w >> (8 * k)) & 0xffu);
#else
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7650u + (unsigned)k)) - 8388608.0f;
#endif
}

template <bool COUNT>
PTB_DEV bool trav_run8(Trav& t, int* stack_i, const uint4* __restrict__ nodes8, const float4* __restrict__ tris8, int budget,
                       TravCounters* cnt) {
    uint2* stack = reinterpret_cast<uint2*>(stack_i);   // callers align the array to 16 bytes
    const unsigned r = ray_octant(t.id);
    const bool negx = t.id.x < 0.0f, negy = t.id.y < 0.0f, negz = t.id.z < 0.0f;
    // 1 / d clamped to +-2^100 for the box tests: an axis-parallel ray (1 / d = inf) would turn q * a + b into NaN; with a huge
    // finite slope the planes of a slab the origin lies in are at -huge / +huge (pass) and both at +-huge otherwise (miss)
    const float idx = fminf(fmaxf(t.id.x, -1.2676506e30f), 1.2676506e30f), idy = fminf(fmaxf(t.id.y, -1.2676506e30f), 1.2676506e30f),
                idz = fminf(fmaxf(t.id.z, -1.2676506e30f), 1.2676506e30f);
    unsigned gbase = (unsigned)t.node, ghits = t.grp;
    while (budget > 0) {
        unsigned tbits = 0u, tbase = 0u;
        while (budget > 0) {
            if (ghits <= 0x00ffffffu) {   // the current group has no child left
                if (t.sp == 0) { t.node = PTB_TRAV_SENTINEL; return true; }
                const uint2 g = stack[--t.sp];
                gbase = g.x; ghits = g.y;
            }
            const int bit = 31 - __clz(ghits);
            ghits &= ~(1u << bit);
            const unsigned slot = (unsigned)(bit - 24) ^ r;
            const unsigned ni = gbase + (unsigned)__popc(ghits & 0xffu & ((1u << slot) - 1u));
            if (ghits > 0x00ffffffu) stack[t.sp++] = make_uint2(gbase, ghits);
            const uint4* np = nodes8 + (size_t)ni * 5;
            const uint4 q0 = __ldg(np + 0), q1 = __ldg(np + 1), q2 = __ldg(np + 2), q3 = __ldg(np + 3), q4 = __ldg(np + 4);
            if (COUNT) cnt->nodes++;
            const unsigned ew = q0.w;
            // a = grid step / d (exact: a power of two times 1 / d), [bn, bf] encloses (p - o) / d
            const float ax = ex_mul(__uint_as_float((ew & 0xffu) << 23), idx), ay = ex_mul(__uint_as_float(((ew >> 8) & 0xffu) << 23), idy),
                        az = ex_mul(__uint_as_float(((ew >> 16) & 0xffu) << 23), idz);
            const float px = __uint_as_float(q0.x), py = __uint_as_float(q0.y), pz = __uint_as_float(q0.z);
            const float rdx = __fsub_rd(px, t.o.x), rux = __fsub_ru(px, t.o.x), rdy = __fsub_rd(py, t.o.y), ruy = __fsub_ru(py, t.o.y),
                        rdz = __fsub_rd(pz, t.o.z), ruz = __fsub_ru(pz, t.o.z);
            const float bnx = __fmul_rd(negx ? rux : rdx, idx), bfx = __fmul_ru(negx ? rdx : rux, idx);
            const float bny = __fmul_rd(negy ? ruy : rdy, idy), bfy = __fmul_ru(negy ? rdy : ruy, idy);
            const float bnz = __fmul_rd(negz ? ruz : rdz, idz), bfz = __fmul_ru(negz ? rdz : ruz, idz);
            const unsigned r4 = r * 0x01010101u;
            unsigned hitmask = 0u;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                // slots 4h .. 4h + 3: near / far plane bytes by ray direction
                const unsigned lox = h ? q2.y : q2.x, loy = h ? q2.w : q2.z, loz = h ? q3.y : q3.x;
                const unsigned hix = h ? q3.w : q3.z, hiy = h ? q4.y : q4.x, hiz = h ? q4.w : q4.z;
                const unsigned nx = negx ? hix : lox, fx = negx ? lox : hix;
                const unsigned ny = negy ? hiy : loy, fy = negy ? loy : hiy;
                const unsigned nz = negz ? hiz : loz, fz = negz ? loz : hiz;
                const unsigned meta4 = h ? q1.w : q1.z;
                const unsigned is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
                const unsigned inner_mask4 = (is_inner4 >> 4) * 0xffu;
                const unsigned bit_index4 = (meta4 ^ (r4 & inner_mask4)) & 0x1f1f1f1fu;
                const unsigned child_bits4 = (meta4 >> 5) & 0x07070707u;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float tn = fmaxf(fmaxf(__fmaf_rd(u8f(nx, k), ax, bnx), __fmaf_rd(u8f(ny, k), ay, bny)),
                                           fmaxf(__fmaf_rd(u8f(nz, k), az, bnz), t.tminp));
                    const float tf = ex_mul(fminf(fminf(__fmaf_ru(u8f(fx, k), ax, bfx), __fmaf_ru(u8f(fy, k), ay, bfy)),
                                                  __fmaf_ru(u8f(fz, k), az, bfz)), 1.000001f);
                    if (tn <= tf && tn <= t.best.t) hitmask |= ((child_bits4 >> (8 * k)) & 0xffu) << ((bit_index4 >> (8 * k)) & 0xffu);
                }
            }
            gbase = q1.x; ghits = (hitmask & 0xff000000u) | (ew >> 24);
            tbase = q1.y; tbits = hitmask & 0x00ffffffu;
            --budget;
            if (tbits) break;
        }
        if (tbits) {
            do {
                const int j = __ffs((int)tbits) - 1;
                tbits &= tbits - 1u;
                const float4* tp = tris8 + (size_t)(tbase + (unsigned)j) * 3;
                const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
                if (COUNT) cnt->tris++;
                float th, b1, b2;
                if (ray_tri(t.o, t.rs, mk3(a), mk3(b), mk3(c), t.tmin, t.tmax, &th, &b1, &b2)) {
                    const int prim = __float_as_int(a.w);
                    if (th < t.best.t || (th == t.best.t && t.best.prim >= 0 && prim < t.best.prim)) {
                        t.best.t = th; t.best.b1 = b1; t.best.b2 = b2; t.best.prim = prim;
                    }
                }
            } while (tbits);
            budget -= 2;
        }
        if (ghits <= 0x00ffffffu && t.sp == 0) { t.node = PTB_TRAV_SENTINEL; return true; }
    }
    t.node = (int)gbase; t.grp = ghits;
    return false;
}

// WIDTH = 2, 4, 8: the tree the kernel was instantiated for; 0: by what the scene was built with (uniform over the launch)
template <int WIDTH>
PTB_DEV void trav_begin_any(Trav& t, int* stack, const ptbv::SceneView& s, float3 o, float3 d, float tmin, float tmax) {
    if (WIDTH == 8 || (WIDTH == 0 && s.nodes8)) trav_begin8(t, o, d, tmin, tmax);
    else trav_begin(t, stack, o, d, tmin, tmax);
}

template <bool COUNT, int WIDTH>
PTB_DEV bool trav_run_any(Trav& t, int* stack, const ptbv::SceneView& s, int budget, TravCounters* cnt) {
    if (WIDTH == 8 || (WIDTH == 0 && s.nodes8)) return trav_run8<COUNT>(t, stack, s.nodes8, s.tris8, budget, cnt);
    if (WIDTH == 4 || (WIDTH == 0 && s.nodes4)) return trav_run4<COUNT>(t, stack, s.nodes4, s.tris, budget, cnt);
    return trav_run<COUNT>(t, stack, s.nodes, s.tris, budget, cnt);
}

template <bool COUNT>
PTB_DEV HitRec bvh_closest_hit(const ptbv::SceneView& s, float3 o, float3 d, float tmin, float tmax, TravCounters* cnt) {
    __align__(16) int stack[PTB_BVH_STACK];
    Trav t;
    trav_begin_any<0>(t, stack, s, o, d, tmin, tmax);
    while (!trav_run_any<COUNT, 0>(t, stack, s, 1 << 20, cnt)) {}
    return t.best;
}

}  // namespace PTB_NS
