// bvh8.cuh -- the 8-wide, quantised copy of the BVH ("nodes8"): node format and the collapse of the 2-wide tree into it.
//
// Part of what replaces optixAccelBuild (optixSphere.cpp:917-967); the layout follows the compressed wide BVH of
// Ylitie, Karras, Laine, "Efficient Incoherent Ray Traversal on GPUs Through Compressed Wide BVHs" (HPG 2017), restated:
//
//   node = 5 x uint4 = 80 B
//     q0 = ( p.x, p.y, p.z (float bits),  ex | ey << 8 | ez << 16 | imask << 24 )
//     q1 = ( child_base, tri_base, meta[0..3], meta[4..7] )
//     q2 = ( qlo.x[0..3], qlo.x[4..7], qlo.y[0..3], qlo.y[4..7] )      one byte per slot
//     q3 = ( qlo.z[0..3], qlo.z[4..7], qhi.x[0..3], qhi.x[4..7] )
//     q4 = ( qhi.y[0..3], qhi.y[4..7], qhi.z[0..3], qhi.z[4..7] )
//   p = low corner of the node's box; e* = biased exponents of the per-axis grid step s = 2^(e - 127) with 255 s >= extent;
//   the box of the child in slot k is [p + qlo[k] s, p + qhi[k] s] -- rounded OUTWARDS from the 2-wide tree's child box, so a
//   ray that enters the original box enters the quantised one.
//   Internal children are stored contiguously: the child in slot k is node child_base + popc(imask & ((1 << k) - 1)).
//   Triangles of the node's leaf children are contiguous in tris8 (a re-ordered copy of the leaf triangles): a node holds
//   at most 24 (8 leaves of <= 3), meta[k] = (unary count << 5) | offset for a leaf (count 1..3 -> 001 / 011 / 111),
//   (1 << 5) | (24 + k) for an internal child, 0 for an empty slot (whose box is inverted: qlo = 255, qhi = 0).
//   Slots are assigned by octant affinity (slot bit set on an axis <=> the child lies towards + on that axis), so that the
//   traversal can order the children a ray hits by (slot ^ ray octant) instead of sorting distances.
//
// collapse8_node and collapse4_node (the greedy 4-wide collapse, uncompressed 128-byte nodes of bvh.cuh) are plain C++
// (host + device): the build kernels (bvh_build.cu: k_collapse8_level, k_collapse4_level) and the CPU test harness
// (tests/host/bvh8_host.cpp) run the same code.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define PTB8_HD __host__ __device__ __forceinline__
#else
#define PTB8_HD inline
#endif

namespace ptb8 {

struct WorkItem { int wide, bin; };   // 8-wide node to write, 2-wide node it is expanded from

// error bits (sticky, reported by the builder)
enum { ERR_LEAF_TOO_BIG = 1, ERR_BAD_BOX = 2 };

PTB8_HD uint32_t f2u(float f) { union { float f; uint32_t u; } c; c.f = f; return c.u; }
PTB8_HD int f2i(float f) { return (int)f2u(f); }

// boxes and codes of the two children of 2-wide node `b` (layout in bvh.cuh)
PTB8_HD void load2(const float4* nodes2, int b, float lo0[3], float hi0[3], int* c0, float lo1[3], float hi1[3], int* c1) {
    const float4 n0 = nodes2[(size_t)b * 4 + 0], n1 = nodes2[(size_t)b * 4 + 1], n2 = nodes2[(size_t)b * 4 + 2], n3 = nodes2[(size_t)b * 4 + 3];
    lo0[0] = n0.x; hi0[0] = n0.y; lo0[1] = n0.z; hi0[1] = n0.w; lo0[2] = n2.x; hi0[2] = n2.y;
    lo1[0] = n1.x; hi1[0] = n1.y; lo1[1] = n1.z; hi1[1] = n1.w; lo1[2] = n2.z; hi1[2] = n2.w;
    *c0 = f2i(n3.x); *c1 = f2i(n3.y);
}

// 4-wide collapse of 2-wide node `b`, greedy by surface area (bvh_build.cu: k_collapse4_level; layout of the 128-byte node in bvh.cuh).
// Writes nodes4[b]; calls push(child) for every internal child (a 2-wide node index: where that child's 4-wide node will be stored).
template <class Push>
PTB8_HD void collapse4_node(const float4* nodes2, int b, float4* nodes4, Push& push) {
    float lo[4][3], hi[4][3]; int code[4];
    int m = 2;
    load2(nodes2, b, lo[0], hi[0], &code[0], lo[1], hi[1], &code[1]);
    while (m < 4) {
        int best = -1; float best_a = -1.0f;
        for (int k = 0; k < m; ++k) {
            if (code[k] < 0) continue;
            const float dx = hi[k][0] - lo[k][0], dy = hi[k][1] - lo[k][1], dz = hi[k][2] - lo[k][2];
            const float a = dx * dy + dy * dz + dz * dx;
            if (a > best_a) { best_a = a; best = k; }
        }
        if (best < 0) break;
        const int c = code[best];
        load2(nodes2, c, lo[best], hi[best], &code[best], lo[m], hi[m], &code[m]);
        ++m;
    }
    union { uint32_t u; float f; } qn; qn.u = 0x7fc00000u;
    for (int k = m; k < 4; ++k) { for (int d = 0; d < 3; ++d) { lo[k][d] = qn.f; hi[k][d] = qn.f; } code[k] = -1; }
    for (int k = 0; k < m; ++k) if (code[k] >= 0) push(code[k]);
    union { int i; float f; } c0, c1, c2, c3; c0.i = code[0]; c1.i = code[1]; c2.i = code[2]; c3.i = code[3];
    float4* o = nodes4 + (size_t)b * 8;
    o[0] = make_float4(lo[0][0], lo[1][0], lo[2][0], lo[3][0]); o[1] = make_float4(hi[0][0], hi[1][0], hi[2][0], hi[3][0]);
    o[2] = make_float4(lo[0][1], lo[1][1], lo[2][1], lo[3][1]); o[3] = make_float4(hi[0][1], hi[1][1], hi[2][1], hi[3][1]);
    o[4] = make_float4(lo[0][2], lo[1][2], lo[2][2], lo[3][2]); o[5] = make_float4(hi[0][2], hi[1][2], hi[2][2], hi[3][2]);
    o[6] = make_float4(c0.f, c1.f, c2.f, c3.f);
    o[7] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
}

// Alloc: uint32_t nodes(uint32_t n), uint32_t tris(uint32_t n) hand out consecutive indices; void push(WorkItem) appends
// to the next level's queue; void error(int bits).
// Returns the number of internal children.
template <class Alloc>
PTB8_HD int collapse8_node(const float4* nodes2, const float4* tris2, WorkItem item, uint4* nodes8, float4* tris8, Alloc& alloc) {
    float lo[8][3], hi[8][3];
    int code[8];
    int n = 2;
    load2(nodes2, item.bin, lo[0], hi[0], &code[0], lo[1], hi[1], &code[1]);
    // greedy: open the internal child with the largest surface area until eight slots are used
    while (n < 8) {
        int best = -1; float best_a = -1.0f;
        for (int k = 0; k < n; ++k) {
            if (code[k] < 0) continue;
            const float dx = hi[k][0] - lo[k][0], dy = hi[k][1] - lo[k][1], dz = hi[k][2] - lo[k][2];
            const float a = dx * dy + dy * dz + dz * dx;
            if (a > best_a) { best_a = a; best = k; }
        }
        if (best < 0) break;
        const int b = code[best];
        load2(nodes2, b, lo[best], hi[best], &code[best], lo[n], hi[n], &code[n]);
        ++n;
    }
    // node box
    float nlo[3], nhi[3];
    for (int d = 0; d < 3; ++d) {
        nlo[d] = lo[0][d]; nhi[d] = hi[0][d];
        for (int k = 1; k < n; ++k) { nlo[d] = fminf(nlo[d], lo[k][d]); nhi[d] = fmaxf(nhi[d], hi[k][d]); }
        if (!(nhi[d] >= nlo[d]) || !(fabsf(nlo[d]) < 1e30f) || !(fabsf(nhi[d]) < 1e30f)) alloc.error(ERR_BAD_BOX);
    }
    // grid step per axis: the smallest power of two s with 255 s >= extent (exponent clamped to what keeps s * (1 / d) normal)
    int ebias[3]; double step[3];
    for (int d = 0; d < 3; ++d) {
        const double ext = (double)nhi[d] - (double)nlo[d];
        int e = -60;
        if (ext > 0.0) { int fe; (void)frexp(ext / 255.0, &fe); e = fe; }   // ext / 255 = m 2^fe, 0.5 <= m < 1  =>  2^fe > ext / 255
        if (e < -60) e = -60;
        if (e > 100) e = 100;
        while (ldexp(255.0, e) < ext && e < 100) ++e;
        ebias[d] = e + 127; step[d] = ldexp(1.0, e);
    }
    // slot assignment by octant affinity: repeatedly take the (child, slot) pair with the largest dot(centre offset, slot sign)
    int slot_of[8]; int child_in[8];
    for (int k = 0; k < 8; ++k) { slot_of[k] = -1; child_in[k] = -1; }
    float off[8][3];
    for (int k = 0; k < n; ++k) for (int d = 0; d < 3; ++d) off[k][d] = (lo[k][d] + hi[k][d]) - (nlo[d] + nhi[d]);
    for (int round = 0; round < n; ++round) {
        int bk = -1, bs = -1; float bc = 0.0f;
        for (int k = 0; k < n; ++k) {
            if (slot_of[k] >= 0) continue;
            for (int s = 0; s < 8; ++s) {
                if (child_in[s] >= 0) continue;
                const float c = ((s & 1) ? off[k][0] : -off[k][0]) + ((s & 2) ? off[k][1] : -off[k][1]) + ((s & 4) ? off[k][2] : -off[k][2]);
                if (bk < 0 || c > bc) { bk = k; bs = s; bc = c; }
            }
        }
        slot_of[bk] = bs; child_in[bs] = bk;
    }
    // children in slot order: internal ones get consecutive node indices, leaves consecutive triangle ranges
    int n_internal = 0, n_tris = 0;
    for (int s = 0; s < 8; ++s) {
        const int k = child_in[s];
        if (k < 0) continue;
        if (code[k] >= 0) ++n_internal;
        else {
            const int cnt = ((~code[k]) & 7) + 1;
            if (cnt > 3) { alloc.error(ERR_LEAF_TOO_BIG); }
            n_tris += cnt;
        }
    }
    if (n_tris > 24) { alloc.error(ERR_LEAF_TOO_BIG); n_tris = 24; }
    const uint32_t child_base = n_internal ? alloc.nodes((uint32_t)n_internal) : 0u;
    const uint32_t tri_base = n_tris ? alloc.tris((uint32_t)n_tris) : 0u;
    uint32_t imask = 0, meta_w[2] = {0u, 0u};
    uint32_t qw[6][2] = {{0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}, {0u, 0u}};  // qlo.x, qlo.y, qlo.z, qhi.x, qhi.y, qhi.z
    int rank = 0, toff = 0;
    for (int s = 0; s < 8; ++s) {
        const int k = child_in[s];
        uint32_t meta = 0u, ql[3] = {255u, 255u, 255u}, qh[3] = {0u, 0u, 0u};
        if (k >= 0) {
            for (int d = 0; d < 3; ++d) {
                const double rl = ((double)lo[k][d] - (double)nlo[d]) / step[d], rh = ((double)hi[k][d] - (double)nlo[d]) / step[d];
                double fl_ = floor(rl), ch_ = ceil(rh);
                if (fl_ < 0.0) fl_ = 0.0;
                if (fl_ > 255.0) fl_ = 255.0;
                if (ch_ > 255.0) ch_ = 255.0;
                if (ch_ < 0.0) ch_ = 0.0;
                // exact check in double (p and q s are exactly representable; the sum is, too, for sane coordinates)
                while (fl_ > 0.0 && (double)nlo[d] + fl_ * step[d] > (double)lo[k][d]) fl_ -= 1.0;
                while (ch_ < 255.0 && (double)nlo[d] + ch_ * step[d] < (double)hi[k][d]) ch_ += 1.0;
                if ((double)nlo[d] + ch_ * step[d] < (double)hi[k][d]) alloc.error(ERR_BAD_BOX);
                ql[d] = (uint32_t)fl_; qh[d] = (uint32_t)ch_;
            }
            if (code[k] >= 0) {
                imask |= 1u << s;
                meta = (1u << 5) | (24u + (uint32_t)s);
                WorkItem w; w.wide = (int)(child_base + (uint32_t)rank); w.bin = code[k];
                alloc.push(w);
                ++rank;
            } else {
                const int first = (~code[k]) >> 3;
                int cnt = ((~code[k]) & 7) + 1;
                if (cnt > 3) cnt = 3;
                if (toff + cnt > 24) cnt = 24 - toff;
                meta = (((1u << cnt) - 1u) << 5) | (uint32_t)toff;
                for (int j = 0; j < cnt; ++j)
                    for (int v = 0; v < 3; ++v) tris8[(size_t)(tri_base + (uint32_t)(toff + j)) * 3 + v] = tris2[(size_t)(first + j) * 3 + v];
                toff += cnt;
            }
        }
        const int w = s >> 2, sh = 8 * (s & 3);
        meta_w[w] |= meta << sh;
        for (int d = 0; d < 3; ++d) { qw[d][w] |= ql[d] << sh; qw[3 + d][w] |= qh[d] << sh; }
    }
    uint4* o = nodes8 + (size_t)item.wide * 5;
    o[0] = make_uint4(f2u(nlo[0]), f2u(nlo[1]), f2u(nlo[2]), (uint32_t)ebias[0] | ((uint32_t)ebias[1] << 8) | ((uint32_t)ebias[2] << 16) | (imask << 24));
    o[1] = make_uint4(child_base, tri_base, meta_w[0], meta_w[1]);
    o[2] = make_uint4(qw[0][0], qw[0][1], qw[1][0], qw[1][1]);
    o[3] = make_uint4(qw[2][0], qw[2][1], qw[3][0], qw[3][1]);
    o[4] = make_uint4(qw[4][0], qw[4][1], qw[5][0], qw[5][1]);
    return n_internal;
}

}  // namespace ptb8
