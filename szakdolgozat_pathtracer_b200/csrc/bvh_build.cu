// bvh_build.cu -- device BVH build for sm_100a.
//
// Replaces optixAccelComputeMemoryUsage / optixAccelBuild / optixAccelCompact
// (optixSphere.cpp:917-967).  Only the input contract is the reference's
// (optixSphere.cpp:862-913: float3 positions at 16-byte stride, 3 per triangle,
// no index buffer); the algorithm is ours:
//   1. k_tri_bounds      per-triangle AABB + centroid, scene bounds by warp
//                        shuffle reduction + one atomic per warp
//   2. k_morton          63-bit (or 30-bit) Morton code of the centroid
//   3. radix sort        own 8 (4) x 8-bit LSD passes (histogram / scan / stable
//                        scatter with __match_any_sync ranking)
//   4. k_hierarchy       Karras 2012 radix tree, one thread per internal node
//   5. k_refit           bottom-up AABBs with one atomic arrival flag per node
//   6. treelet SAH       (bvh_refine.cuh) binned-SAH rebuild of every subtree
//                        of <= treelet_size triangles, in place
//   7. k_emit            64-byte two-child nodes + 48-byte leaf-ordered
//                        triangles (layout in bvh.cuh); subtrees of
//                        <= max_leaf_size triangles collapse into one leaf
// All kernels are HBM-streaming integer/float work; grids are sized from N.
#include <cuda_runtime.h>

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "bvh_build.h"
#include "bvh8.cuh"
#include "bvh_refine.cuh"
#include "device_math.cuh"

namespace ptb {

namespace {

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { err = std::string(#call) + ": " + cudaGetErrorString(e_); return false; } } while (0)

// float atomic min/max through the ordered-int trick
__device__ __forceinline__ void atomic_min_f(float* addr, float v) {
    if (v >= 0.0f) atomicMin((int*)addr, __float_as_int(v)); else atomicMax((unsigned int*)addr, __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
    if (v >= 0.0f) atomicMax((int*)addr, __float_as_int(v)); else atomicMin((unsigned int*)addr, __float_as_uint(v));
}

// scene_bounds: [0..2] = lo, [3..5] = hi of the triangle AABBs, [6..8]/[9..11] of the centroids
__global__ void k_tri_bounds(const float4* __restrict__ verts, uint32_t n, float4* __restrict__ tri_lo,
                             float4* __restrict__ tri_hi, float* __restrict__ scene_bounds) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float clo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, chi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    if (i < n) {
        const float4 a = verts[(size_t)i * 3 + 0], b = verts[(size_t)i * 3 + 1], c = verts[(size_t)i * 3 + 2];
        lo[0] = fminf(a.x, fminf(b.x, c.x)); lo[1] = fminf(a.y, fminf(b.y, c.y)); lo[2] = fminf(a.z, fminf(b.z, c.z));
        hi[0] = fmaxf(a.x, fmaxf(b.x, c.x)); hi[1] = fmaxf(a.y, fmaxf(b.y, c.y)); hi[2] = fmaxf(a.z, fmaxf(b.z, c.z));
        tri_lo[i] = make_float4(lo[0], lo[1], lo[2], 0.0f);
        tri_hi[i] = make_float4(hi[0], hi[1], hi[2], 0.0f);
        for (int k = 0; k < 3; ++k) clo[k] = chi[k] = 0.5f * (lo[k] + hi[k]);
    }
    for (int off = 16; off > 0; off >>= 1) {
        for (int k = 0; k < 3; ++k) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], off));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], off));
            clo[k] = fminf(clo[k], __shfl_xor_sync(0xffffffffu, clo[k], off));
            chi[k] = fmaxf(chi[k], __shfl_xor_sync(0xffffffffu, chi[k], off));
        }
    }
    // block reduction in shared memory, then 12 atomics per BLOCK (per-warp atomics on 12 addresses cost 1.1 ms at 4.6 M triangles)
    __shared__ float red[8][12];
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if (lane == 0u) for (int k = 0; k < 3; ++k) { red[warp][k] = lo[k]; red[warp][3 + k] = hi[k]; red[warp][6 + k] = clo[k]; red[warp][9 + k] = chi[k]; }
    __syncthreads();
    if (threadIdx.x < 12) {
        const bool is_min = threadIdx.x < 3 || (threadIdx.x >= 6 && threadIdx.x < 9);
        float v = red[0][threadIdx.x];
        for (unsigned w = 1; w < (blockDim.x >> 5); ++w) v = is_min ? fminf(v, red[w][threadIdx.x]) : fmaxf(v, red[w][threadIdx.x]);
        if (is_min) atomic_min_f(scene_bounds + threadIdx.x, v); else atomic_max_f(scene_bounds + threadIdx.x, v);
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__device__ __forceinline__ uint64_t expand_bits21(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x001f00000000ffffull;
    v = (v | (v << 16)) & 0x001f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

// Morton code of the centroid inside the CENTROID bounds, each axis normalised by its own extent.  bits = 10: the classic
// 30-bit code (default); bits = 21: a 63-bit code for scenes whose meshes need more than 1024 cells per axis (ties between
// identical keys fall back to an index split, Karras 2012 section 4).  On the BASELINE scenes both give the same trees.
__global__ void k_morton(const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi, uint32_t n,
                         const float* __restrict__ scene_bounds, int bits, int cubic, float huge_frac, uint64_t* __restrict__ keys,
                         uint32_t* __restrict__ vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 lo = tri_lo[i], hi = tri_hi[i];
    const float c[3] = {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
    const float cells = (float)(1u << bits);
    // cubic = 1: one cell size for all three axes (the largest centroid extent).  Measured on C2/C4/C5 (profiles/r1_bvh_probe.md):
    // per-axis normalisation gives the better trees for the reference's slab-shaped scenes (floor 400 x 400, meshes ~2 high),
    // so it is the default; the cubic grid is kept for experiments (PTB_MORTON_CUBIC=1).
    const float ext_max = fmaxf(scene_bounds[9] - scene_bounds[6], fmaxf(scene_bounds[10] - scene_bounds[7], scene_bounds[11] - scene_bounds[8]));
    uint32_t q[3];
    for (int k = 0; k < 3; ++k) {
        const float clo = scene_bounds[6 + k];
        const float ext = cubic ? ext_max : scene_bounds[9 + k] - clo;
        float u = ext > 0.0f ? (c[k] - clo) / ext : 0.0f;
        u = fminf(fmaxf(u * cells, 0.0f), cells - 1.0f);
        q[k] = (uint32_t)u;
    }
    uint64_t key;
    if (bits <= 10) key = (uint64_t)((expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]));
    else key = (expand_bits21(q[0]) << 2) | (expand_bits21(q[1]) << 1) | expand_bits21(q[2]);
    // Huge primitives go to the root.  A triangle whose box is a sizeable fraction of the whole scene (the reference's
    // 400 x 400 floor quad, optixSphere.cpp:598-646) would otherwise sit deep inside the Morton order and inflate the
    // boxes of all its ancestors, so that every ray walks down to it.  Bit 63 separates the two groups: it is the first
    // split of the radix tree, the huge ones become one small subtree next to the root (usually a single leaf) and the
    // rest of the scene gets a tree with tight boxes.  (C2: 7.9 -> see profiles/ nodes per segment.)
    if (huge_frac > 0.0f) {
        const float sx = scene_bounds[3] - scene_bounds[0], sy = scene_bounds[4] - scene_bounds[1], sz = scene_bounds[5] - scene_bounds[2];
        const float tx = hi.x - lo.x, ty = hi.y - lo.y, tz = hi.z - lo.z;
        const float scene_area = sx * sy + sy * sz + sz * sx, tri_area = tx * ty + ty * tz + tz * tx;
        if (!(tri_area >= huge_frac * scene_area)) key |= 1ull << 63;
    }
    keys[i] = key;
    vals[i] = i;
}

// ---- LSD radix sort, 8 bits per pass ------------------------------------------------
// One warp owns a tile of SORT_TILE consecutive keys.  Ranking inside a 32-key
// chunk uses __match_any_sync, so the scatter is stable.
#define SORT_TILE 2048u

__global__ void k_sort_hist(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ hist,
                            uint32_t n_tiles) {
    __shared__ uint32_t h[256];
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    for (uint32_t b = lane; b < 256; b += 32) h[b] = 0;
    __syncwarp();
    const uint32_t begin = tile * SORT_TILE, end = min(begin + SORT_TILE, n);
    for (uint32_t i = begin + lane; i < end; i += 32) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    __syncwarp();
    for (uint32_t b = lane; b < 256; b += 32) hist[(size_t)b * n_tiles + tile] = h[b];  // bin-major
}

// Exclusive scan of the bin-major histogram hist[256][n_tiles], two steps:
//   k_sort_scan_bins    one block per bin: exclusive scan over that bin's tiles in place, bin total -> bin_total[bin]
//   k_sort_scan_totals  one warp-sized problem: exclusive scan of the 256 bin totals -> bin_base[bin]
// The scatter adds bin_base[digit] to the per-tile offset.  (A single-block scan of all 256 * n_tiles entries took
// 0.53 ms per pass at 4.6 M triangles.)
__global__ void __launch_bounds__(256) k_sort_scan_bins(uint32_t* __restrict__ hist, uint32_t n_tiles, uint32_t* __restrict__ bin_total) {
    __shared__ uint32_t warp_sums[8];
    __shared__ uint32_t carry;
    uint32_t* h = hist + (size_t)blockIdx.x * n_tiles;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < n_tiles; base += 256u) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < n_tiles ? h[i] : 0u;
        uint32_t x = v;
        for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((int)lane >= off) x += y; }
        if (lane == 31u) warp_sums[warp] = x;
        __syncthreads();
        uint32_t warp_off = 0, total = 0;
        for (uint32_t w = 0; w < 8u; ++w) { const uint32_t sw = warp_sums[w]; if (w < warp) warp_off += sw; total += sw; }
        const uint32_t c = carry;
        if (i < n_tiles) h[i] = c + warp_off + x - v;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + total;
        __syncthreads();
    }
    if (threadIdx.x == 0) bin_total[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(256) k_sort_scan_totals(const uint32_t* __restrict__ bin_total, uint32_t* __restrict__ bin_base) {
    __shared__ uint32_t warp_sums[8];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const uint32_t v = bin_total[threadIdx.x];
    uint32_t x = v;
    for (int off = 1; off < 32; off <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, off); if ((int)lane >= off) x += y; }
    if (lane == 31u) warp_sums[warp] = x;
    __syncthreads();
    uint32_t warp_off = 0;
    for (uint32_t w = 0; w < warp; ++w) warp_off += warp_sums[w];
    bin_base[threadIdx.x] = warp_off + x - v;
}

__global__ void k_sort_scatter(const uint64_t* __restrict__ keys_in, const uint32_t* __restrict__ vals_in, uint32_t n,
                               int shift, const uint32_t* __restrict__ hist, const uint32_t* __restrict__ bin_base, uint32_t n_tiles,
                               uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out) {
    __shared__ uint32_t offs[256];
    const uint32_t tile = blockIdx.x, lane = threadIdx.x;
    for (uint32_t b = lane; b < 256; b += 32) offs[b] = hist[(size_t)b * n_tiles + tile] + bin_base[b];
    __syncwarp();
    const uint32_t begin = tile * SORT_TILE, end = min(begin + SORT_TILE, n);
    for (uint32_t base = begin; base < end; base += 32) {
        const uint32_t i = base + lane;
        const bool active = i < end;
        const uint64_t key = active ? keys_in[i] : 0ull;
        const uint32_t val = active ? vals_in[i] : 0u;
        const uint32_t digit = active ? ((uint32_t)(key >> shift) & 255u) : 256u;  // 256 = inactive group
        const unsigned peers = __match_any_sync(0xffffffffu, digit);
        const uint32_t rank = (uint32_t)__popc(peers & ((1u << lane) - 1u));
        uint32_t dst = 0;
        if (active) dst = offs[digit] + rank;
        __syncwarp();
        if (active && rank == 0u) offs[digit] += (uint32_t)__popc(peers);
        __syncwarp();
        if (active) { keys_out[dst] = key; vals_out[dst] = val; }
    }
}

// ---- Karras 2012 ----------------------------------------------------------------------
__device__ __forceinline__ int delta(const uint64_t* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    const uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + __clz((uint32_t)i ^ (uint32_t)j);
    return __clzll((long long)(a ^ b));
}

// child encoding in the build tree: >= 0 internal node index, < 0 leaf ~index
__global__ void k_hierarchy(const uint64_t* __restrict__ keys, int n, int2* __restrict__ children,
                            int2* __restrict__ ranges, int* __restrict__ node_parent, int* __restrict__ leaf_parent) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0, t = l;
    do {
        t = (t + 1) >> 1;
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    int left, right;
    if (first == gamma) { left = ~gamma; leaf_parent[gamma] = i; } else { left = gamma; node_parent[gamma] = i; }
    if (last == gamma + 1) { right = ~(gamma + 1); leaf_parent[gamma + 1] = i; } else { right = gamma + 1; node_parent[gamma + 1] = i; }
    children[i] = make_int2(left, right);
    ranges[i] = make_int2(first, last);
    if (i == 0) node_parent[0] = -1;
}

// Leaf boxes are padded so that the conservative slab test of the traversal can
// never reject a box whose triangle passes the watertight test: the edge
// functions of that test are evaluated relative to the ray origin, so their
// rounding error scales with the scene extent, not with the triangle.
__global__ void k_leaf_boxes(const float4* __restrict__ tri_lo, const float4* __restrict__ tri_hi,
                             const uint32_t* __restrict__ sorted_vals, uint32_t n, const float* __restrict__ scene_bounds,
                             float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = sorted_vals[i];
    float4 lo = tri_lo[p], hi = tri_hi[p];
    const float ex = scene_bounds[3] - scene_bounds[0], ey = scene_bounds[4] - scene_bounds[1], ez = scene_bounds[5] - scene_bounds[2];
    const float diag = sqrtf(ex * ex + ey * ey + ez * ez);
    float mag = fmaxf(fmaxf(fabsf(lo.x), fabsf(hi.x)), fmaxf(fmaxf(fabsf(lo.y), fabsf(hi.y)), fmaxf(fabsf(lo.z), fabsf(hi.z))));
    const float pad = diag * 2.384185791015625e-7f + mag * 9.5367431640625e-7f + 1e-30f;  // 2^-22, 2^-20
    leaf_lo[i] = make_float4(lo.x - pad, lo.y - pad, lo.z - pad, 0.0f);
    leaf_hi[i] = make_float4(hi.x + pad, hi.y + pad, hi.z + pad, 0.0f);
}

__global__ void k_refit(const int2* __restrict__ children, const int* __restrict__ node_parent,
                        const int* __restrict__ leaf_parent, const float4* __restrict__ leaf_lo,
                        const float4* __restrict__ leaf_hi, int n, float4* __restrict__ node_lo, float4* __restrict__ node_hi,
                        unsigned int* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = leaf_parent[i];
    while (node >= 0) {
        if (atomicAdd(&flags[node], 1u) == 0u) break;  // first arrival: the sibling subtree is not finished yet
        __threadfence();
        const int2 ch = children[node];
        // volatile-style reads through L2: the sibling's box was written by another SM
        const float4 alo = ch.x < 0 ? leaf_lo[~ch.x] : __ldcg(&node_lo[ch.x]);
        const float4 ahi = ch.x < 0 ? leaf_hi[~ch.x] : __ldcg(&node_hi[ch.x]);
        const float4 blo = ch.y < 0 ? leaf_lo[~ch.y] : __ldcg(&node_lo[ch.y]);
        const float4 bhi = ch.y < 0 ? leaf_hi[~ch.y] : __ldcg(&node_hi[ch.y]);
        node_lo[node] = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
        node_hi[node] = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
        __threadfence();
        node = node_parent[node];
    }
}

__device__ __forceinline__ float half_area(float4 lo, float4 hi) {
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// Final layout.  A child whose collapse flag is set becomes one leaf made of its (contiguous) triangle range;
// nodes below it stay unused.
// stats: [0] live nodes, [1] leaves, sah: sum of area-weighted costs (Ct = Ci = 1).
// block-wide sum of the per-node statistics, one atomic triple per block (three atomics per node on the same addresses
// took 2.6 ms at 4.6 M triangles).  Must be reached by every thread of the block.
__device__ __forceinline__ void emit_stats(unsigned int nodes, unsigned int leaves, float cost, unsigned int* stats, float* sah) {
    __shared__ unsigned int s_nodes[8], s_leaves[8];
    __shared__ float s_cost[8];
    for (int off = 16; off > 0; off >>= 1) {
        nodes += __shfl_xor_sync(0xffffffffu, nodes, off); leaves += __shfl_xor_sync(0xffffffffu, leaves, off);
        cost += __shfl_xor_sync(0xffffffffu, cost, off);
    }
    const unsigned warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31u) == 0u) { s_nodes[warp] = nodes; s_leaves[warp] = leaves; s_cost[warp] = cost; }
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int a = 0, b = 0; float c = 0.0f;
        for (unsigned w = 0; w < (blockDim.x >> 5); ++w) { a += s_nodes[w]; b += s_leaves[w]; c += s_cost[w]; }
        if (a) { atomicAdd(&stats[0], a); atomicAdd(&stats[1], b); atomicAdd(sah, c); }
    }
}

__global__ void k_emit_nodes(const int2* __restrict__ children, const int2* __restrict__ ranges,
                             const int* __restrict__ node_parent, const float4* __restrict__ leaf_lo,
                             const float4* __restrict__ leaf_hi, const float4* __restrict__ node_lo,
                             const float4* __restrict__ node_hi, const unsigned char* __restrict__ collapse, int n,
                             float4* __restrict__ out_nodes, unsigned int* __restrict__ stats, float* __restrict__ sah) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int nodes_here = 0, leaves_here = 0; float cost = 0.0f;
    if (i < n - 1) {
        const bool live = (i == 0) || !collapse[i];
        if (!live) {
            // keep the slot well defined (never referenced)
            const float qnan = __int_as_float(0x7fc00000);
            out_nodes[(size_t)i * 4 + 0] = out_nodes[(size_t)i * 4 + 1] = out_nodes[(size_t)i * 4 + 2] = make_float4(qnan, qnan, qnan, qnan);
            out_nodes[(size_t)i * 4 + 3] = make_float4(__int_as_float(-1), __int_as_float(-1), 0.0f, 0.0f);
        } else {
            const int2 ch = children[i];
            float4 lo[2], hi[2];
            int code[2];
            for (int c = 0; c < 2; ++c) {
                const int child = c ? ch.y : ch.x;
                if (child < 0) {
                    lo[c] = leaf_lo[~child]; hi[c] = leaf_hi[~child];
                    code[c] = ~(((~child) << 3) | 0);
                    leaves_here++; cost += half_area(lo[c], hi[c]) * 1.0f;
                } else {
                    lo[c] = node_lo[child]; hi[c] = node_hi[child];
                    const int2 cr = ranges[child];
                    const int cnt = cr.y - cr.x + 1;
                    if (collapse[child]) { code[c] = ~((cr.x << 3) | (cnt - 1)); leaves_here++; cost += half_area(lo[c], hi[c]) * (float)cnt; }
                    else code[c] = child;
                }
            }
            out_nodes[(size_t)i * 4 + 0] = make_float4(lo[0].x, hi[0].x, lo[0].y, hi[0].y);
            out_nodes[(size_t)i * 4 + 1] = make_float4(lo[1].x, hi[1].x, lo[1].y, hi[1].y);
            out_nodes[(size_t)i * 4 + 2] = make_float4(lo[0].z, hi[0].z, lo[1].z, hi[1].z);
            out_nodes[(size_t)i * 4 + 3] = make_float4(__int_as_float(code[0]), __int_as_float(code[1]), 0.0f, 0.0f);
            cost += half_area(node_lo[i], node_hi[i]);  // one traversal step for this node
            nodes_here = 1;
        }
    }
    emit_stats(nodes_here, leaves_here, cost, stats, sah);  // every thread of the block gets here
}

__global__ void k_emit_tris(const float4* __restrict__ verts, const uint32_t* __restrict__ sorted_vals, uint32_t n,
                            float4* __restrict__ out_tris) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = sorted_vals[i];
    float4 a = verts[(size_t)p * 3 + 0], b = verts[(size_t)p * 3 + 1], c = verts[(size_t)p * 3 + 2];
    a.w = __int_as_float((int)p); b.w = 0.0f; c.w = 0.0f;
    out_tris[(size_t)i * 3 + 0] = a; out_tris[(size_t)i * 3 + 1] = b; out_tris[(size_t)i * 3 + 2] = c;
}

// ---- 4-wide collapse -----------------------------------------------------------------------------------------------
// Greedy by surface area, one level of the 4-wide tree per launch (the same scheme as the 8-wide collapse below): a work item is
// the 2-wide node a 4-wide node starts from; it holds that node's two children and opens the internal child with the largest
// box until four slots are used (a fixed "every even-depth node absorbs its children" rule visits 13-17 % more nodes per ray).
// The 4-wide node is stored at the index of the 2-wide node it starts from, so child codes stay what they are in the 2-wide
// tree (>= 0: node index, < 0: leaf code).  Layout (bvh.cuh): 8 x float4 = 128 B:
//   lo.x[4], hi.x[4], lo.y[4], hi.y[4], lo.z[4], hi.z[4], codes[4], pad; unused slots have NaN boxes (never entered).
// Fewer DEPENDENT node fetches per ray is what pays once the tree is larger than L2.
struct Push4 {   // internal children of one 4-wide node: up to four entries, appended to the next level's queue in one piece
    int c[4]; int n = 0;
    __device__ void operator()(int child) { c[n++] = child; }
};
__global__ void k_collapse4_level(const float4* __restrict__ nodes2, const int* __restrict__ in, unsigned int n_in, int* __restrict__ out,
                                  unsigned int* __restrict__ out_count, float4* __restrict__ nodes4) {
    const unsigned int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n_in) return;
    Push4 push;
    ptb8::collapse4_node(nodes2, in[w], nodes4, push);
    if (push.n) {
        unsigned int at = atomicAdd(out_count, (unsigned int)push.n);
        for (int k = 0; k < push.n; ++k) out[at++] = push.c[k];
    }
}

// ---- 8-wide quantised collapse (bvh8.cuh) -----------------------------------------------------------------------------
// One level of the 8-wide tree per launch: every work item (8-wide node index, 2-wide node it starts from) opens 2-wide
// nodes greedily by surface area until it holds eight children, quantises their boxes, copies the triangles of its leaf
// children into one contiguous range and queues its internal children (which get consecutive node indices) for the next
// level.  counters8: [0] nodes allocated, [1] triangles allocated, [2] error bits, [3 + (level & 1)] length of the queue
// being written.
struct Alloc8 {
    unsigned int* counters; ptb8::WorkItem* next; unsigned int* next_count;
    __device__ uint32_t nodes(uint32_t n) { return atomicAdd(&counters[0], n); }
    __device__ uint32_t tris(uint32_t n) { return atomicAdd(&counters[1], n); }
    __device__ void push(ptb8::WorkItem w) { next[atomicAdd(next_count, 1u)] = w; }
    __device__ void error(int bits) { atomicOr(&counters[2], (unsigned int)bits); }
};
__global__ void k_collapse8_level(const float4* __restrict__ nodes2, const float4* __restrict__ tris2, const ptb8::WorkItem* __restrict__ in,
                                  unsigned int n_in, ptb8::WorkItem* __restrict__ out, unsigned int* __restrict__ out_count,
                                  unsigned int* __restrict__ counters8, uint4* __restrict__ nodes8, float4* __restrict__ tris8) {
    const unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_in) return;
    Alloc8 al; al.counters = counters8; al.next = out; al.next_count = out_count;
    ptb8::collapse8_node(nodes2, tris2, in[i], nodes8, tris8, al);
}

// One device allocation for all the builder's temporaries, carved into 256-byte aligned pieces.
struct Scratch {
    char* base = nullptr; size_t used = 0, cap = 0;
    ~Scratch() { cudaFree(base); }
    static size_t pad(size_t b) { return (b + 255) & ~(size_t)255; }
    template <typename T> void reserve(size_t count) { cap += pad(std::max<size_t>(count, 1) * sizeof(T)); }
    bool commit(std::string& err) {
        cudaError_t e = cudaMalloc((void**)&base, cap ? cap : 256);
        if (e != cudaSuccess) { err = std::string("cudaMalloc (BVH scratch): ") + cudaGetErrorString(e); return false; }
        return true;
    }
    template <typename T> T* take(size_t count) { T* p = (T*)(base + used); used += pad(std::max<size_t>(count, 1) * sizeof(T)); return p; }
};

}  // namespace

void free_bvh(DeviceBvh& b);

namespace {
struct EventPair {  // destroyed on every return path of the build
    cudaEvent_t a = nullptr, b = nullptr;
    ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

bool build_bvh_impl(const float4* d_verts, uint32_t n, const ptb_build_cfg& cfg, cudaStream_t stream, DeviceBvh& out,
                    ptb_build_stats& stats, std::string& err) {
    memset(&stats, 0, sizeof(stats));
    stats.num_triangles = n;
    // the 8-wide quantised tree holds at most 24 triangles per node: leaves of at most 3
    const bool want8 = n >= 2 && cfg.bvh_width == 8;
    const int max_leaf_cfg = cfg.max_leaf_size < 1 ? 1 : (cfg.max_leaf_size > 8 ? 8 : cfg.max_leaf_size);
    const int max_leaf = want8 && max_leaf_cfg > 3 ? 3 : max_leaf_cfg;
    const bool refine = cfg.sah_refine != 0;
    if (n >= (1u << 28)) { err = "BVH build: more than 2^28 triangles"; return false; }

    const uint32_t n_nodes = n >= 2 ? n - 1 : 1;
    float4* d_nodes = nullptr; float4* d_tris = nullptr;
    CK(cudaMalloc((void**)&d_nodes, (size_t)n_nodes * 64));
    if (cudaMalloc((void**)&d_tris, (size_t)std::max(n, 1u) * 48) != cudaSuccess) { cudaFree(d_nodes); err = "cudaMalloc (BVH triangles) failed"; return false; }
    out.nodes = d_nodes; out.tris = d_tris; out.n_nodes = n_nodes; out.n_tris = n;
    // 4-wide copy of the tree (k_collapse4): asked for explicitly, or automatically once the 2-wide tree outgrows what stays
    // cache-resident next to the path state (measured: C4 with 4.6 M triangles gains, C2 with 16 K does not)
    const bool wide = n >= 2 && (cfg.bvh_width == 4 || (cfg.bvh_width == 0 && n >= PTB_WIDE_BVH_MIN_TRIS));
    out.nodes4 = nullptr;
    if (wide && cudaMalloc((void**)&out.nodes4, (size_t)n_nodes * 128) != cudaSuccess) { cudaFree(d_nodes); cudaFree(d_tris); out.nodes = out.tris = nullptr; err = "cudaMalloc (4-wide BVH nodes) failed"; return false; }

    EventPair evs;
    CK(cudaEventCreate(&evs.a)); CK(cudaEventCreate(&evs.b));
    const cudaEvent_t ev0 = evs.a, ev1 = evs.b;

    if (n < 2) {
        // Degenerate scenes: a root whose missing children have NaN boxes (never entered).
        const float qnan = __builtin_nanf("");
        float h[16];
        for (int i = 0; i < 12; ++i) h[i] = qnan;
        int codes[4] = {-1, -1, 0, 0};
        memcpy(h + 12, codes, 16);
        if (n == 1) {
            float v[12];
            CK(cudaMemcpyAsync(v, d_verts, 48, cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            float lo[3], hi[3];
            for (int k = 0; k < 3; ++k) { lo[k] = std::min(v[k], std::min(v[4 + k], v[8 + k])); hi[k] = std::max(v[k], std::max(v[4 + k], v[8 + k])); }
            float mag = 0.0f; for (int k = 0; k < 3; ++k) mag = std::max(mag, std::max(fabsf(lo[k]), fabsf(hi[k])));
            const float pad = mag * 1e-5f + 1e-30f;
            h[0] = lo[0] - pad; h[1] = hi[0] + pad; h[2] = lo[1] - pad; h[3] = hi[1] + pad; h[8] = lo[2] - pad; h[9] = hi[2] + pad;
            int zero = 0; memcpy(&v[3], &zero, 4); v[7] = 0.0f; v[11] = 0.0f;
            CK(cudaMemcpyAsync(d_tris, v, 48, cudaMemcpyHostToDevice, stream));
        }
        CK(cudaMemcpyAsync(d_nodes, h, 64, cudaMemcpyHostToDevice, stream));
        CK(cudaStreamSynchronize(stream));
        stats.num_nodes = 1; stats.num_leaves = n; stats.max_depth = 1; stats.bvh_bytes = 64 + (uint64_t)n * 48;
        stats.bvh_width = 2;
        if (out.nodes4) { cudaFree(out.nodes4); out.nodes4 = nullptr; }
        return true;
    }

    Scratch sc;
    float4 *tri_lo, *tri_hi, *leaf_lo, *leaf_hi, *node_lo, *node_hi;
    float* scene_bounds; uint64_t* keys[2]; uint32_t *vals[2], *hist; int2 *children, *ranges; int *node_parent, *leaf_parent;
    unsigned int *flags, *counters, *counters8; float* sah; unsigned char* collapse; int* treelets; uint32_t* bin_total;
    const uint32_t n_tiles = (n + SORT_TILE - 1) / SORT_TILE;
    for (int k = 0; k < 6; ++k) sc.reserve<float4>(n);
    sc.reserve<float>(12); sc.reserve<uint64_t>(n); sc.reserve<uint64_t>(n); sc.reserve<uint32_t>(n); sc.reserve<uint32_t>(n);
    sc.reserve<uint32_t>((size_t)256 * n_tiles); sc.reserve<int2>(n); sc.reserve<int2>(n); sc.reserve<int>(n); sc.reserve<int>(n);
    sc.reserve<unsigned int>(n); sc.reserve<unsigned int>(8); sc.reserve<unsigned int>(8); sc.reserve<float>(1); sc.reserve<unsigned char>(n); sc.reserve<int>(n);
    sc.reserve<uint32_t>(512);
    if (!sc.commit(err)) return false;
    tri_lo = sc.take<float4>(n); tri_hi = sc.take<float4>(n); leaf_lo = sc.take<float4>(n); leaf_hi = sc.take<float4>(n);
    node_lo = sc.take<float4>(n); node_hi = sc.take<float4>(n);
    scene_bounds = sc.take<float>(12); keys[0] = sc.take<uint64_t>(n); keys[1] = sc.take<uint64_t>(n);
    vals[0] = sc.take<uint32_t>(n); vals[1] = sc.take<uint32_t>(n); hist = sc.take<uint32_t>((size_t)256 * n_tiles);
    children = sc.take<int2>(n); ranges = sc.take<int2>(n); node_parent = sc.take<int>(n); leaf_parent = sc.take<int>(n);
    flags = sc.take<unsigned int>(n); counters = sc.take<unsigned int>(8); counters8 = sc.take<unsigned int>(8); sah = sc.take<float>(1);
    collapse = sc.take<unsigned char>(n); treelets = sc.take<int>(n); bin_total = sc.take<uint32_t>(512);
    // the timed region starts here: kernels and memsets of the build only, no allocation
    CK(cudaEventRecord(ev0, stream));

    const float init_bounds[12] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    CK(cudaMemcpyAsync(scene_bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, stream));
    CK(cudaMemsetAsync(flags, 0, (size_t)n * sizeof(unsigned int), stream));
    CK(cudaMemsetAsync(counters, 0, 8 * sizeof(unsigned int), stream));
    CK(cudaMemsetAsync(sah, 0, sizeof(float), stream));

    const uint32_t B = 256, G = (n + B - 1) / B;
    k_tri_bounds<<<G, B, 0, stream>>>(d_verts, n, tri_lo, tri_hi, scene_bounds);
    const int morton_bits = cfg.morton_bits == 63 ? 21 : 10;
    static const int env_cubic = getenv("PTB_MORTON_CUBIC") ? atoi(getenv("PTB_MORTON_CUBIC")) : 0;  // experiments only
    static const float env_huge = getenv("PTB_HUGE_FRAC") ? (float)atof(getenv("PTB_HUGE_FRAC")) : -1.0f;  // experiments only
    const float huge_frac = env_huge >= 0.0f ? env_huge : 0.0625f;
    const bool huge_bit = huge_frac > 0.0f;
    k_morton<<<G, B, 0, stream>>>(tri_lo, tri_hi, n, scene_bounds, morton_bits, env_cubic, huge_frac, keys[0], vals[0]);
    int cur = 0;
    // 8-bit LSD passes over the bits in use: 30-bit codes need 4, 63-bit codes 8; the huge-primitive flag (bit 63) adds
    // one pass over the top byte to the 30-bit case
    const int n_pass = morton_bits == 10 ? (huge_bit ? 5 : 4) : 8;
    for (int pass = 0; pass < n_pass; ++pass) {
        const int shift = (morton_bits == 10 && pass == 4) ? 56 : pass * 8;
        k_sort_hist<<<n_tiles, 32, 0, stream>>>(keys[cur], n, shift, hist, n_tiles);
        k_sort_scan_bins<<<256, 256, 0, stream>>>(hist, n_tiles, bin_total);
        k_sort_scan_totals<<<1, 256, 0, stream>>>(bin_total, bin_total + 256);
        k_sort_scatter<<<n_tiles, 32, 0, stream>>>(keys[cur], vals[cur], n, shift, hist, bin_total + 256, n_tiles, keys[cur ^ 1], vals[cur ^ 1]);
        cur ^= 1;
    }
    k_hierarchy<<<G, B, 0, stream>>>(keys[cur], (int)n, children, ranges, node_parent, leaf_parent);
    k_leaf_boxes<<<G, B, 0, stream>>>(tri_lo, tri_hi, vals[cur], n, scene_bounds, leaf_lo, leaf_hi);
    k_refit<<<G, B, 0, stream>>>(children, node_parent, leaf_parent, leaf_lo, leaf_hi, (int)n, node_lo, node_hi, flags);
    k_mark_collapse<<<G, B, 0, stream>>>(ranges, (int)n, max_leaf, collapse);
    if (refine) {
        // binned-SAH rebuild of every subtree of <= treelet_size triangles (bvh_refine.cuh); counters[3] = #treelets
        int tmax = cfg.treelet_size > 0 ? cfg.treelet_size : PTB_TREELET_MAX;
        if (tmax > PTB_TREELET_MAX) tmax = PTB_TREELET_MAX;
        if (tmax < 8) tmax = 8;
        k_find_treelets<<<G, B, 0, stream>>>(ranges, node_parent, (int)n, tmax, treelets, counters + 3);
        const uint32_t rb = std::min<uint32_t>((n + 1) / 2, 148u * 16u);
        k_refine_treelets<<<rb, 32 * PTB_REFINE_WARPS, 0, stream>>>(treelets, counters + 3, max_leaf, children, ranges, node_parent, leaf_parent,
                                                                   leaf_lo, leaf_hi, node_lo, node_hi, vals[cur], collapse);
    }
    k_tree_depth<<<G, B, 0, stream>>>(node_parent, leaf_parent, collapse, (int)n, counters + 2);
    k_emit_nodes<<<G, B, 0, stream>>>(children, ranges, node_parent, leaf_lo, leaf_hi, node_lo, node_hi, collapse, (int)n, d_nodes, counters, sah);
    k_emit_tris<<<G, B, 0, stream>>>(d_verts, vals[cur], n, d_tris);
    CK(cudaGetLastError());
    uint32_t levels4 = 0, n_nodes4 = 0;
    if (out.nodes4) {
        // level-synchronous greedy collapse; the queues reuse builder scratch that is dead by now (keys: 8 B per triangle each)
        int* q[2] = {reinterpret_cast<int*>(keys[0]), reinterpret_cast<int*>(keys[1])};
        unsigned int* c4 = counters8;   // [3 + parity]: length of the queue being written
        const int root = 0;
        CK(cudaMemcpyAsync(q[0], &root, sizeof(root), cudaMemcpyHostToDevice, stream));
        unsigned int n_in = 1;
        while (n_in > 0 && levels4 < 200) {
            const int w = (int)(levels4 & 1u);
            n_nodes4 += n_in;
            CK(cudaMemsetAsync(c4 + 3 + (w ^ 1), 0, sizeof(unsigned int), stream));
            k_collapse4_level<<<(n_in + 127u) / 128u, 128, 0, stream>>>(d_nodes, q[w], n_in, q[w ^ 1], c4 + 3 + (w ^ 1), out.nodes4);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&n_in, c4 + 3 + (w ^ 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            ++levels4;
        }
    }
    uint32_t levels8 = 0;
    if (want8) {
        // level-synchronous collapse; the queues reuse builder scratch that is dead by now (keys: 8 B per triangle each)
        uint4* d_nodes8 = nullptr; float4* d_tris8 = nullptr;
        if (cudaMalloc((void**)&d_nodes8, (size_t)n_nodes * 80) != cudaSuccess || cudaMalloc((void**)&d_tris8, (size_t)n * 48) != cudaSuccess) {
            cudaFree(d_nodes8); err = "cudaMalloc (8-wide BVH) failed"; return false;
        }
        out.nodes8 = d_nodes8; out.tris8 = d_tris8;
        ptb8::WorkItem* q[2] = {reinterpret_cast<ptb8::WorkItem*>(keys[0]), reinterpret_cast<ptb8::WorkItem*>(keys[1])};
        unsigned int* c8 = counters8;
        const unsigned int init8[5] = {1u, 0u, 0u, 0u, 0u};
        const ptb8::WorkItem root = {0, 0};
        CK(cudaMemcpyAsync(c8, init8, sizeof(init8), cudaMemcpyHostToDevice, stream));
        CK(cudaMemcpyAsync(q[0], &root, sizeof(root), cudaMemcpyHostToDevice, stream));
        unsigned int n_in = 1;
        while (n_in > 0 && levels8 < 200) {
            const int w = (int)(levels8 & 1u);
            CK(cudaMemsetAsync(c8 + 3 + (w ^ 1), 0, sizeof(unsigned int), stream));
            k_collapse8_level<<<(n_in + 127u) / 128u, 128, 0, stream>>>(d_nodes, d_tris, q[w], n_in, q[w ^ 1], c8 + 3 + (w ^ 1), c8, d_nodes8, d_tris8);
            CK(cudaGetLastError());
            CK(cudaMemcpyAsync(&n_in, c8 + 3 + (w ^ 1), sizeof(unsigned int), cudaMemcpyDeviceToHost, stream));
            CK(cudaStreamSynchronize(stream));
            ++levels8;
        }
    }
    CK(cudaEventRecord(ev1, stream));

    unsigned int h_counters[4]; float h_sah = 0.0f; float4 root_lo, root_hi;
    CK(cudaMemcpyAsync(h_counters, counters, sizeof(h_counters), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&h_sah, sah, sizeof(float), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&root_lo, node_lo, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaMemcpyAsync(&root_hi, node_hi, sizeof(float4), cudaMemcpyDeviceToHost, stream));
    CK(cudaStreamSynchronize(stream));
    float ms = 0.0f;
    CK(cudaEventElapsedTime(&ms, ev0, ev1));

    const float dx = root_hi.x - root_lo.x, dy = root_hi.y - root_lo.y, dz = root_hi.z - root_lo.z;
    const float root_area = dx * dy + dy * dz + dz * dx;
    stats.num_nodes = h_counters[0]; stats.num_leaves = h_counters[1]; stats.max_depth = h_counters[2];
    stats.sah_cost = root_area > 0.0f ? h_sah / root_area : 0.0f;
    // The same statistic without the huge-primitive leaf under the root (the floor quad dominates sah_cost: its box IS the
    // scene's): cost of the OTHER child's subtree relative to that child's own box.  Equal to sah_cost when the root has no leaf child.
    stats.sah_cost_mesh = stats.sah_cost;
    {
        float h_root[16];
        CK(cudaMemcpyAsync(h_root, d_nodes, 64, cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        int codes[2]; memcpy(codes, h_root + 12, 8);
        auto half_area_of = [&](int c) {
            const float lx = c ? h_root[4] : h_root[0], hx = c ? h_root[5] : h_root[1], ly = c ? h_root[6] : h_root[2], hy = c ? h_root[7] : h_root[3];
            const float lz = c ? h_root[10] : h_root[8], hz = c ? h_root[11] : h_root[9];
            const float ex = hx - lx, ey = hy - ly, ez = hz - lz;
            return ex * ey + ey * ez + ez * ex;
        };
        for (int c = 0; c < 2; ++c) {
            if (codes[c] < 0 && codes[1 - c] >= 0) {
                const float a_leaf = half_area_of(c), a_mesh = half_area_of(1 - c);
                const float cnt = (float)(((~codes[c]) & 7) + 1);
                if (a_mesh > 0.0f) stats.sah_cost_mesh = (h_sah - root_area - a_leaf * cnt) / a_mesh;
            }
        }
    }
    stats.build_ms = ms;
    stats.bvh_bytes = (uint64_t)n_nodes * 64 + (uint64_t)n * 48 + (out.nodes4 ? (uint64_t)n_nodes * 128 : 0);
    // the 4-wide traversal pushes up to three entries per level of the collapsed tree
    if (out.nodes4 && (levels4 + 1) * 3 + 2 >= PTB_BVH_MAX_DEPTH) { cudaFree(out.nodes4); out.nodes4 = nullptr; }
    (void)n_nodes4;
    stats.bvh_width = out.nodes4 ? 4 : 2;  // what the traversal kernels will walk
    if (out.nodes8) {
        unsigned int h8[3] = {0u, 0u, 0u};
        CK(cudaMemcpyAsync(h8, counters8, sizeof(h8), cudaMemcpyDeviceToHost, stream));
        CK(cudaStreamSynchronize(stream));
        // a stack entry per level of the 8-wide tree (uint2 entries in the PTB_BVH_STACK ints); any collapse error
        // (a leaf of more than 3 triangles, a non-finite box) falls back to the narrower tree
        if (h8[2] != 0u || levels8 + 2 >= PTB_BVH_MAX_DEPTH / 2 || h8[1] != n || h8[0] > n_nodes) {
            cudaFree(out.nodes8); cudaFree(out.tris8); out.nodes8 = nullptr; out.tris8 = nullptr;
        } else {
            out.n_nodes8 = h8[0];
            // give back what the worst-case allocation did not need
            uint4* tight = nullptr;
            if (h8[0] < n_nodes && cudaMalloc((void**)&tight, (size_t)h8[0] * 80) == cudaSuccess) {
                CK(cudaMemcpyAsync(tight, out.nodes8, (size_t)h8[0] * 80, cudaMemcpyDeviceToDevice, stream));
                CK(cudaStreamSynchronize(stream));
                cudaFree(out.nodes8); out.nodes8 = tight;
            }
            stats.bvh_width = 8;
            stats.bvh_bytes += (uint64_t)h8[0] * 80 + (uint64_t)n * 48;
            stats.num_nodes8 = h8[0];
        }
    }
    if (stats.max_depth >= PTB_BVH_MAX_DEPTH) {
        err = "BVH build: tree depth " + std::to_string(stats.max_depth) + " exceeds the traversal stack (" + std::to_string(PTB_BVH_MAX_DEPTH) + ")";
        return false;
    }
    return true;
}
}  // namespace

// Owns its outputs until it succeeds: a failed build leaves `out` empty (events and scratch are RAII).
bool build_bvh(const float4* d_verts, uint32_t n, const ptb_build_cfg& cfg, cudaStream_t stream, DeviceBvh& out,
               ptb_build_stats& stats, std::string& err) {
    out = DeviceBvh();
    const bool ok = build_bvh_impl(d_verts, n, cfg, stream, out, stats, err);
    if (!ok) free_bvh(out);
    return ok;
}

void free_bvh(DeviceBvh& b) {
    if (b.nodes) cudaFree(b.nodes);
    if (b.nodes4) cudaFree(b.nodes4);
    if (b.nodes8) cudaFree(b.nodes8);
    if (b.tris8) cudaFree(b.tris8);
    if (b.tris) cudaFree(b.tris);
    b = DeviceBvh();
}

}  // namespace ptb
