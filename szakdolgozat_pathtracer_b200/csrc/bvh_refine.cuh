// bvh_refine.cuh -- binned-SAH refinement of the LBVH, in place.
//
// The Karras tree orders triangles along the Morton curve, so every subtree covers a CONTIGUOUS range [a,b] of the
// sorted triangles and owns the internal-node slots a+1..b-1 plus its own root slot (a or b).  A "treelet" is a
// maximal subtree with at most PTB_TREELET_MAX triangles.  One warp rebuilds one treelet top-down with the binned
// surface-area heuristic (16 bins x 3 axes, Ct = Ci = 1) entirely in shared memory:
//   * lanes 0..15 / 16..31 own one bin of axis 0 / 1 (then axis 2) and scan the node's triangles;
//   * 45 (axis, plane) candidates are evaluated by the lanes, the best one is found with shuffles;
//   * the triangle order is partitioned with ballots (stable), so child ranges stay contiguous;
//   * a node with <= max_leaf triangles becomes a leaf when that is cheaper than its best split.
// The rebuilt nodes reuse the treelet's own slots; unused slots are marked dead.  The levels above the treelets keep
// the LBVH topology.  Traversal results do not depend on the tree (hit rule in bvh.cuh), only its cost does.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptb {

#ifndef PTB_TREELET_MAX
#define PTB_TREELET_MAX 256
#endif
#define PTB_SAH_BINS 16
#ifndef PTB_REFINE_WARPS
#define PTB_REFINE_WARPS 2   // warps (= treelets in flight) per block; the shared-memory footprint is ~49 B per treelet triangle and warp
#endif

struct TreeletTask { unsigned short begin, end; int slot; int depth; };

// SAH splits can be very uneven (1 : n-1); below this treelet-local depth the rebuild switches to median splits, so a
// treelet adds at most PTB_TREELET_SAH_DEPTH + log2(PTB_TREELET_MAX) levels and the tree always fits the traversal stack.
#define PTB_TREELET_SAH_DEPTH 24

struct TreeletShared {
    float lo[3][PTB_TREELET_MAX], hi[3][PTB_TREELET_MAX];
    unsigned int vals[PTB_TREELET_MAX];
    unsigned short perm[PTB_TREELET_MAX], tmp[PTB_TREELET_MAX];
    int bin_lo[3][PTB_SAH_BINS][3], bin_hi[3][PTB_SAH_BINS][3];  // order-preserving integer images of floats (sah_key)
    unsigned int bin_cnt[3][PTB_SAH_BINS];
    TreeletTask stack[PTB_TREELET_MAX];
};

// collapse[i] = 1: node i is not an internal node of the final tree (its parent, if live, points at a leaf made of
// ranges[i]); default for the plain LBVH: every subtree of <= max_leaf triangles collapses.
__global__ void k_mark_collapse(const int2* __restrict__ ranges, int n, int max_leaf, unsigned char* __restrict__ collapse) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 r = ranges[i];
    collapse[i] = (i != 0 && r.y - r.x + 1 <= max_leaf) ? 1 : 0;
}

__global__ void k_find_treelets(const int2* __restrict__ ranges, const int* __restrict__ node_parent, int n, int treelet_max,
                                int* __restrict__ list, unsigned int* __restrict__ count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int2 r = ranges[i];
    if (r.y - r.x + 1 > treelet_max) return;
    const int p = node_parent[i];
    if (p >= 0) { const int2 pr = ranges[p]; if (pr.y - pr.x + 1 <= treelet_max) return; }
    list[atomicAdd(count, 1u)] = i;
}

// order-preserving map float -> int (and back; it is an involution on the bit pattern): integer atomicMin / atomicMax
// then order the floats
__device__ __forceinline__ int sah_key(float f) { const int b = __float_as_int(f); return b >= 0 ? b : b ^ 0x7fffffff; }
__device__ __forceinline__ float sah_unkey(int k) { return __int_as_float(k >= 0 ? k : k ^ 0x7fffffff); }

__device__ __forceinline__ float box_half_area(const float* lo, const float* hi) {
    const float dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return dx * dy + dy * dz + dz * dx;
}

__global__ void __launch_bounds__(32 * PTB_REFINE_WARPS)
k_refine_treelets(const int* __restrict__ list, const unsigned int* __restrict__ count, int max_leaf, int2* __restrict__ children,
                  int2* __restrict__ ranges, int* __restrict__ node_parent, int* __restrict__ leaf_parent,
                  float4* __restrict__ leaf_lo, float4* __restrict__ leaf_hi, float4* __restrict__ node_lo,
                  float4* __restrict__ node_hi, uint32_t* __restrict__ sorted_vals, unsigned char* __restrict__ collapse) {
    __shared__ TreeletShared shared[PTB_REFINE_WARPS];
    const unsigned lane = threadIdx.x & 31u;
    TreeletShared& sh = shared[threadIdx.x >> 5];
    const unsigned int n_treelets = *count;
    const unsigned int total_warps = gridDim.x * PTB_REFINE_WARPS;
    for (unsigned int tl = blockIdx.x * PTB_REFINE_WARPS + (threadIdx.x >> 5); tl < n_treelets; tl += total_warps) {
        const int root = list[tl];
        const int2 rr = ranges[root];
        const int a = rr.x, k = rr.y - rr.x + 1;
        __syncwarp();
        for (int i = (int)lane; i < k; i += 32) {
            const float4 l = leaf_lo[a + i], h = leaf_hi[a + i];
            sh.lo[0][i] = l.x; sh.lo[1][i] = l.y; sh.lo[2][i] = l.z;
            sh.hi[0][i] = h.x; sh.hi[1][i] = h.y; sh.hi[2][i] = h.z;
            sh.vals[i] = sorted_vals[a + i];
            sh.perm[i] = (unsigned short)i;
        }
        // every slot of the treelet except its root starts dead
        for (int s = a + 1 + (int)lane; s <= a + k - 2; s += 32) { collapse[s] = 1; ranges[s] = make_int2(0, -1); }
        int next_free = a + 1;
        int sp = 0;
        if (lane == 0) { sh.stack[0].begin = 0; sh.stack[0].end = (unsigned short)k; sh.stack[0].slot = root; sh.stack[0].depth = 0; }
        sp = 1;
        __syncwarp();
        while (sp > 0) {
            --sp;
            const int begin = sh.stack[sp].begin, end = sh.stack[sp].end, slot = sh.stack[sp].slot, depth = sh.stack[sp].depth;
            const int n = end - begin;
            __syncwarp();
            // node box and centroid box
            float nlo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, nhi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            float clo[3] = {3.4e38f, 3.4e38f, 3.4e38f}, chi[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
            for (int j = begin + (int)lane; j < end; j += 32) {
                const int id = sh.perm[j];
                for (int ax = 0; ax < 3; ++ax) {
                    const float l = sh.lo[ax][id], h = sh.hi[ax][id], c = 0.5f * (l + h);
                    nlo[ax] = fminf(nlo[ax], l); nhi[ax] = fmaxf(nhi[ax], h);
                    clo[ax] = fminf(clo[ax], c); chi[ax] = fmaxf(chi[ax], c);
                }
            }
            for (int off = 16; off > 0; off >>= 1)
                for (int ax = 0; ax < 3; ++ax) {
                    nlo[ax] = fminf(nlo[ax], __shfl_xor_sync(0xffffffffu, nlo[ax], off));
                    nhi[ax] = fmaxf(nhi[ax], __shfl_xor_sync(0xffffffffu, nhi[ax], off));
                    clo[ax] = fminf(clo[ax], __shfl_xor_sync(0xffffffffu, clo[ax], off));
                    chi[ax] = fmaxf(chi[ax], __shfl_xor_sync(0xffffffffu, chi[ax], off));
                }
            if (lane == 0) {
                node_lo[slot] = make_float4(nlo[0], nlo[1], nlo[2], 0.0f);
                node_hi[slot] = make_float4(nhi[0], nhi[1], nhi[2], 0.0f);
                ranges[slot] = make_int2(a + begin, a + end - 1);
            }
            const float node_area = box_half_area(nlo, nhi);
            // binning: the lanes stride over the node's triangles and fold them into the 3 x 16 bins with shared-memory
            // atomics (min / max on order-preserving integer images of the floats, so the result is exact and does not
            // depend on the order of arrival)
            for (int e = (int)lane; e < 3 * PTB_SAH_BINS; e += 32) {
                const int ax = e / PTB_SAH_BINS, bn = e % PTB_SAH_BINS;
                sh.bin_cnt[ax][bn] = 0u;
                for (int d = 0; d < 3; ++d) { sh.bin_lo[ax][bn][d] = 0x7f7fffff; sh.bin_hi[ax][bn][d] = sah_key(-3.4e38f); }
            }
            float scale[3];
            for (int ax = 0; ax < 3; ++ax) { const float ext = chi[ax] - clo[ax]; scale[ax] = ext > 0.0f ? (float)PTB_SAH_BINS / ext : 0.0f; }
            __syncwarp();
            for (int j = begin + (int)lane; j < end; j += 32) {
                const int id = sh.perm[j];
                int kl[3], kh[3];
                for (int d = 0; d < 3; ++d) { kl[d] = sah_key(sh.lo[d][id]); kh[d] = sah_key(sh.hi[d][id]); }
                for (int ax = 0; ax < 3; ++ax) {
                    const float c = 0.5f * (sh.lo[ax][id] + sh.hi[ax][id]);
                    int bn = (int)((c - clo[ax]) * scale[ax]);
                    bn = bn < 0 ? 0 : (bn > PTB_SAH_BINS - 1 ? PTB_SAH_BINS - 1 : bn);
                    atomicAdd(&sh.bin_cnt[ax][bn], 1u);
                    for (int d = 0; d < 3; ++d) { atomicMin(&sh.bin_lo[ax][bn][d], kl[d]); atomicMax(&sh.bin_hi[ax][bn][d], kh[d]); }
                }
            }
            __syncwarp();
            // 45 candidates (axis, plane): split after bin `plane`.  Per axis, lanes 0..15 hold the running union of bins
            // 0..lane (prefix scan), lanes 16..31 the union of bins (31 - lane)..15 (suffix scan); lane p < 15 then pairs
            // its left side with the right side of lane 30 - p.
            float best_cost = 3.4e38f; int best_cand = -1; unsigned int best_nl = 0;
            for (int ax = 0; ax < 3; ++ax) {
                const int bn = lane < 16u ? (int)lane : 31 - (int)lane;
                float bl[3], bh[3];
                for (int d = 0; d < 3; ++d) { bl[d] = sah_unkey(sh.bin_lo[ax][bn][d]); bh[d] = sah_unkey(sh.bin_hi[ax][bn][d]); }
                unsigned int cnt = sh.bin_cnt[ax][bn];
                for (int off = 1; off < 16; off <<= 1) {
                    const unsigned int oc = __shfl_up_sync(0xffffffffu, cnt, off, 16);
                    float ol[3], oh[3];
                    for (int d = 0; d < 3; ++d) { ol[d] = __shfl_up_sync(0xffffffffu, bl[d], off, 16); oh[d] = __shfl_up_sync(0xffffffffu, bh[d], off, 16); }
                    if ((int)(lane & 15u) >= off) { cnt += oc; for (int d = 0; d < 3; ++d) { bl[d] = fminf(bl[d], ol[d]); bh[d] = fmaxf(bh[d], oh[d]); } }
                }
                const float my_area = cnt ? box_half_area(bl, bh) : 0.0f;
                const int partner = 30 - (int)lane;  // valid for lanes 0..14
                const float r_area = __shfl_sync(0xffffffffu, my_area, partner & 31);
                const unsigned int r_cnt = __shfl_sync(0xffffffffu, cnt, partner & 31);
                if (lane < (unsigned)(PTB_SAH_BINS - 1) && chi[ax] - clo[ax] > 0.0f && cnt > 0u && r_cnt > 0u) {
                    const float cost = my_area * (float)cnt + r_area * (float)r_cnt;
                    const int cand = ax * (PTB_SAH_BINS - 1) + (int)lane;
                    if (cost < best_cost) { best_cost = cost; best_cand = cand; best_nl = cnt; }
                }
            }
            for (int off = 16; off > 0; off >>= 1) {
                const float oc = __shfl_xor_sync(0xffffffffu, best_cost, off);
                const int ocand = __shfl_xor_sync(0xffffffffu, best_cand, off);
                const unsigned int onl = __shfl_xor_sync(0xffffffffu, best_nl, off);
                if (ocand >= 0 && (best_cand < 0 || oc < best_cost || (oc == best_cost && ocand < best_cand))) { best_cost = oc; best_cand = ocand; best_nl = onl; }
            }
            if (depth >= PTB_TREELET_SAH_DEPTH) best_cand = -1;  // median splits from here on (bounded depth)
            const float leaf_cost = node_area * (float)n;
            const float split_cost = best_cand >= 0 ? node_area + best_cost : 3.4e38f;
            if (n <= max_leaf && slot != 0 && (best_cand < 0 || leaf_cost <= split_cost)) {
                if (lane == 0) collapse[slot] = 1;
                for (int j = begin + (int)lane; j < end; j += 32) leaf_parent[a + j] = slot;  // keeps the depth walk exact
                continue;
            }
            if (lane == 0) collapse[slot] = 0;
            int mid;
            if (best_cand >= 0) {
                const int ax = best_cand / (PTB_SAH_BINS - 1), plane = best_cand % (PTB_SAH_BINS - 1);
                const float ext = chi[ax] - clo[ax];
                const float scale = (float)PTB_SAH_BINS / ext;
                int nleft = 0, nright = 0;
                for (int base = begin; base < end; base += 32) {
                    const int j = base + (int)lane;
                    const bool act = j < end;
                    int id = 0; bool left = false;
                    if (act) {
                        id = sh.perm[j];
                        const float c = 0.5f * (sh.lo[ax][id] + sh.hi[ax][id]);
                        int b = (int)((c - clo[ax]) * scale);
                        b = b < 0 ? 0 : (b > PTB_SAH_BINS - 1 ? PTB_SAH_BINS - 1 : b);
                        left = b <= plane;
                    }
                    const unsigned lm = __ballot_sync(0xffffffffu, act && left), rm = __ballot_sync(0xffffffffu, act && !left);
                    const unsigned lt = (1u << lane) - 1u;
                    if (act) {
                        const int dst = left ? begin + nleft + __popc(lm & lt) : begin + (int)best_nl + nright + __popc(rm & lt);
                        sh.tmp[dst] = (unsigned short)id;
                    }
                    nleft += __popc(lm); nright += __popc(rm);
                }
                __syncwarp();
                for (int j = begin + (int)lane; j < end; j += 32) sh.perm[j] = sh.tmp[j];
                mid = begin + (int)best_nl;
            } else {
                mid = begin + n / 2;  // all centroids coincide: object median in Morton order
            }
            __syncwarp();
            // children
            int child[2];
            const int cb[2] = {begin, mid}, ce[2] = {mid, end};
            for (int c = 0; c < 2; ++c) {
                const int cn = ce[c] - cb[c];
                if (cn == 1) {
                    child[c] = ~(a + cb[c]);
                    if (lane == 0) leaf_parent[a + cb[c]] = slot;
                } else {
                    const int cslot = next_free++;
                    child[c] = cslot;
                    if (lane == 0) {
                        node_parent[cslot] = slot;
                        sh.stack[sp].begin = (unsigned short)cb[c]; sh.stack[sp].end = (unsigned short)ce[c]; sh.stack[sp].slot = cslot;
                        sh.stack[sp].depth = depth + 1;
                    }
                    sp++;
                }
            }
            if (lane == 0) children[slot] = make_int2(child[0], child[1]);
            __syncwarp();
        }
        // the new triangle order of the treelet
        __syncwarp();
        for (int i = (int)lane; i < k; i += 32) {
            const int id = sh.perm[i];
            leaf_lo[a + i] = make_float4(sh.lo[0][id], sh.lo[1][id], sh.lo[2][id], 0.0f);
            leaf_hi[a + i] = make_float4(sh.hi[0][id], sh.hi[1][id], sh.hi[2][id], 0.0f);
            sorted_vals[a + i] = sh.vals[id];
        }
        __syncwarp();
    }
}

// depth of the final tree (levels of live internal nodes above the deepest leaf)
__global__ void k_tree_depth(const int* __restrict__ node_parent, const int* __restrict__ leaf_parent,
                             const unsigned char* __restrict__ collapse, int n, unsigned int* __restrict__ max_depth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    unsigned int depth = 0;
    for (int p = leaf_parent[i]; p >= 0; p = node_parent[p]) if (!collapse[p]) depth++;
    atomicMax(max_depth, depth);
}

}  // namespace ptb
