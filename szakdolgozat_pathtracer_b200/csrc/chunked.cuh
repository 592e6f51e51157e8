// chunked.cuh -- the block-local wavefront: the stages of kernels.cuh run over CHUNKS of the path pool.
//
// Why (profiles/r1_v2_*): with global queues appended by atomics, the slot order of a queue degrades with
// every bounce, so k_shade / k_miss gather their 16-byte path-state records from half-used 32-byte sectors all
// over a pool that is larger than L2 (1.6 GB at 8 subframes): both kernels sat at ~45 % of HBM peak, stalled on
// long-scoreboard, with an L2 hit rate below 30 % even for the 2 MB of geometry.
//
// Layout: one status byte per slot (ST_TRACE / ST_HIT / ST_MISS / ST_DONE).  A block owns a CHUNK of consecutive slots
// (256 threads x 1..8 slots per thread; 2048 slots for large launches, fewer for small ones so that they still fill the
// chip).  Every stage first compacts the slots of its chunk that are in the wanted state(s) into ascending lists in
// shared memory (status bytes read as words, popcount + one packed warp scan + block scan; ascending lists make the
// state accesses that follow coalesced and whole-sector), then runs the stage body over the list with all lanes busy:
//   trace        lanes fetch rays from the shared list dynamically (one shared-memory atomic per warp) and advance
//                them in quanta (trav_run / trav_run4)
//   shade, miss  closest hit or environment lookup, then Russian roulette + accumulation + path regeneration
// Two drivers share the stage bodies:
//   k_chunk_trace / k_chunk_shade / k_chunk_miss   one kernel per stage and wavefront iteration (pipeline 2)
//   k_chunk_fused                                  a block alternates {TRACE} -> trace and {HIT, MISS} -> shade+miss (one
//                                                  merged stage) over ITS chunk until every pixel of the chunk has
//                                                  finished its samples: one launch, the chunk's path state is re-read
//                                                  out of L2 between stages (pipeline 3, the default).
// Path state is read and written with evict-first hints (kernels.cuh: ldp / stp).
// Counters (segments, hits, misses) are summed per block and added to the context totals once per block.
#pragma once
#include "kernels.cuh"

namespace PTB_NS {

#ifndef PTB_CHUNK_THREADS        // threads per block.  Measured on C2 / C2 close camera / C5 (profiles/r2_experiments.md):
#if PTB_FAST                     //   exact build: 256 > 128 (25.2 vs 26.0 ms);  fast build: 128 > 256 > 64 (19.4 / 20.3 / 20.9 ms):
#define PTB_CHUNK_THREADS 128    //   with the leaner shading code the stage barriers weigh more, and a barrier over 4 warps
#else                            //   waits less than one over 8
#define PTB_CHUNK_THREADS 256
#endif
#endif
#ifndef PTB_CHUNK_SPT
#define PTB_CHUNK_SPT 8          // default slots per thread (1, 2, 4, 8 or 16)
#endif
#define PTB_CHUNK (PTB_CHUNK_THREADS * PTB_CHUNK_SPT)  // slots per block at the default SPT
#ifndef PTB_SHADE_DYNAMIC
#define PTB_SHADE_DYNAMIC 1      // 1: the shade + miss stage hands out its list in warp-sized batches (see chunk_stage_shade_miss):
#endif                           //    +1.4 % exact, +2..3 % fast
#ifndef PTB_TRACE_TWO_LISTS
#define PTB_TRACE_TWO_LISTS 1    // 1: the fused kernel lists bounce rays and freshly started camera rays separately, so that warps
#endif                           // of the trace stage mostly hold rays of one kind (camera rays of neighbouring pixels are coherent):
                                 // +2.6 % on C2 (both builds), +3..5 % on the close cameras and C5
#ifndef PTB_STATUS_SMEM          // 1: the fused kernel keeps its chunk's status bytes in shared memory: +1 % at 256 threads per block,
#define PTB_STATUS_SMEM (!PTB_FAST)  // -1.7 % at 128 (eight blocks' worth of shared memory takes the next L1 carve-out step)
#endif
#ifndef PTB_MINB_WIDE            // fused kernel, launches that fill the chip: resident 128-thread units per SM the register budget
#if PTB_FAST                     // is sized for (8 -> 64 registers, 9 -> 56, 10 -> 48; the exact build's 256-thread blocks: 8 -> 64).
#define PTB_MINB_WIDE 9          // Measured with the early-store / late-load shade stage: fast 9 > 8 > 10 (18.39 / 18.66 / 19.03 ms),
#else                            // exact 8 > 10 > 9 (23.37 / 23.94 / 24.11 ms)
#define PTB_MINB_WIDE 8
#endif
#endif
// The chunk helpers are templates on SPT = slots per thread (chunk = PTB_CHUNK_THREADS * SPT slots): the fused kernel uses
// smaller chunks for small launches (a 600 x 400 frame has 117 chunks of 2048 slots -- less than one block per SM -- but
// 938 chunks of 256), everything else uses the default through the aliases below.

enum SlotStatus : unsigned char { ST_DONE = 0, ST_TRACE = 1, ST_HIT = 2, ST_MISS = 3, ST_TRACE_NEW = 4 };  // _NEW: a fresh camera ray (fused kernel only)

template <int SPT = PTB_CHUNK_SPT>
struct ChunkSharedT {
    static constexpr unsigned int CHUNK = PTB_CHUNK_THREADS * SPT;
    unsigned short list[PTB_CHUNK_THREADS * SPT];  // slot offsets inside the chunk, ascending
    unsigned int warp_sums[PTB_CHUNK_THREADS / 32];
    unsigned int warp_sums_b[PTB_CHUNK_THREADS / 32];
    unsigned int n;                  // list length
    unsigned int next;               // dynamic fetch cursor of the trace stage
    unsigned int count[4];           // per-block totals: segments, hits, misses, -
};
using ChunkShared = ChunkSharedT<>;

// Compacts the offsets of the chunk's slots whose status == want into sh.list (ascending).  Block-uniform result.
// bit k of the result is set when status byte k of this thread's PTB_CHUNK_SPT slots equals want (slots beyond n_slots
// read as ST_DONE; `want` is never ST_DONE)
// status bytes of this thread's SPT slots, packed little-endian into 64-bit words (slots beyond n_slots read as ST_DONE)
template <int SPT>
PTB_DEV void chunk_status_words(const unsigned char* __restrict__ status, uint32_t first, uint32_t n_slots, unsigned long long* words) {
    if (SPT >= 8) {
#pragma unroll
        for (int w = 0; w < (SPT + 7) / 8; ++w) {
            const uint32_t f0 = first + 8u * w;
            unsigned long long bytes = 0ull;
#if PTB_STATUS_SMEM
            if (f0 + 8u <= n_slots) bytes = *reinterpret_cast<const unsigned long long*>(status + f0);  // generic: shared or global
#else
            if (f0 + 8u <= n_slots) bytes = __ldcs(reinterpret_cast<const unsigned long long*>(status + f0));
#endif
            else if (f0 < n_slots) for (uint32_t k = 0; f0 + k < n_slots; ++k) bytes |= (unsigned long long)status[f0 + k] << (8u * k);
            words[w] = bytes;
        }
    } else {
        unsigned long long bytes = 0ull;
        if (first + (uint32_t)SPT <= n_slots) {
            if (SPT == 4) bytes = *reinterpret_cast<const unsigned int*>(status + first);
            else if (SPT == 2) bytes = *reinterpret_cast<const unsigned short*>(status + first);
            else bytes = status[first];
        } else if (first < n_slots) for (uint32_t k = 0; first + k < n_slots; ++k) bytes |= (unsigned long long)status[first + k] << (8u * k);
        words[0] = bytes;
    }
}
// bit k of the result is set when status byte k equals want (`want` is never ST_DONE, so padding never matches)
template <int SPT>
PTB_DEV unsigned int chunk_match(const unsigned long long* words, unsigned char want) {
    unsigned int m = 0;
#pragma unroll
    for (int k = 0; k < SPT; ++k) m |= (((unsigned int)(words[k >> 3] >> (8 * (k & 7))) & 0xffu) == (unsigned int)want ? 1u : 0u) << k;
    return m;
}

// Compacts the offsets of the chunk's slots whose status == want into sh.list (ascending).  Block-uniform result.
template <int SPT>
PTB_DEV unsigned int chunk_build_list(ChunkSharedT<SPT>& sh, const unsigned char* __restrict__ status, uint32_t base,
                                      uint32_t n_slots, unsigned char want) {
    const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long words[(SPT + 7) / 8];
    chunk_status_words<SPT>(status, base + tid * (unsigned int)SPT, n_slots, words);
    unsigned int match = chunk_match<SPT>(words, want);
    const unsigned int mine = (unsigned int)__popc(match);
    unsigned int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, off); if ((int)lane >= off) incl += y; }
    if (lane == 31u) sh.warp_sums[warp] = incl;
    __syncthreads();
    unsigned int warp_off = 0, total = 0;
#pragma unroll
    for (unsigned int w = 0; w < PTB_CHUNK_THREADS / 32; ++w) { const unsigned int v = sh.warp_sums[w]; if (w < warp) warp_off += v; total += v; }
    unsigned int pos = warp_off + incl - mine;
    unsigned int m = match;
    while (m) { const int k = __ffs(m) - 1; m &= m - 1u; sh.list[pos++] = (unsigned short)(tid * (unsigned int)SPT + (unsigned int)k); }
    if (tid == 0) { sh.n = total; sh.next = 0; }
    __syncthreads();
    return total;
}

// One pass over the chunk's status bytes builds TWO lists: slots in state want_a ascending from the front of sh.list,
// slots in state want_b ascending at its back (sh.list[CHUNK - nb ..)); want_b = 0xff matches nothing.  Block-uniform.
template <int SPT>
PTB_DEV void chunk_build_two(ChunkSharedT<SPT>& sh, const unsigned char* __restrict__ status, uint32_t base, uint32_t n_slots,
                             unsigned char want_a, unsigned char want_b, unsigned int* na, unsigned int* nb) {
    const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long words[(SPT + 7) / 8];
    chunk_status_words<SPT>(status, base + tid * (unsigned int)SPT, n_slots, words);
    unsigned int ma = chunk_match<SPT>(words, want_a), mb = chunk_match<SPT>(words, want_b);
    // one packed inclusive warp scan for both counts (each <= 512 per warp: 16 bits are plenty)
    const unsigned int mine = (unsigned int)__popc(ma) | ((unsigned int)__popc(mb) << 16);
    unsigned int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const unsigned int y = __shfl_up_sync(0xffffffffu, incl, off); if ((int)lane >= off) incl += y; }
    if (lane == 31u) { sh.warp_sums[warp] = incl & 0xffffu; sh.warp_sums_b[warp] = incl >> 16; }
    __syncthreads();
    unsigned int off_a = 0, tot_a = 0, off_b = 0, tot_b = 0;
#pragma unroll
    for (unsigned int w = 0; w < PTB_CHUNK_THREADS / 32; ++w) {
        const unsigned int va = sh.warp_sums[w], vb = sh.warp_sums_b[w];
        if (w < warp) { off_a += va; off_b += vb; }
        tot_a += va; tot_b += vb;
    }
    unsigned int pa = off_a + (incl & 0xffffu) - (mine & 0xffffu);
    unsigned int pb = ChunkSharedT<SPT>::CHUNK - tot_b + off_b + (incl >> 16) - (mine >> 16);
    while (ma) { const int k = __ffs(ma) - 1; ma &= ma - 1u; sh.list[pa++] = (unsigned short)(tid * (unsigned int)SPT + (unsigned int)k); }
    while (mb) { const int k = __ffs(mb) - 1; mb &= mb - 1u; sh.list[pb++] = (unsigned short)(tid * (unsigned int)SPT + (unsigned int)k); }
    if (tid == 0) { sh.n = tot_a; sh.next = 0; }
    __syncthreads();
    *na = tot_a; *nb = tot_b;
}

// ---- stage bodies over one chunk ---------------------------------------------------------------------------
template <bool COUNT, int QUANTUM, int SPT, int WIDTH = 0>
PTB_DEV void chunk_stage_trace(ChunkSharedT<SPT>& sh, const SceneView& s, const FrameView& f, const PathView& p,
                               unsigned char* __restrict__ status, uint32_t base, unsigned int n, bool first_iteration,
                               TravCounters& tc, uint32_t sbase = 0, unsigned int n_front = 0xffffffffu) {
    // the list is [0, n_front) at the front of sh.list and the remaining n - n_front entries at its back (chunk_build_two)
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    __align__(16) int stack[PTB_BVH_STACK];
    Trav t;
    t.node = PTB_TRAV_SENTINEL; t.sp = 0; t.grp = 0u; t.best.prim = -1; t.best.t = 0.0f; t.best.b1 = 0.0f; t.best.b2 = 0.0f;
    uint32_t slot = 0;
    bool have = false, exhausted = false;
    unsigned int hits = 0;
    for (;;) {
        __syncwarp();
        if (!exhausted) {
            const unsigned need = __ballot_sync(0xffffffffu, !have);
            if (need) {
                const int leader = __ffs(need) - 1;
                const unsigned int cnt = (unsigned int)__popc(need);
                unsigned int b0 = 0;
                if ((int)lane == leader) b0 = atomicAdd(&sh.next, cnt);
                b0 = __shfl_sync(0xffffffffu, b0, leader);
                const unsigned int idx = b0 + (unsigned int)__popc(need & lt_mask);
                if (!have && idx < n) {
                    slot = base + sh.list[idx < n_front ? idx : ChunkSharedT<SPT>::CHUNK - n + idx];
                    const float4 o4 = ldp(&p.ray_o[slot]), d4 = ldp(&p.ray_d[slot]);
                    trav_begin_any<WIDTH>(t, stack, s, mk3(o4), mk3(d4), f.tmin, f.tmax);
                    have = true;
                }
                if (b0 + cnt >= n) exhausted = true;  // warp-uniform
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (have && trav_run_any<COUNT, WIDTH>(t, stack, s, QUANTUM, &tc)) {
            have = false;
            stp(&p.hit[slot], make_float4(t.best.t, t.best.b1, t.best.b2, __int_as_float(t.best.prim)));
            const bool is_hit = t.best.prim >= 0;
            status[slot - sbase] = is_hit ? ST_HIT : ST_MISS;
            hits += is_hit ? 1u : 0u;
            if (first_iteration && f.aux_primary && slot < f.n_pixels) f.aux_primary[(size_t)image_row(f, slot / f.W) * f.W + slot % f.W] = t.best.prim;
        }
    }
    for (int off = 16; off > 0; off >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, off);
    if (lane == 0u && hits) atomicAdd(&sh.count[1], hits);
}

template <int SPT>
PTB_DEV void chunk_stage_shade(ChunkSharedT<SPT>& sh, const SceneView& s, const FrameView& f, const PathView& p,
                               unsigned char* __restrict__ status, uint32_t base, unsigned int n, unsigned int list_off = 0) {
    for (unsigned int i = threadIdx.x; i < n; i += PTB_CHUNK_THREADS) {
        const uint32_t slot = base + sh.list[list_off + i];
        const float4 o4 = ldp(&p.ray_o[slot]), d4 = ldp(&p.ray_d[slot]), h4 = ldp(&p.hit[slot]), as = ldp(&p.atten_seed[slot]);
        const uint4 mi = ldp(&p.misc[slot]);
        Bounce b;
        b.atten = mk3(as); b.seed = __float_as_uint(as.w);
        const int depth = (int)mi.y;
        closest_hit(s, f, __float_as_int(h4.w), h4.y, h4.z, h4.x, mk3(o4), mk3(d4), depth, b);
        status[slot] = after_segment(f, p, slot, b, mi.x, depth, mi.z) ? ST_TRACE : ST_DONE;
    }
}

template <int SPT>
PTB_DEV void chunk_stage_miss(ChunkSharedT<SPT>& sh, const SceneView& s, const FrameView& f, const PathView& p,
                              unsigned char* __restrict__ status, uint32_t base, unsigned int n, unsigned int list_off = 0) {
    for (unsigned int i = threadIdx.x; i < n; i += PTB_CHUNK_THREADS) {
        const uint32_t slot = base + sh.list[list_off + i];
        const float4 d4 = ldp(&p.ray_d[slot]), as = ldp(&p.atten_seed[slot]);
        const uint4 mi = ldp(&p.misc[slot]);
        const float3 ray_dir = normalize(mk3(d4));
        const float u = 0.5f + AR_DIVC(det_atan2f(ray_dir.z, ray_dir.x), 2.0f * PTB_PI_F);
        const float v = 0.5f - AR_DIVC(det_asinf(ray_dir.y), PTB_PI_F);
        const float4 hdr = sample_env(s.env, s.env_w, s.env_h, u, v);
        Bounce b;
        b.atten = mk3(as); b.seed = __float_as_uint(as.w);
        b.radiance = mk3(0.0f) + b.atten * mk3(hdr);
        b.origin = mk3(0.0f); b.direction = mk3(0.0f);
        b.done = 1;
        status[slot] = after_segment(f, p, slot, b, mi.x, (int)mi.y, mi.z) ? ST_TRACE : ST_DONE;
    }
}

// Shade and miss as ONE pass over both lists of chunk_build_two (hits at the front, misses at the back): the raygen-side
// end of segment (after_segment: Russian roulette, accumulation, path regeneration) exists once in the instruction
// stream, there is one stage barrier less per iteration, and the per-stage rounding loss (a list of n items keeps
// ceil(n / threads) rounds busy) is paid once instead of twice.  Only the warp that straddles the hit/miss boundary
// executes both bodies.
template <int SPT>
PTB_DEV void chunk_stage_shade_miss(ChunkSharedT<SPT>& sh, const SceneView& s, const FrameView& f, const PathView& p,
                                    unsigned char* __restrict__ status, uint32_t base, unsigned int n_hit, unsigned int n_miss,
                                    uint32_t sbase = 0, bool mark_new = false) {
    const unsigned int total = n_hit + n_miss;
#if PTB_SHADE_DYNAMIC
    // warps take 32 consecutive items at a time from the list cursor (reset by chunk_build_two): a warp that drew cheap items
    // (misses, Russian-roulette kills) goes on to the next batch instead of waiting at the stage barrier
    for (;;) {
        unsigned int b0 = 0;
        if ((threadIdx.x & 31u) == 0u) b0 = atomicAdd(&sh.next, 32u);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (b0 >= total) break;
        const unsigned int i = b0 + (threadIdx.x & 31u);
        if (i >= total) continue;
#else
    for (unsigned int i = threadIdx.x; i < total; i += PTB_CHUNK_THREADS) {
#endif
        const bool is_hit = i < n_hit;
        const uint32_t slot = base + sh.list[is_hit ? i : ChunkSharedT<SPT>::CHUNK - total + i];
        // Hits: only the words the hit shader needs are read up front (payload seed, depth); the attenuation and the raygen-side
        // words are read AFTER it and applied to what it returns for an attenuation of 1 (x * 1 and 0 + x are exact, so the
        // values are the same).  With the bounce ray stored early (closest_hit<true>) this takes ~12 live values out of the BSDF
        // code: spills of the 64-register build 200 -> 76 B stores, 116 -> 52 B loads; +1 % fast, +4 % exact, and the fast build
        // then runs best at 56 registers / 9 blocks per SM (PTB_MINB_WIDE): +2.3 to +3.6 % in total (profiles/r2_experiments.md).
        const float4 d4 = ldp(&p.ray_d[slot]);
        float4 as; uint4 mi;
        Bounce b;
        if (is_hit) {
            b.seed = __ldcs(reinterpret_cast<const unsigned int*>(&p.atten_seed[slot]) + 3);
            const int depth_h = (int)__ldcs(reinterpret_cast<const unsigned int*>(&p.misc[slot]) + 1);
            b.atten = mk3(1.0f);
            const float4 o4 = ldp(&p.ray_o[slot]), h4 = ldp(&p.hit[slot]);
            closest_hit<true>(s, f, __float_as_int(h4.w), h4.y, h4.z, h4.x, mk3(o4), mk3(d4), depth_h, b, &p.ray_o[slot], &p.ray_d[slot]);
            as = ldp(&p.atten_seed[slot]); mi = ldp(&p.misc[slot]);
            b.atten = mk3(as) * b.atten;
            b.radiance = mk3(as) * b.radiance;
        } else {
            as = ldp(&p.atten_seed[slot]); mi = ldp(&p.misc[slot]);
            b.atten = mk3(as); b.seed = __float_as_uint(as.w);
            const float3 ray_dir = normalize(mk3(d4));
            const float u = 0.5f + AR_DIVC(det_atan2f(ray_dir.z, ray_dir.x), 2.0f * PTB_PI_F);
            const float v = 0.5f - AR_DIVC(det_asinf(ray_dir.y), PTB_PI_F);
            const float4 hdr = sample_env(s.env, s.env_w, s.env_h, u, v);
            b.radiance = mk3(0.0f) + b.atten * mk3(hdr);
            b.origin = mk3(0.0f); b.direction = mk3(0.0f);
            b.done = 1;
        }
        const int next = after_segment<true>(f, p, slot, b, mi.x, (int)mi.y, mi.z);
        status[slot - sbase] = next == 0 ? ST_DONE : (next == 2 && mark_new ? ST_TRACE_NEW : ST_TRACE);
    }
}

template <int SPT>
PTB_DEV void chunk_flush_counts(ChunkSharedT<SPT>& sh, unsigned long long* __restrict__ totals, unsigned long long* trav_stats,
                                const TravCounters& tc, bool count) {
    __syncthreads();
    if (threadIdx.x == 0) {
        if (sh.count[0]) atomicAdd(&totals[0], (unsigned long long)sh.count[0]);
        if (sh.count[1]) atomicAdd(&totals[1], (unsigned long long)sh.count[1]);
        if (sh.count[0] - sh.count[1]) atomicAdd(&totals[2], (unsigned long long)(sh.count[0] - sh.count[1]));
    }
    if (count) {
        unsigned long long a = tc.nodes, b = tc.tris;
        for (int off = 16; off > 0; off >>= 1) { a += __shfl_xor_sync(0xffffffffu, a, off); b += __shfl_xor_sync(0xffffffffu, b, off); }
        if ((threadIdx.x & 31u) == 0u) { atomicAdd(&trav_stats[0], a); atomicAdd(&trav_stats[1], b); }
    }
}

// ---- drivers ------------------------------------------------------------------------------------------------
// camera rays of sample 0 (the raygen stage) for the chunked pool: same as k_raygen_init + status
__global__ void __launch_bounds__(256) k_chunk_raygen(FrameView f, PathView p, unsigned char* __restrict__ status) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_slots) return;
    const uint32_t pix = i % f.n_pixels, sub = i / f.n_pixels;
    const uint32_t ix = pix % f.W, iy = image_row(f, pix / f.W);
    stp(&p.pixsum[i], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
    if (iy >= f.H) { status[i] = ST_DONE; return; }  // padding rows of the last interleaved strip
    uint32_t seed = iy * f.W + ix + ((uint32_t)f.subframe + sub) * f.W * f.H;  // cu:316
    float3 o, d;
    start_sample(f, ix, iy, seed, o, d);
    stp(&p.ray_o[i], make_float4(o.x, o.y, o.z, 0.0f));
    stp(&p.ray_d[i], make_float4(d.x, d.y, d.z, 0.0f));
    stp(&p.atten_seed[i], make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed)));
    stp(&p.misc[i], make_uint4(seed, (uint32_t)f.max_depth, 0u, __float_as_uint(-1.0f)));  // .w: no BSDF pdf yet (linear.cuh)
    status[i] = ST_TRACE;
}

template <bool COUNT, int QUANTUM>
__global__ void __launch_bounds__(PTB_CHUNK_THREADS) k_chunk_trace(SceneView s, FrameView f, PathView p, unsigned char* status,
                                                                  unsigned long long* totals, unsigned long long* trav_stats, int iter) {
    __shared__ ChunkShared sh;
    const uint32_t base = blockIdx.x * PTB_CHUNK;
    if (threadIdx.x < 4) sh.count[threadIdx.x] = 0;
    const unsigned int n = chunk_build_list(sh, status, base, p.n_slots, ST_TRACE);
    if (n == 0) return;
    if (threadIdx.x == 0) sh.count[0] = n;
    TravCounters tc; tc.nodes = 0; tc.tris = 0;
    chunk_stage_trace<COUNT, QUANTUM>(sh, s, f, p, status, base, n, iter == 0, tc);
    chunk_flush_counts(sh, totals, trav_stats, tc, COUNT);
}

__global__ void __launch_bounds__(PTB_CHUNK_THREADS) k_chunk_shade(SceneView s, FrameView f, PathView p, unsigned char* status) {
    __shared__ ChunkShared sh;
    const uint32_t base = blockIdx.x * PTB_CHUNK;
    const unsigned int n = chunk_build_list(sh, status, base, p.n_slots, ST_HIT);
    if (n) chunk_stage_shade(sh, s, f, p, status, base, n);
}

__global__ void __launch_bounds__(PTB_CHUNK_THREADS) k_chunk_miss(SceneView s, FrameView f, PathView p, unsigned char* status) {
    __shared__ ChunkShared sh;
    const uint32_t base = blockIdx.x * PTB_CHUNK;
    const unsigned int n = chunk_build_list(sh, status, base, p.n_slots, ST_MISS);
    if (n) chunk_stage_miss(sh, s, f, p, status, base, n);
}

// One block = one chunk, from the first camera ray to the last sample of its pixels.  Two list passes per wavefront
// iteration (one code copy, alternating phases): {TRACE, BUSY} before the trace stage, {HIT, MISS} before shade + miss.
// totals[3] is not touched here (launch count is added by k_fold_counters' sibling on the host path).
// WIDTH: the tree the trace stage walks (2, 4, 8: one instantiation each, so that the 2-wide kernel does not carry the wide
// traversals' code; 0: decided at run time, the instrumented variant).
template <bool COUNT, int QUANTUM, int MINB, int SPT, int WIDTH>
__global__ void __launch_bounds__(PTB_CHUNK_THREADS, (MINB * 128 + PTB_CHUNK_THREADS - 1) / PTB_CHUNK_THREADS) k_chunk_fused(SceneView s, FrameView f, PathView p, unsigned char* status,
                                                                  unsigned long long* totals, unsigned long long* trav_stats,
                                                                  unsigned int* max_iters_seen) {
    __shared__ ChunkSharedT<SPT> sh;
    const uint32_t base = blockIdx.x * ChunkSharedT<SPT>::CHUNK;
    if (threadIdx.x < 4) sh.count[threadIdx.x] = 0;
    TravCounters tc; tc.nodes = 0; tc.tris = 0;
    unsigned int iter = 0;
#if PTB_STATUS_SMEM
    // the chunk's status bytes live in shared memory for the whole life of the block (read once from what raygen wrote):
    // the list pass at the top of every phase and the status stores of the stages never leave the SM
    __shared__ __align__(8) unsigned char st_sh[ChunkSharedT<SPT>::CHUNK];
    for (unsigned int k = threadIdx.x; k < ChunkSharedT<SPT>::CHUNK; k += PTB_CHUNK_THREADS) st_sh[k] = base + k < p.n_slots ? status[base + k] : (unsigned char)ST_DONE;
    __syncthreads();
    unsigned char* const stp_ = st_sh;
    const uint32_t sbase = base, sn = ChunkSharedT<SPT>::CHUNK;   // padding reads as ST_DONE, so the list pass needs no bound
#else
    unsigned char* const stp_ = status;
    const uint32_t sbase = 0, sn = p.n_slots;
#endif
    for (unsigned int phase = 0;; phase ^= 1u) {
        unsigned int na, nb;
#if PTB_TRACE_TWO_LISTS
        chunk_build_two(sh, stp_, base - sbase, sn, phase ? ST_HIT : ST_TRACE, phase ? ST_MISS : ST_TRACE_NEW, &na, &nb);
#else
        chunk_build_two(sh, stp_, base - sbase, sn, phase ? ST_HIT : ST_TRACE, phase ? ST_MISS : (unsigned char)0xff, &na, &nb);
#endif
        if (phase == 0u) {
            if (na + nb == 0u) break;  // every pixel of the chunk has finished its samples
            if (threadIdx.x == 0) sh.count[0] += na + nb;
            chunk_stage_trace<COUNT, QUANTUM, SPT, WIDTH>(sh, s, f, p, stp_, base, na + nb, iter == 0, tc, sbase, na);
            ++iter;
        } else {
            chunk_stage_shade_miss(sh, s, f, p, stp_, base, na, nb, sbase, PTB_TRACE_TWO_LISTS != 0);
        }
        __syncthreads();  // status / hit records of this chunk are block-visible from here on
    }
    chunk_flush_counts(sh, totals, trav_stats, tc, COUNT);
    if (threadIdx.x == 0) atomicMax(max_iters_seen, iter);
}

// ---- host-side launchers shared by both builds of the kernels (renderer.cu calls ptb::..., fast_kernels.cu wraps
// ptb_fast::...).  Grid sizes are derived HERE from the slot count, because the two builds may use different block sizes
// (PTB_CHUNK_THREADS is a per-translation-unit constant). ------------------------------------------------------------
struct ChunkLaunch {
    SceneView s; FrameView f; PathView p;
    unsigned char* status; unsigned long long* totals; unsigned long long* trav_stats; unsigned int* max_iters;
    int num_sms;
    int spt_request;   // slots per thread asked for by the caller: 0 = by launch size, else 1, 2, 4 or 8
    int count;         // 1: count nodes visited / triangles tested (2048-slot chunks only)
};

inline void launch_chunk_raygen(const ChunkLaunch& a, cudaStream_t st) {
    k_chunk_raygen<<<(a.p.n_slots + 255u) / 256u, 256, 0, st>>>(a.f, a.p, a.status);
}

// Chunk size by launch size: 8 slots per thread when that still gives 8 waves of blocks, otherwise fewer slots per thread so
// that a small frame (the reference's 600 x 400 / 1600 x 1200 launches of one subframe) spreads over the whole chip instead
// of running its per-pixel sample chains on a few warps per SM (profiles/r1_experiments.md: 600 x 400 4.66 -> 2.50 ms).
inline void launch_chunk_fused(const ChunkLaunch& a, cudaStream_t st) {
    const uint32_t slots = a.p.n_slots;
    const uint32_t full = (uint32_t)a.num_sms * (1024u / PTB_CHUNK_THREADS);  // resident blocks at 64 registers
    int spt = 8;
    while (spt > 1 && (slots + PTB_CHUNK_THREADS * (uint32_t)spt - 1u) / (PTB_CHUNK_THREADS * (uint32_t)spt) < 8u * full) spt >>= 1;
    if (a.spt_request) spt = a.spt_request;
    if (a.count) spt = 8;  // the counting variant is instantiated for the largest chunks only: size the grid for it
    const uint32_t chunk = PTB_CHUNK_THREADS * (uint32_t)spt;
    const uint32_t chunks = (slots + chunk - 1u) / chunk;
    // 64 registers once there are enough chunks to keep that many blocks busy, the unconstrained ~80-register build for
    // launches that cannot fill the chip anyway
    const bool wide = chunks >= full;
    const int width = a.s.nodes8 ? 8 : (a.s.nodes4 ? 4 : 2);
#define PTB_CF_LAUNCH(COUNT, MINB, SPT, WIDTH) k_chunk_fused<COUNT, PTB_TRACE_QUANTUM, MINB, SPT, WIDTH><<<chunks, PTB_CHUNK_THREADS, 0, st>>>(a.s, a.f, a.p, a.status, a.totals, a.trav_stats, a.max_iters)
#define PTB_CF_BY_SPT(MINB, WIDTH) do { if (spt == 8) PTB_CF_LAUNCH(false, MINB, 8, WIDTH); else if (spt == 4) PTB_CF_LAUNCH(false, MINB, 4, WIDTH); \
                                        else if (spt == 2) PTB_CF_LAUNCH(false, MINB, 2, WIDTH); else PTB_CF_LAUNCH(false, MINB, 1, WIDTH); } while (0)
#define PTB_CF_BY_WIDTH(MINB) do { if (width == 8) PTB_CF_BY_SPT(MINB, 8); else if (width == 4) PTB_CF_BY_SPT(MINB, 4); else PTB_CF_BY_SPT(MINB, 2); } while (0)
    if (a.count) PTB_CF_LAUNCH(true, 5, 8, 0);
    else if (wide) PTB_CF_BY_WIDTH(PTB_MINB_WIDE);
    else PTB_CF_BY_WIDTH(5);
#undef PTB_CF_BY_WIDTH
#undef PTB_CF_BY_SPT
#undef PTB_CF_LAUNCH
}

// one wavefront iteration of the stage-kernel pipeline (pipeline 2): 0 = trace, 1 = shade, 2 = miss
inline void launch_chunk_stage(const ChunkLaunch& a, int stage, int iter, cudaStream_t st) {
    const uint32_t chunks = (a.p.n_slots + PTB_CHUNK - 1u) / PTB_CHUNK;
    if (stage == 0) {
        if (a.count) k_chunk_trace<true, PTB_TRACE_QUANTUM><<<chunks, PTB_CHUNK_THREADS, 0, st>>>(a.s, a.f, a.p, a.status, a.totals, a.trav_stats, iter);
        else k_chunk_trace<false, PTB_TRACE_QUANTUM><<<chunks, PTB_CHUNK_THREADS, 0, st>>>(a.s, a.f, a.p, a.status, a.totals, a.trav_stats, iter);
    } else if (stage == 1) k_chunk_shade<<<chunks, PTB_CHUNK_THREADS, 0, st>>>(a.s, a.f, a.p, a.status);
    else k_chunk_miss<<<chunks, PTB_CHUNK_THREADS, 0, st>>>(a.s, a.f, a.p, a.status);
}

}  // namespace PTB_NS
