// pool.cuh -- the persistent block-local wavefront ("path pool"): pipeline 4 (an option; the default is pipeline 3,
// chunked.cuh, which is 9 % faster on the 64-spp headline launch).
//
// Why (profiles/r1_pool.md): in chunked.cuh a block is married to 2048 fixed slots until the LAST of them has finished
// its samples.  On C2 a slot needs 10.3 segments on average but the slowest slot of a chunk needs ~50, so a block spends
// most of its 53 iterations with a list that is 10-20 % full: per-iteration costs (list passes, barriers, the
// ceil(n / threads) rounding of every stage, the drain of the trace stage) are paid for a handful of rays, lane
// utilisation sits at 21 of 32, and every block pays its own tail.
//
// Here the blocks are PERSISTENT (grid = what fits on the chip) and own P = PTB_CHUNK *positions*, not slots.  A position
// whose slot has finished all its samples immediately takes the next unstarted slot of the launch (slot ids are handed
// out from one global counter in per-block batches, guided self-scheduling), using the very same code path that starts
// the next sample of a pixel.  Lists therefore stay full until the launch runs out of slots and there is ONE tail per
// launch instead of one per chunk.  Consequences for memory: path state is indexed by (block, position), so the pool is
// grid x P x 80 B (~95 MB on a B200: L2-sized) regardless of resolution and batch; only the per-slot sample sums
// (16 B, written once) scale with the frame.
//
// Stages per iteration (two list passes, chunk_build_two):
//   pass {TRACE}            -> trace: dynamic ray fetch from the shared list, traversal in quanta (trav_run)
//   pass {HIT | MISS, FREE} -> shade / miss / (re)start, then the raygen-side end of segment, one code copy:
//                              Russian roulette, accumulate, next sample of the pixel or next slot of the launch
// A slot's arithmetic is the same as in every other pipeline (closest_hit, sample_env, start_sample, trav_run are
// shared), and no result depends on which block or iteration processes a slot: output is bit-identical.
#pragma once
#include "chunked.cuh"

namespace PTB_NS {

enum PoolStatus : unsigned char { ST_FREE = 5, ST_IDLE = 6 };  // FREE: wants a new slot; IDLE: the launch has no more slots

struct PoolShared : ChunkShared {
    unsigned char status[PTB_CHUNK];   // per position
    uint32_t slot[PTB_CHUNK];          // launch-wide slot id (subframe * n_pixels + pixel) held by each position
    unsigned int batch_next, batch_end;  // this block's batch of unstarted slot ids
    unsigned int pool_dry;             // the launch-wide counter has run past n_slots
};

struct PoolView {
    float4* out_pixsum;        // per slot: sum of its finished samples (what k_resolve folds into the accumulator)
    unsigned int* next_slot;   // launch-wide cursor, zeroed by the host
    uint32_t n_slots;
    uint32_t grid;             // blocks of the launch (for the batch size)
};

// two lists in one pass over the block's status bytes (shared memory): a = want_a at the front, b = want_b1 or want_b2 at the back
PTB_DEV void pool_build_two(PoolShared& sh, unsigned char want_a, unsigned char want_b1, unsigned char want_b2,
                            unsigned int* na, unsigned int* nb) {
    const unsigned int tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    unsigned long long words[PTB_CHUNK_SPT / 8];
#pragma unroll
    for (int w = 0; w < PTB_CHUNK_SPT / 8; ++w) words[w] = *reinterpret_cast<const unsigned long long*>(sh.status + tid * PTB_CHUNK_SPT + 8 * w);
    unsigned int ma = chunk_match<PTB_CHUNK_SPT>(words, want_a), mb = chunk_match<PTB_CHUNK_SPT>(words, want_b1);
    if (want_b2 != want_b1) mb |= chunk_match<PTB_CHUNK_SPT>(words, want_b2);
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
    unsigned int pb = (unsigned int)PTB_CHUNK - tot_b + off_b + (incl >> 16) - (mine >> 16);
    while (ma) { const int k = __ffs(ma) - 1; ma &= ma - 1u; sh.list[pa++] = (unsigned short)(tid * (unsigned int)PTB_CHUNK_SPT + (unsigned int)k); }
    while (mb) { const int k = __ffs(mb) - 1; mb &= mb - 1u; sh.list[pb++] = (unsigned short)(tid * (unsigned int)PTB_CHUNK_SPT + (unsigned int)k); }
    if (tid == 0) { sh.n = tot_a; sh.next = 0; }
    __syncthreads();
    *na = tot_a; *nb = tot_b;
}

template <bool COUNT, int QUANTUM>
PTB_DEV void pool_stage_trace(PoolShared& sh, const SceneView& s, const FrameView& f, const PathView& p, uint32_t gbase,
                              unsigned int n, int* stack, TravCounters& tc) {
    const unsigned lane = threadIdx.x & 31u;
    const unsigned lt_mask = (1u << lane) - 1u;
    Trav t;
    t.node = PTB_TRAV_SENTINEL; t.sp = 0; t.grp = 0u; t.best.prim = -1; t.best.t = 0.0f; t.best.b1 = 0.0f; t.best.b2 = 0.0f;
    unsigned int pos = 0;
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
                    pos = sh.list[idx];
                    const float4 o4 = p.ray_o[gbase + pos], d4 = p.ray_d[gbase + pos];
                    trav_begin_any<0>(t, stack, s, mk3(o4), mk3(d4), f.tmin, f.tmax);
                    have = true;
                }
                if (b0 + cnt >= n) exhausted = true;  // warp-uniform
            }
        }
        if (!__any_sync(0xffffffffu, have)) break;
        if (have && trav_run_any<COUNT, 0>(t, stack, s, QUANTUM, &tc)) {
            have = false;
            p.hit[gbase + pos] = make_float4(t.best.t, t.best.b1, t.best.b2, __int_as_float(t.best.prim));
            const bool is_hit = t.best.prim >= 0;
            sh.status[pos] = is_hit ? ST_HIT : ST_MISS;
            hits += is_hit ? 1u : 0u;
            if (f.aux_primary) {  // primary-hit ids (tests): the camera ray of sample 0 of subframe 0
                const uint4 mi = p.misc[gbase + pos];
                const uint32_t slot = sh.slot[pos];
                if (mi.z == 0u && (int)mi.y == f.max_depth && slot < f.n_pixels)
                    f.aux_primary[(size_t)image_row(f, slot / f.W) * f.W + slot % f.W] = t.best.prim;
            }
        }
    }
    for (int off = 16; off > 0; off >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, off);
    if (lane == 0u && hits) atomicAdd(&sh.count[1], hits);
}

// next unstarted slot of the launch for one lane, out of the block's batch; false when the batch is empty
PTB_DEV bool pool_take_slot(PoolShared& sh, uint32_t* slot) {
    const unsigned int id = atomicAdd(&sh.batch_next, 1u);
    if (id >= sh.batch_end) return false;
    *slot = id;
    return true;
}

// shade / miss / start, then the raygen side of the segment end (cu:376-395) and the regeneration of the position
PTB_DEV void pool_stage_shade_miss(PoolShared& sh, const SceneView& s, const FrameView& f, const PathView& p, const PoolView& pv,
                                   uint32_t gbase, unsigned int n_hit, unsigned int n_other) {
    const unsigned int total = n_hit + n_other;
    for (unsigned int i = threadIdx.x; i < total; i += PTB_CHUNK_THREADS) {
        const bool is_hit = i < n_hit;
        const unsigned int pos = sh.list[is_hit ? i : (unsigned int)PTB_CHUNK - total + i];
        const uint32_t g = gbase + pos;
        const bool is_free = !is_hit && sh.status[pos] == ST_FREE;
        uint32_t slot = sh.slot[pos];
        uint32_t seed_rg = 0, sample = 0;
        int depth = 0;
        Bounce b;
        bool new_slot = is_free;   // the position needs a fresh slot
        bool again = false;        // the position has a ray to trace afterwards
        float4 sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        if (!is_free) {
            const float4 d4 = p.ray_d[g], as = p.atten_seed[g];
            const uint4 mi = p.misc[g];
            seed_rg = mi.x; depth = (int)mi.y; sample = mi.z;
            b.atten = mk3(as); b.seed = __float_as_uint(as.w);
            if (is_hit) {
                const float4 o4 = p.ray_o[g], h4 = p.hit[g];
                closest_hit(s, f, __float_as_int(h4.w), h4.y, h4.z, h4.x, mk3(o4), mk3(d4), depth, b);
            } else {
                const float3 ray_dir = normalize(mk3(d4));
                const float u = 0.5f + det_atan2f(ray_dir.z, ray_dir.x) / (2.0f * PTB_PI_F);
                const float v = 0.5f - det_asinf(ray_dir.y) / PTB_PI_F;
                const float4 hdr = sample_env(s.env, s.env_w, s.env_h, u, v);
                b.radiance = mk3(0.0f) + b.atten * mk3(hdr);
                b.origin = mk3(0.0f); b.direction = mk3(0.0f);
                b.done = 1;
            }
            // after_segment (kernels.cuh) with the regeneration step left to the common tail below
            const float pr = fmaxf(b.atten.x, fmaxf(b.atten.y, b.atten.z));
            bool done = b.done != 0;
            if (!done) done = myrnd(seed_rg) > pr;  // short-circuit: no draw when payload.done
            if (!done) {
                p.ray_o[g] = make_float4(b.origin.x, b.origin.y, b.origin.z, 0.0f);
                p.ray_d[g] = make_float4(b.direction.x, b.direction.y, b.direction.z, 0.0f);
                p.atten_seed[g] = make_float4(b.atten.x, b.atten.y, b.atten.z, __uint_as_float(b.seed));
                p.misc[g] = make_uint4(seed_rg, (uint32_t)(depth - 1), sample, 0u);
                sh.status[pos] = ST_TRACE;
                continue;
            }
            const float3 path_rgb = pr > 0.0f ? b.radiance / pr : mk3(0.0f);  // cu:384-387 (rule R4)
            sum = p.pixsum[g];
            sum.x = sum.x + path_rgb.x; sum.y = sum.y + path_rgb.y; sum.z = sum.z + path_rgb.z;
            sample += 1u;
            if (sample >= (uint32_t)f.spp) {  // the slot is finished: publish its sum, the position takes a new slot
                pv.out_pixsum[slot] = sum;
                new_slot = true;
            }
        }
        if (new_slot) {
            again = pool_take_slot(sh, &slot);
            if (!again) { sh.status[pos] = sh.pool_dry ? ST_IDLE : ST_FREE; continue; }
            sh.slot[pos] = slot;
            sample = 0u;
            sum = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            const uint32_t pix0 = slot % f.n_pixels, sub = slot / f.n_pixels;
            const uint32_t iy0 = image_row(f, pix0 / f.W);
            if (iy0 >= f.H) { pv.out_pixsum[slot] = sum; sh.status[pos] = ST_FREE; continue; }  // padding row of the last interleaved strip
            seed_rg = iy0 * f.W + pix0 % f.W + ((uint32_t)f.subframe + sub) * f.W * f.H;  // cu:316
        }
        // next camera ray of this position's pixel (cu:326-360): sample `sample`, stream state seed_rg
        p.pixsum[g] = sum;
        const uint32_t pix = fd_mod(slot, f.div_pixels);
        const uint32_t prow = fd_div(pix, f.div_w);
        float3 o, d;
        start_sample(f, pix - prow * f.W, image_row(f, prow), seed_rg, o, d);
        p.ray_o[g] = make_float4(o.x, o.y, o.z, 0.0f);
        p.ray_d[g] = make_float4(d.x, d.y, d.z, 0.0f);
        p.atten_seed[g] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed_rg));
        p.misc[g] = make_uint4(seed_rg, (uint32_t)f.max_depth, sample, 0u);
        sh.status[pos] = ST_TRACE;
    }
}

// thread 0: refill the block's batch of slot ids when it is empty (guided self-scheduling: batches shrink towards the end
// of the launch so that the last ones are spread over all blocks)
PTB_DEV void pool_refill_batch(PoolShared& sh, const PoolView& pv) {
    if (sh.pool_dry || sh.batch_next < sh.batch_end) return;
    const unsigned int seen = *((volatile unsigned int*)pv.next_slot);
    const unsigned int remaining = seen < pv.n_slots ? pv.n_slots - seen : 0u;
    unsigned int want = remaining / (2u * pv.grid);
    want = want < 64u ? 64u : (want > (unsigned int)PTB_CHUNK ? (unsigned int)PTB_CHUNK : want);
    const unsigned int b0 = atomicAdd(pv.next_slot, want);
    if (b0 >= pv.n_slots) { sh.pool_dry = 1u; sh.batch_next = 0u; sh.batch_end = 0u; return; }
    sh.batch_next = b0;
    sh.batch_end = b0 + want < pv.n_slots ? b0 + want : pv.n_slots;
}

template <bool COUNT, int QUANTUM, int MINB>
__global__ void __launch_bounds__(PTB_CHUNK_THREADS, (MINB * 128) / PTB_CHUNK_THREADS) k_pool_fused(SceneView s, FrameView f, PathView p, PoolView pv,
                                                                  unsigned long long* totals, unsigned long long* trav_stats,
                                                                  unsigned int* max_iters_seen) {
    __shared__ __align__(16) PoolShared sh;
    const uint32_t gbase = blockIdx.x * PTB_CHUNK;
    if (threadIdx.x < 4) sh.count[threadIdx.x] = 0;
    if (threadIdx.x == 0) { sh.batch_next = 0u; sh.batch_end = 0u; sh.pool_dry = 0u; }
    for (unsigned int i = threadIdx.x; i < PTB_CHUNK; i += PTB_CHUNK_THREADS) { sh.status[i] = ST_FREE; sh.slot[i] = 0u; }
    TravCounters tc; tc.nodes = 0; tc.tris = 0;
    __align__(16) int stack[PTB_BVH_STACK];
    __syncthreads();
    unsigned int iter = 0;
    for (;; ++iter) {
        unsigned int na, nb;
        pool_build_two(sh, ST_TRACE, 0xffu, 0xffu, &na, &nb);
        if (na) {
            if (threadIdx.x == 0) sh.count[0] += na;
            pool_stage_trace<COUNT, QUANTUM>(sh, s, f, p, gbase, na, stack, tc);
        }
        if (threadIdx.x == 0) pool_refill_batch(sh, pv);
        __syncthreads();  // hit records / status of the traced rays and the batch are block-visible
        pool_build_two(sh, ST_HIT, ST_MISS, ST_FREE, &na, &nb);
        if (na + nb == 0u) break;  // nothing traced, nothing free: every position is idle (the launch is out of slots)
        pool_stage_shade_miss(sh, s, f, p, pv, gbase, na, nb);
        __syncthreads();
    }
    chunk_flush_counts(sh, totals, trav_stats, tc, COUNT);
    if (threadIdx.x == 0) atomicMax(max_iters_seen, iter);
}

}  // namespace PTB_NS
