// fast_kernels.cu -- the block-local wavefront (chunked.cuh) compiled a second time with fast arithmetic:
// nvcc -fmad=true (FMA contraction), PTB_FAST = 1 (MUFU reciprocal / square root / sin / cos in the shading code,
// device_math.cuh), namespace ptb_fast.  Camera rays, traversal set-up and the ray-triangle test are written with
// un-contractable IEEE intrinsics (ex_*), so primary-hit IDs are the exact build's; images agree with the oracle to the
// tolerance tests/test_gpu_fast_mode.py states.  Selected per launch with ptb_render_cfg.arith_mode = 1.
#define PTB_NS ptb_fast
#define PTB_FAST 1
#include "chunked.cuh"
#include "fast_api.h"

namespace ptb_fast_api {
static ptb_fast::ChunkLaunch convert(const ChunkLaunchArgs& a) {
    ptb_fast::ChunkLaunch c;
    c.s = a.s; c.f = a.f; c.p = a.p;
    c.status = a.status; c.totals = a.totals; c.trav_stats = a.trav_stats; c.max_iters = a.max_iters;
    c.num_sms = a.num_sms; c.spt_request = a.spt_request; c.count = a.count;
    return c;
}
void raygen(const ChunkLaunchArgs& a, cudaStream_t st) { ptb_fast::launch_chunk_raygen(convert(a), st); }
void fused(const ChunkLaunchArgs& a, cudaStream_t st) { ptb_fast::launch_chunk_fused(convert(a), st); }
void stage(const ChunkLaunchArgs& a, int stage_id, int iter, cudaStream_t st) { ptb_fast::launch_chunk_stage(convert(a), stage_id, iter, st); }
}  // namespace ptb_fast_api
