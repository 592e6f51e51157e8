// fast_api.h -- entry points of the fast-arithmetic build of the wavefront kernels (fast_kernels.cu, namespace ptb_fast).
// The argument struct is rebuilt from the shared views (views.cuh) on the other side, so no type of one kernel
// namespace crosses into the other.
#pragma once
#include <cuda_runtime.h>
#include "views.cuh"

namespace ptb_fast_api {
struct ChunkLaunchArgs {
    ptbv::SceneView s; ptbv::FrameView f; ptbv::PathView p;
    unsigned char* status; unsigned long long* totals; unsigned long long* trav_stats; unsigned int* max_iters;
    int num_sms, spt_request, count;
};
void raygen(const ChunkLaunchArgs& a, cudaStream_t st);
void fused(const ChunkLaunchArgs& a, cudaStream_t st);
void stage(const ChunkLaunchArgs& a, int stage, int iter, cudaStream_t st);
}  // namespace ptb_fast_api
