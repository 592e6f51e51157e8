// bvh_build.h -- device BVH object and its builder (bvh_build.cu).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/ptb.h"

#define PTB_WIDE_BVH_MIN_TRIS 1000000u  // bvh_width = 0 (auto): 4-wide traversal from this many triangles on (C5 with 0.33 M: -1 %, C4 with 4.6 M: +4 %)
#define PTB_BVH_MAX_DEPTH 128 // traversal stack entries (bvh.cuh: PTB_BVH_STACK)

namespace ptb {

struct DeviceBvh {
    float4* nodes = nullptr;  // n_nodes x 4 float4 (64 B each), layout in bvh.cuh
    float4* nodes4 = nullptr; // optional 4-wide copy: n_nodes x 8 float4 (128 B each, indexed like `nodes`), layout in bvh.cuh
    float4* tris = nullptr;   // n_tris x 3 float4 (48 B each), leaf order
    uint4* nodes8 = nullptr;  // optional 8-wide quantised copy: n_nodes8 x 5 uint4 (80 B each), layout in bvh8.cuh
    float4* tris8 = nullptr;  // its triangles (the leaf triangles again, contiguous per 8-wide node)
    uint32_t n_nodes = 0, n_tris = 0, n_nodes8 = 0;
};

// d_verts: 3 float4 per triangle (the reference's vertex buffer, optixSphere.cpp:871-894).
bool build_bvh(const float4* d_verts, uint32_t n_tris, const ptb_build_cfg& cfg, cudaStream_t stream, DeviceBvh& out,
               ptb_build_stats& stats, std::string& err);
void free_bvh(DeviceBvh& b);

}  // namespace ptb
