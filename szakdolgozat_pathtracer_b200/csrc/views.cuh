// views.cuh -- plain-old-data views of the scene, the frame and the path pool that the host side hands to the kernels.
// They live in their own namespace because the kernels are compiled twice into two namespaces (PTB_NS = ptb: exact
// arithmetic, the parity-checked build; PTB_NS = ptb_fast: fast arithmetic, fast_kernels.cu) and both take the same
// argument structs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ptbv {

struct DevTexture { const void* data; int w, h; int fmt; };  // fmt: 0 none, 1 RGBA8, 2 float4
struct DevMaterial {
    DevTexture tex[4];  // albedo, roughness, normal, metallic
    float emission[3], diffuse[3], specular[3];
    float roughness;
    int metallic;
    int transparent;
};

struct SceneView {
    const float4* nodes; const float4* nodes4; const float4* tris;  // nodes4: 4-wide copy of the tree or nullptr
    const uint4* nodes8; const float4* tris8;                       // 8-wide quantised copy + its triangle order (bvh8.cuh) or nullptr
    const float4* verts; const float4* normals; const float2* uvs; const uint32_t* mat_ids;
    const DevMaterial* mats;
    const float4* env; int env_w, env_h;
};

// Exact unsigned division by a launch constant d for x < 2^31: q = (x * m) >> sh with m = ceil(2^sh / d),
// sh = 31 + ceil(log2 d) (the error m*d - 2^sh is < d <= 2^(sh-31), so it cannot carry into the quotient for x < 2^31).
// One 64-bit multiply instead of the ~20-instruction software division; every sample start needs three of them.
struct FastDiv { unsigned long long m; uint32_t sh, d; };

struct FrameView {
    uint32_t W, H;
    FastDiv div_w, div_pixels;  // by W and by n_pixels
    uint32_t row0;       // first image row rendered by this launch (row band; 0 for the whole frame)
    uint32_t il_n, il_r, il_h;  // il_n > 1: interleaved strips of il_h rows, this launch renders strips il_r, il_r + il_n, ...
    uint32_t n_pixels;   // W * rows of the band
    int n_subframes;     // subframes rendered by this launch as ONE wavefront (slot = sub * n_pixels + pixel)
    int subframe, dof;   // subframe = index of the first one
    float3 eye, U, V, Wv;
    int spp, max_depth;
    float tmin, tmax, dof_blur, focus_dist, nmap_strength;
    float exposure_scale, inv_gamma, contrast;
    int accumulate_mode, write_frame;
    float4* accum; uchar4* frame; int* aux_primary;
};

// Path pool, structure of arrays, one entry per pixel slot, 16-byte records.
struct PathView {
    float4* ray_o;       // origin.xyz, -
    float4* ray_d;       // direction.xyz, -
    float4* hit;         // t, b1, b2, prim (int bits)
    float4* atten_seed;  // attenuation.xyz, payload.seed (uint bits)
    uint4* misc;         // raygen seed, depth (int), sample index, -
    float4* pixsum;      // sum of finished samples .xyz
    uint32_t n_slots;
};

}  // namespace ptbv
