/* include/ptb.h -- C ABI of the B200-native path-tracing core ("ptb").
 *
 * Drop-in boundary for the render path of safardani/szakdolgozat-pathtracer.
 * The reference has no plugin/FFI interface: its integrator sits behind the
 * OptiX host API as called from main() (optixSphere.cpp:754-1538).  Each entry
 * point below replaces one group of those calls and cites it.  Plain pointers
 * and sizes only; no CUDA, OptiX or torch types appear in the signatures
 * (streams are passed as void* = cudaStream_t).
 *
 * Conventions
 *  - every function returns PTB_OK (0) or a PTB_ERR_* code; the message is
 *    available from ptb_last_error() (thread-local).  The reference throws
 *    from CUDA_CHECK/OPTIX_CHECK and exits 1 (optixSphere.cpp:1532-1537).
 *  - one context is used from one host thread at a time, like the reference's
 *    single-threaded render loop (optixSphere.cpp:1390-1437).
 *  - ptb_launch() is asynchronous on the given stream, like optixLaunch.  All
 *    launches of one context share its path pool and counters: ONE launch in
 *    flight per context (issue the next one on the same stream, or synchronise
 *    first); use one context per stream for concurrent launches.
 *  - a handle from ptb_accel_build() names the upload of the scene AS IT WAS;
 *    ptb_launch() fails with PTB_ERR_INVALID when the scene was modified since
 *    (materials, environment, geometry) until it is built again.
 *  - there is NO CPU fallback: every compute entry point fails with
 *    PTB_ERR_NO_DEVICE when no CUDA device is usable.
 */
#ifndef PTB_H
#define PTB_H
#include <stddef.h>
#include <stdint.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define PTB_OK 0
#define PTB_ERR_INVALID 1
#define PTB_ERR_IO 2
#define PTB_ERR_CUDA 3
#define PTB_ERR_UNSUPPORTED 4
#define PTB_ERR_NO_DEVICE 5

#define PTB_PIPELINE_QUEUES 1
#define PTB_PIPELINE_CHUNK_STAGES 2
#define PTB_PIPELINE_CHUNK_FUSED 3
#define PTB_PIPELINE_POOL_FUSED 4
#define PTB_PIPELINE_DEFAULT PTB_PIPELINE_CHUNK_FUSED

#define PTB_ARITH_EXACT 0
#define PTB_ARITH_FAST 1

/* ---- data layouts shared with the reference (optixSphere.h) ---------------
 * Layout-identical to the CUDA vector types the reference uses; the sizes and
 * offsets are pinned by tests against oracle/_ref/ref_probe. */
typedef struct ptb_float2 { float x, y; } ptb_float2;
typedef struct ptb_float3 { float x, y, z; } ptb_float3;
#if defined(__cplusplus)
typedef struct alignas(16) ptb_float4 { float x, y, z, w; } ptb_float4;
typedef struct alignas(4) ptb_uchar4 { unsigned char x, y, z, w; } ptb_uchar4;
#else
typedef struct ptb_float4 { _Alignas(16) float x; float y, z, w; } ptb_float4;
typedef struct ptb_uchar4 { _Alignas(4) unsigned char x; unsigned char y, z, w; } ptb_uchar4;
#endif

/* optixSphere.h:2-7 (128 bytes) */
typedef struct ptb_TriangleData {
    ptb_float4 v0, v1, v2, n0, n1, n2;
    ptb_float2 uv0, uv1, uv2;
} ptb_TriangleData;

/* optixSphere.h:10-31 (120 bytes).  frame_buffer / accum_buffer are DEVICE
 * pointers owned by the caller (or by a ptb_output); handle comes from
 * ptb_accel_build(). */
typedef struct ptb_Params {
    unsigned int image_width;
    unsigned int image_height;
    int origin_x;
    int origin_y;
    int subframe_index;
    ptb_uchar4* frame_buffer;
    ptb_float4* accum_buffer;
    bool dof;
    ptb_float3 eye, U, V, W;
    ptb_TriangleData* triangles; /* unused by the reference's device code */
    unsigned int num_triangles;  /* unused by the reference's device code */
    unsigned long long handle;   /* OptixTraversableHandle in the reference */
} ptb_Params;

/* optixSphere.h:58-63 */
typedef struct ptb_MissData {
    ptb_float4* hdr_image_data;
    int width, height;
} ptb_MissData;

/* optixSphere.h:67-102 (168 bytes).  In ptb_scene_set_materials() the texture
 * pointers are HOST pointers to float4 texels (the reference's upload format,
 * optixSphere.cpp:364-380); texcoords/vertices/normals are ignored (the scene
 * owns the geometry). */
typedef struct ptb_HitGroupData {
    ptb_float4* albedo_texture_data; int tex_width, tex_height; bool has_texture;
    ptb_float4* roughness_texture_data; int roughness_width, roughness_height; bool has_roughness_map;
    ptb_float4* normal_texture_data; int normal_width, normal_height; bool has_normal_map;
    ptb_float4* metallic_texture_data; int metallic_width, metallic_height; bool has_metallic_map;
    ptb_float2* texcoords;
    ptb_float4* vertices;
    ptb_float4* normals;
    ptb_float3 emission_color, diffuse_color, specular;
    float roughness;
    bool metallic;
    bool transparent;
} ptb_HitGroupData;

/* ---- opaque objects ---------------------------------------------------------- */
typedef struct ptb_context ptb_context;
typedef struct ptb_scene ptb_scene;
typedef struct ptb_output ptb_output;

/* ---- render configuration: the reference's compile-time literals ------------ */
typedef struct ptb_render_cfg {
    int32_t spp_per_launch;   /* optixSphere.cu:323   sample_batch_count = 10 */
    int32_t max_depth;        /* optixSphere.cu:360   payload.depth = 20 */
    float tmin, tmax;         /* optixSphere.cu:368-369   0.01, 1e16 */
    float dof_blur;           /* optixSphere.cu:285   0.01 */
    float focus_dist;         /* optixSphere.cu:329   1.0 */
    float nmap_strength;      /* optixSphere.cu:697   0.4 */
    float exposure;           /* optixSphere.cu:412   -0.5 */
    float gamma;              /* optixSphere.cu:425   2.2 */
    float contrast;           /* optixSphere.cu:432   1.25 */
    int32_t accumulate_mode;  /* 0: running average by subframe_index (optixSphere.cu:403-409);
                                 1: accum += launch mean (sample-split multi-GPU; resolve with ptb_resolve) */
    int32_t write_frame;      /* 1: tonemap into Params.frame_buffer (optixSphere.cu:411-435); 0: skip */
    int32_t env_importance_sampling; /* 0 = the reference estimator (BSDF sampling only; the parity-checked path).
                                 1 = OPTIONAL MODE BEYOND THE REFERENCE: linear estimator with next-event estimation of the
                                     environment through a prebuilt luminance*sin(theta) CDF, shadow rays and MIS;
                                 2 = the same linear estimator with BSDF sampling only (A/B baseline for mode 1).
                                 Modes 1/2 cannot reproduce the reference's expectation (its sample value is divided by
                                 max(attenuation), optixSphere.cu:382-387) and are validated on their own. */
    int32_t count_traversal;  /* 1: also count BVH nodes visited / triangles tested (slower) */
    int32_t profile_stages;   /* 1: bracket every stage kernel with CUDA events (ptb_launch_get_stage_ms) */
    int32_t subframes_per_launch; /* default 1.  n > 1: this one call renders subframes subframe_index .. +n-1 as ONE
                                 wavefront of n*W*H path slots (results are bit-identical to n consecutive calls).
                                 Keeps the GPU full through the tail of the per-pixel sample chains.  The pool is
                                 bounded by max_pool_bytes: larger launches run as batches of subframes. */
    int32_t pipeline;         /* 0 = default.  1: global ray queues, one kernel per stage and iteration;
                                 2: block-local wavefront over chunks of the path pool, one kernel per stage and
                                    iteration; 3: the same stages fused into one kernel per launch (default);
                                 4: persistent path pool (blocks own positions, slots are handed out from one counter;
                                    path state does not grow with the frame).  All produce bit-identical results. */
    int32_t row_begin, row_end; /* 0, 0 = the whole frame.  Otherwise only image rows [row_begin, row_end) are rendered (and only
                                 those rows of accum/frame/aux are touched): tile partitioning for multi-GPU single-pass frames.
                                 Every pixel is still seeded by its full-frame coordinates, so bands tile bit-identically. */
    int32_t row_interleave_count, row_interleave_index, row_interleave_height; /* count > 1: the frame is cut into strips of
                                 `height` rows and this launch renders strips index, index + count, ... (load-balanced tile
                                 partitioning: sky strips are cheap, strips over the mesh are not).  Chunked pipelines only. */
    int32_t* aux_primary_hit; /* optional DEVICE int32[W*H]: primitive hit by the first segment of sample 0, -1 = miss */
    int32_t chunk_slots_per_thread; /* pipeline 3: 0 = chosen from the launch size (default); 1, 2, 4 or 8 = slots per thread of
                                 a 256-thread block, i.e. 256 .. 2048 path slots per chunk.  Results do not depend on it. */
    int32_t arith_mode;       /* PTB_ARITH_EXACT (0, default): one fixed IEEE operation order, no FMA contraction, software
                                 sin/cos: accumulation buffer and frame are BIT-IDENTICAL to the CPU oracle.
                                 PTB_ARITH_FAST (1): the reference's own build mode (--use_fast_math):
                                 FMA contraction and MUFU reciprocal / square root / sin / cos in the shading code.  Camera
                                 rays and ray-triangle tests stay exact, so primary-hit IDs are bit-identical; images agree
                                 within the RMSE bound tests/test_gpu_fast_mode.py states.  Pipelines 2 and 3 only. */
    int64_t max_pool_bytes;   /* bound of the path-state pool (97 B per resident path slot); 0 = 2 GiB.  A launch of
                                 subframes_per_launch * W * H slots that would exceed it is rendered as consecutive batches
                                 of subframes inside ptb_launch(): same result bit for bit, bounded memory. */
    int32_t overlap_lanes;    /* consecutive SMALL launches (the reference's render loop issues one subframe of 600x400 .. 1600x1200
                                 pixels per optixLaunch, optixSphere.cpp:1390-1437) end in a long thin tail: every pixel's samples
                                 are one sequential chain.  0 (default) = automatic: in the reference's accumulate mode (accumulate_mode
                                 0) a launch of the default pipeline with fewer than 9 M path slots (subframes x pixels) renders on one
                                 of 4 internal streams with its own path pool, so that the tail of launch k overlaps the start of
                                 launch k + 1 (600x400: 2.15 -> 0.73 ms per launch); the accumulate / tonemap stage stays on the
                                 caller's stream, in call order -- buffers and results are exactly those of serial launches.
                                 1 = off; 2..4 = that many lanes (any accumulate mode).  Every lane in use owns a path pool of
                                 the launch's size (97 B per slot: 4 x 23 MB at 600x400, 4 x 186 MB at 1600x1200).  The render of such a launch does not wait for
                                 work the caller enqueued on `stream` before the call (only the accumulate stage does); launches
                                 with aux_primary_hit, counters, stage profiling or stream capture never overlap. */
    int32_t reserved0;
} ptb_render_cfg;

typedef struct ptb_launch_stats {
    uint64_t segments;     /* closest-hit ray casts = traceRadiance calls (optixSphere.cu:364) */
    uint64_t paths;        /* camera samples started */
    uint64_t hits, misses;
    uint64_t nodes_visited, tris_tested; /* only when count_traversal */
    uint32_t iterations;   /* wavefront iterations executed */
    uint32_t kernel_launches;
} ptb_launch_stats;

typedef struct ptb_build_cfg {
    int32_t max_leaf_size; /* triangles per leaf after refinement (1..8), default 4 */
    int32_t sah_refine;    /* 1: binned-SAH refinement of the LBVH (default); 0: plain LBVH */
    int32_t sah_bins;      /* default 16 */
    int32_t treelet_size;  /* primitives per refinement treelet, default 256 (the maximum; larger values are clamped) */
    int32_t morton_bits;   /* 30 (default: 10 bits per axis) or 63 (21 bits per axis) */
    int32_t bvh_width;     /* 0 (default): 4-wide traversal for large scenes, 2-wide otherwise; 2, 4 or 8 force one
                              (8: quantised 80-byte nodes, leaves of at most 3 triangles).  The hit rule does not depend on
                              the tree, so results are identical. */
} ptb_build_cfg;

typedef struct ptb_build_stats {
    uint32_t num_triangles, num_nodes, num_leaves, max_depth;
    float sah_cost;        /* sum(area*cost)/root area, Ct=1 Ci=1 */
    float build_ms;        /* device time of the build (CUDA events) */
    uint64_t bvh_bytes;
    uint32_t bvh_width;    /* 2, 4 or 8: the tree the traversal kernels walk (ptb_build_cfg.bvh_width = 0 picks by scene size) */
    float sah_cost_mesh;   /* sah_cost of the subtree beside the huge-primitive leaf under the root (the reference's floor quad
                              spans the scene and dominates sah_cost), relative to that subtree's own box; = sah_cost without such a leaf */
    uint32_t num_nodes8;   /* nodes of the 8-wide quantised tree (0 unless bvh_width == 8) */
} ptb_build_stats;

typedef struct ptb_material_info {
    float emission_color[3], diffuse_color[3], specular[3];
    float roughness;
    int32_t metallic, transparent;
    int32_t has_albedo, albedo_w, albedo_h;
    int32_t has_roughness, roughness_w, roughness_h;
    int32_t has_normal, normal_w, normal_h;
    int32_t has_metallic, metallic_w, metallic_h;
} ptb_material_info;

/* ---- errors -------------------------------------------------------------------- */
const char* ptb_last_error(void);
const char* ptb_version(void);

/* ---- context: optixInit / optixDeviceContextCreate (optixSphere.cpp:798-812),
 *      optixDeviceContextDestroy (optixSphere.cpp:1529) ------------------------- */
int ptb_context_create(int device, ptb_context** out);
void ptb_context_destroy(ptb_context* ctx);
int ptb_context_synchronize(ptb_context* ctx, void* stream);

/* ---- scene load: createSceneGeometry(files, scale) + flatten
 *      (optixSphere.cpp:400-649, 845-858).  Host-only; needs no device.
 *      Materials follow the <stem>_{albedo,roughness,normal,metallic}.png
 *      convention (optixSphere.cpp:522-546); untextured files get a random
 *      material drawn from std::mt19937(material_seed) (the reference seeds it
 *      from std::random_device, optixSphere.cpp:141-148).  A 2-triangle floor is
 *      appended at the lowest vertex (optixSphere.cpp:598-646). */
int ptb_scene_load_obj(const char* const* files, int n_files, float scale, uint32_t material_seed, ptb_scene** out);
/* raw variant: triangles as the reference holds them before flattening */
int ptb_scene_create(const ptb_TriangleData* tris, uint32_t n_tris, const uint32_t* mat_ids, ptb_scene** out);
/* the reference's asset-free demo scene (createSceneGeometry with loadFromFile = false, optixSphere.cpp:295-353,
 * 650-751): a 20x20 ground quad + three UV spheres (16 stacks x 32 slices, radius 1) with four fixed materials.
 * Dead code in the reference (loadFromFile is hard-coded true, optixSphere.cpp:829); kept as a regression scene. */
int ptb_scene_create_demo(ptb_scene** out);
/* hit-group table: optixSphere.cpp:1196-1261 */
int ptb_scene_set_materials(ptb_scene* scene, const ptb_HitGroupData* mats, int n);
/* environment: sutil::loadImage(exr) + MissData (optixSphere.cpp:835-836, 1161-1188) */
int ptb_scene_set_env_file(ptb_scene* scene, const char* path);
int ptb_scene_set_env_pixels(ptb_scene* scene, const float* rgba, int w, int h);
void ptb_scene_destroy(ptb_scene* scene);

/* host-side read-back (parity tests, tools) */
uint32_t ptb_scene_num_triangles(const ptb_scene* scene);
int ptb_scene_num_materials(const ptb_scene* scene);
int ptb_scene_copy_triangles(const ptb_scene* scene, ptb_TriangleData* out, uint32_t cap);
int ptb_scene_copy_material_ids(const ptb_scene* scene, uint32_t* out, uint32_t cap);
int ptb_scene_get_material(const ptb_scene* scene, int index, ptb_material_info* out);
/* kind: 0 albedo, 1 roughness, 2 normal, 3 metallic; out = w*h*4 floats (texel = byte/255.0f) */
int ptb_scene_copy_texture(const ptb_scene* scene, int material, int kind, float* out_rgba, size_t cap_floats);
int ptb_scene_env_size(const ptb_scene* scene, int* w, int* h);
int ptb_scene_copy_env(const ptb_scene* scene, float* out_rgba, size_t cap_floats);

/* ---- acceleration structure: optixAccelComputeMemoryUsage / optixAccelBuild /
 *      optixAccelCompact (optixSphere.cpp:860-968).  Uploads the scene to the
 *      context's device and builds the BVH there (LBVH from Morton codes, then
 *      binned-SAH refinement).  The handle goes into Params.handle. */
void ptb_default_build_cfg(ptb_build_cfg* cfg);
int ptb_accel_build(ptb_context* ctx, ptb_scene* scene, const ptb_build_cfg* cfg, void* stream,
                    unsigned long long* handle, ptb_build_stats* stats);
/* copies the flattened BVH back: nodes = 16 floats per node (64 B), tris = 12 floats per leaf-ordered triangle */
int ptb_accel_read(ptb_context* ctx, unsigned long long handle, float* nodes, uint32_t cap_nodes,
                   float* tris, uint32_t cap_tris, uint32_t* n_nodes, uint32_t* n_tris);
/* the 8-wide quantised copy (ptb_build_cfg.bvh_width = 8): nodes8 = 20 words per node (80 B), tris8 = 12 floats per triangle in
 * the 8-wide tree's order; *n_nodes8 = 0 when the handle has no such tree */
int ptb_accel_read8(ptb_context* ctx, unsigned long long handle, uint32_t* nodes8, uint32_t cap_nodes8,
                    float* tris8, uint32_t cap_tris, uint32_t* n_nodes8, uint32_t* n_tris);

/* ---- camera: sutil::Camera::UVWFrame as used by handleCameraUpdate /
 *      configureCamera (optixSphere.cpp:102-120, 238-247) ---------------------- */
void ptb_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy_deg, float aspect,
                    float U[3], float V[3], float W[3]);
/* fills eye/U/V/W with the reference camera: eye (0,2,6) -> (0,0,0), up +Y, fovY 50 */
void ptb_params_default_camera(ptb_Params* params);

/* ---- launch: Params upload + optixLaunch(pipeline, stream, d_param, sizeof(Params),
 *      &sbt, width, height, 1) (optixSphere.cpp:1403-1418, 1470-1479) ---------- */
void ptb_default_render_cfg(ptb_render_cfg* cfg);
int ptb_launch(ptb_context* ctx, const ptb_Params* params, const ptb_render_cfg* cfg, void* stream);
/* synchronises the stream of the last launch and returns its counters */
int ptb_launch_get_stats(ptb_context* ctx, ptb_launch_stats* out);
/* running totals over every launch of this context since the last reset, summed on the device at the
 * end of each launch (no host synchronisation inside ptb_launch): out[0] segments, [1] hits, [2] misses,
 * [3] launches.  Synchronises the last stream. */
int ptb_context_get_totals(ptb_context* ctx, uint64_t out[4], int reset);
/* device time per stage of the last launch that had profile_stages = 1, in ms:
 * out[0] raygen, [1] traversal, [2] shade, [3] miss, [4] accumulate/tonemap, [5] whole launch */
int ptb_launch_get_stage_ms(ptb_context* ctx, float out[6]);
/* accumulate/tonemap stage on its own (optixSphere.cu:401-435): frame = tonemap(accum * scale).
 * Used after a multi-GPU reduce of sum-mode accumulators. */
int ptb_resolve(ptb_context* ctx, const ptb_float4* accum, ptb_float4* accum_out, ptb_uchar4* frame, uint32_t n_pixels,
                float scale, const ptb_render_cfg* cfg, void* stream);
/* Multi-GPU accumulate/tonemap fused with its exchange over NVLink peer memory (reduce-scatter ->
 * tonemap -> gather, as ONE kernel per rank, no NCCL on the data path).  accums[k] is rank k's sum-mode accumulator
 * (own memory or a peer mapping from ptb_ipc_open); the kernel sums them in rank order for pixels
 * [first_pixel, first_pixel + n_pixels), scales, and stores the float4 result to accum_out and the tonemapped pixel to
 * frame -- both may be peer pointers (e.g. rank 0's buffers), which is the gather.  The caller orders it after every
 * rank's rendering (a stream-ordered barrier) and must barrier again before the root reads the frame. */
int ptb_resolve_peers(ptb_context* ctx, const ptb_float4* const* accums, int n_ranks, ptb_float4* accum_out, ptb_uchar4* frame,
                      uint32_t first_pixel, uint32_t n_pixels, float scale, const ptb_render_cfg* cfg, void* stream);
/* The same with the frame's earlier history: prev_accum (may be null, may alias accum_out) holds the mean of prev_weight
 * earlier subframes and enters the sum first as prev_accum * prev_weight; scale is then 1 / (prev_weight + new subframes).
 * This is how ptb_multi_launch() continues the reference's running average (optixSphere.cu:403-409) across launches. */
int ptb_resolve_peers_accumulate(ptb_context* ctx, const ptb_float4* const* accums, int n_ranks, const ptb_float4* prev_accum,
                                 float prev_weight, ptb_float4* accum_out, ptb_uchar4* frame, uint32_t first_pixel,
                                 uint32_t n_pixels, float scale, const ptb_render_cfg* cfg, void* stream);
/* ---- cross-rank ordering without NCCL (one process per GPU, e.g. under torchrun).  Every rank owns a flag block
 *      (ptb_peer_flags_create; export it with ptb_ipc_export, map the peers' with ptb_ipc_open).  Epochs only grow.
 *        ptb_peer_signal(kind 0)  stream-ordered "everything I enqueued so far (my rendering) is done": stores epoch into
 *                                 arrive[my_rank] of EVERY rank's block (peer stores over NVLink);
 *        ptb_resolve_peers_sync   ptb_resolve_peers_accumulate whose blocks first wait (on the device) until all n ranks have
 *                                 signalled `epoch` in my_flags, and whose last block stores `epoch` into done[my_rank] of
 *                                 the ROOT's block (root_flags, may be null);
 *        ptb_peer_wait(kind 1)    on the root: device-side wait until done[r] >= epoch for all r: the frame is complete.
 *      With double-buffered accumulators this is the only ordering a step needs.  A wait that sees no
 *      signal for 4 s gives up and sets the block's error word (ptb_peer_flags_error) instead of hanging the GPU. */
int ptb_peer_flags_create(ptb_context* ctx, uint32_t** flags);   /* release with ptb_device_free */
int ptb_peer_signal(ptb_context* ctx, uint32_t* const* flag_blocks, int n_ranks, int my_rank, int kind, uint32_t epoch, void* stream);
int ptb_peer_wait(ptb_context* ctx, uint32_t* my_flags, int kind, int n_ranks, uint32_t epoch, void* stream);
int ptb_peer_flags_error(ptb_context* ctx, const uint32_t* flags, void* stream, int* timed_out);
int ptb_resolve_peers_sync(ptb_context* ctx, const ptb_float4* const* accums, int n_ranks, int my_rank, uint32_t* my_flags,
                           uint32_t* root_flags, uint32_t epoch, const ptb_float4* prev_accum, float prev_weight,
                           ptb_float4* accum_out, ptb_uchar4* frame, uint32_t first_pixel, uint32_t n_pixels, float scale,
                           const ptb_render_cfg* cfg, void* stream);
/* CUDA IPC plumbing for the above (one process per GPU): export a cudaMalloc'ed buffer, open a peer's, close it */
int ptb_ipc_export(ptb_context* ctx, const void* device_ptr, unsigned char handle[64]);
int ptb_ipc_open(ptb_context* ctx, const unsigned char handle[64], void** device_ptr);
int ptb_ipc_close(ptb_context* ctx, void* device_ptr);
/* batch closest-hit query on the built BVH (device arrays of float3 / outputs) */
int ptb_trace_rays(ptb_context* ctx, unsigned long long handle, const float* d_origins, const float* d_dirs, uint32_t n,
                   float tmin, float tmax, int32_t* d_prim, float* d_t, float* d_b1, float* d_b2, void* stream);

/* ---- multi-GPU context: the n-device counterpart of optixDeviceContextCreate + the render loop's optixLaunch
 *      (optixSphere.cpp:798-812, 1390-1437; SURVEY.md section 8b "multi-GPU is internal", 8e).  ONE host process drives all
 *      devices; the scene is replicated (ptb_multi_accel_build uploads it and builds a BVH on every device); one
 *      ptb_multi_launch() renders the cfg->subframes_per_launch subframes that start at params->subframe_index:
 *        PTB_SPLIT_SAMPLES  contiguous blocks of subframes per device (global subframe indices seed the RNG,
 *                           optixSphere.cu:316) into per-device sum accumulators, then one fused reduce-scatter ->
 *                           accumulate -> tonemap -> gather kernel per device over peer memory (NVLink).  The frame's mean
 *                           is continued as (old mean * subframe_index + sum of the new launch means) / (subframe_index + K):
 *                           equal to the reference's running average up to rounding (sum order), not bitwise.
 *        PTB_SPLIT_TILES    16-row strips dealt round-robin; every device writes its strips straight into the root's
 *                           buffers through peer pointers: bit-identical to a single-GPU launch.
 *      params->accum_buffer / frame_buffer live on the ROOT device (index 0: allocate them through ptb_multi_context(m, 0));
 *      params->handle is ignored.  The call is asynchronous; ptb_multi_synchronize() waits for every device.  The same
 *      device may be listed more than once (several contexts on one GPU: how the single-GPU tests exercise this path). */
typedef struct ptb_multi ptb_multi;
#define PTB_SPLIT_SAMPLES 0
#define PTB_SPLIT_TILES 1
int ptb_multi_create(const int* devices, int n_devices, ptb_multi** out);
void ptb_multi_destroy(ptb_multi* m);
int ptb_multi_device_count(const ptb_multi* m);
ptb_context* ptb_multi_context(ptb_multi* m, int index);
void* ptb_multi_stream(ptb_multi* m, int index);   /* cudaStream_t the launches of device `index` are issued on */
int ptb_multi_accel_build(ptb_multi* m, ptb_scene* scene, const ptb_build_cfg* cfg, ptb_build_stats* stats);
int ptb_multi_launch(ptb_multi* m, const ptb_Params* params, const ptb_render_cfg* cfg, int split);
int ptb_multi_synchronize(ptb_multi* m);
/* segments, hits, misses, launches summed over the devices (ptb_context_get_totals of every context) */
int ptb_multi_get_totals(ptb_multi* m, uint64_t out[4], int reset);

/* ---- output buffer: sutil::CUDAOutputBuffer<uchar4> (optixSphere.cpp:1284,
 *      1376-1382, 1401, 1419, 1484-1486, 256) ----------------------------------- */
int ptb_output_create(ptb_context* ctx, uint32_t width, uint32_t height, ptb_output** out);
int ptb_output_resize(ptb_output* ob, uint32_t width, uint32_t height);
ptb_uchar4* ptb_output_map(ptb_output* ob);       /* device pointer */
void ptb_output_unmap(ptb_output* ob, void* stream);
const ptb_uchar4* ptb_output_host_ptr(ptb_output* ob); /* device->host copy, then host pointer */
uint32_t ptb_output_width(const ptb_output* ob);
uint32_t ptb_output_height(const ptb_output* ob);
void ptb_output_destroy(ptb_output* ob);

/* ---- device memory helpers (the reference calls cudaMalloc/cudaMemcpy directly,
 *      e.g. accum buffer optixSphere.cpp:1298-1301) ------------------------------ */
int ptb_device_alloc(ptb_context* ctx, size_t bytes, void** out);
int ptb_device_free(ptb_context* ctx, void* p);
int ptb_device_memset(ptb_context* ctx, void* p, int value, size_t bytes, void* stream);
int ptb_copy_to_device(ptb_context* ctx, void* dst, const void* src, size_t bytes, void* stream);
int ptb_copy_to_host(ptb_context* ctx, void* dst, const void* src, size_t bytes, void* stream);   /* returns when the copy has finished */
/* the same without waiting: dst must be page-locked host memory; the caller orders it (stream / event / ptb_context_synchronize) */
int ptb_copy_to_host_async(ptb_context* ctx, void* dst, const void* src, size_t bytes, void* stream);

/* ---- image files: sutil::loadImage / sutil::saveImage (optixSphere.cpp:359, 836,
 *      1483-1489).  PNG (8/16-bit gray, gray+alpha, RGB, RGBA, palette) -> RGBA8 with
 *      stb_image STBI_rgb_alpha semantics; EXR scanline NONE/ZIPS/ZIP, HALF/FLOAT
 *      -> float4 (missing A = 1).  Buffers are released with ptb_free(). */
int ptb_image_load_rgba8(const char* path, uint8_t** pixels, int* w, int* h);
int ptb_image_load_float4(const char* path, float** pixels, int* w, int* h);
/* .png or .ppm by extension; rows are written bottom-up (row 0 of the frame buffer is
 * the bottom image row, optixSphere.cu:332,400) when flip_y != 0 */
int ptb_save_image(const char* path, const ptb_uchar4* pixels, int w, int h, int flip_y);
/* the float4 accumulation buffer (HOST copy of Params.accum_buffer, optixSphere.cpp:1298-1301) as a raw file for parity
 * tools: 16-byte header {"PTBA", 1, w, h} + w*h float4, row 0 = bottom image row.  ptb_load_accum_raw() reads it back
 * (release with ptb_free()). */
int ptb_save_accum_raw(const char* path, const ptb_float4* accum, int w, int h);
int ptb_load_accum_raw(const char* path, float** accum, int* w, int* h);
void ptb_free(void* p);
/* The OBJ reader on its own (what tinyobj::LoadObj(triangulate=true) hands to the reference,
 * optixSphere.cpp:431, 447-515): per face vertex, in file/fan order, 10 words =
 * vx vy vz nx ny nz tx ty (float32, raw: unscaled, normals not normalised) + has_normal, has_texcoord (int32).
 * *records is released with ptb_free(). */
int ptb_obj_read(const char* path, void** records, uint64_t* n_face_vertices);

/* ---- roofline denominators measured on the device itself (the L2 peak is not in
 *      MEASURED_PEAKS.json).  Streams `bytes` of device memory through every SM with 16-byte loads, `iters` times,
 *      and returns the read bandwidth in GB/s: bytes <= 32 MiB stays L2-resident (L2 -> SM peak), bytes >> 126 MiB
 *      measures HBM reads. */
int ptb_microbench_read(ptb_context* ctx, size_t bytes, int iters, double* gb_per_s);

/* ---- device self-test hooks used by the parity tests --------------------------- */
/* op: 0 rng (in: seed as uint bits -> out: next seed bits, u), 1 sincos (x -> s, c),
 *     2 atan2 (y, x -> r), 3 asin (x -> r), 4 texel widening (b -> b/255 two ways), 5 pow (x, y -> r).
 *     in/out are HOST arrays. */
/* n samples of the scene's environment CDF: xi = n x 2 uniforms (HOST) -> out = n x 4 (direction xyz, solid-angle pdf) */
int ptb_test_env_sample(ptb_context* ctx, unsigned long long handle, const float* xi, uint32_t n, float* out);
int ptb_test_device_math(ptb_context* ctx, int op, const float* in, int in_stride, float* out, int out_stride, uint32_t n);

#ifdef __cplusplus
}
#endif
#endif /* PTB_H */
