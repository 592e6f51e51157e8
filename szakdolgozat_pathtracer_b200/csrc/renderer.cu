// renderer.cu -- device half of the C ABI (include/ptb.h): context, scene upload,
// BVH build entry, the wavefront launch loop, output buffers.
//
// One ptb_launch() = one optixLaunch of the reference (optixSphere.cpp:1403-1418):
// every pixel renders spp_per_launch samples, the result is folded into
// Params.accum_buffer and tonemapped into Params.frame_buffer.  Internally
// (ptb_render_cfg.pipeline):
//
//   3 (default)  k_chunk_raygen, k_chunk_fused (all wavefront iterations in one kernel), k_resolve
//   4            k_pool_fused (persistent path pool, camera rays included), k_resolve
//   2            k_chunk_raygen, spp*(max_depth+1) x { k_chunk_trace, k_chunk_shade, k_chunk_miss }, k_resolve
//   1            k_raygen_init, spp*(max_depth+1) x { k_trace, k_shade, k_miss } over global queues, k_resolve
//
// All kernels of a launch go to the caller's stream without any host
// synchronisation; counters live in device memory and are folded into the
// context's running totals by k_fold_totals at the end of the launch.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "bvh_build.h"
#include "fast_api.h"
#include "host.h"
#include "linear.cuh"
#include "pool.cuh"

namespace ptb {

struct DeviceScene {
    int device = 0;
    uint64_t revision = 0;
    uint32_t n_tris = 0;
    float4* verts = nullptr; float4* normals = nullptr; float2* uvs = nullptr; uint32_t* mat_ids = nullptr;
    DevMaterial* mats = nullptr; int n_mats = 0;
    std::vector<void*> textures;
    float4* env = nullptr; int env_w = 0, env_h = 0;
    float *cdf_marginal = nullptr, *cdf_conditional = nullptr, *cdf_row_weight = nullptr; float cdf_total = 0.0f;
    DeviceBvh bvh;
    unsigned long long handle = 0;
    ptb_context* owner = nullptr;
    const ptb_scene* source = nullptr;  // the host scene this upload was made from (it owns this object)
};

void free_device_scene_buffers(DeviceScene* d) {
    cudaFree(d->verts); cudaFree(d->normals); cudaFree(d->uvs); cudaFree(d->mat_ids); cudaFree(d->mats); cudaFree(d->env);
    cudaFree(d->cdf_marginal); cudaFree(d->cdf_conditional); cudaFree(d->cdf_row_weight);
    d->cdf_marginal = d->cdf_conditional = d->cdf_row_weight = nullptr;
    for (void* t : d->textures) cudaFree(t);
    d->textures.clear();
    d->verts = d->normals = nullptr; d->uvs = nullptr; d->mat_ids = nullptr; d->mats = nullptr; d->env = nullptr;
    free_bvh(d->bvh);
}

}  // namespace ptb

using namespace ptb;

struct ptb_context {
    int device = 0;
    int num_sms = 148;
    // path pool + queues, sized for the largest frame seen
    uint32_t pool_slots = 0;
    float4 *ray_o = nullptr, *ray_d = nullptr, *hit = nullptr, *atten_seed = nullptr, *pixsum = nullptr;
    uint4* misc = nullptr;
    uint32_t *q_trace[2] = {nullptr, nullptr}, *q_hit = nullptr, *q_miss = nullptr;
    uint32_t* counters = nullptr; uint32_t counters_cap = 0;  // in iterations
    unsigned long long* trav_stats = nullptr;
    unsigned long long* totals = nullptr;  // segments, hits, misses, launches since the last reset
    unsigned long long* launch_totals = nullptr;  // the same for the launch in flight ([3]: iterations of the fused pipeline)
    unsigned char* status = nullptr;       // one byte per slot (chunked pipelines)
    float4* out_pixsum = nullptr; uint32_t out_pixsum_slots = 0;  // per-slot sample sums of the pool pipeline
    // linear estimator (env_importance_sampling != 0): pending shadow rays, allocated on first use
    float4 *shadow_o = nullptr, *shadow_d = nullptr, *shadow_c = nullptr; unsigned char* shadow_flag = nullptr; uint32_t shadow_slots = 0;
    int last_pipeline = 0;
    bool keep_launch_totals = false;  // set by ptb_launch while it renders the later batches of a bounded-pool launch
    bool defer_fold = false;          // ... and while a later batch will fold the launch counters into the running totals
    float stage_sum[6] = {0, 0, 0, 0, 0, 0}; bool stage_sum_valid = false;  // stage times of the last batched, profiled launch
    // last launch, for ptb_launch_get_stats
    cudaStream_t last_stream = nullptr;
    uint32_t last_iters = 0, last_kernels = 0;
    uint64_t last_paths = 0;
    bool last_counted = false;
    // stage profiling (profile_stages): events[0] start, then 3 per iteration, then after resolve
    std::vector<cudaEvent_t> events;
    uint32_t prof_iters = 0;  // iterations of the last profiled launch, 0 = none
    std::map<unsigned long long, DeviceScene*> scenes;
    unsigned long long next_handle = 1;
    // launch overlap (ptb_render_cfg.overlap_lanes): lane 0's path pool is the members above, lanes 1.. keep theirs here and
    // are swapped in for the duration of a ptb_launch call (LaneGuard); every lane has its render stream and two events
    struct LanePool {
        uint32_t pool_slots = 0;
        float4 *ray_o = nullptr, *ray_d = nullptr, *hit = nullptr, *atten_seed = nullptr, *pixsum = nullptr;
        uint4* misc = nullptr;
        uint32_t *q_trace[2] = {nullptr, nullptr}, *q_hit = nullptr, *q_miss = nullptr;
        uint32_t* counters = nullptr; uint32_t counters_cap = 0;
        unsigned long long *trav_stats = nullptr, *launch_totals = nullptr;
        unsigned char* status = nullptr;
    };
    struct LaneSync { cudaStream_t stream = nullptr; cudaEvent_t render_done = nullptr, resolve_done = nullptr; bool used = false; };
    LanePool extra_lanes[3];
    LaneSync lane_sync[4];
    int next_lane = 0, last_lane = 0;
};

namespace {
// exchanges the path pool of lane `lane` (> 0) with the context's own members; a second call swaps back
void swap_lane(ptb_context* c, int lane) {
    if (lane <= 0) return;
    ptb_context::LanePool& l = c->extra_lanes[lane - 1];
    std::swap(c->pool_slots, l.pool_slots);
    std::swap(c->ray_o, l.ray_o); std::swap(c->ray_d, l.ray_d); std::swap(c->hit, l.hit); std::swap(c->atten_seed, l.atten_seed);
    std::swap(c->pixsum, l.pixsum); std::swap(c->misc, l.misc);
    std::swap(c->q_trace[0], l.q_trace[0]); std::swap(c->q_trace[1], l.q_trace[1]); std::swap(c->q_hit, l.q_hit); std::swap(c->q_miss, l.q_miss);
    std::swap(c->counters, l.counters); std::swap(c->counters_cap, l.counters_cap);
    std::swap(c->trav_stats, l.trav_stats); std::swap(c->launch_totals, l.launch_totals); std::swap(c->status, l.status);
}
struct LaneGuard {
    ptb_context* c; int lane;
    LaneGuard(ptb_context* c_, int lane_) : c(c_), lane(lane_) { swap_lane(c, lane); }
    ~LaneGuard() { swap_lane(c, lane); }
};
}  // namespace

struct ptb_output {
    ptb_context* ctx = nullptr;
    uint32_t w = 0, h = 0;
    uchar4* d_pixels = nullptr;
    std::vector<ptb_uchar4> host;
};

namespace ptb {
void free_device_scene(DeviceScene* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    if (d->owner) d->owner->scenes.erase(d->handle);
    free_device_scene_buffers(d);
    delete d;
}
}  // namespace ptb

namespace {

#define PTB_SLOT_BYTES 97ull  // path state per slot: six 16-byte records + one status byte

int fail(int code, const std::string& msg) { set_error(msg); return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(PTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)

template <typename T>
cudaError_t upload(T** dst, const void* src, size_t bytes, cudaStream_t st) {
    cudaError_t e = cudaMalloc((void**)dst, bytes ? bytes : 16);
    if (e != cudaSuccess) return e;
    if (bytes) e = cudaMemcpyAsync(*dst, src, bytes, cudaMemcpyHostToDevice, st);
    return e;
}

void free_pool(ptb_context* c) {
    cudaFree(c->ray_o); cudaFree(c->ray_d); cudaFree(c->hit); cudaFree(c->atten_seed); cudaFree(c->pixsum); cudaFree(c->misc);
    cudaFree(c->q_trace[0]); cudaFree(c->q_trace[1]); cudaFree(c->q_hit); cudaFree(c->q_miss); cudaFree(c->status);
    cudaFree(c->shadow_o); cudaFree(c->shadow_d); cudaFree(c->shadow_c); cudaFree(c->shadow_flag);
    c->shadow_o = c->shadow_d = c->shadow_c = nullptr; c->shadow_flag = nullptr; c->shadow_slots = 0;
    cudaFree(c->out_pixsum); c->out_pixsum = nullptr; c->out_pixsum_slots = 0;
    c->status = nullptr;
    c->ray_o = c->ray_d = c->hit = c->atten_seed = c->pixsum = nullptr; c->misc = nullptr;
    c->q_trace[0] = c->q_trace[1] = c->q_hit = c->q_miss = nullptr;
    c->pool_slots = 0;
}

int ensure_pool(ptb_context* c, uint32_t slots, uint32_t iters) {
    if (slots > c->pool_slots) {
        free_pool(c);
        const size_t n = slots;
        CU(cudaMalloc((void**)&c->ray_o, n * 16)); CU(cudaMalloc((void**)&c->ray_d, n * 16));
        CU(cudaMalloc((void**)&c->hit, n * 16)); CU(cudaMalloc((void**)&c->atten_seed, n * 16));
        CU(cudaMalloc((void**)&c->pixsum, n * 16)); CU(cudaMalloc((void**)&c->misc, n * 16));
        CU(cudaMalloc((void**)&c->q_trace[0], n * 4)); CU(cudaMalloc((void**)&c->q_trace[1], n * 4));
        CU(cudaMalloc((void**)&c->q_hit, n * 4)); CU(cudaMalloc((void**)&c->q_miss, n * 4));
        CU(cudaMalloc((void**)&c->status, n + 64));
        c->pool_slots = slots;
    }
    if (iters + 2 > c->counters_cap) {
        cudaFree(c->counters); c->counters = nullptr; c->counters_cap = 0;
        CU(cudaMalloc((void**)&c->counters, (size_t)(iters + 2) * 4 * sizeof(uint32_t)));
        c->counters_cap = iters + 2;
    }
    if (!c->trav_stats) CU(cudaMalloc((void**)&c->trav_stats, 2 * sizeof(unsigned long long)));
    if (!c->totals) {
        CU(cudaMalloc((void**)&c->totals, 4 * sizeof(unsigned long long)));
        CU(cudaMemset(c->totals, 0, 4 * sizeof(unsigned long long)));
    }
    if (!c->launch_totals) CU(cudaMalloc((void**)&c->launch_totals, 4 * sizeof(unsigned long long)));
    return PTB_OK;
}

SceneView scene_view(const DeviceScene* d) {
    SceneView s;
    s.nodes = d->bvh.nodes; s.nodes4 = d->bvh.nodes4; s.tris = d->bvh.tris;
    s.nodes8 = d->bvh.nodes8; s.tris8 = d->bvh.tris8;
    s.verts = d->verts; s.normals = d->normals; s.uvs = d->uvs; s.mat_ids = d->mat_ids; s.mats = d->mats;
    s.env = d->env; s.env_w = d->env_w; s.env_h = d->env_h;
    return s;
}

DeviceScene* find_scene(ptb_context* ctx, unsigned long long handle) {
    auto it = ctx->scenes.find(handle);
    return it == ctx->scenes.end() ? nullptr : it->second;
}

}  // namespace

extern "C" {

int ptb_context_create(int device, ptb_context** out) {
    if (!out) return fail(PTB_ERR_INVALID, "ptb_context_create: null out");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(PTB_ERR_NO_DEVICE, std::string("no CUDA device is usable (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count 0") +
                                           "); ptb has no CPU path");
    if (device < 0 || device >= count) return fail(PTB_ERR_INVALID, "ptb_context_create: device index out of range");
    CU(cudaSetDevice(device));
    CU(cudaFree(0));  // the reference initialises the runtime the same way (optixSphere.cpp:801)
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    ptb_context* c = new ptb_context();
    c->device = device; c->num_sms = prop.multiProcessorCount;
    *out = c;
    return PTB_OK;
}

void ptb_context_destroy(ptb_context* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    // scenes stay owned by their ptb_scene; just detach them
    for (auto& kv : ctx->scenes) kv.second->owner = nullptr;
    for (int lane = 1; lane < 4; ++lane) {   // the extra lanes' pools
        swap_lane(ctx, lane);
        free_pool(ctx);
        cudaFree(ctx->counters); cudaFree(ctx->trav_stats); cudaFree(ctx->launch_totals);
        ctx->counters = nullptr; ctx->counters_cap = 0; ctx->trav_stats = nullptr; ctx->launch_totals = nullptr;
        swap_lane(ctx, lane);
    }
    for (auto& l : ctx->lane_sync) {
        if (l.stream) cudaStreamDestroy(l.stream);
        if (l.render_done) cudaEventDestroy(l.render_done);
        if (l.resolve_done) cudaEventDestroy(l.resolve_done);
    }
    free_pool(ctx);
    cudaFree(ctx->counters); cudaFree(ctx->trav_stats); cudaFree(ctx->totals); cudaFree(ctx->launch_totals);
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    delete ctx;
}

int ptb_context_synchronize(ptb_context* ctx, void* stream) {
    if (!ctx) return fail(PTB_ERR_INVALID, "ptb_context_synchronize: null context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return PTB_OK;
}

int ptb_accel_build(ptb_context* ctx, ptb_scene* scene, const ptb_build_cfg* cfg_in, void* stream_, unsigned long long* handle,
                    ptb_build_stats* stats_out) {
    if (!ctx || !scene || !handle) return fail(PTB_ERR_INVALID, "ptb_accel_build: bad arguments");
    if (scene->env.empty()) return fail(PTB_ERR_INVALID, "ptb_accel_build: the scene has no environment map (ptb_scene_set_env_*)");
    if (scene->mats.empty()) return fail(PTB_ERR_INVALID, "ptb_accel_build: the scene has no materials");
    cudaStream_t st = (cudaStream_t)stream_;
    CU(cudaSetDevice(ctx->device));
    ptb_build_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else ptb_default_build_cfg(&cfg);

    // a scene can be built on several contexts (multi-GPU replicas); a rebuild replaces THIS context's upload only
    for (size_t i = 0; i < scene->devs.size();) {
        if (scene->devs[i]->owner == ctx || scene->devs[i]->owner == nullptr) { free_device_scene(scene->devs[i]); scene->devs.erase(scene->devs.begin() + i); }
        else ++i;
    }
    CU(cudaSetDevice(ctx->device));  // freeing another device's stale upload switched the current device
    DeviceScene* d = new DeviceScene();
    d->device = ctx->device; d->revision = scene->revision; d->source = scene;
    const uint32_t n = (uint32_t)scene->tris.size();
    d->n_tris = n;

    // Flatten to the reference's vertex/normal/texcoord arrays, 3 entries per triangle (optixSphere.cpp:845-858).
    std::vector<ptb_float4> verts((size_t)n * 3), normals((size_t)n * 3);
    std::vector<ptb_float2> uvs((size_t)n * 3);
    {
        auto flatten = [&](uint32_t lo, uint32_t hi) {
            for (uint32_t i = lo; i < hi; ++i) {
                const ptb_TriangleData& t = scene->tris[i];
                verts[(size_t)i * 3] = t.v0; verts[(size_t)i * 3 + 1] = t.v1; verts[(size_t)i * 3 + 2] = t.v2;
                normals[(size_t)i * 3] = t.n0; normals[(size_t)i * 3 + 1] = t.n1; normals[(size_t)i * 3 + 2] = t.n2;
                uvs[(size_t)i * 3] = t.uv0; uvs[(size_t)i * 3 + 1] = t.uv1; uvs[(size_t)i * 3 + 2] = t.uv2;
            }
        };
        unsigned nt = n >= (1u << 18) ? std::thread::hardware_concurrency() : 1u;  // a 1.2 GB shuffle at 4.6 M triangles
        if (nt == 0) nt = 1;
        if (nt > 16) nt = 16;
        std::vector<std::thread> pool;
        const uint32_t per = (n + nt - 1) / nt;
        for (unsigned k = 1; k < nt; ++k) pool.emplace_back(flatten, std::min(n, k * per), std::min(n, (k + 1) * per));
        flatten(0, std::min(n, per));
        for (std::thread& th : pool) th.join();
    }
    cudaError_t e = cudaSuccess;
    auto ok = [&](cudaError_t r) { if (e == cudaSuccess) e = r; return r == cudaSuccess; };
    ok(upload(&d->verts, verts.data(), verts.size() * sizeof(ptb_float4), st));
    ok(upload(&d->normals, normals.data(), normals.size() * sizeof(ptb_float4), st));
    ok(upload(&d->uvs, uvs.data(), uvs.size() * sizeof(ptb_float2), st));
    ok(upload(&d->mat_ids, scene->mat_ids.data(), scene->mat_ids.size() * sizeof(uint32_t), st));
    ok(upload(&d->env, scene->env.data(), scene->env.size() * sizeof(float), st));
    d->env_w = scene->env_w; d->env_h = scene->env_h;

    // hit-group table (optixSphere.cpp:1196-1261); every material owns its textures
    std::vector<DevMaterial> dm(scene->mats.size());
    for (size_t i = 0; i < scene->mats.size() && e == cudaSuccess; ++i) {
        const Material& m = scene->mats[i];
        DevMaterial& o = dm[i];
        memset(&o, 0, sizeof(o));
        for (int c = 0; c < 3; ++c) { o.emission[c] = m.emission_color[c]; o.diffuse[c] = m.diffuse_color[c]; o.specular[c] = m.specular[c]; }
        o.roughness = m.roughness; o.metallic = m.metallic ? 1 : 0; o.transparent = m.transparent ? 1 : 0;
        for (int k = 0; k < TEX_COUNT; ++k) {
            const Texture& t = m.tex[k];
            if (!t.has) continue;
            void* p = nullptr;
            const size_t bytes = t.is_float ? t.rgba32f.size() * sizeof(float) : t.rgba8.size();
            if (!ok(upload((char**)&p, t.is_float ? (const void*)t.rgba32f.data() : (const void*)t.rgba8.data(), bytes, st))) break;
            d->textures.push_back(p);
            o.tex[k].data = p; o.tex[k].w = t.w; o.tex[k].h = t.h; o.tex[k].fmt = t.is_float ? 2 : 1;
        }
    }
    ok(upload(&d->mats, dm.data(), dm.size() * sizeof(DevMaterial), st));
    d->n_mats = (int)dm.size();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);  // host staging vectors go out of scope below
    if (e != cudaSuccess) { free_device_scene_buffers(d); delete d; return fail(PTB_ERR_CUDA, std::string("scene upload: ") + cudaGetErrorString(e)); }

    // environment CDF for the optional light-sampling mode (env_cdf.cuh): luminance * sin(theta), with a floor of
    // 1 % of the mean luminance so that no direction with non-zero radiance has zero density
    {
        const int ew = d->env_w, eh = d->env_h;
        double mean = 0.0;
        for (size_t i = 0; i < scene->env.size(); i += 4) mean += 0.2126 * scene->env[i] + 0.7152 * scene->env[i + 1] + 0.0722 * scene->env[i + 2];
        mean /= (double)ew * eh;
        cudaError_t ce = cudaMalloc((void**)&d->cdf_marginal, (size_t)(eh + 1) * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMalloc((void**)&d->cdf_conditional, (size_t)eh * (ew + 1) * sizeof(float));
        if (ce == cudaSuccess) ce = cudaMalloc((void**)&d->cdf_row_weight, (size_t)(eh + 1) * sizeof(float));
        if (ce == cudaSuccess) {
            k_env_row_cdf<<<eh, 256, 0, st>>>(d->env, ew, eh, (float)(0.01 * mean) + 1e-12f, d->cdf_conditional, d->cdf_row_weight);
            k_env_marginal<<<1, 256, 0, st>>>(eh, d->cdf_row_weight, d->cdf_marginal, d->cdf_row_weight + eh);
            k_env_normalize_rows<<<eh, 256, 0, st>>>(ew, d->cdf_conditional, d->cdf_row_weight);
            ce = cudaMemcpyAsync(&d->cdf_total, d->cdf_row_weight + eh, sizeof(float), cudaMemcpyDeviceToHost, st);
            if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
        }
        if (ce != cudaSuccess) { free_device_scene_buffers(d); delete d; return fail(PTB_ERR_CUDA, std::string("environment CDF: ") + cudaGetErrorString(ce)); }
    }

    ptb_build_stats stats;
    std::string err;
    bool built = build_bvh(d->verts, n, cfg, st, d->bvh, stats, err);
    if (!built && cfg.sah_refine) {
        // an SAH treelet can in principle grow deeper than the traversal stack: fall back to the plain LBVH (depth <= 62)
        if (getenv("PTB_VERBOSE")) fprintf(stderr, "ptb_accel_build: refined build failed (%s), falling back to the plain LBVH\n", err.c_str());
        free_bvh(d->bvh);
        cfg.sah_refine = 0;
        built = build_bvh(d->verts, n, cfg, st, d->bvh, stats, err);
    }
    if (!built) { free_device_scene_buffers(d); delete d; return fail(PTB_ERR_CUDA, err); }
    d->handle = ctx->next_handle++;
    d->owner = ctx;
    ctx->scenes[d->handle] = d;
    scene->devs.push_back(d);
    *handle = d->handle;
    if (stats_out) *stats_out = stats;
    return PTB_OK;
}

int ptb_accel_read(ptb_context* ctx, unsigned long long handle, float* nodes, uint32_t cap_nodes, float* tris, uint32_t cap_tris,
                   uint32_t* n_nodes, uint32_t* n_tris) {
    if (!ctx) return fail(PTB_ERR_INVALID, "ptb_accel_read: null context");
    DeviceScene* d = find_scene(ctx, handle);
    if (!d) return fail(PTB_ERR_INVALID, "ptb_accel_read: unknown handle");
    CU(cudaSetDevice(ctx->device));
    if (n_nodes) *n_nodes = d->bvh.n_nodes;
    if (n_tris) *n_tris = d->bvh.n_tris;
    if (nodes) {
        if (cap_nodes < d->bvh.n_nodes) return fail(PTB_ERR_INVALID, "ptb_accel_read: node buffer too small");
        CU(cudaMemcpy(nodes, d->bvh.nodes, (size_t)d->bvh.n_nodes * 64, cudaMemcpyDeviceToHost));
    }
    if (tris) {
        if (cap_tris < d->bvh.n_tris) return fail(PTB_ERR_INVALID, "ptb_accel_read: triangle buffer too small");
        CU(cudaMemcpy(tris, d->bvh.tris, (size_t)d->bvh.n_tris * 48, cudaMemcpyDeviceToHost));
    }
    return PTB_OK;
}

int ptb_accel_read8(ptb_context* ctx, unsigned long long handle, uint32_t* nodes8, uint32_t cap_nodes8, float* tris8, uint32_t cap_tris,
                    uint32_t* n_nodes8, uint32_t* n_tris) {
    if (!ctx) return fail(PTB_ERR_INVALID, "ptb_accel_read8: null context");
    DeviceScene* d = find_scene(ctx, handle);
    if (!d) return fail(PTB_ERR_INVALID, "ptb_accel_read8: unknown handle");
    CU(cudaSetDevice(ctx->device));
    const uint32_t nn = d->bvh.nodes8 ? d->bvh.n_nodes8 : 0u, nt = d->bvh.nodes8 ? d->bvh.n_tris : 0u;
    if (n_nodes8) *n_nodes8 = nn;
    if (n_tris) *n_tris = nt;
    if (nodes8 && nn) {
        if (cap_nodes8 < nn) return fail(PTB_ERR_INVALID, "ptb_accel_read8: node buffer too small");
        CU(cudaMemcpy(nodes8, d->bvh.nodes8, (size_t)nn * 80, cudaMemcpyDeviceToHost));
    }
    if (tris8 && nt) {
        if (cap_tris < nt) return fail(PTB_ERR_INVALID, "ptb_accel_read8: triangle buffer too small");
        CU(cudaMemcpy(tris8, d->bvh.tris8, (size_t)nt * 48, cudaMemcpyDeviceToHost));
    }
    return PTB_OK;
}

int ptb_launch(ptb_context* ctx, const ptb_Params* P, const ptb_render_cfg* cfg_in, void* stream_) {
    if (!ctx || !P) return fail(PTB_ERR_INVALID, "ptb_launch: bad arguments");
    ptb_render_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else ptb_default_render_cfg(&cfg);
    if (P->image_width == 0 || P->image_height == 0) return fail(PTB_ERR_INVALID, "ptb_launch: empty image");
    if ((uint64_t)P->image_width * P->image_height > 0x7fffffffull) return fail(PTB_ERR_INVALID, "ptb_launch: image too large");
    if (!P->accum_buffer) return fail(PTB_ERR_INVALID, "ptb_launch: Params.accum_buffer is null");
    if (cfg.write_frame && !P->frame_buffer) return fail(PTB_ERR_INVALID, "ptb_launch: Params.frame_buffer is null (set write_frame = 0 to skip tonemapping)");
    if (cfg.spp_per_launch < 1 || cfg.max_depth < 0 || cfg.max_depth > 1000) return fail(PTB_ERR_INVALID, "ptb_launch: bad spp_per_launch / max_depth");
    if (cfg.env_importance_sampling < 0 || cfg.env_importance_sampling > 2) return fail(PTB_ERR_INVALID, "ptb_launch: env_importance_sampling must be 0, 1 or 2");
    DeviceScene* d = find_scene(ctx, P->handle);
    if (!d) return fail(PTB_ERR_INVALID, "ptb_launch: Params.handle does not name a built acceleration structure");
    if (d->source && d->source->revision != d->revision)
        return fail(PTB_ERR_INVALID, "ptb_launch: the scene was modified (materials, environment or geometry) after ptb_accel_build; build it again");
    cudaStream_t st = (cudaStream_t)stream_;
    CU(cudaSetDevice(ctx->device));

    const int n_sub = cfg.subframes_per_launch < 1 ? 1 : cfg.subframes_per_launch;
    uint32_t row0 = 0, rows = P->image_height;
    if (cfg.row_begin != 0 || cfg.row_end != 0) {
        if (cfg.row_begin < 0 || cfg.row_end <= cfg.row_begin || (uint32_t)cfg.row_end > P->image_height)
            return fail(PTB_ERR_INVALID, "ptb_launch: bad row band");
        row0 = (uint32_t)cfg.row_begin; rows = (uint32_t)(cfg.row_end - cfg.row_begin);
    }
    uint32_t il_n = 0, il_r = 0, il_h = 0;
    if (cfg.row_interleave_count > 1) {
        if (cfg.row_begin != 0 || cfg.row_end != 0 || cfg.row_interleave_index < 0 || cfg.row_interleave_index >= cfg.row_interleave_count ||
            cfg.row_interleave_height < 1)
            return fail(PTB_ERR_INVALID, "ptb_launch: bad row interleave");
        il_n = (uint32_t)cfg.row_interleave_count; il_r = (uint32_t)cfg.row_interleave_index; il_h = (uint32_t)cfg.row_interleave_height;
        const uint32_t strips = (P->image_height + il_h - 1u) / il_h;                 // strips in the frame
        const uint32_t mine = strips > il_r ? (strips - il_r + il_n - 1u) / il_n : 0u;  // strips il_r, il_r + il_n, ...
        rows = mine * il_h;                                                             // the last one may be padded
        if (rows == 0) return PTB_OK;  // more ranks than strips: nothing to render
    }
    const uint32_t n_pixels = P->image_width * rows;
    // The path pool is bounded: a launch whose n_sub * pixels slots would need more than max_pool_bytes of path state
    // (97 B per slot) is rendered as consecutive batches of subframes, each one wavefront -- bit-identical to the single
    // wavefront (and to n_sub separate launches); counters of the batches add up.
    {
        const uint64_t cap_bytes = cfg.max_pool_bytes > 0 ? (uint64_t)cfg.max_pool_bytes : (2ull << 30);
        const uint64_t px = n_pixels;   // pixels this launch renders (a band or a set of strips renders fewer than W * H)
        uint64_t max_sub = cap_bytes / (px * PTB_SLOT_BYTES);
        if (max_sub < 1) max_sub = 1;
        if ((uint64_t)n_sub > max_sub && cfg.pipeline != PTB_PIPELINE_POOL_FUSED) {
            uint32_t kernels = 0; uint64_t paths = 0;
            float stage_sum[6] = {0, 0, 0, 0, 0, 0};
            for (int first = 0; first < n_sub; first += (int)max_sub) {
                ptb_Params p2 = *P; p2.subframe_index = P->subframe_index + first;
                ptb_render_cfg c2 = cfg;
                c2.subframes_per_launch = n_sub - first < (int)max_sub ? n_sub - first : (int)max_sub;
                if (first > 0) c2.aux_primary_hit = nullptr;   // primary hits are those of the launch's first subframe
                ctx->keep_launch_totals = first > 0;                        // later batches add to the first one's launch counters ...
                ctx->defer_fold = first + (int)max_sub < n_sub;             // ... which are folded into the running totals once, by the last
                const int rc = ptb_launch(ctx, &p2, &c2, stream_);
                ctx->keep_launch_totals = false; ctx->defer_fold = false;
                if (rc != PTB_OK) return rc;
                kernels += ctx->last_kernels; paths += ctx->last_paths;
                if (cfg.profile_stages) {  // stage times of a batched launch are the sums over its batches (synchronises: profiling only)
                    float ms[6];
                    ctx->stage_sum_valid = false;
                    const int prc = ptb_launch_get_stage_ms(ctx, ms);
                    if (prc != PTB_OK) return prc;
                    for (int k = 0; k < 6; ++k) stage_sum[k] += ms[k];
                }
            }
            ctx->last_kernels = kernels; ctx->last_paths = paths;
            if (cfg.profile_stages) { for (int k = 0; k < 6; ++k) ctx->stage_sum[k] = stage_sum[k]; ctx->stage_sum_valid = true; }
            return PTB_OK;
        }
    }
    if ((uint64_t)n_pixels * (uint64_t)n_sub > 0x7fffffffull) return fail(PTB_ERR_INVALID, "ptb_launch: subframes_per_launch * pixels too large");
    const uint32_t slots = n_pixels * (uint32_t)n_sub;
    const uint32_t iters = (uint32_t)cfg.spp_per_launch * (uint32_t)(cfg.max_depth + 1);

    static const int env_pipe = getenv("PTB_PIPELINE") ? atoi(getenv("PTB_PIPELINE")) : -1;  // experiments only
    int pipeline = cfg.pipeline > 0 ? cfg.pipeline : (env_pipe > 0 ? env_pipe : PTB_PIPELINE_DEFAULT);
    if (pipeline < PTB_PIPELINE_QUEUES || pipeline > PTB_PIPELINE_POOL_FUSED) return fail(PTB_ERR_INVALID, "ptb_launch: unknown pipeline");
    if (cfg.env_importance_sampling) pipeline = PTB_PIPELINE_CHUNK_STAGES;  // the linear modes run as chunked stage kernels
    if (cfg.arith_mode != PTB_ARITH_EXACT && cfg.arith_mode != PTB_ARITH_FAST) return fail(PTB_ERR_INVALID, "ptb_launch: arith_mode must be 0 (exact) or 1 (fast)");
    const bool fast = cfg.arith_mode == PTB_ARITH_FAST;
    if (fast && (cfg.env_importance_sampling || (pipeline != PTB_PIPELINE_CHUNK_FUSED && pipeline != PTB_PIPELINE_CHUNK_STAGES)))
        return fail(PTB_ERR_UNSUPPORTED, "ptb_launch: arith_mode = 1 (fast) exists for the chunked pipelines (2, 3) of the reference estimator only");
    if (il_n > 1 && pipeline == PTB_PIPELINE_QUEUES) return fail(PTB_ERR_UNSUPPORTED, "ptb_launch: row interleave needs a chunked pipeline (2, 3 or 4)");

    // pool pipeline: persistent blocks own PTB_CHUNK positions each; path state is per position, not per slot
    const uint32_t pool_blocks_needed = (slots + PTB_CHUNK - 1u) / PTB_CHUNK;
    const bool pool_wide = pool_blocks_needed >= (uint32_t)ctx->num_sms * (1024u / PTB_CHUNK_THREADS);
    const uint32_t pool_cap = (uint32_t)ctx->num_sms * (pool_wide ? 1024u / PTB_CHUNK_THREADS : 640u / PTB_CHUNK_THREADS);
    const uint32_t pool_grid = pool_blocks_needed < pool_cap ? pool_blocks_needed : pool_cap;
    const bool use_pool = pipeline == PTB_PIPELINE_POOL_FUSED;

    // Launch overlap (ptb_render_cfg.overlap_lanes): a small launch (fewer than 9 M path slots) renders on one of the context's internal
    // streams with that lane's own path pool, so that its thin tail (every pixel's samples are one sequential chain) runs beside
    // the start of the next launch; accumulate / tonemap (k_resolve) and the counter fold stay on the caller's stream in call
    // order.  Same kernels, same buffers, same results.
    // automatic: four lanes for the reference's own accumulate mode (the viewer loop); the sum mode of the multi-GPU exchanges
    // keeps serial launches unless the caller asks for lanes
    const int n_lanes = cfg.overlap_lanes == 0 ? (cfg.accumulate_mode == 0 ? 4 : 1) : cfg.overlap_lanes;
    if (n_lanes < 1 || n_lanes > 4) return fail(PTB_ERR_INVALID, "ptb_launch: overlap_lanes must be 0 (automatic), 1 (off), 2, 3 or 4");
    bool overlap = n_lanes > 1 && pipeline == PTB_PIPELINE_CHUNK_FUSED && slots < 9000000u && !cfg.profile_stages &&
                   !cfg.count_traversal && !cfg.aux_primary_hit && !cfg.env_importance_sampling && !ctx->keep_launch_totals && !ctx->defer_fold;
    if (overlap) {
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); overlap = false; }
    }
    int lane = 0;
    if (overlap) { lane = ctx->next_lane % n_lanes; ctx->next_lane = (lane + 1) % n_lanes; }
    ptb_context::LaneSync& ls = ctx->lane_sync[lane];
    if (!ls.resolve_done) {
        CU(cudaEventCreateWithFlags(&ls.render_done, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ls.resolve_done, cudaEventDisableTiming));
    }
    if (overlap && !ls.stream) CU(cudaStreamCreateWithFlags(&ls.stream, cudaStreamNonBlocking));
    const cudaStream_t rst = overlap ? ls.stream : st;   // where the path tracing of this launch runs
    LaneGuard lane_guard(ctx, lane);                      // lanes 1.. bring their own pool for the duration of this call
    // the lane's previous accumulate stage (on whatever stream it ran) has to be done with the pool before it is overwritten
    if (ls.used) CU(cudaStreamWaitEvent(rst, ls.resolve_done, 0));

    int rc = ensure_pool(ctx, use_pool ? pool_grid * PTB_CHUNK : slots, iters);
    if (rc != PTB_OK) return rc;
    if (use_pool && slots > ctx->out_pixsum_slots) {
        cudaFree(ctx->out_pixsum); ctx->out_pixsum = nullptr; ctx->out_pixsum_slots = 0;
        CU(cudaMalloc((void**)&ctx->out_pixsum, (size_t)slots * 16));
        ctx->out_pixsum_slots = slots;
    }

    FrameView f;
    auto make_fastdiv = [](uint32_t d) {
        FastDiv fd; fd.d = d;
        uint32_t L = 0; while ((1ull << L) < (unsigned long long)d) ++L;
        fd.sh = 31u + L;
        fd.m = (unsigned long long)((((unsigned __int128)1 << fd.sh) + d - 1) / d);
        return fd;
    };
    f.div_w = make_fastdiv(P->image_width); f.div_pixels = make_fastdiv(n_pixels);
    f.W = P->image_width; f.H = P->image_height; f.row0 = row0; f.il_n = il_n; f.il_r = il_r; f.il_h = il_h; f.n_pixels = n_pixels; f.n_subframes = n_sub; f.subframe = P->subframe_index; f.dof = P->dof ? 1 : 0;
    f.eye = make_float3(P->eye.x, P->eye.y, P->eye.z); f.U = make_float3(P->U.x, P->U.y, P->U.z);
    f.V = make_float3(P->V.x, P->V.y, P->V.z); f.Wv = make_float3(P->W.x, P->W.y, P->W.z);
    f.spp = cfg.spp_per_launch; f.max_depth = cfg.max_depth; f.tmin = cfg.tmin; f.tmax = cfg.tmax;
    f.dof_blur = cfg.dof_blur; f.focus_dist = cfg.focus_dist; f.nmap_strength = cfg.nmap_strength;
    f.exposure_scale = exp2f(cfg.exposure); f.inv_gamma = 1.0f / cfg.gamma; f.contrast = cfg.contrast;
    f.accumulate_mode = cfg.accumulate_mode; f.write_frame = cfg.write_frame;
    f.accum = (float4*)P->accum_buffer; f.frame = (uchar4*)P->frame_buffer; f.aux_primary = cfg.aux_primary_hit;

    PathView p;
    p.ray_o = ctx->ray_o; p.ray_d = ctx->ray_d; p.hit = ctx->hit; p.atten_seed = ctx->atten_seed; p.misc = ctx->misc;
    p.pixsum = ctx->pixsum; p.n_slots = slots;
    QueueView q;
    q.trace[0] = ctx->q_trace[0]; q.trace[1] = ctx->q_trace[1]; q.hit = ctx->q_hit; q.miss = ctx->q_miss;
    q.counters = ctx->counters; q.trav_stats = ctx->trav_stats;
    const SceneView s = scene_view(d);

    CU(cudaMemsetAsync(ctx->counters, 0, (size_t)(iters + 2) * 4 * sizeof(uint32_t), rst));
    if (!ctx->keep_launch_totals) {  // later batches of one launch add to the counters of the first
        CU(cudaMemsetAsync(ctx->trav_stats, 0, 2 * sizeof(unsigned long long), rst));
        CU(cudaMemsetAsync(ctx->launch_totals, 0, 4 * sizeof(unsigned long long), rst));
    }
    const uint32_t pix_blocks = (slots + 255u) / 256u;
    const bool prof = cfg.profile_stages != 0;
    if (prof) {
        const size_t need_ev = (size_t)iters * 3 + 4;
        while (ctx->events.size() < need_ev) { cudaEvent_t e; CU(cudaEventCreate(&e)); ctx->events.push_back(e); }
        CU(cudaEventRecord(ctx->events[0], rst));
    }
    ctx->prof_iters = 0;
    ctx->stage_sum_valid = false;
    uint32_t launches = 0;
    uint32_t prof_iters = iters;
    if (cfg.env_importance_sampling) {
        // optional mode beyond the reference (linear.cuh): trace -> shade(+NEE) -> shadow -> miss, one kernel per stage
        if (slots > ctx->shadow_slots) {
            cudaFree(ctx->shadow_o); cudaFree(ctx->shadow_d); cudaFree(ctx->shadow_c); cudaFree(ctx->shadow_flag);
            ctx->shadow_o = ctx->shadow_d = ctx->shadow_c = nullptr; ctx->shadow_flag = nullptr; ctx->shadow_slots = 0;
            CU(cudaMalloc((void**)&ctx->shadow_o, (size_t)slots * 16)); CU(cudaMalloc((void**)&ctx->shadow_d, (size_t)slots * 16));
            CU(cudaMalloc((void**)&ctx->shadow_c, (size_t)slots * 16)); CU(cudaMalloc((void**)&ctx->shadow_flag, (size_t)slots + 64));
            ctx->shadow_slots = slots;
        }
        CU(cudaMemsetAsync(ctx->shadow_flag, 0, (size_t)slots + 64, rst));
        LinearView lv;
        lv.cdf.marginal = d->cdf_marginal; lv.cdf.conditional = d->cdf_conditional; lv.cdf.row_weight = d->cdf_row_weight;
        lv.cdf.total = d->cdf_total; lv.cdf.w = d->env_w; lv.cdf.h = d->env_h;
        lv.shadow_o = ctx->shadow_o; lv.shadow_d = ctx->shadow_d; lv.shadow_c = ctx->shadow_c; lv.shadow_flag = ctx->shadow_flag;
        lv.nee = cfg.env_importance_sampling == 1 ? 1 : 0;
        const uint32_t chunks = (slots + PTB_CHUNK - 1u) / PTB_CHUNK;
        k_chunk_raygen<<<pix_blocks, 256, 0, rst>>>(f, p, ctx->status);
        if (prof) CU(cudaEventRecord(ctx->events[1], rst));
        launches = 1;
        for (uint32_t it = 0; it < iters; ++it) {
            k_chunk_trace<false, PTB_TRACE_QUANTUM><<<chunks, PTB_CHUNK_THREADS, 0, rst>>>(s, f, p, ctx->status, ctx->launch_totals, ctx->trav_stats, (int)it);
            if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + 0], rst));
            k_chunk_shade_linear<<<chunks, PTB_CHUNK_THREADS, 0, rst>>>(s, f, p, lv, ctx->status);
            if (lv.nee) { k_chunk_shadow<<<chunks, PTB_CHUNK_THREADS, 0, rst>>>(s, f, p, lv); launches += 1; }
            if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + 1], rst));
            k_chunk_miss_linear<<<chunks, PTB_CHUNK_THREADS, 0, rst>>>(s, f, p, lv, ctx->status);
            if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + 2], rst));
            launches += 3;
        }
    } else if (pipeline == PTB_PIPELINE_QUEUES) {
        // global queues, one kernel per stage and iteration (kernels.cuh)
        k_raygen_init<<<pix_blocks, 256, 0, rst>>>(f, p, q);
        if (prof) CU(cudaEventRecord(ctx->events[1], rst));
        const uint32_t need = (slots + 127u) / 128u;
        const uint32_t cap = (uint32_t)ctx->num_sms * 32u;
        const uint32_t grid = need < cap ? need : cap;
        const uint32_t tcap = (uint32_t)ctx->num_sms * 10u;  // k_trace is persistent
        const uint32_t tgrid = need < tcap ? need : tcap;
        launches = 1;
        for (uint32_t it = 0; it < iters; ++it) {
            if (cfg.count_traversal) k_trace<true, PTB_TRACE_QUANTUM><<<tgrid, 128, 0, rst>>>(s, f, p, q, (int)it);
            else k_trace<false, PTB_TRACE_QUANTUM><<<tgrid, 128, 0, rst>>>(s, f, p, q, (int)it);
            if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + 0], rst));
            k_shade<<<grid, 128, 0, rst>>>(s, f, p, q, (int)it);
            if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + 1], rst));
            k_miss<<<grid, 128, 0, rst>>>(s, f, p, q, (int)it);
            if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + 2], rst));
            launches += 3;
        }
        k_fold_counters<<<1, 256, 0, rst>>>(ctx->counters, iters, ctx->launch_totals);
        launches += 1;
    } else if (use_pool) {
        // persistent block-local wavefront (pool.cuh): one kernel, camera rays included
        unsigned int* max_iters = (unsigned int*)(ctx->launch_totals + 3);
        PoolView pv;
        pv.out_pixsum = ctx->out_pixsum; pv.next_slot = max_iters + 1;  // both words zeroed by the memset above
        pv.n_slots = slots; pv.grid = pool_grid;
        PathView ps = p; ps.n_slots = pool_grid * PTB_CHUNK;
        if (prof) CU(cudaEventRecord(ctx->events[1], rst));  // no separate raygen kernel
        if (cfg.count_traversal) k_pool_fused<true, PTB_TRACE_QUANTUM, 5><<<pool_grid, PTB_CHUNK_THREADS, 0, rst>>>(s, f, ps, pv, ctx->launch_totals, ctx->trav_stats, max_iters);
        else if (pool_wide) k_pool_fused<false, PTB_TRACE_QUANTUM, 8><<<pool_grid, PTB_CHUNK_THREADS, 0, rst>>>(s, f, ps, pv, ctx->launch_totals, ctx->trav_stats, max_iters);
        else k_pool_fused<false, PTB_TRACE_QUANTUM, 5><<<pool_grid, PTB_CHUNK_THREADS, 0, rst>>>(s, f, ps, pv, ctx->launch_totals, ctx->trav_stats, max_iters);
        launches = 1;
        prof_iters = 0;
        if (prof) {
            CU(cudaEventRecord(ctx->events[2], rst)); CU(cudaEventRecord(ctx->events[3], rst)); CU(cudaEventRecord(ctx->events[4], rst));
            prof_iters = 1;
        }
        p.pixsum = ctx->out_pixsum;  // what k_resolve folds
    } else {
        // block-local wavefront over chunks of the path pool (chunked.cuh), in the exact build of the kernels (namespace
        // ptb) or in the fast-arithmetic one (fast_kernels.cu, cfg.arith_mode = 1)
        ChunkLaunch cl;
        cl.s = s; cl.f = f; cl.p = p; cl.status = ctx->status; cl.totals = ctx->launch_totals; cl.trav_stats = ctx->trav_stats;
        cl.max_iters = (unsigned int*)(ctx->launch_totals + 3);
        cl.count = cfg.count_traversal ? 1 : 0; cl.num_sms = ctx->num_sms;
        static const int env_spt = getenv("PTB_SPT") ? atoi(getenv("PTB_SPT")) : 0;  // experiments only
        cl.spt_request = cfg.chunk_slots_per_thread ? cfg.chunk_slots_per_thread : env_spt;
        if (cl.spt_request != 0 && cl.spt_request != 8 && cl.spt_request != 4 && cl.spt_request != 2 && cl.spt_request != 1)
            return fail(PTB_ERR_INVALID, "ptb_launch: chunk_slots_per_thread must be 0, 1, 2, 4 or 8");
        ptb_fast_api::ChunkLaunchArgs fa;
        if (fast) {
            fa.s = s; fa.f = f; fa.p = p; fa.status = cl.status; fa.totals = cl.totals; fa.trav_stats = cl.trav_stats; fa.max_iters = cl.max_iters;
            fa.num_sms = cl.num_sms; fa.spt_request = cl.spt_request; fa.count = cl.count;
            ptb_fast_api::raygen(fa, rst);
        } else launch_chunk_raygen(cl, rst);
        if (prof) CU(cudaEventRecord(ctx->events[1], rst));
        launches = 1;
        if (pipeline == PTB_PIPELINE_CHUNK_STAGES) {
            for (uint32_t it = 0; it < iters; ++it) {
                for (int stage = 0; stage < 3; ++stage) {
                    if (fast) ptb_fast_api::stage(fa, stage, (int)it, rst); else launch_chunk_stage(cl, stage, (int)it, rst);
                    if (prof) CU(cudaEventRecord(ctx->events[2 + (size_t)it * 3 + stage], rst));
                }
                launches += 3;
            }
        } else {
            if (fast) ptb_fast_api::fused(fa, rst); else launch_chunk_fused(cl, rst);
            launches += 1;
            prof_iters = 0;
            if (prof) {  // a single kernel: everything between raygen and resolve is reported as "trace"
                CU(cudaEventRecord(ctx->events[2], rst)); CU(cudaEventRecord(ctx->events[3], rst)); CU(cudaEventRecord(ctx->events[4], rst));
                prof_iters = 1;
            }
        }
    }
    if (overlap) { CU(cudaEventRecord(ls.render_done, rst)); CU(cudaStreamWaitEvent(st, ls.render_done, 0)); }
    k_resolve<<<(n_pixels + 255u) / 256u, 256, 0, st>>>(f, p);
    if (!ctx->defer_fold) k_fold_totals<<<1, 32, 0, st>>>(ctx->launch_totals, ctx->totals);
    CU(cudaEventRecord(ls.resolve_done, st));
    ls.used = true;
    ctx->last_lane = lane;
    launches += 2;
    if (prof) { CU(cudaEventRecord(ctx->events[2 + (size_t)prof_iters * 3], st)); ctx->prof_iters = prof_iters; }
    CU(cudaGetLastError());
    ctx->last_stream = st; ctx->last_iters = iters; ctx->last_kernels = launches; ctx->last_pipeline = pipeline;
    ctx->last_paths = (uint64_t)slots * (uint64_t)cfg.spp_per_launch;  // slots already counts the batched subframes
    ctx->last_counted = cfg.count_traversal != 0;
    return PTB_OK;
}

int ptb_launch_get_stats(ptb_context* ctx, ptb_launch_stats* out) {
    if (!ctx || !out) return fail(PTB_ERR_INVALID, "ptb_launch_get_stats: bad arguments");
    memset(out, 0, sizeof(*out));
    if (!ctx->last_iters) return fail(PTB_ERR_INVALID, "ptb_launch_get_stats: no launch yet");
    CU(cudaSetDevice(ctx->device));
    LaneGuard lane_guard(ctx, ctx->last_lane);   // the counters of the lane the last launch ran on
    std::vector<uint32_t> h((size_t)(ctx->last_iters + 1) * 4);
    CU(cudaMemcpyAsync(h.data(), ctx->counters, h.size() * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->last_stream));
    unsigned long long tv[2] = {0, 0}, lt[4] = {0, 0, 0, 0};
    CU(cudaMemcpyAsync(tv, ctx->trav_stats, sizeof(tv), cudaMemcpyDeviceToHost, ctx->last_stream));
    CU(cudaMemcpyAsync(lt, ctx->launch_totals, sizeof(lt), cudaMemcpyDeviceToHost, ctx->last_stream));
    CU(cudaStreamSynchronize(ctx->last_stream));
    out->segments = lt[0]; out->hits = lt[1]; out->misses = lt[2];
    uint32_t used = 0;
    if (ctx->last_pipeline == PTB_PIPELINE_CHUNK_FUSED || ctx->last_pipeline == PTB_PIPELINE_POOL_FUSED) used = (uint32_t)(lt[3] & 0xffffffffull);
    else for (uint32_t it = 0; it < ctx->last_iters; ++it) if (h[(size_t)it * 4 + 0]) used = it + 1;
    out->paths = ctx->last_paths; out->iterations = used; out->kernel_launches = ctx->last_kernels;
    if (ctx->last_counted) { out->nodes_visited = tv[0]; out->tris_tested = tv[1]; }
    return PTB_OK;
}

int ptb_context_get_totals(ptb_context* ctx, uint64_t out[4], int reset) {
    if (!ctx || !out) return fail(PTB_ERR_INVALID, "ptb_context_get_totals: bad arguments");
    for (int i = 0; i < 4; ++i) out[i] = 0;
    if (!ctx->totals) return PTB_OK;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));
    unsigned long long h[4];
    CU(cudaMemcpy(h, ctx->totals, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out[i] = h[i];
    if (reset) CU(cudaMemset(ctx->totals, 0, sizeof(h)));
    return PTB_OK;
}

int ptb_launch_get_stage_ms(ptb_context* ctx, float out[6]) {
    if (!ctx || !out) return fail(PTB_ERR_INVALID, "ptb_launch_get_stage_ms: bad arguments");
    if (ctx->stage_sum_valid) { for (int i = 0; i < 6; ++i) out[i] = ctx->stage_sum[i]; return PTB_OK; }  // a batched launch: sums
    if (!ctx->prof_iters) return fail(PTB_ERR_INVALID, "ptb_launch_get_stage_ms: the last launch did not set profile_stages");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->last_stream));
    for (int i = 0; i < 6; ++i) out[i] = 0.0f;
    float ms = 0.0f;
    const std::vector<cudaEvent_t>& ev = ctx->events;
    CU(cudaEventElapsedTime(&ms, ev[0], ev[1])); out[0] = ms;
    for (uint32_t it = 0; it < ctx->prof_iters; ++it) {
        const size_t b = 2 + (size_t)it * 3;
        CU(cudaEventElapsedTime(&ms, ev[b - 1], ev[b])); out[1] += ms;
        CU(cudaEventElapsedTime(&ms, ev[b], ev[b + 1])); out[2] += ms;
        CU(cudaEventElapsedTime(&ms, ev[b + 1], ev[b + 2])); out[3] += ms;
    }
    const size_t last = 2 + (size_t)ctx->prof_iters * 3;
    CU(cudaEventElapsedTime(&ms, ev[last - 1], ev[last])); out[4] = ms;
    CU(cudaEventElapsedTime(&ms, ev[0], ev[last])); out[5] = ms;
    return PTB_OK;
}

int ptb_resolve(ptb_context* ctx, const ptb_float4* accum, ptb_float4* accum_out, ptb_uchar4* frame, uint32_t n_pixels, float scale,
                const ptb_render_cfg* cfg_in, void* stream_) {
    if (!ctx || !accum || !n_pixels) return fail(PTB_ERR_INVALID, "ptb_resolve: bad arguments");
    ptb_render_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else ptb_default_render_cfg(&cfg);
    CU(cudaSetDevice(ctx->device));
    k_resolve_scaled<<<(n_pixels + 255u) / 256u, 256, 0, (cudaStream_t)stream_>>>((const float4*)accum, (float4*)accum_out, (uchar4*)frame,
                                                                                 n_pixels, scale, exp2f(cfg.exposure), 1.0f / cfg.gamma, cfg.contrast);
    CU(cudaGetLastError());
    return PTB_OK;
}

int ptb_resolve_peers(ptb_context* ctx, const ptb_float4* const* accums, int n_ranks, ptb_float4* accum_out, ptb_uchar4* frame,
                      uint32_t first_pixel, uint32_t n_pixels, float scale, const ptb_render_cfg* cfg_in, void* stream_) {
    return ptb_resolve_peers_accumulate(ctx, accums, n_ranks, nullptr, 0.0f, accum_out, frame, first_pixel, n_pixels, scale, cfg_in, stream_);
}

int ptb_resolve_peers_accumulate(ptb_context* ctx, const ptb_float4* const* accums, int n_ranks, const ptb_float4* prev_accum, float prev_weight,
                                 ptb_float4* accum_out, ptb_uchar4* frame, uint32_t first_pixel, uint32_t n_pixels, float scale,
                                 const ptb_render_cfg* cfg_in, void* stream_) {
    return ptb_resolve_peers_sync(ctx, accums, n_ranks, 0, nullptr, nullptr, 0u, prev_accum, prev_weight, accum_out, frame, first_pixel, n_pixels, scale,
                                  cfg_in, stream_);
}

int ptb_resolve_peers_sync(ptb_context* ctx, const ptb_float4* const* accums, int n_ranks, int my_rank, uint32_t* my_flags, uint32_t* root_flags,
                           uint32_t epoch, const ptb_float4* prev_accum, float prev_weight, ptb_float4* accum_out, ptb_uchar4* frame,
                           uint32_t first_pixel, uint32_t n_pixels, float scale, const ptb_render_cfg* cfg_in, void* stream_) {
    if (!ctx || !accums || n_ranks < 1 || n_ranks > PTB_MAX_RANKS) return fail(PTB_ERR_INVALID, "ptb_resolve_peers: bad arguments");
    if (my_rank < 0 || my_rank >= n_ranks || (root_flags && !my_flags)) return fail(PTB_ERR_INVALID, "ptb_resolve_peers_sync: bad rank / flag blocks");
    ptb_render_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else ptb_default_render_cfg(&cfg);
    PeerAccums pa;
    pa.n = n_ranks;
    for (int k = 0; k < n_ranks; ++k) { if (!accums[k]) return fail(PTB_ERR_INVALID, "ptb_resolve_peers: null accumulator"); pa.a[k] = (const float4*)accums[k]; }
    PeerSync sy;
    sy.my_flags = my_flags; sy.root_done = root_flags ? root_flags + PTB_FLAG_DONE + my_rank : nullptr; sy.n = n_ranks; sy.epoch = epoch;
    CU(cudaSetDevice(ctx->device));
    // with a done signal the kernel must run even for an empty slice (one block: the signal itself)
    const uint32_t blocks = n_pixels ? (n_pixels + 255u) / 256u : (root_flags ? 1u : 0u);
    if (blocks) k_resolve_peers<<<blocks, 256, 0, (cudaStream_t)stream_>>>(pa, sy, (const float4*)prev_accum, prev_weight, (float4*)accum_out, (uchar4*)frame, first_pixel, n_pixels,
                                                                                                scale, exp2f(cfg.exposure), 1.0f / cfg.gamma, cfg.contrast);
    CU(cudaGetLastError());
    return PTB_OK;
}

int ptb_peer_signal(ptb_context* ctx, uint32_t* const* flag_blocks, int n_ranks, int my_rank, int kind, uint32_t epoch, void* stream_) {
    if (!ctx || !flag_blocks || n_ranks < 1 || n_ranks > PTB_MAX_RANKS || my_rank < 0 || my_rank >= n_ranks || (kind != 0 && kind != 1))
        return fail(PTB_ERR_INVALID, "ptb_peer_signal: bad arguments");
    PeerFlags pf; pf.n = n_ranks;
    for (int k = 0; k < n_ranks; ++k) { if (!flag_blocks[k]) return fail(PTB_ERR_INVALID, "ptb_peer_signal: null flag block"); pf.f[k] = flag_blocks[k]; }
    CU(cudaSetDevice(ctx->device));
    k_peer_signal<<<1, 32, 0, (cudaStream_t)stream_>>>(pf, my_rank, kind ? PTB_FLAG_DONE : 0, epoch);
    CU(cudaGetLastError());
    return PTB_OK;
}

int ptb_peer_wait(ptb_context* ctx, uint32_t* my_flags, int kind, int n_ranks, uint32_t epoch, void* stream_) {
    if (!ctx || !my_flags || n_ranks < 1 || n_ranks > PTB_MAX_RANKS || (kind != 0 && kind != 1)) return fail(PTB_ERR_INVALID, "ptb_peer_wait: bad arguments");
    CU(cudaSetDevice(ctx->device));
    k_peer_wait<<<1, 32, 0, (cudaStream_t)stream_>>>(my_flags, kind ? PTB_FLAG_DONE : 0, n_ranks, epoch);
    CU(cudaGetLastError());
    return PTB_OK;
}

int ptb_peer_flags_create(ptb_context* ctx, uint32_t** flags) {
    if (!ctx || !flags) return fail(PTB_ERR_INVALID, "ptb_peer_flags_create: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc((void**)flags, PTB_FLAG_WORDS * sizeof(uint32_t)));
    CU(cudaMemset(*flags, 0, PTB_FLAG_WORDS * sizeof(uint32_t)));
    return PTB_OK;
}

int ptb_peer_flags_error(ptb_context* ctx, const uint32_t* flags, void* stream_, int* timed_out) {
    if (!ctx || !flags || !timed_out) return fail(PTB_ERR_INVALID, "ptb_peer_flags_error: bad arguments");
    CU(cudaSetDevice(ctx->device));
    uint32_t w = 0;
    CU(cudaMemcpyAsync(&w, flags + PTB_FLAG_ERROR, sizeof(w), cudaMemcpyDeviceToHost, (cudaStream_t)stream_));
    CU(cudaStreamSynchronize((cudaStream_t)stream_));
    *timed_out = w != 0;
    return PTB_OK;
}

int ptb_ipc_export(ptb_context* ctx, const void* device_ptr, unsigned char handle[64]) {
    if (!ctx || !device_ptr || !handle) return fail(PTB_ERR_INVALID, "ptb_ipc_export: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, const_cast<void*>(device_ptr)));
    memcpy(handle, &h, 64);
    return PTB_OK;
}
int ptb_ipc_open(ptb_context* ctx, const unsigned char handle[64], void** device_ptr) {
    if (!ctx || !handle || !device_ptr) return fail(PTB_ERR_INVALID, "ptb_ipc_open: bad arguments");
    CU(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(device_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return PTB_OK;
}
int ptb_ipc_close(ptb_context* ctx, void* device_ptr) {
    if (!ctx || !device_ptr) return fail(PTB_ERR_INVALID, "ptb_ipc_close: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaIpcCloseMemHandle(device_ptr));
    return PTB_OK;
}

int ptb_trace_rays(ptb_context* ctx, unsigned long long handle, const float* d_origins, const float* d_dirs, uint32_t n, float tmin,
                   float tmax, int32_t* d_prim, float* d_t, float* d_b1, float* d_b2, void* stream_) {
    if (!ctx || !d_origins || !d_dirs) return fail(PTB_ERR_INVALID, "ptb_trace_rays: bad arguments");
    DeviceScene* d = find_scene(ctx, handle);
    if (!d) return fail(PTB_ERR_INVALID, "ptb_trace_rays: unknown handle");
    CU(cudaSetDevice(ctx->device));
    if (n) k_trace_rays<<<(n + 127u) / 128u, 128, 0, (cudaStream_t)stream_>>>(scene_view(d), d_origins, d_dirs, n, tmin, tmax, d_prim, d_t, d_b1, d_b2);
    CU(cudaGetLastError());
    return PTB_OK;
}

// ---- output buffer -------------------------------------------------------------------------
int ptb_output_create(ptb_context* ctx, uint32_t width, uint32_t height, ptb_output** out) {
    if (!ctx || !out || !width || !height) return fail(PTB_ERR_INVALID, "ptb_output_create: bad arguments");
    CU(cudaSetDevice(ctx->device));
    ptb_output* o = new ptb_output();
    o->ctx = ctx;
    int rc = ptb_output_resize(o, width, height);
    if (rc != PTB_OK) { delete o; return rc; }
    *out = o;
    return PTB_OK;
}
int ptb_output_resize(ptb_output* ob, uint32_t width, uint32_t height) {
    if (!ob || !width || !height) return fail(PTB_ERR_INVALID, "ptb_output_resize: bad arguments");
    CU(cudaSetDevice(ob->ctx->device));
    if (ob->d_pixels) { cudaFree(ob->d_pixels); ob->d_pixels = nullptr; }
    CU(cudaMalloc((void**)&ob->d_pixels, (size_t)width * height * sizeof(uchar4)));
    ob->w = width; ob->h = height; ob->host.clear();
    return PTB_OK;
}
ptb_uchar4* ptb_output_map(ptb_output* ob) { return ob ? (ptb_uchar4*)ob->d_pixels : nullptr; }
void ptb_output_unmap(ptb_output* ob, void* stream) {
    if (!ob) return;
    cudaSetDevice(ob->ctx->device);
    cudaStreamSynchronize((cudaStream_t)stream);  // CUDAOutputBuffer::unmap synchronises its stream
}
const ptb_uchar4* ptb_output_host_ptr(ptb_output* ob) {
    if (!ob) return nullptr;
    cudaSetDevice(ob->ctx->device);
    ob->host.resize((size_t)ob->w * ob->h);
    if (cudaMemcpy(ob->host.data(), ob->d_pixels, ob->host.size() * sizeof(uchar4), cudaMemcpyDeviceToHost) != cudaSuccess) {
        set_error("ptb_output_host_ptr: device to host copy failed");
        return nullptr;
    }
    return ob->host.data();
}
uint32_t ptb_output_width(const ptb_output* ob) { return ob ? ob->w : 0; }
uint32_t ptb_output_height(const ptb_output* ob) { return ob ? ob->h : 0; }
void ptb_output_destroy(ptb_output* ob) {
    if (!ob) return;
    cudaSetDevice(ob->ctx->device);
    cudaFree(ob->d_pixels);
    delete ob;
}

// ---- device memory helpers -----------------------------------------------------------------
int ptb_device_alloc(ptb_context* ctx, size_t bytes, void** out) {
    if (!ctx || !out) return fail(PTB_ERR_INVALID, "ptb_device_alloc: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMalloc(out, bytes ? bytes : 16));
    return PTB_OK;
}
int ptb_device_free(ptb_context* ctx, void* p) {
    if (!ctx) return fail(PTB_ERR_INVALID, "ptb_device_free: null context");
    CU(cudaSetDevice(ctx->device));
    CU(cudaFree(p));
    return PTB_OK;
}
int ptb_device_memset(ptb_context* ctx, void* p, int value, size_t bytes, void* stream) {
    if (!ctx || !p) return fail(PTB_ERR_INVALID, "ptb_device_memset: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(p, value, bytes, (cudaStream_t)stream));
    return PTB_OK;
}
int ptb_copy_to_device(ptb_context* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    if (!ctx || !dst || !src) return fail(PTB_ERR_INVALID, "ptb_copy_to_device: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return PTB_OK;
}
int ptb_copy_to_host_async(ptb_context* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    if (!ctx || !dst || !src) return fail(PTB_ERR_INVALID, "ptb_copy_to_host_async: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return PTB_OK;
}
int ptb_copy_to_host(ptb_context* ctx, void* dst, const void* src, size_t bytes, void* stream) {
    if (!ctx || !dst || !src) return fail(PTB_ERR_INVALID, "ptb_copy_to_host: bad arguments");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return PTB_OK;
}

int ptb_microbench_read(ptb_context* ctx, size_t bytes, int iters, double* gb_per_s) {
    if (!ctx || !gb_per_s || bytes < 4096 || iters < 1) return fail(PTB_ERR_INVALID, "ptb_microbench_read: bad arguments");
    CU(cudaSetDevice(ctx->device));
    float4* buf = nullptr; float* sink = nullptr;
    const size_t n_vec = bytes / 16;
    CU(cudaMalloc((void**)&buf, n_vec * 16));
    if (cudaMalloc((void**)&sink, 16) != cudaSuccess) { cudaFree(buf); return fail(PTB_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemset(buf, 0, n_vec * 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int grid = ctx->num_sms * 8;
    k_microbench_read<<<grid, 256>>>(buf, n_vec, 2, sink);  // warm-up: brings the buffer into L2
    cudaEventRecord(e0);
    k_microbench_read<<<grid, 256>>>(buf, n_vec, iters, sink);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    float ms = 0.0f;
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(sink);
    if (e != cudaSuccess || ms <= 0.0f) return fail(PTB_ERR_CUDA, std::string("ptb_microbench_read: ") + cudaGetErrorString(e));
    *gb_per_s = (double)n_vec * 16.0 * (double)iters / ((double)ms * 1e-3) / 1e9;
    return PTB_OK;
}

int ptb_test_env_sample(ptb_context* ctx, unsigned long long handle, const float* xi, uint32_t n, float* out) {
    if (!ctx || !xi || !out) return fail(PTB_ERR_INVALID, "ptb_test_env_sample: bad arguments");
    DeviceScene* d = find_scene(ctx, handle);
    if (!d) return fail(PTB_ERR_INVALID, "ptb_test_env_sample: unknown handle");
    CU(cudaSetDevice(ctx->device));
    float *d_xi = nullptr, *d_out = nullptr;
    CU(cudaMalloc((void**)&d_xi, (size_t)n * 8 + 16));
    if (cudaMalloc((void**)&d_out, (size_t)n * 16 + 16) != cudaSuccess) { cudaFree(d_xi); return fail(PTB_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemcpy(d_xi, xi, (size_t)n * 8, cudaMemcpyHostToDevice);
    EnvCdf cdf;
    cdf.marginal = d->cdf_marginal; cdf.conditional = d->cdf_conditional; cdf.row_weight = d->cdf_row_weight; cdf.total = d->cdf_total;
    cdf.w = d->env_w; cdf.h = d->env_h;
    if (n) k_env_sample_test<<<(n + 255u) / 256u, 256>>>(cdf, d_xi, n, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n * 16, cudaMemcpyDeviceToHost);
    cudaFree(d_xi); cudaFree(d_out);
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, std::string("ptb_test_env_sample: ") + cudaGetErrorString(e));
    return PTB_OK;
}

int ptb_test_device_math(ptb_context* ctx, int op, const float* in, int in_stride, float* out, int out_stride, uint32_t n) {
    if (!ctx || !in || !out || in_stride < 1 || out_stride < 1 || op < 0 || op > 5) return fail(PTB_ERR_INVALID, "ptb_test_device_math: bad arguments");
    CU(cudaSetDevice(ctx->device));
    float *d_in = nullptr, *d_out = nullptr;
    CU(cudaMalloc((void**)&d_in, (size_t)n * in_stride * sizeof(float) + 16));
    if (cudaMalloc((void**)&d_out, (size_t)n * out_stride * sizeof(float) + 16) != cudaSuccess) { cudaFree(d_in); return fail(PTB_ERR_CUDA, "cudaMalloc failed"); }
    cudaMemcpy(d_in, in, (size_t)n * in_stride * sizeof(float), cudaMemcpyHostToDevice);
    cudaMemset(d_out, 0, (size_t)n * out_stride * sizeof(float));
    if (n) k_test_math<<<(n + 255u) / 256u, 256>>>(op, d_in, in_stride, d_out, out_stride, n);
    cudaError_t e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaMemcpy(out, d_out, (size_t)n * out_stride * sizeof(float), cudaMemcpyDeviceToHost);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, std::string("ptb_test_device_math: ") + cudaGetErrorString(e));
    return PTB_OK;
}

}  // extern "C"
