// multi.cpp -- multi-GPU rendering behind the C ABI: ONE host process drives n devices.
//
// The reference renders on one GPU (one OptiX device context, optixSphere.cpp:798-812; render loop 1390-1437).  A
// ptb_multi is the n-device counterpart of that context: the scene is replicated on every device
// (ptb_multi_accel_build), and one ptb_multi_launch() renders the K = subframes_per_launch subframes of a launch by
//   PTB_SPLIT_SAMPLES  dealing contiguous blocks of subframes to the devices (each keeps its GLOBAL subframe indices,
//                      which seed the RNG, optixSphere.cu:316) into per-device sum accumulators, then ONE fused
//                      reduce-scatter -> accumulate -> tonemap -> gather kernel per device over peer memory
//                      (k_resolve_peers: peer loads of the other devices' accumulators over NVLink, peer stores into the
//                      root's float4 / uchar4 buffers);
//   PTB_SPLIT_TILES    dealing 16-row strips of the frame round-robin to the devices; every device renders its strips
//                      with the reference's own accumulate mode straight into the root's buffers through peer pointers
//                      (36 B per pixel over NVLink), bit-identical to a single-GPU launch.
// Ordering is by CUDA events between the per-device streams of this one process; there is no NCCL and no host
// synchronisation inside a launch.  This layer uses only the single-GPU C ABI (include/ptb.h) plus the CUDA runtime.
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "host.h"

using namespace ptb;

struct ptb_multi {
    int n = 0;
    std::vector<int> device;
    std::vector<ptb_context*> ctx;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev_render, ev_resolve;
    std::vector<unsigned long long> handle;
    std::vector<void*> local_accum;  // per device: sum-mode accumulator of the sample split (n_pixels float4)
    size_t local_pixels = 0;
    bool resolve_pending = false;    // ev_resolve[] of the previous launch are recorded and not yet waited for
    int strip_rows = 16;
};

namespace {
int fail(int code, const std::string& msg) { set_error(msg); return code; }
#define CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return fail(PTB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define PT(call) do { int rc_ = (call); if (rc_ != PTB_OK) return rc_; } while (0)

void free_local(ptb_multi* m) {
    for (int g = 0; g < m->n; ++g)
        if (g < (int)m->local_accum.size() && m->local_accum[g]) { cudaSetDevice(m->device[g]); cudaFree(m->local_accum[g]); m->local_accum[g] = nullptr; }
    m->local_pixels = 0;
}
}  // namespace

extern "C" {

int ptb_multi_create(const int* devices, int n_devices, ptb_multi** out) {
    if (!devices || n_devices < 1 || n_devices > 16 || !out) return fail(PTB_ERR_INVALID, "ptb_multi_create: bad arguments (1..16 devices)");
    ptb_multi* m = new ptb_multi();
    m->n = n_devices;
    m->device.assign(devices, devices + n_devices);
    m->ctx.assign(n_devices, nullptr); m->stream.assign(n_devices, nullptr);
    m->ev_render.assign(n_devices, nullptr); m->ev_resolve.assign(n_devices, nullptr);
    m->handle.assign(n_devices, 0ull); m->local_accum.assign(n_devices, nullptr);
    int rc = PTB_OK;
    for (int g = 0; g < n_devices && rc == PTB_OK; ++g) {
        rc = ptb_context_create(devices[g], &m->ctx[g]);
        if (rc != PTB_OK) break;
        cudaError_t e = cudaSetDevice(devices[g]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&m->stream[g], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_render[g], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&m->ev_resolve[g], cudaEventDisableTiming);
        if (e != cudaSuccess) rc = fail(PTB_ERR_CUDA, std::string("ptb_multi_create: ") + cudaGetErrorString(e));
    }
    // every device reads every other device's accumulator and writes the root's buffers: peer access both ways
    for (int g = 0; g < n_devices && rc == PTB_OK; ++g)
        for (int h = 0; h < n_devices && rc == PTB_OK; ++h) {
            if (devices[g] == devices[h]) continue;
            int can = 0;
            cudaError_t e = cudaDeviceCanAccessPeer(&can, devices[g], devices[h]);
            if (e != cudaSuccess || !can) { rc = fail(PTB_ERR_UNSUPPORTED, "ptb_multi_create: device " + std::to_string(devices[g]) + " cannot access the memory of device " + std::to_string(devices[h]) + " (no NVLink / PCIe peer path)"); break; }
            cudaSetDevice(devices[g]);
            e = cudaDeviceEnablePeerAccess(devices[h], 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) rc = fail(PTB_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
        }
    if (rc != PTB_OK) { ptb_multi_destroy(m); return rc; }
    *out = m;
    return PTB_OK;
}

void ptb_multi_destroy(ptb_multi* m) {
    if (!m) return;
    for (int g = 0; g < m->n; ++g) {
        cudaSetDevice(m->device[g]);
        if (m->stream[g]) cudaStreamSynchronize(m->stream[g]);
    }
    free_local(m);
    for (int g = 0; g < m->n; ++g) {
        cudaSetDevice(m->device[g]);
        if (m->ev_render[g]) cudaEventDestroy(m->ev_render[g]);
        if (m->ev_resolve[g]) cudaEventDestroy(m->ev_resolve[g]);
        if (m->stream[g]) cudaStreamDestroy(m->stream[g]);
        if (m->ctx[g]) ptb_context_destroy(m->ctx[g]);
    }
    delete m;
}

int ptb_multi_device_count(const ptb_multi* m) { return m ? m->n : 0; }
ptb_context* ptb_multi_context(ptb_multi* m, int index) { return (m && index >= 0 && index < m->n) ? m->ctx[index] : nullptr; }
void* ptb_multi_stream(ptb_multi* m, int index) { return (m && index >= 0 && index < m->n) ? (void*)m->stream[index] : nullptr; }

int ptb_multi_accel_build(ptb_multi* m, ptb_scene* scene, const ptb_build_cfg* cfg, ptb_build_stats* stats) {
    if (!m || !scene) return fail(PTB_ERR_INVALID, "ptb_multi_accel_build: bad arguments");
    // the build is replicated, not sharded: every device uploads the scene and builds its own BVH
    for (int g = 0; g < m->n; ++g) {
        ptb_build_stats st;
        PT(ptb_accel_build(m->ctx[g], scene, cfg, m->stream[g], &m->handle[g], &st));
        if (g == 0 && stats) *stats = st;
    }
    return PTB_OK;
}

int ptb_multi_synchronize(ptb_multi* m) {
    if (!m) return fail(PTB_ERR_INVALID, "ptb_multi_synchronize: null");
    for (int g = 0; g < m->n; ++g) { CU(cudaSetDevice(m->device[g])); CU(cudaStreamSynchronize(m->stream[g])); }
    m->resolve_pending = false;
    return PTB_OK;
}

int ptb_multi_launch(ptb_multi* m, const ptb_Params* params, const ptb_render_cfg* cfg_in, int split) {
    if (!m || !params) return fail(PTB_ERR_INVALID, "ptb_multi_launch: bad arguments");
    if (split != PTB_SPLIT_SAMPLES && split != PTB_SPLIT_TILES) return fail(PTB_ERR_INVALID, "ptb_multi_launch: split must be PTB_SPLIT_SAMPLES or PTB_SPLIT_TILES");
    ptb_render_cfg cfg;
    if (cfg_in) cfg = *cfg_in; else ptb_default_render_cfg(&cfg);
    if (!m->handle[0]) return fail(PTB_ERR_INVALID, "ptb_multi_launch: no scene has been built (ptb_multi_accel_build)");
    // a row band (a crop of the frame) is the caller's business under the sample split; the tile split owns the row partition
    if (cfg.row_interleave_count > 1 || (split == PTB_SPLIT_TILES && (cfg.row_begin || cfg.row_end)))
        return fail(PTB_ERR_INVALID, "ptb_multi_launch: the row partition is chosen by the split mode");
    if (cfg.accumulate_mode != 0) return fail(PTB_ERR_INVALID, "ptb_multi_launch: accumulate_mode must be 0 (the running average of the reference)");
    const int K = cfg.subframes_per_launch < 1 ? 1 : cfg.subframes_per_launch;
    const int n = m->n;
    // the previous launch's resolve kernels read the local accumulators and write the root's buffers: order after them
    auto wait_previous = [&](int g) -> cudaError_t {
        if (!m->resolve_pending) return cudaSuccess;
        for (int h = 0; h < n; ++h) { cudaError_t e = cudaStreamWaitEvent(m->stream[g], m->ev_resolve[h], 0); if (e != cudaSuccess) return e; }
        return cudaSuccess;
    };
    if (n == 1) {
        ptb_Params p = *params; p.handle = m->handle[0];
        return ptb_launch(m->ctx[0], &p, &cfg, m->stream[0]);
    }
    if (split == PTB_SPLIT_TILES) {
        // every device renders strips g, g + n, ... of the frame into the ROOT's buffers (peer pointers): same arithmetic per
        // pixel as one GPU rendering the whole frame, so the result is bit-identical; no reduction, the stores are the gather
        for (int g = 0; g < n; ++g) {
            CU(cudaSetDevice(m->device[g]));
            CU(wait_previous(g));
            ptb_Params p = *params; p.handle = m->handle[g];
            ptb_render_cfg c = cfg;
            c.row_interleave_count = n; c.row_interleave_index = g; c.row_interleave_height = m->strip_rows;
            if (c.pipeline == PTB_PIPELINE_QUEUES) c.pipeline = 0;
            PT(ptb_launch(m->ctx[g], &p, &c, m->stream[g]));
            CU(cudaEventRecord(m->ev_resolve[g], m->stream[g]));
        }
        m->resolve_pending = true;
        return PTB_OK;
    }
    // ---- sample split -------------------------------------------------------------------------------------------
    const size_t n_pixels = (size_t)params->image_width * params->image_height;
    if (n_pixels == 0 || !params->accum_buffer) return fail(PTB_ERR_INVALID, "ptb_multi_launch: empty image or null accum_buffer");
    if (cfg.write_frame && !params->frame_buffer) return fail(PTB_ERR_INVALID, "ptb_multi_launch: Params.frame_buffer is null (set write_frame = 0 to skip tonemapping)");
    if (n_pixels != m->local_pixels) {
        PT(ptb_multi_synchronize(m));
        free_local(m);
        for (int g = 0; g < n; ++g) { CU(cudaSetDevice(m->device[g])); CU(cudaMalloc(&m->local_accum[g], n_pixels * sizeof(ptb_float4))); }
        m->local_pixels = n_pixels;
    }
    const int per = (K + n - 1) / n;  // contiguous blocks: device g renders subframes [g * per, min(K, (g + 1) * per))
    for (int g = 0; g < n; ++g) {
        CU(cudaSetDevice(m->device[g]));
        CU(wait_previous(g));
        CU(cudaMemsetAsync(m->local_accum[g], 0, n_pixels * sizeof(ptb_float4), m->stream[g]));
        const int lo = g * per < K ? g * per : K, hi = (g + 1) * per < K ? (g + 1) * per : K;
        if (hi > lo) {
            ptb_Params p = *params;
            p.handle = m->handle[g]; p.subframe_index = params->subframe_index + lo;
            p.accum_buffer = (ptb_float4*)m->local_accum[g]; p.frame_buffer = nullptr;
            ptb_render_cfg c = cfg;
            c.accumulate_mode = 1; c.write_frame = 0; c.subframes_per_launch = hi - lo; c.aux_primary_hit = (g == 0) ? cfg.aux_primary_hit : nullptr;
            PT(ptb_launch(m->ctx[g], &p, &c, m->stream[g]));
        }
        CU(cudaEventRecord(m->ev_render[g], m->stream[g]));
    }
    // fused exchange: device g reduces + tonemaps pixels [g * slice, ...) out of ALL accumulators and stores them into the
    // root's buffers; the running average continues from what the root's accumulator holds (weight = subframe_index)
    std::vector<const ptb_float4*> accs(n);
    for (int g = 0; g < n; ++g) accs[g] = (const ptb_float4*)m->local_accum[g];
    const size_t slice = (n_pixels + n - 1) / n;
    const float prev_weight = (float)params->subframe_index;
    const float scale = 1.0f / (float)(params->subframe_index + K);
    for (int g = 0; g < n; ++g) {
        CU(cudaSetDevice(m->device[g]));
        for (int h = 0; h < n; ++h) if (h != g) CU(cudaStreamWaitEvent(m->stream[g], m->ev_render[h], 0));
        const size_t first = (size_t)g * slice < n_pixels ? (size_t)g * slice : n_pixels;
        const size_t count = first + slice <= n_pixels ? slice : n_pixels - first;
        PT(ptb_resolve_peers_accumulate(m->ctx[g], accs.data(), n, params->subframe_index > 0 ? params->accum_buffer : nullptr, prev_weight,
                                        params->accum_buffer, cfg.write_frame ? params->frame_buffer : nullptr, (uint32_t)first, (uint32_t)count, scale, &cfg, m->stream[g]));
        CU(cudaEventRecord(m->ev_resolve[g], m->stream[g]));
    }
    m->resolve_pending = true;
    return PTB_OK;
}

int ptb_multi_get_totals(ptb_multi* m, uint64_t out[4], int reset) {
    if (!m || !out) return fail(PTB_ERR_INVALID, "ptb_multi_get_totals: bad arguments");
    for (int i = 0; i < 4; ++i) out[i] = 0;
    PT(ptb_multi_synchronize(m));
    for (int g = 0; g < m->n; ++g) {
        uint64_t t[4];
        PT(ptb_context_get_totals(m->ctx[g], t, reset));
        for (int i = 0; i < 4; ++i) out[i] += t[i];
    }
    return PTB_OK;
}

}  // extern "C"
