// oracle/ref_shim/ref_driver.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the reference's device file, unmodified and from where it lies
// (REF_CU = /root/reference/optixSphere.cu, passed by oracle/Makefile), into a
// host library and drives its own entry points __raygen__rg /
// __closesthit__radiance / __miss__radiance pixel by pixel.  Exposes the same
// orc_* C interface as the restated oracle (oracle.h).
//
// Because the reference hard-codes its literals (10 samples per launch, depth
// 20, tmin/tmax, DoF constants, normal-map strength, tonemap constants:
// optixSphere.cu:323,360,368-369,285,329,697,412,425,432) this library only
// accepts the default OrcConfig; anything else returns 2.
#include REF_CU

#include <chrono>
#include <mutex>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../oracle.h"
#include "../oracle_isect.h"

namespace refshim {
Global g;
thread_local Tls tls;
}

namespace {
struct TraceUser { orc::TriSoup soup; const orc::CpuBvh* bvh; };

refshim::TraceResult trace_cb(const void* user, const float o[3], const float d[3], float tmin, float tmax) {
    const TraceUser* u = (const TraceUser*)user;
    orc::v3 oo = orc::mk3(o[0], o[1], o[2]), dd = orc::mk3(d[0], d[1], d[2]);
    orc::Hit h = u->bvh ? u->bvh->closest(u->soup, oo, dd, tmin, tmax) : orc::closest_brute(u->soup, oo, dd, tmin, tmax);
    refshim::TraceResult r = {h.prim, h.t, h.b1, h.b2};
    return r;
}

std::mutex g_mu;
const float* g_bvh_key = nullptr; uint32_t g_bvh_n = 0; orc::CpuBvh g_bvh;

bool is_default(const OrcConfig& c) {
    return c.spp_per_launch == 10 && c.max_depth == 20 && c.tmin == 0.01f && c.tmax == 1e16f && c.dof_blur == 0.01f &&
           c.focus_dist == 1.0f && c.nmap_strength == 0.4f && c.exposure == -0.5f && c.gamma == 2.2f &&
           c.contrast == 1.25f && c.accumulate_sum == 0;
}
}  // namespace

extern "C" {

void orc_default_config(OrcConfig* cfg) {
    cfg->spp_per_launch = 10; cfg->max_depth = 20; cfg->tmin = 0.01f; cfg->tmax = 1e16f;
    cfg->dof_blur = 0.01f; cfg->focus_dist = 1.0f; cfg->nmap_strength = 0.4f;
    cfg->exposure = -0.5f; cfg->gamma = 2.2f; cfg->contrast = 1.25f;
    cfg->sat_cuda = 0; cfg->use_bvh = 1; cfg->threads = 0; cfg->accumulate_sum = 0;
}

int orc_render(const OrcScene* scene, const OrcParams* P, const OrcConfig* cfg, float* accum, uint8_t* frame,
               int32_t* primary_hit, OrcStats* stats, int32_t x0, int32_t y0, int32_t x1, int32_t y1) {
    if (!scene || !P || !cfg || !accum || !frame) return 1;
    if (!is_default(*cfg)) return 2;
    std::lock_guard<std::mutex> lk(g_mu);  // `params` and refshim::g are process globals

    TraceUser tu; tu.soup.verts = scene->vertices; tu.soup.n = scene->num_tris; tu.bvh = nullptr;
    if (cfg->use_bvh) {
        if (g_bvh_key != scene->vertices || g_bvh_n != scene->num_tris) {
            g_bvh.build(tu.soup); g_bvh_key = scene->vertices; g_bvh_n = scene->num_tris;
        }
        tu.bvh = &g_bvh;
    }

    // Hit-group table exactly as the reference fills it (optixSphere.cpp:1196-1261),
    // except that each material points at its own textures (oracle rule R5).
    std::vector<HitGroupData> hg((size_t)scene->num_mats);
    for (int i = 0; i < scene->num_mats; ++i) {
        const OrcMaterial& m = scene->mats[i];
        HitGroupData& h = hg[(size_t)i];
        memset(&h, 0, sizeof(h));
        h.vertices = (float4*)scene->vertices; h.normals = (float4*)scene->normals; h.texcoords = (float2*)scene->texcoords;
        h.emission_color = float3{m.emission_color[0], m.emission_color[1], m.emission_color[2]};
        h.diffuse_color = float3{m.diffuse_color[0], m.diffuse_color[1], m.diffuse_color[2]};
        h.specular = float3{m.specular[0], m.specular[1], m.specular[2]};
        h.roughness = m.roughness_value; h.metallic = m.metallic_flag != 0; h.transparent = m.transparent_flag != 0;
        h.albedo_texture_data = m.albedo.has ? (float4*)m.albedo.rgba : nullptr;
        h.tex_width = m.albedo.has ? m.albedo.w : 0; h.tex_height = m.albedo.has ? m.albedo.h : 0; h.has_texture = m.albedo.has != 0;
        h.roughness_texture_data = m.roughness.has ? (float4*)m.roughness.rgba : nullptr;
        h.roughness_width = m.roughness.has ? m.roughness.w : 0; h.roughness_height = m.roughness.has ? m.roughness.h : 0;
        h.has_roughness_map = m.roughness.has != 0;
        h.normal_texture_data = m.normal.has ? (float4*)m.normal.rgba : nullptr;
        h.normal_width = m.normal.has ? m.normal.w : 0; h.normal_height = m.normal.has ? m.normal.h : 0;
        h.has_normal_map = m.normal.has != 0;
        h.metallic_texture_data = m.metallic.has ? (float4*)m.metallic.rgba : nullptr;
        h.metallic_width = m.metallic.has ? m.metallic.w : 0; h.metallic_height = m.metallic.has ? m.metallic.h : 0;
        h.has_metallic_map = m.metallic.has != 0;
    }
    MissData ms; ms.hdr_image_data = (float4*)scene->env_rgba; ms.width = scene->env_w; ms.height = scene->env_h;
    RayGenData rg; memset(&rg, 0, sizeof(rg));

    // Params exactly as main() fills it (optixSphere.cpp:1293-1308, 238-247).
    memset(&params, 0, sizeof(params));
    params.image_width = P->width; params.image_height = P->height;
    params.origin_x = (int)P->width / 2; params.origin_y = (int)P->height / 2;
    params.subframe_index = P->subframe_index;
    params.frame_buffer = (uchar4*)frame; params.accum_buffer = (float4*)accum;
    params.dof = P->dof != 0;
    params.eye = float3{P->eye[0], P->eye[1], P->eye[2]};
    params.U = float3{P->U[0], P->U[1], P->U[2]};
    params.V = float3{P->V[0], P->V[1], P->V[2]};
    params.W = float3{P->W[0], P->W[1], P->W[2]};
    params.handle = 1;

    refshim::g.trace = trace_cb; refshim::g.user = &tu;
    refshim::g.raygen_data = &rg; refshim::g.miss_data = &ms;
    refshim::g.hitgroup_base = hg.data(); refshim::g.hitgroup_stride = sizeof(HitGroupData);
    refshim::g.mat_ids = scene->mat_ids; refshim::g.primary_hit = primary_hit; refshim::g.width = P->width;

    int nthreads = cfg->threads;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
    unsigned long long seg = 0, hits = 0, misses = 0, hung = 0;
    auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : seg, hits, misses, hung)
    for (int32_t y = y0; y < y1; ++y) {
        refshim::Tls& t = refshim::tls;
        t.segments = t.hits = t.misses = 0;
        for (int32_t x = x0; x < x1; ++x) {
            t.idx = uint3{(unsigned)x, (unsigned)y, 0u}; t.dim = uint3{P->width, P->height, 1u};
            t.sbt = refshim::g.raygen_data; t.segs_this_pixel = 0; t.segs_this_path = 0;
            try { __raygen__rg(); } catch (refshim::PathHang&) { hung++; }
        }
        seg += t.segments; hits += t.hits; misses += t.misses;
    }
    auto t1 = std::chrono::steady_clock::now();
    if (stats) {
        stats->segments = seg; stats->hits = hits; stats->misses = misses;
        stats->paths = (uint64_t)(x1 - x0) * (uint64_t)(y1 - y0) * 10u;
        stats->seconds = std::chrono::duration<double>(t1 - t0).count(); stats->threads = nthreads;
    }
    return hung ? 3 : 0;
}

int32_t orc_closest_hit(const OrcScene* scene, const float org[3], const float dir[3], float tmin, float tmax,
                        int32_t use_bvh, float* t, float* b1, float* b2) {
    orc::TriSoup s; s.verts = scene->vertices; s.n = scene->num_tris;
    orc::v3 o = orc::mk3(org[0], org[1], org[2]), d = orc::mk3(dir[0], dir[1], dir[2]);
    orc::Hit h;
    if (use_bvh) { orc::CpuBvh b; b.build(s); h = b.closest(s, o, d, tmin, tmax); }
    else h = orc::closest_brute(s, o, d, tmin, tmax);
    if (t) *t = h.t;
    if (b1) *b1 = h.b1;
    if (b2) *b2 = h.b2;
    return h.prim;
}

// The reference's own RNG (cu:24-35), called directly.
uint32_t orc_rng_next(uint32_t seed, int32_t, float* u) {
    unsigned int s = seed;
    float r = myrnd(s);
    if (u) *u = r;
    return s;
}
void orc_sincos(float x, float* s, float* c) { *s = sinf(x); *c = cosf(x); }
float orc_atan2(float y, float x) { return atan2f(y, x); }
float orc_asin(float x) { return asinf(x); }
float orc_pow(float x, float y) { return powf(x, y); }
void orc_tonemap_pixel(const float rgb[3], const OrcConfig*, uint8_t out[4]) {
    // cu:411-435 inlined in __raygen__rg; not separately callable.  Unsupported here.
    (void)rgb; out[0] = out[1] = out[2] = out[3] = 0;
}
void orc_sample_texture(const OrcTexture* tex, float u, float v, float out[4]) {
    float4 c = sampleTexture((float4*)tex->rgba, tex->w, tex->h, u, v);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
void orc_sample_env(const float* env, int32_t w, int32_t h, const float dir[3], float out[4]) {
    float3 d = normalize(float3{dir[0], dir[1], dir[2]});
    float u = 0.5f + atan2f(d.z, d.x) / (2.0f * M_PIf);
    float v = 0.5f - asinf(d.y) / M_PIf;
    float4 c = sampleHDRI((float4*)env, w, h, u, v);
    out[0] = c.x; out[1] = c.y; out[2] = c.z; out[3] = c.w;
}
void orc_camera_uvw(const float*, const float*, const float*, float, float, float*, float*, float*) {}
const char* orc_impl_name(void) { return "reference-host-shim"; }

}  // extern "C"
