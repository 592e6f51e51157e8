// capi_host.cpp -- the host-only half of the C ABI (include/ptb.h): errors, scene
// objects, camera, defaults, image files.  The device half is renderer.cu.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "host.h"

using namespace ptb;

namespace {

int fail(int code, const std::string& msg) { set_error(msg); return code; }

bool ends_with_ci(const std::string& s, const char* suffix) {
    size_t n = strlen(suffix);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; ++i) if (tolower((unsigned char)s[s.size() - n + i]) != suffix[i]) return false;
    return true;
}

// Keep the 8-bit form when every texel is exactly byte/255.0f.
void texture_from_float4(Texture& t, const ptb_float4* src, int w, int h) {
    t = Texture();
    if (!src || w <= 0 || h <= 0) return;
    t.has = true; t.w = w; t.h = h;
    const size_t n = (size_t)w * h * 4;
    const float* f = (const float*)src;
    t.rgba8.resize(n);
    bool exact = true;
    for (size_t i = 0; i < n && exact; ++i) {
        float v = f[i];
        if (!(v >= 0.0f && v <= 1.0f)) { exact = false; break; }
        int b = (int)(v * 255.0f + 0.5f);
        if ((float)b / 255.0f != v) exact = false;
        t.rgba8[i] = (uint8_t)b;
    }
    if (!exact) {
        t.rgba8.clear(); t.is_float = true;
        t.rgba32f.assign(f, f + n);
    }
}

}  // namespace

extern "C" {

const char* ptb_last_error(void) { return get_error(); }
const char* ptb_version(void) { return "ptb 0.1 (sm_100a wavefront path tracer)"; }

void ptb_default_render_cfg(ptb_render_cfg* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->spp_per_launch = 10; cfg->max_depth = 20; cfg->tmin = 0.01f; cfg->tmax = 1e16f;
    cfg->dof_blur = 0.01f; cfg->focus_dist = 1.0f; cfg->nmap_strength = 0.4f;
    cfg->exposure = -0.5f; cfg->gamma = 2.2f; cfg->contrast = 1.25f;
    cfg->accumulate_mode = 0; cfg->write_frame = 1; cfg->env_importance_sampling = 0; cfg->count_traversal = 0; cfg->profile_stages = 0; cfg->subframes_per_launch = 1; cfg->pipeline = 0; cfg->row_begin = 0; cfg->row_end = 0; cfg->row_interleave_count = 0; cfg->row_interleave_index = 0; cfg->row_interleave_height = 0;
    cfg->aux_primary_hit = nullptr;
}

void ptb_default_build_cfg(ptb_build_cfg* cfg) {
    if (!cfg) return;
    cfg->max_leaf_size = 4; cfg->sah_refine = 1; cfg->sah_bins = 16; cfg->treelet_size = 256; cfg->morton_bits = 30; cfg->bvh_width = 0;
}

int ptb_scene_load_obj(const char* const* files, int n_files, float scale, uint32_t material_seed, ptb_scene** out) {
    if (!files || n_files < 0 || !out) return fail(PTB_ERR_INVALID, "ptb_scene_load_obj: bad arguments");
    std::vector<std::string> names;
    for (int i = 0; i < n_files; ++i) {
        if (!files[i]) return fail(PTB_ERR_INVALID, "ptb_scene_load_obj: null file name");
        names.push_back(files[i]);
    }
    ptb_scene* s = new ptb_scene();
    std::string err;
    if (!build_scene_from_obj(names, scale, material_seed, *s, err)) { delete s; return fail(PTB_ERR_IO, err); }
    *out = s;
    return PTB_OK;
}

int ptb_scene_create(const ptb_TriangleData* tris, uint32_t n_tris, const uint32_t* mat_ids, ptb_scene** out) {
    if ((!tris && n_tris) || !out) return fail(PTB_ERR_INVALID, "ptb_scene_create: bad arguments");
    ptb_scene* s = new ptb_scene();
    s->tris.assign(tris, tris + n_tris);
    if (mat_ids) s->mat_ids.assign(mat_ids, mat_ids + n_tris); else s->mat_ids.assign(n_tris, 0u);
    // one neutral material until ptb_scene_set_materials() is called
    Material m; for (int c = 0; c < 3; ++c) { m.diffuse_color[c] = 0.5f; m.specular[c] = 0.5f; }
    m.roughness = 0.4f;
    uint32_t max_id = 0; for (uint32_t id : s->mat_ids) if (id > max_id) max_id = id;
    s->mats.assign((size_t)max_id + 1, m);
    s->revision++;
    *out = s;
    return PTB_OK;
}

// createSceneGeometry, loadFromFile == false (optixSphere.cpp:650-751) with generateSphereMesh (295-353).
int ptb_scene_create_demo(ptb_scene** out) {
    if (!out) return fail(PTB_ERR_INVALID, "ptb_scene_create_demo: null out");
    ptb_scene* s = new ptb_scene();
    auto f4 = [](float x, float y, float z, float w) { ptb_float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; };
    auto push = [&](ptb_float4 a, ptb_float4 b, ptb_float4 c, ptb_float4 na, ptb_float4 nb, ptb_float4 nc, uint32_t mat) {
        ptb_TriangleData t; memset(&t, 0, sizeof(t));
        t.v0 = a; t.v1 = b; t.v2 = c; t.n0 = na; t.n1 = nb; t.n2 = nc;
        s->tris.push_back(t); s->mat_ids.push_back(mat);
    };
    const float colors[4][3] = {{0.5f, 0.5f, 0.5f}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    const float spec[4][3] = {{1, 1, 1}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int i = 0; i < 4; ++i) {
        Material m;
        for (int c = 0; c < 3; ++c) { m.diffuse_color[c] = colors[i][c]; m.specular[c] = spec[i][c]; m.emission_color[c] = colors[i][c] * 0.0f; }
        m.roughness = i == 0 ? 0.8f : 0.0f;
        s->mats.push_back(m);
    }
    const float ps = 10.0f;
    const ptb_float4 v0 = f4(-ps, 0, -ps, 1), v1 = f4(-ps, 0, ps, 1), v2 = f4(ps, 0, -ps, 1), v3 = f4(ps, 0, ps, 1), gn = f4(0, 1, 0, 0);
    push(v0, v1, v2, gn, gn, gn, 0);
    push(v2, v1, v3, gn, gn, gn, 0);
    const float centers[3][3] = {{-3, 1, 0}, {0, 1, 0}, {3, 1, 0}};
    const int stacks = 16, slices = 32;
    const float radius = 1.0f;
    for (int sp = 0; sp < 3; ++sp) {
        std::vector<ptb_float4> verts, norms;
        for (int i = 0; i <= stacks; ++i) {
            const float phi = (float)(3.14159265358979323846 * i / stacks);
            const float y = radius * cosf(phi), r = radius * sinf(phi);
            for (int j = 0; j <= slices; ++j) {
                const float theta = (float)(2.0f * 3.14159265358979323846 * j / slices);
                const float x = r * cosf(theta), z = r * sinf(theta);
                verts.push_back(f4(centers[sp][0] + x, centers[sp][1] + y, centers[sp][2] + z, 1.0f));
                const float inv = 1.0f / sqrtf(x * x + y * y + z * z);
                norms.push_back(f4(x * inv, y * inv, z * inv, 0.0f));
            }
        }
        for (int i = 0; i < stacks; ++i)
            for (int j = 0; j < slices; ++j) {
                const int first = i * (slices + 1) + j, second = first + slices + 1;
                push(verts[first], verts[second], verts[first + 1], norms[first], norms[second], norms[first + 1], 1u + (uint32_t)sp);
                push(verts[first + 1], verts[second], verts[second + 1], norms[first + 1], norms[second], norms[second + 1], 1u + (uint32_t)sp);
            }
    }
    s->revision++;
    *out = s;
    return PTB_OK;
}

int ptb_scene_set_materials(ptb_scene* scene, const ptb_HitGroupData* mats, int n) {
    if (!scene || !mats || n <= 0) return fail(PTB_ERR_INVALID, "ptb_scene_set_materials: bad arguments");
    for (uint32_t id : scene->mat_ids) if (id >= (uint32_t)n) return fail(PTB_ERR_INVALID, "ptb_scene_set_materials: a triangle references a material beyond the table");
    std::vector<Material> table((size_t)n);
    for (int i = 0; i < n; ++i) {
        const ptb_HitGroupData& h = mats[i];
        Material& m = table[(size_t)i];
        m.emission_color[0] = h.emission_color.x; m.emission_color[1] = h.emission_color.y; m.emission_color[2] = h.emission_color.z;
        m.diffuse_color[0] = h.diffuse_color.x; m.diffuse_color[1] = h.diffuse_color.y; m.diffuse_color[2] = h.diffuse_color.z;
        m.specular[0] = h.specular.x; m.specular[1] = h.specular.y; m.specular[2] = h.specular.z;
        m.roughness = h.roughness; m.metallic = h.metallic; m.transparent = h.transparent;
        // setMaterialProperty (optixSphere.cu:601) needs both the flag and a non-null pointer
        if (h.has_texture && h.albedo_texture_data) texture_from_float4(m.tex[TEX_ALBEDO], h.albedo_texture_data, h.tex_width, h.tex_height);
        if (h.has_roughness_map && h.roughness_texture_data) texture_from_float4(m.tex[TEX_ROUGHNESS], h.roughness_texture_data, h.roughness_width, h.roughness_height);
        if (h.has_normal_map && h.normal_texture_data) texture_from_float4(m.tex[TEX_NORMAL], h.normal_texture_data, h.normal_width, h.normal_height);
        if (h.has_metallic_map && h.metallic_texture_data) texture_from_float4(m.tex[TEX_METALLIC], h.metallic_texture_data, h.metallic_width, h.metallic_height);
    }
    scene->mats.swap(table);
    scene->revision++;
    return PTB_OK;
}

int ptb_scene_set_env_file(ptb_scene* scene, const char* path) {
    if (!scene || !path) return fail(PTB_ERR_INVALID, "ptb_scene_set_env_file: bad arguments");
    std::string err;
    std::vector<float> px; int w = 0, h = 0;
    if (ends_with_ci(path, ".exr")) {
        if (!load_exr_float4(path, px, w, h, err)) return fail(PTB_ERR_IO, err);
    } else {
        std::vector<uint8_t> b;
        if (!load_png_rgba8(path, b, w, h, err)) return fail(PTB_ERR_IO, err);
        px.resize(b.size());
        for (size_t i = 0; i < b.size(); ++i) px[i] = b[i] / 255.0f;
    }
    scene->env.swap(px); scene->env_w = w; scene->env_h = h; scene->revision++;
    return PTB_OK;
}

int ptb_scene_set_env_pixels(ptb_scene* scene, const float* rgba, int w, int h) {
    if (!scene || !rgba || w <= 0 || h <= 0) return fail(PTB_ERR_INVALID, "ptb_scene_set_env_pixels: bad arguments");
    scene->env.assign(rgba, rgba + (size_t)w * h * 4); scene->env_w = w; scene->env_h = h; scene->revision++;
    return PTB_OK;
}

void ptb_scene_destroy(ptb_scene* scene) {
    if (!scene) return;
    for (DeviceScene* d : scene->devs) free_device_scene(d);
    delete scene;
}

uint32_t ptb_scene_num_triangles(const ptb_scene* scene) { return scene ? (uint32_t)scene->tris.size() : 0; }
int ptb_scene_num_materials(const ptb_scene* scene) { return scene ? (int)scene->mats.size() : 0; }

int ptb_scene_copy_triangles(const ptb_scene* scene, ptb_TriangleData* out, uint32_t cap) {
    if (!scene || !out || cap < scene->tris.size()) return fail(PTB_ERR_INVALID, "ptb_scene_copy_triangles: bad arguments");
    memcpy(out, scene->tris.data(), scene->tris.size() * sizeof(ptb_TriangleData));
    return PTB_OK;
}
int ptb_scene_copy_material_ids(const ptb_scene* scene, uint32_t* out, uint32_t cap) {
    if (!scene || !out || cap < scene->mat_ids.size()) return fail(PTB_ERR_INVALID, "ptb_scene_copy_material_ids: bad arguments");
    memcpy(out, scene->mat_ids.data(), scene->mat_ids.size() * sizeof(uint32_t));
    return PTB_OK;
}
int ptb_scene_get_material(const ptb_scene* scene, int index, ptb_material_info* out) {
    if (!scene || !out || index < 0 || index >= (int)scene->mats.size()) return fail(PTB_ERR_INVALID, "ptb_scene_get_material: bad arguments");
    const Material& m = scene->mats[(size_t)index];
    memset(out, 0, sizeof(*out));
    for (int c = 0; c < 3; ++c) { out->emission_color[c] = m.emission_color[c]; out->diffuse_color[c] = m.diffuse_color[c]; out->specular[c] = m.specular[c]; }
    out->roughness = m.roughness; out->metallic = m.metallic; out->transparent = m.transparent;
    out->has_albedo = m.tex[0].has; out->albedo_w = m.tex[0].w; out->albedo_h = m.tex[0].h;
    out->has_roughness = m.tex[1].has; out->roughness_w = m.tex[1].w; out->roughness_h = m.tex[1].h;
    out->has_normal = m.tex[2].has; out->normal_w = m.tex[2].w; out->normal_h = m.tex[2].h;
    out->has_metallic = m.tex[3].has; out->metallic_w = m.tex[3].w; out->metallic_h = m.tex[3].h;
    return PTB_OK;
}
int ptb_scene_copy_texture(const ptb_scene* scene, int material, int kind, float* out, size_t cap_floats) {
    if (!scene || !out || material < 0 || material >= (int)scene->mats.size() || kind < 0 || kind >= TEX_COUNT)
        return fail(PTB_ERR_INVALID, "ptb_scene_copy_texture: bad arguments");
    const Texture& t = scene->mats[(size_t)material].tex[kind];
    if (!t.has) return fail(PTB_ERR_INVALID, "ptb_scene_copy_texture: material has no such texture");
    const size_t n = (size_t)t.w * t.h * 4;
    if (cap_floats < n) return fail(PTB_ERR_INVALID, "ptb_scene_copy_texture: buffer too small");
    if (t.is_float) memcpy(out, t.rgba32f.data(), n * sizeof(float));
    else for (size_t i = 0; i < n; ++i) out[i] = t.rgba8[i] / 255.0f;  // optixSphere.cpp:370-373
    return PTB_OK;
}
int ptb_scene_env_size(const ptb_scene* scene, int* w, int* h) {
    if (!scene) return fail(PTB_ERR_INVALID, "ptb_scene_env_size: null scene");
    if (w) *w = scene->env_w;
    if (h) *h = scene->env_h;
    return PTB_OK;
}
int ptb_scene_copy_env(const ptb_scene* scene, float* out, size_t cap_floats) {
    if (!scene || !out || cap_floats < scene->env.size()) return fail(PTB_ERR_INVALID, "ptb_scene_copy_env: bad arguments");
    memcpy(out, scene->env.data(), scene->env.size() * sizeof(float));
    return PTB_OK;
}

// sutil::Camera::UVWFrame (OptiX SDK 8.0.0, restated): W is not normalised (it
// carries the focal length), |V| = |W| tan(fovY/2), |U| = |V| * aspect.
void ptb_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy_deg, float aspect,
                    float U[3], float V[3], float W[3]) {
    float w[3] = {lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
    float wlen = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    float u[3] = {w[1] * up[2] - w[2] * up[1], w[2] * up[0] - w[0] * up[2], w[0] * up[1] - w[1] * up[0]};
    float inv = 1.0f / sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    for (int i = 0; i < 3; ++i) u[i] *= inv;
    float v[3] = {u[1] * w[2] - u[2] * w[1], u[2] * w[0] - u[0] * w[2], u[0] * w[1] - u[1] * w[0]};
    inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int i = 0; i < 3; ++i) v[i] *= inv;
    const float vlen = wlen * tanf(0.5f * fovy_deg * 3.14159265358979323846f / 180.0f);
    const float ulen = vlen * aspect;
    for (int i = 0; i < 3; ++i) { U[i] = u[i] * ulen; V[i] = v[i] * vlen; W[i] = w[i]; }
}

// configureCamera (optixSphere.cpp:102-120) + handleCameraUpdate (238-247)
void ptb_params_default_camera(ptb_Params* p) {
    if (!p) return;
    const float eye[3] = {0.0f, 2.0f, 6.0f}, lookat[3] = {0.0f, 0.0f, 0.0f}, up[3] = {0.0f, 1.0f, 0.0f};
    float U[3], V[3], W[3];
    float aspect = (float)p->image_width / (float)p->image_height;
    ptb_camera_uvw(eye, lookat, up, 50.0f, aspect, U, V, W);
    p->eye.x = eye[0]; p->eye.y = eye[1]; p->eye.z = eye[2];
    p->U.x = U[0]; p->U.y = U[1]; p->U.z = U[2];
    p->V.x = V[0]; p->V.y = V[1]; p->V.z = V[2];
    p->W.x = W[0]; p->W.y = W[1]; p->W.z = W[2];
}

int ptb_image_load_rgba8(const char* path, uint8_t** pixels, int* w, int* h) {
    if (!path || !pixels || !w || !h) return fail(PTB_ERR_INVALID, "ptb_image_load_rgba8: bad arguments");
    std::vector<uint8_t> px; std::string err;
    if (!load_png_rgba8(path, px, *w, *h, err)) return fail(PTB_ERR_IO, err);
    *pixels = (uint8_t*)malloc(px.size());
    if (!*pixels) return fail(PTB_ERR_INVALID, "out of memory");
    memcpy(*pixels, px.data(), px.size());
    return PTB_OK;
}
int ptb_image_load_float4(const char* path, float** pixels, int* w, int* h) {
    if (!path || !pixels || !w || !h) return fail(PTB_ERR_INVALID, "ptb_image_load_float4: bad arguments");
    std::vector<float> px; std::string err;
    if (!load_exr_float4(path, px, *w, *h, err)) return fail(PTB_ERR_IO, err);
    *pixels = (float*)malloc(px.size() * sizeof(float));
    if (!*pixels) return fail(PTB_ERR_INVALID, "out of memory");
    memcpy(*pixels, px.data(), px.size() * sizeof(float));
    return PTB_OK;
}
int ptb_save_image(const char* path, const ptb_uchar4* pixels, int w, int h, int flip_y) {
    if (!path || !pixels || w <= 0 || h <= 0) return fail(PTB_ERR_INVALID, "ptb_save_image: bad arguments");
    std::string err; bool ok;
    if (ends_with_ci(path, ".ppm")) ok = save_ppm_rgb8(path, (const uint8_t*)pixels, w, h, flip_y != 0, err);
    else if (ends_with_ci(path, ".png")) ok = save_png_rgba8(path, (const uint8_t*)pixels, w, h, flip_y != 0, err);
    else return fail(PTB_ERR_UNSUPPORTED, "ptb_save_image: only .png and .ppm are supported");
    return ok ? PTB_OK : fail(PTB_ERR_IO, err);
}
// Raw float4 accumulation buffer for parity tools (last row): 16-byte header
// {"PTBA", version 1, width, height} (little-endian uint32) followed by width*height float4 texels, row 0 first
// (the bottom image row, optixSphere.cu:332,400), exactly as the buffer lies in memory.
int ptb_save_accum_raw(const char* path, const ptb_float4* accum, int w, int h) {
    if (!path || !accum || w <= 0 || h <= 0) return fail(PTB_ERR_INVALID, "ptb_save_accum_raw: bad arguments");
    FILE* f = fopen(path, "wb");
    if (!f) return fail(PTB_ERR_IO, std::string("ptb_save_accum_raw: cannot open ") + path);
    const uint32_t hdr[4] = {0x41425450u /* "PTBA" */, 1u, (uint32_t)w, (uint32_t)h};
    const size_t n = (size_t)w * (size_t)h;
    const bool ok = fwrite(hdr, sizeof(hdr), 1, f) == 1 && fwrite(accum, sizeof(ptb_float4), n, f) == n;
    if (fclose(f) != 0 || !ok) return fail(PTB_ERR_IO, std::string("ptb_save_accum_raw: write failed: ") + path);
    return PTB_OK;
}
int ptb_load_accum_raw(const char* path, float** accum, int* w, int* h) {
    if (!path || !accum || !w || !h) return fail(PTB_ERR_INVALID, "ptb_load_accum_raw: bad arguments");
    FILE* f = fopen(path, "rb");
    if (!f) return fail(PTB_ERR_IO, std::string("ptb_load_accum_raw: cannot open ") + path);
    uint32_t hdr[4] = {0, 0, 0, 0};
    if (fread(hdr, sizeof(hdr), 1, f) != 1 || hdr[0] != 0x41425450u || hdr[1] != 1u || hdr[2] == 0 || hdr[3] == 0 ||
        (uint64_t)hdr[2] * hdr[3] > 0x7fffffffull) {
        fclose(f);
        return fail(PTB_ERR_IO, std::string("ptb_load_accum_raw: not a PTBA version 1 file: ") + path);
    }
    const size_t n = (size_t)hdr[2] * hdr[3];
    float* px = (float*)malloc(n * 16);
    if (!px) { fclose(f); return fail(PTB_ERR_INVALID, "out of memory"); }
    const bool ok = fread(px, 16, n, f) == n;
    fclose(f);
    if (!ok) { free(px); return fail(PTB_ERR_IO, std::string("ptb_load_accum_raw: truncated file: ") + path); }
    *accum = px; *w = (int)hdr[2]; *h = (int)hdr[3];
    return PTB_OK;
}
void ptb_free(void* p) { free(p); }

int ptb_obj_read(const char* path, void** records, uint64_t* n_face_vertices) {
    if (!path || !records || !n_face_vertices) return fail(PTB_ERR_INVALID, "ptb_obj_read: bad arguments");
    ObjMesh mesh; std::string err;
    if (!load_obj(path, mesh, err)) return fail(PTB_ERR_IO, err);
    const size_t n = mesh.indices.size();
    uint32_t* out = (uint32_t*)calloc(n ? n : 1, 40);
    if (!out) return fail(PTB_ERR_INVALID, "out of memory");
    for (size_t i = 0; i < n; ++i) {
        const ObjIndex& ix = mesh.indices[i];
        float rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        int32_t flags[2] = {0, 0};
        for (int k = 0; k < 3; ++k) rec[k] = mesh.v[3 * (size_t)ix.v + k];
        if (ix.vn >= 0) { flags[0] = 1; for (int k = 0; k < 3; ++k) rec[3 + k] = mesh.vn[3 * (size_t)ix.vn + k]; }
        if (!mesh.vt.empty() && ix.vt >= 0) { flags[1] = 1; rec[6] = mesh.vt[2 * (size_t)ix.vt]; rec[7] = mesh.vt[2 * (size_t)ix.vt + 1]; }
        memcpy(out + i * 10, rec, 32);
        memcpy(out + i * 10 + 8, flags, 8);
    }
    *records = out; *n_face_vertices = n;
    return PTB_OK;
}

}  // extern "C"
