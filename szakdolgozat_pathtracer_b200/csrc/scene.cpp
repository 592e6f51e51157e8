// scene.cpp -- host scene build: createSceneGeometry(loadFromFile=true) and the
// hit-group/material table of the reference (optixSphere.cpp:69-90, 355-392,
// 400-649, 1196-1261), restated behind the C ABI.  No CUDA in this file.
//
// Deliberate, documented differences from the reference:
//  * the random-material RNG is seeded explicitly (material_seed) instead of
//    std::random_device (optixSphere.cpp:141-148), and a draw is mapped to
//    [0,1) as (mt19937() >> 8) * 2^-24 rather than through the
//    implementation-defined std::uniform_real_distribution<float>;
//  * every material owns its textures (the reference keeps ONE global device
//    pointer per texture kind, so the last loaded file wins: optixSphere.cpp:395-398);
//  * has_* flags default to false (uninitialised in the reference for
//    untextured files: optixSphere.cpp:518 vs 572-582);
//  * a face vertex without a texcoord index gets uv = (0,0) (the reference
//    indexes attrib.texcoords[-2], optixSphere.cpp:485-489);
//  * the floor triangles get uv = (0,0) (left uninitialised at optixSphere.cpp:620-646).
#include <atomic>
#include <cmath>
#include <cstring>
#include <random>
#include <thread>

#include "host.h"

namespace ptb {

namespace {

thread_local std::string g_error;

struct HostRng {
    std::mt19937 gen;
    explicit HostRng(uint32_t seed) : gen(seed) {}
    float next() { return (float)(gen() >> 8) * (1.0f / 16777216.0f); }  // rnd_f(), optixSphere.cpp:145-148
};

inline ptb_float4 f4(float x, float y, float z, float w) { ptb_float4 r; r.x = x; r.y = y; r.z = z; r.w = w; return r; }

// setUpImageTexture (optixSphere.cpp:355-392) minus the upload: the 8-bit
// texels are kept; the byte/255.0f widening happens where they are read.
void load_texture_if_present(Texture& t, const std::string& filename) {
    t = Texture();
    if (!file_exists(filename)) return;
    std::string err;
    int w = 0, h = 0;
    if (!load_png_rgba8(filename, t.rgba8, w, h, err)) {
        // sutil::loadImage would throw; the reference has no handler for a
        // broken texture either.  Treat it as absent but remember why.
        set_error("texture " + filename + ": " + err);
        t = Texture();
        return;
    }
    t.has = true; t.w = w; t.h = h; t.is_float = false;
}

}  // namespace

void set_error(const std::string& msg) { g_error = msg; }
const char* get_error() { return g_error.c_str(); }

bool build_scene_from_obj(const std::vector<std::string>& files, float scale, uint32_t material_seed, ptb_scene& scene,
                          std::string& err) {
    scene.tris.clear(); scene.mat_ids.clear(); scene.mats.clear();
    HostRng rng(material_seed);
    float minHeight = 10.0f;  // optixSphere.cpp:418

    // Parse and flatten the files concurrently (one task per file, as many threads as the host offers): the result of a
    // file does not depend on the others, and everything order-dependent -- triangle order, material indices, the
    // material RNG, which error is reported -- is settled afterwards in file order.
    struct PerFile { bool ok = false; std::string err; std::vector<ptb_TriangleData> tris; float min_y = 3.4e38f; };
    std::vector<PerFile> per(files.size());
    {
        std::atomic<size_t> next(0);
        auto work = [&]() {
            for (size_t i = next++; i < files.size(); i = next++) {
                PerFile& pf = per[i];
                ObjMesh mesh;
                pf.ok = load_obj(files[i], mesh, pf.err);
                if (!pf.ok) continue;
                const bool have_vt = !mesh.vt.empty();
                pf.tris.reserve(mesh.indices.size() / 3);
                for (size_t t = 0; t + 2 < mesh.indices.size(); t += 3) {
                    ptb_float4 vertices[3], normals[3];
                    ptb_float2 uv[3];
                    for (int v = 0; v < 3; ++v) {
                        const ObjIndex& idx = mesh.indices[t + (size_t)v];
                        float vx = mesh.v[3 * (size_t)idx.v + 0] * scale;
                        float vy = mesh.v[3 * (size_t)idx.v + 1] * scale;
                        float vz = mesh.v[3 * (size_t)idx.v + 2] * scale;
                        vertices[v] = f4(vx, vy, vz, 0.0f);
                        if (idx.vn >= 0) {
                            float nx = mesh.vn[3 * (size_t)idx.vn + 0], ny = mesh.vn[3 * (size_t)idx.vn + 1], nz = mesh.vn[3 * (size_t)idx.vn + 2];
                            float inv = 1.0f / sqrtf(nx * nx + ny * ny + nz * nz);  // sutil normalize()
                            normals[v] = f4(nx * inv, ny * inv, nz * inv, 0.0f);
                        } else {
                            normals[v] = f4(0.0f, 1.0f, 0.0f, 0.0f);  // optixSphere.cpp:480-482
                        }
                        if (have_vt && idx.vt >= 0) { uv[v].x = mesh.vt[2 * (size_t)idx.vt + 0]; uv[v].y = mesh.vt[2 * (size_t)idx.vt + 1]; }
                        else { uv[v].x = 0.0f; uv[v].y = 0.0f; }
                    }
                    for (int v = 0; v < 3; ++v) if (vertices[v].y < pf.min_y) pf.min_y = vertices[v].y;  // optixSphere.cpp:497-499
                    ptb_TriangleData tri;
                    memset(&tri, 0, sizeof(tri));
                    tri.v0 = vertices[0]; tri.v1 = vertices[1]; tri.v2 = vertices[2];
                    tri.n0 = normals[0]; tri.n1 = normals[1]; tri.n2 = normals[2];
                    tri.uv0 = uv[0]; tri.uv1 = uv[1]; tri.uv2 = uv[2];
                    pf.tris.push_back(tri);
                }
            }
        };
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0) nthreads = 1;
        if (nthreads > files.size()) nthreads = (unsigned)files.size();
        std::vector<std::thread> pool;
        for (unsigned k = 1; k < nthreads; ++k) pool.emplace_back(work);
        work();
        for (std::thread& th : pool) th.join();
    }
    for (size_t i = 0; i < files.size(); ++i) {
        if (!per[i].ok) { err = "Failed to load/parse .obj file: " + per[i].err; return false; }  // optixSphere.cpp:441-443
        const size_t startIndex = scene.tris.size();
        scene.tris.insert(scene.tris.end(), per[i].tris.begin(), per[i].tris.end());
        std::vector<ptb_TriangleData>().swap(per[i].tris);
        if (per[i].min_y < minHeight) minHeight = per[i].min_y;

        // optixSphere.cpp:516-582: textures by file-name convention, else a random material.
        Material m;
        const std::string stem = files[i].substr(0, files[i].find_last_of('.'));
        load_texture_if_present(m.tex[TEX_ALBEDO], stem + "_albedo.png");
        load_texture_if_present(m.tex[TEX_ROUGHNESS], stem + "_roughness.png");
        load_texture_if_present(m.tex[TEX_NORMAL], stem + "_normal.png");
        load_texture_if_present(m.tex[TEX_METALLIC], stem + "_metallic.png");
        // The reference draws (color, decider) before it looks at the flags and
        // again inside the untextured branch; keep the same number of draws.
        rng.next(); rng.next(); rng.next(); rng.next();
        float color[3], specular[3], emission, roughness;
        bool metallic;
        if (m.tex[0].has || m.tex[1].has || m.tex[2].has || m.tex[3].has) {
            color[0] = color[1] = color[2] = 0.5f; specular[0] = specular[1] = specular[2] = 0.5f;
            emission = 0.0f; roughness = 0.4f; metallic = false;
        } else {
            color[0] = rng.next(); color[1] = rng.next(); color[2] = rng.next();
            float decider = rng.next();
            specular[0] = color[0]; specular[1] = color[1]; specular[2] = color[2];
            emission = decider < 0.1f ? 100.0f : 0.0f;
            roughness = rng.next();
            metallic = decider > 0.5f && decider < 0.65f;
        }
        for (int c = 0; c < 3; ++c) {
            m.emission_color[c] = color[c] * emission;  // optixSphere.cpp:1209
            m.diffuse_color[c] = color[c];
            m.specular[c] = specular[c];
        }
        m.roughness = roughness; m.metallic = metallic; m.transparent = false;
        const uint32_t matIndex = (uint32_t)scene.mats.size();
        scene.mats.push_back(std::move(m));
        for (size_t t = startIndex; t < scene.tris.size(); ++t) scene.mat_ids.push_back(matIndex);
    }

    // Floor (optixSphere.cpp:598-646): color .2, specular .2, emission 0, roughness .1.
    Material floor;
    for (int c = 0; c < 3; ++c) { floor.emission_color[c] = 0.2f * 0.0f; floor.diffuse_color[c] = 0.2f; floor.specular[c] = 0.2f; }
    floor.roughness = 0.1f; floor.metallic = false; floor.transparent = false;
    const uint32_t floorIndex = (uint32_t)scene.mats.size();
    scene.mats.push_back(floor);
    const float floor_y = minHeight, floor_size = 200.0f;
    const ptb_float4 fv0 = f4(-floor_size, floor_y, -floor_size, 0.0f), fv1 = f4(-floor_size, floor_y, floor_size, 0.0f);
    const ptb_float4 fv2 = f4(floor_size, floor_y, -floor_size, 0.0f), fv3 = f4(floor_size, floor_y, floor_size, 0.0f);
    const ptb_float4 fn = f4(0.0f, 1.0f, 0.0f, 0.0f);
    ptb_TriangleData t1, t2;
    memset(&t1, 0, sizeof(t1)); memset(&t2, 0, sizeof(t2));
    t1.v0 = fv0; t1.v1 = fv1; t1.v2 = fv2; t1.n0 = t1.n1 = t1.n2 = fn;
    t2.v0 = fv2; t2.v1 = fv1; t2.v2 = fv3; t2.n0 = t2.n1 = t2.n2 = fn;
    scene.tris.push_back(t1); scene.mat_ids.push_back(floorIndex);
    scene.tris.push_back(t2); scene.mat_ids.push_back(floorIndex);
    scene.revision++;
    return true;
}

}  // namespace ptb
