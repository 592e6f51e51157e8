// host.h -- host-side (no CUDA) declarations of the ptb library: OBJ reader,
// image files, scene build.  Everything here runs without a GPU.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "../../include/ptb.h"

namespace ptb {

// thread-local error string behind ptb_last_error()
void set_error(const std::string& msg);
const char* get_error();

// ---- obj_loader.cpp ---------------------------------------------------------
struct ObjIndex { int v, vt, vn; };
struct ObjMesh {
    std::vector<float> v, vn, vt;       // 3, 3, 2 floats per element
    std::vector<ObjIndex> indices;      // 3 per triangle, fan-triangulated, file order
};
bool load_obj(const std::string& path, ObjMesh& mesh, std::string& err);

// ---- image_io.cpp -----------------------------------------------------------
bool file_exists(const std::string& path);
// PNG -> RGBA8 (stb_image STBI_rgb_alpha semantics), row 0 = top row of the file
bool load_png_rgba8(const std::string& path, std::vector<uint8_t>& rgba, int& w, int& h, std::string& err);
// EXR (scanline; NONE/RLE/ZIPS/ZIP; HALF/FLOAT/UINT) -> float4, row 0 = top, missing A = 1
bool load_exr_float4(const std::string& path, std::vector<float>& rgba, int& w, int& h, std::string& err);
bool save_png_rgba8(const std::string& path, const uint8_t* rgba, int w, int h, bool flip_y, std::string& err);
bool save_ppm_rgb8(const std::string& path, const uint8_t* rgba, int w, int h, bool flip_y, std::string& err);

// ---- scene.cpp --------------------------------------------------------------
enum TexKind { TEX_ALBEDO = 0, TEX_ROUGHNESS = 1, TEX_NORMAL = 2, TEX_METALLIC = 3, TEX_COUNT = 4 };

// A texture keeps the 8-bit source when every texel is byte/255.0f (the
// reference inflates those to float4 on the host, optixSphere.cpp:364-380; the
// kernels redo the same division, so values are bit-identical at a quarter of
// the footprint).  Anything else stays float4.
struct Texture {
    bool has = false;
    int w = 0, h = 0;
    bool is_float = false;
    std::vector<uint8_t> rgba8;
    std::vector<float> rgba32f;
};

// One entry of the hit-group table (HitGroupData, optixSphere.h:67-102).
struct Material {
    float emission_color[3] = {0, 0, 0};
    float diffuse_color[3] = {0, 0, 0};
    float specular[3] = {0, 0, 0};
    float roughness = 0.0f;
    bool metallic = false;
    bool transparent = false;
    Texture tex[TEX_COUNT];
};

struct DeviceScene;  // renderer.cu

}  // namespace ptb

struct ptb_scene {
    std::vector<ptb_TriangleData> tris;
    std::vector<uint32_t> mat_ids;
    std::vector<ptb::Material> mats;
    std::vector<float> env;  // float4 texels
    int env_w = 0, env_h = 0;
    std::vector<ptb::DeviceScene*> devs;  // owned; one upload per context the scene was built on (ptb_accel_build)
    uint64_t revision = 0;            // bumped by every host-side mutation
};

namespace ptb {
// createSceneGeometry(loadFromFile=true) restated (optixSphere.cpp:400-649).
bool build_scene_from_obj(const std::vector<std::string>& files, float scale, uint32_t material_seed, ptb_scene& scene,
                          std::string& err);
void free_device_scene(DeviceScene* d);  // renderer.cu
}  // namespace ptb
