// oracle/ref_shim/ref_probe.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Two reference-derived pins, compiled against the reference's own headers
// from where they lie (oracle/Makefile, target _ref/ref_probe):
//   ref_probe layout            prints sizeof/offsetof of every struct in
//                               optixSphere.h, using the real CUDA vector types
//   ref_probe obj <file.obj> <out.bin>
//                               loads the OBJ with the reference's vendored
//                               tiny_obj_loader.h (triangulate=true, as
//                               optixSphere.cpp:431 calls it) and dumps, per
//                               face vertex in shape/face order, 8 float32:
//                               vx vy vz nx ny nz tx ty followed by 2 int32
//                               flags (has_normal, has_texcoord).
#include <vector_types.h>
typedef unsigned long long OptixTraversableHandle;
#include "optixSphere.h"

#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"

#include <cstddef>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#define OFF(S, f) printf("  \"%s.%s\": %zu,\n", #S, #f, offsetof(S, f))

static int layout() {
    printf("{\n");
    printf("  \"sizeof.TriangleData\": %zu,\n", sizeof(TriangleData));
    printf("  \"alignof.TriangleData\": %zu,\n", alignof(TriangleData));
    OFF(TriangleData, v0); OFF(TriangleData, v1); OFF(TriangleData, v2);
    OFF(TriangleData, n0); OFF(TriangleData, n1); OFF(TriangleData, n2);
    OFF(TriangleData, uv0); OFF(TriangleData, uv1); OFF(TriangleData, uv2);
    printf("  \"sizeof.Params\": %zu,\n", sizeof(Params));
    OFF(Params, image_width); OFF(Params, image_height); OFF(Params, origin_x); OFF(Params, origin_y);
    OFF(Params, subframe_index); OFF(Params, frame_buffer); OFF(Params, accum_buffer); OFF(Params, dof);
    OFF(Params, eye); OFF(Params, U); OFF(Params, V); OFF(Params, W);
    OFF(Params, triangles); OFF(Params, num_triangles); OFF(Params, handle);
    printf("  \"sizeof.Payload\": %zu,\n", sizeof(Payload));
    printf("  \"sizeof.RayGenData\": %zu,\n", sizeof(RayGenData));
    printf("  \"sizeof.MissData\": %zu,\n", sizeof(MissData));
    OFF(MissData, hdr_image_data); OFF(MissData, width); OFF(MissData, height);
    printf("  \"sizeof.HitGroupData\": %zu,\n", sizeof(HitGroupData));
    OFF(HitGroupData, albedo_texture_data); OFF(HitGroupData, tex_width); OFF(HitGroupData, tex_height);
    OFF(HitGroupData, has_texture);
    OFF(HitGroupData, roughness_texture_data); OFF(HitGroupData, roughness_width); OFF(HitGroupData, roughness_height);
    OFF(HitGroupData, has_roughness_map);
    OFF(HitGroupData, normal_texture_data); OFF(HitGroupData, normal_width); OFF(HitGroupData, normal_height);
    OFF(HitGroupData, has_normal_map);
    OFF(HitGroupData, metallic_texture_data); OFF(HitGroupData, metallic_width); OFF(HitGroupData, metallic_height);
    OFF(HitGroupData, has_metallic_map);
    OFF(HitGroupData, texcoords); OFF(HitGroupData, vertices); OFF(HitGroupData, normals);
    OFF(HitGroupData, emission_color); OFF(HitGroupData, diffuse_color); OFF(HitGroupData, specular);
    OFF(HitGroupData, roughness); OFF(HitGroupData, metallic); OFF(HitGroupData, transparent);
    printf("  \"end\": 0\n}\n");
    return 0;
}

static int dump_obj(const char* path, const char* out) {
    tinyobj::attrib_t attrib;
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    std::string err;
    bool ret = tinyobj::LoadObj(&attrib, &shapes, &materials, &err, path, "");
    if (!ret) { fprintf(stderr, "LoadObj failed: %s\n", err.c_str()); return 1; }
    FILE* f = fopen(out, "wb");
    if (!f) return 1;
    unsigned long long nfv = 0;
    for (const auto& shape : shapes) {
        size_t index_offset = 0;
        for (size_t fi = 0; fi < shape.mesh.num_face_vertices.size(); fi++) {
            int fv = shape.mesh.num_face_vertices[fi];
            if (fv != 3) { index_offset += fv; continue; }
            for (int v = 0; v < 3; v++) {
                tinyobj::index_t idx = shape.mesh.indices[index_offset + v];
                float rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                int flags[2] = {0, 0};
                rec[0] = attrib.vertices[3 * idx.vertex_index + 0];
                rec[1] = attrib.vertices[3 * idx.vertex_index + 1];
                rec[2] = attrib.vertices[3 * idx.vertex_index + 2];
                if (idx.normal_index >= 0) {
                    flags[0] = 1;
                    rec[3] = attrib.normals[3 * idx.normal_index + 0];
                    rec[4] = attrib.normals[3 * idx.normal_index + 1];
                    rec[5] = attrib.normals[3 * idx.normal_index + 2];
                }
                if (!attrib.texcoords.empty() && idx.texcoord_index >= 0) {
                    flags[1] = 1;
                    rec[6] = attrib.texcoords[2 * idx.texcoord_index + 0];
                    rec[7] = attrib.texcoords[2 * idx.texcoord_index + 1];
                }
                fwrite(rec, sizeof(rec), 1, f);
                fwrite(flags, sizeof(flags), 1, f);
                nfv++;
            }
            index_offset += fv;
        }
    }
    fclose(f);
    printf("{\"face_vertices\": %llu, \"triangles\": %llu, \"shapes\": %zu}\n", nfv, nfv / 3, shapes.size());
    return 0;
}

int main(int argc, char** argv) {
    if (argc >= 2 && !strcmp(argv[1], "layout")) return layout();
    if (argc >= 4 && !strcmp(argv[1], "obj")) return dump_obj(argv[2], argv[3]);
    fprintf(stderr, "usage: ref_probe layout | ref_probe obj <file.obj> <out.bin>\n");
    return 2;
}
