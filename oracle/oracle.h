/* oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C interface of the CPU oracle (liboracle.so, built from oracle_pt.cpp) and of
 * the host-compiled reference (oracle/_ref/libref_pt.so, built from
 * /root/reference/optixSphere.cu through ref_shim/).  Both libraries export the
 * same orc_* entry points over the same structs so tests can run one against
 * the other.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load these libraries.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One texture = float4 texels, as the reference uploads them
 * (optixSphere.cpp:364-380: byte/255.0f). */
typedef struct OrcTexture {
    const float* rgba; /* w*h*4 floats, row 0 first */
    int32_t w, h;
    int32_t has;
    int32_t _pad;
} OrcTexture;

/* Mirrors HitGroupData (optixSphere.h:67-102) minus the geometry pointers. */
typedef struct OrcMaterial {
    OrcTexture albedo, roughness, normal, metallic;
    float emission_color[3];
    float diffuse_color[3];
    float specular[3];
    float roughness_value;
    int32_t metallic_flag;
    int32_t transparent_flag;
} OrcMaterial;

typedef struct OrcScene {
    const float* vertices;  /* float4[3*N]  (optixSphere.cpp:845-858) */
    const float* normals;   /* float4[3*N] */
    const float* texcoords; /* float2[3*N] */
    const uint32_t* mat_ids; /* N */
    uint32_t num_tris;
    int32_t num_mats;
    const OrcMaterial* mats;
    const float* env_rgba;  /* float4[env_w*env_h] (MissData, optixSphere.h:58-63) */
    int32_t env_w, env_h;
} OrcScene;

/* Mirrors the fields of Params (optixSphere.h:10-31) the integrator reads. */
typedef struct OrcParams {
    uint32_t width, height;
    int32_t subframe_index;
    int32_t dof;
    float eye[3], U[3], V[3], W[3];
} OrcParams;

/* The reference's compile-time literals, made explicit. */
typedef struct OrcConfig {
    int32_t spp_per_launch; /* optixSphere.cu:323  (10) */
    int32_t max_depth;      /* optixSphere.cu:360  (20) */
    float tmin, tmax;       /* optixSphere.cu:368-369 (0.01, 1e16) */
    float dof_blur;         /* optixSphere.cu:285  (0.01) */
    float focus_dist;       /* optixSphere.cu:329  (1.0) */
    float nmap_strength;    /* optixSphere.cu:697  (0.4) */
    float exposure;         /* optixSphere.cu:412  (-0.5) */
    float gamma;            /* optixSphere.cu:425  (2.2) */
    float contrast;         /* optixSphere.cu:432  (1.25) */
    int32_t sat_cuda;       /* 1: float->uint saturates like CUDA (oracle rule); 0: x86 wrap */
    int32_t use_bvh;        /* 0: brute force closest hit; 1: oracle's own CPU BVH */
    int32_t threads;        /* OpenMP threads, 0 = all */
    int32_t accumulate_sum; /* 0: reference running average (cu:403-409); 1: accum += launch mean */
} OrcConfig;

typedef struct OrcStats {
    uint64_t segments; /* traceRadiance calls (optixSphere.cu:364) */
    uint64_t paths;
    uint64_t hits;
    uint64_t misses;
    double seconds;    /* wall time of the render region */
    int32_t threads;
    int32_t _pad;
} OrcStats;

void orc_default_config(OrcConfig* cfg);

/* Renders pixels [x0,x1) x [y0,y1) of the frame described by params (the
 * rest of accum/frame/primary_hit is left untouched).  accum: float4[W*H]
 * in/out; frame: uchar4[W*H] out; primary_hit: int32[W*H] out or NULL
 * (-1 = miss), the primitive hit by the FIRST segment of sample 0.
 * Returns 0 on success. */
int orc_render(const OrcScene* scene, const OrcParams* params, const OrcConfig* cfg,
               float* accum, uint8_t* frame, int32_t* primary_hit, OrcStats* stats,
               int32_t x0, int32_t y0, int32_t x1, int32_t y1);

/* Closest hit of one ray (brute force or BVH per use_bvh); returns prim or -1. */
int32_t orc_closest_hit(const OrcScene* scene, const float org[3], const float dir[3],
                        float tmin, float tmax, int32_t use_bvh, float* t, float* b1, float* b2);

/* Unit-level entry points for known-answer tests. */
uint32_t orc_rng_next(uint32_t seed, int32_t sat_cuda, float* u); /* returns new state */
void orc_sincos(float x, float* s, float* c);
float orc_atan2(float y, float x);
float orc_asin(float x);
float orc_pow(float x, float y); /* detmath pow of the display transform */
void orc_tonemap_pixel(const float accum_rgb[3], const OrcConfig* cfg, uint8_t out_rgba[4]);
void orc_sample_texture(const OrcTexture* tex, float u, float v, float out[4]);
void orc_sample_env(const float* env_rgba, int32_t w, int32_t h, const float dir[3], float out[4]);
void orc_camera_uvw(const float eye[3], const float lookat[3], const float up[3], float fovy_deg,
                    float aspect, float U[3], float V[3], float W[3]);
const char* orc_impl_name(void); /* "oracle-restatement" or "reference-host-shim" */

#ifdef __cplusplus
}
#endif
#endif
