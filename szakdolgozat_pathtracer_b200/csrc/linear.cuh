// linear.cuh -- the LINEAR estimator with next-event estimation of the environment (ptb_render_cfg.
// env_importance_sampling = 1) and its BSDF-sampling-only twin (= 2, the A/B baseline the tests compare it with).
//
// This is an OPTIONAL MODE BEYOND THE REFERENCE: a standard unidirectional path tracer over
// the reference's material model, in which a sample is the plain sum of throughput-weighted emissions, so the
// environment can be importance-sampled (env_cdf.cuh) with shadow rays and multiple importance sampling.  It shares
// the RNG, camera, textures, traversal and accumulate/tonemap stages with the reference integrator, but it is NOT
// parity-checked against the oracle: the reference's estimator divides each sample by max(attenuation) at
// termination (optixSphere.cu:382-387) and cannot be reproduced by any light-sampling scheme.
//
//   BSDF      f = (1 - m) albedo / pi  +  F D G / (4 |n.wo| |n.wi|)      (D, G, F0 as optixSphere.cu:439-492, 759-761;
//                                                                          F = Schlick(F0, wo.h))
//   sampling  with probability ps = m + (1 - m) Schlick(n.wo, 1.5) (cu:777-780) a GGX half vector, else cosine
//   pdf       ps D (n.h) / (4 wo.h) + (1 - ps) (n.wi) / pi
//   NEE       one environment direction per hit from the CDF, shadow ray, power heuristic against the BSDF pdf
//   miss      environment radiance weighted by the power heuristic against the CDF pdf (weight 1 for camera rays)
//   RR        from the third segment on: survive with q = min(1, max(throughput)), throughput /= q
// Because the estimator is linear, every contribution is added straight into the slot's pixel sum.
#pragma once
#include "chunked.cuh"
#include "env_cdf.cuh"

namespace PTB_NS {

struct LinearView {
    EnvCdf cdf;
    float4* shadow_o;              // origin of the pending shadow ray (the hit point)
    float4* shadow_d;              // direction of the pending shadow ray
    float4* shadow_c;              // its contribution if unoccluded
    unsigned char* shadow_flag;    // 1: this slot has a pending shadow ray
    int nee;                       // 1: next-event estimation + MIS; 0: BSDF sampling only
};

struct SurfacePoint { float3 n, albedo; float roughness, metallic; float3 emission; float3 pos; bool valid; };

// geometry + material evaluation shared with the reference closest hit (cu:631-714), without the RNG side effects
PTB_DEV SurfacePoint surface_at(const SceneView& s, const FrameView& f, int prim_idx, float b1, float b2, float t_hit, float3 ro, float3 rd) {
    SurfacePoint sp;
    const DevMaterial& m = s.mats[__ldg(s.mat_ids + prim_idx)];
    const size_t vo = (size_t)prim_idx * 3;
    const float3 v0 = mk3(__ldg(s.verts + vo)), v1 = mk3(__ldg(s.verts + vo + 1)), v2 = mk3(__ldg(s.verts + vo + 2));
    float3 flat = normalize(cross(v1 - v0, v2 - v0));
    flat = faceforward(flat, -rd, flat);
    const float3 n0 = mk3(__ldg(s.normals + vo)), n1 = mk3(__ldg(s.normals + vo + 1)), n2 = mk3(__ldg(s.normals + vo + 2));
    const float ba = 1.0f - b1 - b2;
    const float2 uv0 = __ldg(s.uvs + vo), uv1 = __ldg(s.uvs + vo + 1), uv2 = __ldg(s.uvs + vo + 2);
    const float uvx = uv0.x * ba + uv1.x * b1 + uv2.x * b2;
    const float uvy = 1.0f - (uv0.y * ba + uv1.y * b1 + uv2.y * b2);
    float3 n = ba * n0 + b1 * n1 + b2 * n2;
    sp.valid = length(n) > 0.01f;
    n = sp.valid ? normalize(n) : flat;
    if (dot(n, rd) > 0.0f) n = flat;
    sp.albedo = material_property(m.tex[0], mk3(m.diffuse[0], m.diffuse[1], m.diffuse[2]), uvx, uvy);
    if (m.tex[2].fmt != 0) {
        float3 nm = normalize(2.0f * material_property(m.tex[2], mk3(0.0f, 1.0f, 0.0f), uvx, uvy) - mk3(1.0f));
        nm = mk3(nm.x, nm.z, nm.y);
        const Onb o(n);
        n = normalize(f.nmap_strength * o.inverse_transform(nm) + (1.0f - f.nmap_strength) * n);
    }
    sp.n = n;
    sp.roughness = clampf(material_property(m.tex[1], mk3(m.roughness), uvx, uvy).x, 0.015f, 0.999f);
    sp.metallic = material_property(m.tex[3], m.metallic ? mk3(1.0f) : mk3(0.0f), uvx, uvy).x;
    sp.emission = mk3(m.emission[0], m.emission[1], m.emission[2]);
    sp.pos = ro + t_hit * rd;
    return sp;
}

PTB_DEV float spec_probability(const SurfacePoint& sp, float3 wo) {
    return sp.metallic + (1.0f - sp.metallic) * Fresnel_Schlick_float(fmaxf(dot(sp.n, wo), 0.0f), 1.5f);
}

// f(wo, wi) * |n.wi| and the sampling pdf of wi
PTB_DEV float3 bsdf_cos(const SurfacePoint& sp, float3 wo, float3 wi, float ps, float* pdf) {
    const float nl = dot(sp.n, wi), nv = dot(sp.n, wo);
    *pdf = 0.0f;
    if (!(nl > 0.0f) || !(nv > 0.0f)) return mk3(0.0f);
    const float alpha = sp.roughness * sp.roughness;
    const float3 h = normalize(wo + wi);
    const float nh = fmaxf(dot(sp.n, h), 1e-10f), vh = fmaxf(dot(wo, h), 1e-10f);
    const float D = D_GGX(sp.n, h, alpha);
    const float G = G_SchlickGGX(alpha, sp.n, wo) * G_SchlickGGX(alpha, sp.n, wi);
    float3 F0 = lerp(mk3(0.04f), sp.albedo, sp.metallic);
    const float3 F = F0 + (mk3(1.0f) - F0) * det_pow5(1.0f - clampf(vh, 0.0f, 1.0f));
    const float3 spec = F * (D * G / (4.0f * nv * nl));
    const float3 diff = sp.albedo * ((1.0f - sp.metallic) * 0.31830988618379067154f);
    *pdf = ps * D * nh / (4.0f * vh) + (1.0f - ps) * nl * 0.31830988618379067154f;
    return (spec + diff) * nl;
}

PTB_DEV float power_heuristic(float a, float b) { const float a2 = a * a, b2 = b * b; return a2 + b2 > 0.0f ? a2 / (a2 + b2) : 0.0f; }

// misc.w holds the pdf of the BSDF sample that produced the current ray (bits), or -1 for camera rays.
// Ends the current sample of `slot` (its contributions are already in pixsum) and starts the next one if any.
PTB_DEV bool linear_next_sample(const FrameView& f, const PathView& p, uint32_t slot, uint32_t seed_rg, uint32_t sample) {
    sample += 1u;
    if (sample >= (uint32_t)f.spp) return false;
    const uint32_t pix = slot % f.n_pixels;
    float3 o, d;
    start_sample(f, pix % f.W, image_row(f, pix / f.W), seed_rg, o, d);
    p.ray_o[slot] = make_float4(o.x, o.y, o.z, 0.0f);
    p.ray_d[slot] = make_float4(d.x, d.y, d.z, 0.0f);
    p.atten_seed[slot] = make_float4(1.0f, 1.0f, 1.0f, __uint_as_float(seed_rg));
    p.misc[slot] = make_uint4(seed_rg, (uint32_t)f.max_depth, sample, __float_as_uint(-1.0f));
    return true;
}

PTB_DEV void add_to_pixel(const PathView& p, uint32_t slot, float3 c) {
    if (!(c.x == c.x) || !(c.y == c.y) || !(c.z == c.z)) return;  // drop NaN contributions
    float4 sum = p.pixsum[slot];
    sum.x += c.x; sum.y += c.y; sum.z += c.z;
    p.pixsum[slot] = sum;
}

PTB_DEV void chunk_stage_shade_linear(ChunkShared& sh, const SceneView& s, const FrameView& f, const PathView& p, const LinearView& lv,
                                      unsigned char* __restrict__ status, uint32_t base, unsigned int n) {
    for (unsigned int i = threadIdx.x; i < n; i += PTB_CHUNK_THREADS) {
        const uint32_t slot = base + sh.list[i];
        const float4 o4 = p.ray_o[slot], d4 = p.ray_d[slot], h4 = p.hit[slot], as = p.atten_seed[slot];
        const uint4 mi = p.misc[slot];
        float3 atten = mk3(as);
        uint32_t seed = __float_as_uint(as.w);
        const int depth = (int)mi.y;
        const float3 rd = mk3(d4), wo = -rd;
        const SurfacePoint sp = surface_at(s, f, __float_as_int(h4.w), h4.y, h4.z, h4.x, mk3(o4), rd);
        bool alive = sp.valid;
        if (alive && length(sp.emission) > 0.0001f) {  // emitters terminate the path, as in the reference (cu:725-731)
            add_to_pixel(p, slot, atten * sp.emission);
            alive = false;
        }
        if (alive && depth <= 0) alive = false;
        float pdf_b = 0.0f;
        float3 wi = mk3(0.0f);
        if (alive) {
            const float ps = spec_probability(sp, wo);
            if (lv.nee) {
                const float x1 = myrnd(seed), x2 = myrnd(seed);
                float pdf_l;
                const float3 wl = env_sample(lv.cdf, x1 * 0.99999994f, x2 * 0.99999994f, &pdf_l);
                float pb;
                const float3 fc = bsdf_cos(sp, wo, wl, ps, &pb);
                if (pdf_l > 0.0f && (fc.x > 0.0f || fc.y > 0.0f || fc.z > 0.0f)) {
                    const float u = 0.5f + atan2f(wl.z, wl.x) * 0.15915494309189533577f;
                    const float v = 0.5f - asinf(fminf(fmaxf(wl.y, -1.0f), 1.0f)) * 0.31830988618379067154f;
                    const float3 Le = mk3(sample_env(s.env, s.env_w, s.env_h, u, v));
                    const float3 c = atten * fc * Le * (power_heuristic(pdf_l, pb) / pdf_l);
                    lv.shadow_o[slot] = make_float4(sp.pos.x, sp.pos.y, sp.pos.z, 0.0f);
                    lv.shadow_d[slot] = make_float4(wl.x, wl.y, wl.z, 0.0f);
                    lv.shadow_c[slot] = make_float4(c.x, c.y, c.z, 0.0f);
                    lv.shadow_flag[slot] = 1;
                }
            }
            // BSDF sample
            const float xl = myrnd(seed), r1 = myrnd(seed), r2 = myrnd(seed);
            const Onb onb(sp.n);
            if (xl < ps) {
                const float3 h = onb.inverse_transform(GGX_importance_sample(r1, r2, sp.roughness * sp.roughness));
                wi = normalize(reflect(rd, h));
            } else {
                wi = normalize(onb.inverse_transform(cosine_sample_hemisphere(r1, r2)));
            }
            const float3 fc = bsdf_cos(sp, wo, wi, ps, &pdf_b);
            if (pdf_b > 0.0f && (fc.x > 0.0f || fc.y > 0.0f || fc.z > 0.0f)) atten = atten * (fc / pdf_b);
            else alive = false;
        }
        uint32_t seed_rg = mi.x;
        if (alive && f.max_depth - depth >= 1) {  // Russian roulette from the third segment on
            const float q = fminf(1.0f, fmaxf(atten.x, fmaxf(atten.y, atten.z)));
            if (!(myrnd(seed_rg) < q)) alive = false; else atten = atten / q;
        }
        bool again;
        if (alive) {
            p.ray_o[slot] = make_float4(sp.pos.x, sp.pos.y, sp.pos.z, 0.0f);
            p.ray_d[slot] = make_float4(wi.x, wi.y, wi.z, 0.0f);
            p.atten_seed[slot] = make_float4(atten.x, atten.y, atten.z, __uint_as_float(seed));
            p.misc[slot] = make_uint4(seed_rg, (uint32_t)(depth - 1), mi.z, __float_as_uint(pdf_b));
            again = true;
        } else {
            again = linear_next_sample(f, p, slot, seed_rg, mi.z);
        }
        status[slot] = again ? ST_TRACE : ST_DONE;
    }
}

// any-hit traversal of the pending shadow rays; unoccluded contributions go to the pixel sum
PTB_DEV void chunk_stage_shadow(ChunkShared& sh, const SceneView& s, const FrameView& f, const PathView& p, const LinearView& lv,
                                uint32_t base, unsigned int n) {
    for (unsigned int i = threadIdx.x; i < n; i += PTB_CHUNK_THREADS) {
        const uint32_t slot = base + sh.list[i];
        const float4 d = lv.shadow_d[slot], c = lv.shadow_c[slot];
        const float3 o = mk3(lv.shadow_o[slot]);
        TravCounters tc;
        const HitRec h = bvh_closest_hit<false>(s, o, mk3(d), f.tmin, f.tmax, &tc);
        if (h.prim < 0) add_to_pixel(p, slot, mk3(c));
        lv.shadow_flag[slot] = 0;
    }
}

PTB_DEV void chunk_stage_miss_linear(ChunkShared& sh, const SceneView& s, const FrameView& f, const PathView& p, const LinearView& lv,
                                     unsigned char* __restrict__ status, uint32_t base, unsigned int n) {
    for (unsigned int i = threadIdx.x; i < n; i += PTB_CHUNK_THREADS) {
        const uint32_t slot = base + sh.list[i];
        const float4 d4 = p.ray_d[slot], as = p.atten_seed[slot];
        const uint4 mi = p.misc[slot];
        const float3 rd = normalize(mk3(d4));
        const float u = 0.5f + atan2f(rd.z, rd.x) * 0.15915494309189533577f;
        const float v = 0.5f - asinf(fminf(fmaxf(rd.y, -1.0f), 1.0f)) * 0.31830988618379067154f;
        const float3 Le = mk3(sample_env(s.env, s.env_w, s.env_h, u, v));
        const float pdf_b = __uint_as_float(mi.w);
        float w = 1.0f;
        if (lv.nee && pdf_b >= 0.0f) w = power_heuristic(pdf_b, env_pdf_uv(lv.cdf, u, v));
        add_to_pixel(p, slot, mk3(as) * Le * w);
        status[slot] = linear_next_sample(f, p, slot, mi.x, mi.z) ? ST_TRACE : ST_DONE;
    }
}

// ---- stage kernels of the linear mode (pipeline 2 layout: one kernel per stage and iteration) ----------------
__global__ void __launch_bounds__(PTB_CHUNK_THREADS) k_chunk_shade_linear(SceneView s, FrameView f, PathView p, LinearView lv, unsigned char* status) {
    __shared__ ChunkShared sh;
    const uint32_t base = blockIdx.x * PTB_CHUNK;
    const unsigned int n = chunk_build_list(sh, status, base, p.n_slots, ST_HIT);
    if (n) chunk_stage_shade_linear(sh, s, f, p, lv, status, base, n);
}
__global__ void __launch_bounds__(PTB_CHUNK_THREADS) k_chunk_shadow(SceneView s, FrameView f, PathView p, LinearView lv) {
    __shared__ ChunkShared sh;
    const uint32_t base = blockIdx.x * PTB_CHUNK;
    const unsigned int n = chunk_build_list(sh, lv.shadow_flag, base, p.n_slots, 1);
    if (n) chunk_stage_shadow(sh, s, f, p, lv, base, n);
}
__global__ void __launch_bounds__(PTB_CHUNK_THREADS) k_chunk_miss_linear(SceneView s, FrameView f, PathView p, LinearView lv, unsigned char* status) {
    __shared__ ChunkShared sh;
    const uint32_t base = blockIdx.x * PTB_CHUNK;
    const unsigned int n = chunk_build_list(sh, status, base, p.n_slots, ST_MISS);
    if (n) chunk_stage_miss_linear(sh, s, f, p, lv, status, base, n);
}

// test hook: n samples of the environment CDF -> direction (xyz) and pdf
__global__ void k_env_sample_test(EnvCdf cdf, const float* __restrict__ xi, uint32_t n, float* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float pdf;
    const float3 d = env_sample(cdf, xi[2 * (size_t)i], xi[2 * (size_t)i + 1], &pdf);
    out[4 * (size_t)i] = d.x; out[4 * (size_t)i + 1] = d.y; out[4 * (size_t)i + 2] = d.z; out[4 * (size_t)i + 3] = pdf;
}

}  // namespace PTB_NS
