// env_cdf.cuh -- importance sampling of the equirect environment map through a prebuilt CDF
// (BASELINE.json north_star stage (d)).
//
// NOT a reference behaviour: the reference samples the BSDF only and its per-sample value is non-linear in the path
// throughput (optixSphere.cu:376-387), so no light sampling can reproduce its expectation.
// This file serves the separate LINEAR estimator of linear.cuh (ptb_render_cfg.env_importance_sampling = 1).
//
// Density: p(texel i,j) proportional to (luminance(i,j) + floor) * sin(theta_j), the solid-angle weight of an
// equirect row.  Stored as a marginal CDF over rows (h + 1 floats) and one conditional CDF per row ((w + 1) floats
// each), both normalised to [0,1].  Texel (i,j) covers u in [i/w,(i+1)/w), v in [j/h,(j+1)/h) with the SAME (u,v)
// parametrisation as the miss lookup (optixSphere.cu:543-544): u = 0.5 + atan2(z,x)/2pi, v = 0.5 - asin(y)/pi.
#pragma once
#include "device_math.cuh"

namespace PTB_NS {

struct EnvCdf {
    const float* marginal;     // [h + 1]
    const float* conditional;  // [h][w + 1]
    const float* row_weight;   // [h]: unnormalised row integrals
    float total;               // sum of row_weight
    int w, h;
};

__device__ __forceinline__ float env_luminance(float4 c) { return 0.2126f * c.x + 0.7152f * c.y + 0.0722f * c.z; }

// one block per row: unnormalised running sums of f = (lum + floor) * sin(theta)
__global__ void __launch_bounds__(256) k_env_row_cdf(const float4* __restrict__ env, int w, int h, float floor_lum,
                                                     float* __restrict__ conditional, float* __restrict__ row_weight) {
    __shared__ float warp_sums[8];
    __shared__ float carry;
    const int row = blockIdx.x;
    const float theta = 3.14159265358979323846f * ((float)row + 0.5f) / (float)h;
    const float st = sinf(theta);
    if (threadIdx.x == 0) { carry = 0.0f; conditional[(size_t)row * (w + 1)] = 0.0f; }
    __syncthreads();
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (int base = 0; base < w; base += 256) {
        const int i = base + (int)threadIdx.x;
        float v = 0.0f;
        if (i < w) v = (fmaxf(env_luminance(env[(size_t)row * w + i]), 0.0f) + floor_lum) * st;
        float x = v;
        for (int off = 1; off < 32; off <<= 1) { const float y = __shfl_up_sync(0xffffffffu, x, off); if ((int)lane >= off) x += y; }
        if (lane == 31u) warp_sums[warp] = x;
        __syncthreads();
        float woff = 0.0f, tot = 0.0f;
        for (unsigned k = 0; k < 8; ++k) { const float s = warp_sums[k]; if (k < warp) woff += s; tot += s; }
        const float c = carry;
        if (i < w) conditional[(size_t)row * (w + 1) + i + 1] = c + woff + x;
        __syncthreads();
        if (threadIdx.x == 0) carry = c + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) row_weight[row] = carry;
}

// marginal CDF over the rows (single block; h is a few thousand at most)
__global__ void __launch_bounds__(256) k_env_marginal(int h, const float* __restrict__ row_weight, float* __restrict__ marginal,
                                                      float* __restrict__ total_out) {
    __shared__ float total;
    if (threadIdx.x == 0) {
        float acc = 0.0f;
        marginal[0] = 0.0f;
        for (int r = 0; r < h; ++r) { acc += row_weight[r]; marginal[r + 1] = acc; }
        total = acc;
        *total_out = acc;
    }
    __syncthreads();
    const float inv_total = total > 0.0f ? 1.0f / total : 0.0f;
    for (int r = threadIdx.x; r <= h; r += blockDim.x) marginal[r] = r == h ? 1.0f : marginal[r] * inv_total;
}

// normalise the conditional CDF of every row to [0,1] (one block per row)
__global__ void __launch_bounds__(256) k_env_normalize_rows(int w, float* __restrict__ conditional, const float* __restrict__ row_weight) {
    const int r = blockIdx.x;
    const float rw = row_weight[r];
    float* row = conditional + (size_t)r * (w + 1);
    for (int i = threadIdx.x; i <= w; i += blockDim.x) row[i] = i == w ? 1.0f : (rw > 0.0f ? row[i] / rw : (float)i / (float)w);
}

// largest index k in [0, n-1] with cdf[k] <= x  (cdf has n + 1 entries, cdf[0] = 0, cdf[n] = 1)
PTB_DEV int cdf_find(const float* __restrict__ cdf, int n, float x) {
    int lo = 0, hi = n;  // invariant: cdf[lo] <= x < cdf[hi] (or hi == n)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= x) lo = mid; else hi = mid;
    }
    return lo;
}

PTB_DEV float3 env_uv_to_dir(float u, float v) {
    const float phi = (u - 0.5f) * 6.28318530717958647692f;
    const float el = (0.5f - v) * 3.14159265358979323846f;  // asin(y)
    float se, ce, sp, cp;
    sincosf(el, &se, &ce);
    sincosf(phi, &sp, &cp);
    return mk3(ce * cp, se, ce * sp);
}

// pdf with respect to solid angle of the direction whose equirect coordinates are (u, v)
PTB_DEV float env_pdf_uv(const EnvCdf& e, float u, float v) {
    int i = (int)(u * (float)e.w), j = (int)(v * (float)e.h);
    i = i < 0 ? 0 : (i >= e.w ? e.w - 1 : i);
    j = j < 0 ? 0 : (j >= e.h ? e.h - 1 : j);
    const float pm = __ldg(e.marginal + j + 1) - __ldg(e.marginal + j);
    const float* row = e.conditional + (size_t)j * (e.w + 1);
    const float pc = __ldg(row + i + 1) - __ldg(row + i);
    const float st = sinf(3.14159265358979323846f * ((float)j + 0.5f) / (float)e.h);
    // p(u,v) = pm * pc * w * h ; d(omega) = 2 pi^2 sin(theta) du dv
    return st > 0.0f ? pm * pc * (float)e.w * (float)e.h / (19.7392088021787172f * st) : 0.0f;
}

PTB_DEV float env_pdf_dir(const EnvCdf& e, float3 d) {
    const float u = 0.5f + atan2f(d.z, d.x) * 0.15915494309189533577f;
    const float v = 0.5f - asinf(fminf(fmaxf(d.y, -1.0f), 1.0f)) * 0.31830988618379067154f;
    return env_pdf_uv(e, u, v);
}

// (xi1, xi2) in [0,1) -> direction, pdf (solid angle)
PTB_DEV float3 env_sample(const EnvCdf& e, float xi1, float xi2, float* pdf) {
    xi1 = fminf(fmaxf(xi1, 0.0f), 0.99999994f); xi2 = fminf(fmaxf(xi2, 0.0f), 0.99999994f);
    const int j = cdf_find(e.marginal, e.h, xi1);
    const float m0 = __ldg(e.marginal + j), m1 = __ldg(e.marginal + j + 1);
    const float fv = m1 > m0 ? (xi1 - m0) / (m1 - m0) : 0.5f;
    const float* row = e.conditional + (size_t)j * (e.w + 1);
    const int i = cdf_find(row, e.w, xi2);
    const float c0 = __ldg(row + i), c1 = __ldg(row + i + 1);
    const float fu = c1 > c0 ? (xi2 - c0) / (c1 - c0) : 0.5f;
    const float u = ((float)i + fu) / (float)e.w, v = ((float)j + fv) / (float)e.h;
    *pdf = env_pdf_uv(e, u, v);
    return env_uv_to_dir(u, v);
}

}  // namespace PTB_NS
