// bvh8_host.cpp -- CPU harness for the wide BVHs: the 4-wide and the 8-wide quantised collapse (csrc/bvh8.cuh) and the traversals
// (csrc/bvh.cuh: trav_run, trav_run4, trav_run8) compiled for the host through cuda_host_shim.h and
// checked against a brute-force loop over all triangles with the same watertight test and tie rule.
// Test infrastructure (run by tests/test_bvh8_host.py); the 2-wide input tree is a median-split tree built here in the
// library's node format.
#include "cuda_host_shim.h"
#define PTB_HOST_SIM 1
#include "../../szakdolgozat_pathtracer_b200/csrc/bvh.cuh"
#include "../../szakdolgozat_pathtracer_b200/csrc/bvh8.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

using namespace ptb;

struct Tri { float v[3][3]; int id; };
struct Box { float lo[3], hi[3]; };

static Box tri_box(const Tri& t, float diag) {
    Box b;
    for (int d = 0; d < 3; ++d) {
        b.lo[d] = std::min(t.v[0][d], std::min(t.v[1][d], t.v[2][d]));
        b.hi[d] = std::max(t.v[0][d], std::max(t.v[1][d], t.v[2][d]));
    }
    float mag = 0.0f;
    for (int d = 0; d < 3; ++d) mag = std::max(mag, std::max(fabsf(b.lo[d]), fabsf(b.hi[d])));
    const float pad = diag * 2.384185791015625e-7f + mag * 9.5367431640625e-7f + 1e-30f;  // bvh_build.cu: k_leaf_boxes
    for (int d = 0; d < 3; ++d) { b.lo[d] -= pad; b.hi[d] += pad; }
    return b;
}
static Box merge(const Box& a, const Box& b) {
    Box r;
    for (int d = 0; d < 3; ++d) { r.lo[d] = std::min(a.lo[d], b.lo[d]); r.hi[d] = std::max(a.hi[d], b.hi[d]); }
    return r;
}

struct Builder {
    std::vector<Tri> tris;           // re-ordered into leaf order while building
    std::vector<float4> nodes;       // 4 float4 per node
    int max_leaf; float diag;
    // returns child code and box of the subtree over [b, e)
    int build(int b, int e, Box* box) {
        if (e - b <= max_leaf) {
            Box bb = tri_box(tris[b], diag);
            for (int i = b + 1; i < e; ++i) bb = merge(bb, tri_box(tris[i], diag));
            *box = bb;
            return ~((b << 3) | (e - b - 1));
        }
        float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
        for (int i = b; i < e; ++i) for (int d = 0; d < 3; ++d) {
            const float c = (tris[i].v[0][d] + tris[i].v[1][d] + tris[i].v[2][d]);
            clo[d] = std::min(clo[d], c); chi[d] = std::max(chi[d], c);
        }
        int ax = 0;
        if (chi[1] - clo[1] > chi[ax] - clo[ax]) ax = 1;
        if (chi[2] - clo[2] > chi[ax] - clo[ax]) ax = 2;
        const int m = (b + e) / 2;
        std::nth_element(tris.begin() + b, tris.begin() + m, tris.begin() + e, [ax](const Tri& x, const Tri& y) {
            return x.v[0][ax] + x.v[1][ax] + x.v[2][ax] < y.v[0][ax] + y.v[1][ax] + y.v[2][ax]; });
        const int me = (int)(nodes.size() / 4);
        nodes.resize(nodes.size() + 4);
        Box b0, b1;
        const int c0 = build(b, m, &b0), c1 = build(m, e, &b1);
        nodes[(size_t)me * 4 + 0] = make_float4(b0.lo[0], b0.hi[0], b0.lo[1], b0.hi[1]);
        nodes[(size_t)me * 4 + 1] = make_float4(b1.lo[0], b1.hi[0], b1.lo[1], b1.hi[1]);
        nodes[(size_t)me * 4 + 2] = make_float4(b0.lo[2], b0.hi[2], b1.lo[2], b1.hi[2]);
        nodes[(size_t)me * 4 + 3] = make_float4(__int_as_float(c0), __int_as_float(c1), 0.0f, 0.0f);
        *box = merge(b0, b1);
        return me;
    }
};

struct HostAlloc {
    uint32_t n_nodes = 1, n_tris = 0; int err = 0;
    std::vector<ptb8::WorkItem> next;
    uint32_t nodes(uint32_t n) { const uint32_t r = n_nodes; n_nodes += n; return r; }
    uint32_t tris(uint32_t n) { const uint32_t r = n_tris; n_tris += n; return r; }
    void push(ptb8::WorkItem w) { next.push_back(w); }
    void error(int b) { err |= b; }
};

static HitRec brute(const std::vector<float4>& tris, float3 o, float3 d, float tmin, float tmax) {
    const float3 id = mk3(ex_div(1.0f, d.x), ex_div(1.0f, d.y), ex_div(1.0f, d.z));
    const RayShear rs = ray_shear(d, id);
    HitRec best; best.t = tmax; best.b1 = 0; best.b2 = 0; best.prim = -1;
    for (size_t i = 0; i < tris.size() / 3; ++i) {
        float th, b1, b2;
        if (ray_tri(o, rs, mk3(tris[i * 3]), mk3(tris[i * 3 + 1]), mk3(tris[i * 3 + 2]), tmin, tmax, &th, &b1, &b2)) {
            const int prim = __float_as_int(tris[i * 3].w);
            if (th < best.t || (th == best.t && best.prim >= 0 && prim < best.prim)) { best.t = th; best.b1 = b1; best.b2 = b2; best.prim = prim; }
        }
    }
    return best;
}

static bool same(const HitRec& a, const HitRec& b) {
    return a.prim == b.prim && __float_as_int(a.t) == __float_as_int(b.t) && __float_as_int(a.b1) == __float_as_int(b.b1) &&
           __float_as_int(a.b2) == __float_as_int(b.b2);
}

int main(int argc, char** argv) {
    const int n_tris = argc > 1 ? atoi(argv[1]) : 4000;
    const int n_rays = argc > 2 ? atoi(argv[2]) : 20000;
    const unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 1u;
    std::mt19937 rng(seed);
    std::uniform_real_distribution<float> U(-1.0f, 1.0f);
    Builder B; B.max_leaf = 3;
    // scene: clustered small triangles, some slivers, a few large ones and a floor quad far larger than the rest
    for (int i = 0; i < n_tris; ++i) {
        Tri t; t.id = i;
        const float cx = 40.0f * U(rng), cy = 3.0f * U(rng) + 3.0f, cz = 40.0f * U(rng);
        const float s = (i % 97 == 0) ? 15.0f : ((i % 11 == 0) ? 0.01f : 0.8f);
        for (int v = 0; v < 3; ++v) { t.v[v][0] = cx + s * U(rng); t.v[v][1] = cy + s * U(rng); t.v[v][2] = cz + s * U(rng); }
        if (i % 53 == 0) t.v[2][1] = t.v[1][1] = t.v[0][1];       // axis-aligned (flat box)
        B.tris.push_back(t);
    }
    { Tri a = {{{-400, 0, -400}, {400, 0, -400}, {400, 0, 400}}, n_tris}, b = {{{-400, 0, -400}, {400, 0, 400}, {-400, 0, 400}}, n_tris + 1};
      B.tris.push_back(a); B.tris.push_back(b); }
    B.diag = sqrtf(800.0f * 800.0f * 2.0f + 30.0f * 30.0f);
    Box root;
    const int rc = B.build(0, (int)B.tris.size(), &root);
    if (rc != 0) { printf("FAIL: root code %d\n", rc); return 1; }
    std::vector<float4> tris2(B.tris.size() * 3);
    for (size_t i = 0; i < B.tris.size(); ++i) {
        const Tri& t = B.tris[i];
        tris2[i * 3 + 0] = make_float4(t.v[0][0], t.v[0][1], t.v[0][2], __int_as_float(t.id));
        tris2[i * 3 + 1] = make_float4(t.v[1][0], t.v[1][1], t.v[1][2], 0.0f);
        tris2[i * 3 + 2] = make_float4(t.v[2][0], t.v[2][1], t.v[2][2], 0.0f);
    }
    const size_t n_nodes2 = B.nodes.size() / 4;
    // ---- collapse into the 8-wide tree, level by level as the build kernel does
    std::vector<uint4> nodes8(n_nodes2 * 5);
    std::vector<float4> tris8(tris2.size());
    HostAlloc al;
    std::vector<ptb8::WorkItem> cur(1); cur[0].wide = 0; cur[0].bin = 0;
    int levels = 0; size_t slots_used = 0, internal_children = 0;
    while (!cur.empty()) {
        al.next.clear();
        for (const ptb8::WorkItem& w : cur) internal_children += (size_t)ptb8::collapse8_node(B.nodes.data(), tris2.data(), w, nodes8.data(), tris8.data(), al);
        cur = al.next; ++levels;
    }
    if (al.err) { printf("FAIL: collapse error bits %d\n", al.err); return 1; }
    if (al.n_tris != tris2.size() / 3) { printf("FAIL: %u triangles in the 8-wide tree, %zu expected\n", al.n_tris, tris2.size() / 3); return 1; }
    for (uint32_t i = 0; i < al.n_nodes; ++i) { const uint4 q1 = nodes8[(size_t)i * 5 + 1]; for (int k = 0; k < 8; ++k) slots_used += (((k < 4 ? q1.z : q1.w) >> (8 * (k & 3))) & 0xffu) ? 1 : 0; }
    // every triangle id appears exactly once in tris8
    { std::vector<int> seen(tris2.size() / 3, 0);
      for (size_t i = 0; i < tris8.size() / 3; ++i) { const int id = __float_as_int(tris8[i * 3].w); if (id < 0 || id >= (int)seen.size() || seen[id]++) { printf("FAIL: triangle order\n"); return 1; } } }

    // ---- and into the 4-wide tree (greedy by surface area), level by level
    std::vector<float4> nodes4(n_nodes2 * 8);
    size_t n_nodes4 = 0; int levels4 = 0;
    {
        struct HostPush { std::vector<int>* q; void operator()(int c) { q->push_back(c); } };
        std::vector<int> cur4(1, 0), next4;
        while (!cur4.empty()) {
            next4.clear();
            HostPush hp{&next4};
            for (int b : cur4) ptb8::collapse4_node(B.nodes.data(), b, nodes4.data(), hp);
            n_nodes4 += cur4.size(); cur4 = next4; ++levels4;
        }
    }

    ptbv::SceneView sv; memset(&sv, 0, sizeof(sv));
    sv.nodes = B.nodes.data(); sv.tris = tris2.data(); sv.nodes8 = nodes8.data(); sv.tris8 = tris8.data();
    ptbv::SceneView sv2 = sv; sv2.nodes8 = nullptr;
    ptbv::SceneView sv4 = sv2; sv4.nodes4 = nodes4.data();

    unsigned long long n2 = 0, t2 = 0, n8 = 0, t8 = 0, n4 = 0, t4 = 0; int bad = 0, hits = 0;
    unsigned long long kn2[16] = {0}, kn8[16] = {0}, kt2[16] = {0}, kt8[16] = {0}; int kc[16] = {0};
    for (int r = 0; r < n_rays; ++r) {
        float3 o = mk3(60.0f * U(rng), 8.0f * U(rng) + 6.0f, 60.0f * U(rng));
        float3 d = mk3(U(rng), U(rng), U(rng));
        const int kind = r % 16;
        if (kind == 1) d.x = 0.0f;                       // axis-parallel rays: 1 / d = inf
        if (kind == 2) { d.y = 0.0f; d.z = -0.0f; }
        if (kind == 3) { d.x = 0.0f; d.z = 0.0f; d.y = -1.0f; }
        if (kind == 4) d.y = 1e-7f * U(rng);             // nearly parallel to the floor
        if (kind == 5) { o = mk3(3000.0f * U(rng), 2000.0f, 3000.0f * U(rng)); d = mk3(-o.x + 30 * U(rng), -o.y, -o.z + 30 * U(rng)); }  // far camera
        if (kind == 6) {                                 // origin on a triangle's vertex (bounce rays start on surfaces)
            const Tri& t = B.tris[(size_t)(rng() % B.tris.size())];
            o = mk3(t.v[0][0], t.v[0][1], t.v[0][2]);
        }
        const float len = sqrtf(d.x * d.x + d.y * d.y + d.z * d.z);
        if (!(len > 0.0f)) continue;
        d = mk3(d.x / len, d.y / len, d.z / len);
        const float tmin = 0.01f, tmax = 1e16f;
        const HitRec hb = brute(tris2, o, d, tmin, tmax);
        TravCounters c2 = {0, 0}, c8 = {0, 0}, c8q = {0, 0};
        const HitRec h2 = bvh_closest_hit<true>(sv2, o, d, tmin, tmax, &c2);
        const HitRec h8 = bvh_closest_hit<true>(sv, o, d, tmin, tmax, &c8);
        TravCounters c4 = {0, 0};
        const HitRec h4 = bvh_closest_hit<true>(sv4, o, d, tmin, tmax, &c4);
        n4 += c4.nodes; t4 += c4.tris;
        // the same ray in quanta of 8 steps, as the wavefront kernels run it
        __attribute__((aligned(16))) int stack[PTB_BVH_STACK];
        Trav t;
        trav_begin_any<8>(t, stack, sv, o, d, tmin, tmax);
        int guard = 0;
        while (!trav_run_any<true, 8>(t, stack, sv, 8, &c8q) && ++guard < 100000) {}
        n2 += c2.nodes; t2 += c2.tris; n8 += c8.nodes; t8 += c8.tris;
        kn2[kind] += c2.nodes; kn8[kind] += c8.nodes; kt2[kind] += c2.tris; kt8[kind] += c8.tris; kc[kind]++;
        hits += hb.prim >= 0;
        if (!same(hb, h2) || !same(hb, h8) || !same(hb, h4) || !same(hb, t.best) || c8q.nodes != c8.nodes || c8q.tris != c8.tris) {
            if (bad < 10) printf("MISMATCH ray %d kind %d: brute prim %d t %.9g | 2-wide %d %.9g | 8-wide %d %.9g | quanta %d %.9g (nodes %u/%u)\n", r, kind, hb.prim, hb.t,
                                 h2.prim, h2.t, h8.prim, h8.t, t.best.prim, t.best.t, c8q.nodes, c8.nodes);
            ++bad;
        }
    }
    printf("triangles %zu  nodes2 %zu  nodes8 %u  levels %d  slots used per node %.2f\n", tris2.size() / 3, n_nodes2, al.n_nodes, levels, (double)slots_used / al.n_nodes);
    printf("rays %d  hits %d  per ray: 2-wide %.2f nodes %.2f tris | 4-wide (%zu nodes, %d levels) %.2f nodes %.2f tris | 8-wide %.2f nodes %.2f tris\n", n_rays, hits,
           (double)n2 / n_rays, (double)t2 / n_rays, n_nodes4, levels4, (double)n4 / n_rays, (double)t4 / n_rays, (double)n8 / n_rays, (double)t8 / n_rays);
    for (int k = 0; k < 8; ++k) if (kc[k]) printf("  ray kind %d: 2-wide %.1f nodes %.1f tris | 8-wide %.1f nodes %.1f tris\n", k, (double)kn2[k] / kc[k], (double)kt2[k] / kc[k], (double)kn8[k] / kc[k], (double)kt8[k] / kc[k]);
    if (bad) { printf("FAIL: %d mismatches\n", bad); return 1; }
    printf("OK\n");
    return 0;
}
