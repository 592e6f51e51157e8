// obj_loader.cpp -- Wavefront OBJ reader for the scene-load row of the hot path
// (SURVEY.md section 8 a13).
//
// The reference calls tinyobj::LoadObj(..., triangulate=true) from the vendored
// tiny_obj_loader.h v1.0.6 (optixSphere.cpp:431) and then walks
// shapes -> faces -> 3 indices (optixSphere.cpp:447-515).  This is an
// independent reader that produces the SAME flattened face-vertex stream:
//   * triangle order = file order of `f` lines, each polygon fanned as
//     (0, k-1, k)                                   (tiny_obj_loader.h:908-931)
//   * indices are 1-based, 0 maps to 0, negative indices are relative to the
//     number of elements parsed so far              (tiny_obj_loader.h:425-429, 691-723)
//   * reals are parsed with the same digit-accumulation procedure as
//     tinyobj's tryParseDouble (tiny_obj_loader.h:474-587) and then narrowed to
//     float, so every coordinate is bit-identical to what the reference sees
//     (pinned by tests/test_scene_load.py against oracle/_ref/ref_probe).
//   * .mtl files are not read: the reference parses them and then ignores the
//     result (optixSphere.cpp:431, `materials` is never used).
#include "host.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>

namespace ptb {

namespace {

inline bool is_digit(char c) { return c >= '0' && c <= '9'; }
inline bool is_space(char c) { return c == ' ' || c == '\t'; }

// Decimal -> double with tinyobj's procedure: integer digits accumulate as
// m = m*10 + d, fraction digits add d * 10^-k (table for k < 8, pow beyond),
// a decimal exponent e is applied as ldexp(m * 5^e, e).
bool parse_double(const char* s, const char* s_end, double* result) {
    if (s >= s_end) return false;
    double mantissa = 0.0;
    int exponent = 0;
    char sign = '+', exp_sign = '+';
    const char* curr = s;
    int read = 0;
    bool more;
    if (*curr == '+' || *curr == '-') { sign = *curr; curr++; }
    else if (!is_digit(*curr)) return false;
    more = (curr != s_end);
    while (more && is_digit(*curr)) {
        mantissa *= 10; mantissa += (int)(*curr - '0');
        curr++; read++; more = (curr != s_end);
    }
    if (read == 0) return false;
    bool has_exp_part = false;
    if (more) {
        if (*curr == '.') {
            static const double neg_pow10[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            curr++; read = 1; more = (curr != s_end);
            while (more && is_digit(*curr)) {
                mantissa += (int)(*curr - '0') * (read < 8 ? neg_pow10[read] : std::pow(10.0, -read));
                read++; curr++; more = (curr != s_end);
            }
            has_exp_part = more;
        } else if (*curr == 'e' || *curr == 'E') {
            has_exp_part = true;
        }
    }
    if (has_exp_part && (*curr == 'e' || *curr == 'E')) {
        curr++;
        more = (curr != s_end);
        if (more && (*curr == '+' || *curr == '-')) { exp_sign = *curr; curr++; }
        else if (!is_digit(*curr)) return false;
        read = 0; more = (curr != s_end);
        while (more && is_digit(*curr)) {
            exponent *= 10; exponent += (int)(*curr - '0');
            curr++; read++; more = (curr != s_end);
        }
        exponent *= (exp_sign == '+' ? 1 : -1);
        if (read == 0) return false;
    }
    *result = (sign == '+' ? 1 : -1) * (exponent ? std::ldexp(mantissa * std::pow(5.0, exponent), exponent) : mantissa);
    return true;
}

float parse_real(const char** token) {
    (*token) += strspn(*token, " \t");
    const char* end = (*token) + strcspn(*token, " \t\r");
    double val = 0.0;
    parse_double(*token, end, &val);
    *token = end;
    return (float)val;
}

inline int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); }

// A face vertex as written in the file: raw 1-based / relative integers, ABSENT where a field is missing.  The
// indices are resolved (fix_index) once the number of elements parsed BEFORE the face is known, which for a chunk in
// the middle of the file is only the case after all chunks have been counted.
const int ABSENT = INT32_MIN;
struct RawTriple { int v, vt, vn; };

// i, i/j, i//k, i/j/k
RawTriple parse_triple(const char** token) {
    RawTriple t = {ABSENT, ABSENT, ABSENT};
    t.v = atoi(*token);
    (*token) += strcspn(*token, "/ \t\r");
    if ((*token)[0] != '/') return t;
    (*token)++;
    if ((*token)[0] == '/') {
        (*token)++;
        t.vn = atoi(*token);
        (*token) += strcspn(*token, "/ \t\r");
        return t;
    }
    t.vt = atoi(*token);
    (*token) += strcspn(*token, "/ \t\r");
    if ((*token)[0] != '/') return t;
    (*token)++;
    t.vn = atoi(*token);
    (*token) += strcspn(*token, "/ \t\r");
    return t;
}

// what one chunk of the file contributes
struct Chunk {
    std::vector<float> v, vn, vt;
    std::vector<RawTriple> corners;     // 3 per triangle, fanned
    std::vector<uint32_t> tri_counts;   // per triangle: elements of THIS chunk parsed before its face line (v, vn, vt)
};

void parse_chunk(const std::string& data, size_t pos, size_t end, Chunk& out) {
    std::vector<RawTriple> face;
    std::string line;
    while (pos < end) {
        // one line, accepting \n, \r\n and \r endings
        size_t e = pos;
        while (e < data.size() && data[e] != '\n' && data[e] != '\r') e++;
        line.assign(data, pos, e - pos);
        if (e < data.size() && data[e] == '\r' && e + 1 < data.size() && data[e + 1] == '\n') e++;
        pos = e + 1;
        if (line.empty()) continue;
        const char* token = line.c_str();
        token += strspn(token, " \t");
        if (token[0] == '\0' || token[0] == '#') continue;
        if (token[0] == 'v' && is_space(token[1])) {
            token += 2;
            float x = parse_real(&token), y = parse_real(&token), z = parse_real(&token);
            out.v.push_back(x); out.v.push_back(y); out.v.push_back(z);
        } else if (token[0] == 'v' && token[1] == 'n' && is_space(token[2])) {
            token += 3;
            float x = parse_real(&token), y = parse_real(&token), z = parse_real(&token);
            out.vn.push_back(x); out.vn.push_back(y); out.vn.push_back(z);
        } else if (token[0] == 'v' && token[1] == 't' && is_space(token[2])) {
            token += 3;
            float x = parse_real(&token), y = parse_real(&token);
            out.vt.push_back(x); out.vt.push_back(y);
        } else if (token[0] == 'f' && is_space(token[1])) {
            token += 2;
            token += strspn(token, " \t");
            face.clear();
            while (token[0] != '\0' && token[0] != '\r' && token[0] != '\n') {
                face.push_back(parse_triple(&token));
                token += strspn(token, " \t\r");
            }
            for (size_t k = 2; k < face.size(); ++k) {
                out.corners.push_back(face[0]); out.corners.push_back(face[k - 1]); out.corners.push_back(face[k]);
                out.tri_counts.push_back((uint32_t)(out.v.size() / 3)); out.tri_counts.push_back((uint32_t)(out.vn.size() / 3));
                out.tri_counts.push_back((uint32_t)(out.vt.size() / 2));
            }
        }
        // g / o / s / usemtl / mtllib carry nothing the reference uses
    }
}

}  // namespace

// The file is cut at line ends into one chunk per host thread (at least 4 MB each); the chunks are parsed concurrently
// and stitched together in file order, so the result is the same as a single sequential pass.
bool load_obj(const std::string& path, ObjMesh& mesh, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "Cannot open file [" + path + "]"; return false; }
    std::string data;
    if (fseek(f, 0, SEEK_END) == 0) { const long sz = ftell(f); if (sz > 0) data.reserve((size_t)sz); fseek(f, 0, SEEK_SET); }
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
    fclose(f);

    unsigned hw = std::thread::hardware_concurrency();
    if (hw == 0) hw = 1;
    size_t min_chunk = 4u << 20;
    if (const char* e = getenv("PTB_OBJ_CHUNK_BYTES")) { min_chunk = (size_t)std::max(1L, atol(e)); hw = 64; }  // tests: force many chunks
    size_t n_chunks = std::min<size_t>(hw, data.size() / min_chunk + 1);
    std::vector<size_t> cut(n_chunks + 1, data.size());
    cut[0] = 0;
    for (size_t k = 1; k < n_chunks; ++k) {
        size_t p = data.size() / n_chunks * k;
        if (p < cut[k - 1]) p = cut[k - 1];
        while (p < data.size() && data[p] != '\n' && data[p] != '\r') p++;  // to the end of the line that is cut
        if (p < data.size() && data[p] == '\r' && p + 1 < data.size() && data[p + 1] == '\n') p++;
        cut[k] = p < data.size() ? p + 1 : data.size();
    }
    std::vector<Chunk> chunks(n_chunks);
    {
        std::vector<std::thread> pool;
        for (size_t k = 1; k < n_chunks; ++k) pool.emplace_back([&, k]() { parse_chunk(data, cut[k], cut[k + 1], chunks[k]); });
        parse_chunk(data, cut[0], cut[1], chunks[0]);
        for (std::thread& th : pool) th.join();
    }

    mesh = ObjMesh();
    std::vector<size_t> ov(n_chunks + 1, 0), ovn(n_chunks + 1, 0), ovt(n_chunks + 1, 0), oc(n_chunks + 1, 0);
    for (size_t k = 0; k < n_chunks; ++k) {
        ov[k + 1] = ov[k] + chunks[k].v.size() / 3; ovn[k + 1] = ovn[k] + chunks[k].vn.size() / 3;
        ovt[k + 1] = ovt[k] + chunks[k].vt.size() / 2; oc[k + 1] = oc[k] + chunks[k].corners.size();
    }
    mesh.v.resize(ov[n_chunks] * 3); mesh.vn.resize(ovn[n_chunks] * 3); mesh.vt.resize(ovt[n_chunks] * 2);
    mesh.indices.resize(oc[n_chunks]);
    auto stitch = [&](size_t k) {
        const Chunk& c = chunks[k];
        if (!c.v.empty()) memcpy(&mesh.v[ov[k] * 3], c.v.data(), c.v.size() * sizeof(float));
        if (!c.vn.empty()) memcpy(&mesh.vn[ovn[k] * 3], c.vn.data(), c.vn.size() * sizeof(float));
        if (!c.vt.empty()) memcpy(&mesh.vt[ovt[k] * 2], c.vt.data(), c.vt.size() * sizeof(float));
        for (size_t i = 0; i < c.corners.size(); ++i) {
            const uint32_t* cnt = &c.tri_counts[(i / 3) * 3];
            const int nv = (int)(ov[k] + cnt[0]), nvn = (int)(ovn[k] + cnt[1]), nvt = (int)(ovt[k] + cnt[2]);
            const RawTriple& r = c.corners[i];
            ObjIndex ix;
            ix.v = fix_index(r.v, nv);
            ix.vt = r.vt == ABSENT ? -1 : fix_index(r.vt, nvt);
            ix.vn = r.vn == ABSENT ? -1 : fix_index(r.vn, nvn);
            mesh.indices[oc[k] + i] = ix;
        }
    };
    {
        std::vector<std::thread> pool;
        for (size_t k = 1; k < n_chunks; ++k) pool.emplace_back(stitch, k);
        stitch(0);
        for (std::thread& th : pool) th.join();
    }
    const int nv = (int)(mesh.v.size() / 3), nvn = (int)(mesh.vn.size() / 3), nvt = (int)(mesh.vt.size() / 2);
    for (const ObjIndex& ix : mesh.indices) {
        if (ix.v < 0 || ix.v >= nv || ix.vn >= nvn || ix.vt >= nvt) {
            err = "face index out of range in [" + path + "]";
            return false;
        }
    }
    return true;
}

}  // namespace ptb
