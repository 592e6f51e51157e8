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

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

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

struct Triple { int v, vt, vn; };

// i, i/j, i//k, i/j/k
Triple parse_triple(const char** token, int nv, int nvn, int nvt) {
    Triple t = {-1, -1, -1};
    t.v = fix_index(atoi(*token), nv);
    (*token) += strcspn(*token, "/ \t\r");
    if ((*token)[0] != '/') return t;
    (*token)++;
    if ((*token)[0] == '/') {
        (*token)++;
        t.vn = fix_index(atoi(*token), nvn);
        (*token) += strcspn(*token, "/ \t\r");
        return t;
    }
    t.vt = fix_index(atoi(*token), nvt);
    (*token) += strcspn(*token, "/ \t\r");
    if ((*token)[0] != '/') return t;
    (*token)++;
    t.vn = fix_index(atoi(*token), nvn);
    (*token) += strcspn(*token, "/ \t\r");
    return t;
}

}  // namespace

bool load_obj(const std::string& path, ObjMesh& mesh, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "Cannot open file [" + path + "]"; return false; }
    std::string data;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) data.append(buf, n);
    fclose(f);

    mesh = ObjMesh();
    std::vector<Triple> face;
    size_t pos = 0;
    std::string line;
    while (pos < data.size()) {
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
            mesh.v.push_back(x); mesh.v.push_back(y); mesh.v.push_back(z);
        } else if (token[0] == 'v' && token[1] == 'n' && is_space(token[2])) {
            token += 3;
            float x = parse_real(&token), y = parse_real(&token), z = parse_real(&token);
            mesh.vn.push_back(x); mesh.vn.push_back(y); mesh.vn.push_back(z);
        } else if (token[0] == 'v' && token[1] == 't' && is_space(token[2])) {
            token += 3;
            float x = parse_real(&token), y = parse_real(&token);
            mesh.vt.push_back(x); mesh.vt.push_back(y);
        } else if (token[0] == 'f' && is_space(token[1])) {
            token += 2;
            token += strspn(token, " \t");
            face.clear();
            while (token[0] != '\0' && token[0] != '\r' && token[0] != '\n') {
                face.push_back(parse_triple(&token, (int)(mesh.v.size() / 3), (int)(mesh.vn.size() / 3), (int)(mesh.vt.size() / 2)));
                token += strspn(token, " \t\r");
            }
            for (size_t k = 2; k < face.size(); ++k) {
                const Triple tri[3] = {face[0], face[k - 1], face[k]};
                for (int c = 0; c < 3; ++c) {
                    ObjIndex ix; ix.v = tri[c].v; ix.vt = tri[c].vt; ix.vn = tri[c].vn;
                    mesh.indices.push_back(ix);
                }
            }
        }
        // g / o / s / usemtl / mtllib carry nothing the reference uses
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
