// image_io.cpp -- the on-disk image formats either side of the path
// (SURVEY.md section 8 f2): what sutil::loadImage / sutil::saveImage do for the
// reference (optixSphere.cpp:359, 836, 1483-1489).  The SDK delegates to
// stb_image (PNG) and tinyexr (EXR); neither is available here, so these are
// independent decoders on top of zlib with the same observable results:
//   PNG -> RGBA8 as stbi_load(..., STBI_rgb_alpha): gray -> (g,g,g,255),
//          gray+alpha -> (g,g,g,a), RGB -> (r,g,b,255), palette expanded,
//          tRNS colour keys -> alpha 0, 16-bit samples -> high byte,
//          1/2/4-bit gray scaled to 0..255.
//   EXR -> float4 as tinyexr LoadEXR: channels R,G,B,(A); missing A = 1;
//          a single luminance channel is replicated; HALF widened exactly.
#include "host.h"

#include <zlib.h>

#include <atomic>
#include <thread>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

namespace ptb {

namespace {

bool read_file(const std::string& path, std::vector<uint8_t>& out, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    fseek(f, 0, SEEK_END);
    long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    if (sz < 0) { fclose(f); err = "cannot size " + path; return false; }
    out.resize((size_t)sz);
    size_t got = sz ? fread(out.data(), 1, (size_t)sz, f) : 0;
    fclose(f);
    if (got != (size_t)sz) { err = "short read on " + path; return false; }
    return true;
}

inline uint32_t be32(const uint8_t* p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }

bool zlib_inflate(const uint8_t* src, size_t n, std::vector<uint8_t>& dst, size_t expected, std::string& err) {
    dst.resize(expected);
    uLongf dl = (uLongf)expected;
    int rc = uncompress(dst.data(), &dl, src, (uLong)n);
    if (rc != Z_OK) { err = "zlib inflate failed (" + std::to_string(rc) + ")"; return false; }
    dst.resize(dl);
    return true;
}

inline int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}

// Undo the per-scanline filters of one (sub)image in place; returns packed rows.
bool png_unfilter(const uint8_t* in, size_t in_size, int w, int h, int bits_per_pixel, std::vector<uint8_t>& rows,
                  std::string& err) {
    const size_t stride = ((size_t)w * (size_t)bits_per_pixel + 7) / 8;
    const int bpp = std::max(1, bits_per_pixel / 8);
    if (in_size < (stride + 1) * (size_t)h) { err = "PNG: not enough image data"; return false; }
    rows.assign(stride * (size_t)h, 0);
    for (int y = 0; y < h; ++y) {
        const uint8_t* src = in + (stride + 1) * (size_t)y;
        uint8_t* cur = rows.data() + stride * (size_t)y;
        const uint8_t* prev = y ? cur - stride : nullptr;
        const int ft = src[0];
        src++;
        for (size_t i = 0; i < stride; ++i) {
            int a = i >= (size_t)bpp ? cur[i - bpp] : 0;
            int b = prev ? prev[i] : 0;
            int c = (prev && i >= (size_t)bpp) ? prev[i - bpp] : 0;
            int v = src[i];
            switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: err = "PNG: bad filter type"; return false;
            }
            cur[i] = (uint8_t)v;
        }
    }
    return true;
}

}  // namespace

bool file_exists(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    fclose(f);
    return true;
}

bool load_png_rgba8(const std::string& path, std::vector<uint8_t>& rgba, int& w, int& h, std::string& err) {
    std::vector<uint8_t> file;
    if (!read_file(path, file, err)) return false;
    static const uint8_t sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (file.size() < 8 || memcmp(file.data(), sig, 8) != 0) { err = "not a PNG: " + path; return false; }
    size_t pos = 8;
    int depth = 0, ctype = 0, interlace = 0;
    bool have_ihdr = false;
    std::vector<uint8_t> idat, plte, trns;
    while (pos + 12 <= file.size()) {
        uint32_t len = be32(&file[pos]);
        const uint8_t* type = &file[pos + 4];
        if (pos + 12 + (size_t)len > file.size()) { err = "PNG: truncated chunk"; return false; }
        const uint8_t* body = &file[pos + 8];
        if (!memcmp(type, "IHDR", 4)) {
            if (len < 13) { err = "PNG: bad IHDR"; return false; }
            w = (int)be32(body); h = (int)be32(body + 4);
            depth = body[8]; ctype = body[9]; interlace = body[12];
            if (body[10] != 0 || body[11] != 0) { err = "PNG: unknown compression/filter method"; return false; }
            have_ihdr = true;
        } else if (!memcmp(type, "PLTE", 4)) plte.assign(body, body + len);
        else if (!memcmp(type, "tRNS", 4)) trns.assign(body, body + len);
        else if (!memcmp(type, "IDAT", 4)) idat.insert(idat.end(), body, body + len);
        else if (!memcmp(type, "IEND", 4)) break;
        pos += 12 + (size_t)len;
    }
    if (!have_ihdr || w <= 0 || h <= 0) { err = "PNG: missing IHDR"; return false; }
    int channels;
    switch (ctype) {
        case 0: channels = 1; break;
        case 2: channels = 3; break;
        case 3: channels = 1; break;
        case 4: channels = 2; break;
        case 6: channels = 4; break;
        default: err = "PNG: bad colour type"; return false;
    }
    if (!(depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) { err = "PNG: bad bit depth"; return false; }
    if ((ctype == 2 || ctype == 4 || ctype == 6) && depth < 8) { err = "PNG: bad depth for colour type"; return false; }
    if (ctype == 3 && depth == 16) { err = "PNG: bad depth for palette"; return false; }
    const int bits_pp = channels * depth;

    // Sub-image list: 1 pass, or the 7 Adam7 passes.
    struct Pass { int x0, y0, dx, dy; };
    static const Pass adam7[7] = {{0, 0, 8, 8}, {4, 0, 8, 8}, {0, 4, 4, 8}, {2, 0, 4, 4}, {0, 2, 2, 4}, {1, 0, 2, 2}, {0, 1, 1, 2}};
    static const Pass whole = {0, 0, 1, 1};
    const int npass = interlace ? 7 : 1;
    size_t raw_size = 0;
    for (int p = 0; p < npass; ++p) {
        const Pass& ps = interlace ? adam7[p] : whole;
        int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        raw_size += (((size_t)pw * bits_pp + 7) / 8 + 1) * (size_t)ph;
    }
    std::vector<uint8_t> raw;
    if (!zlib_inflate(idat.data(), idat.size(), raw, raw_size, err)) return false;
    if (raw.size() != raw_size) { err = "PNG: unexpected image data size"; return false; }

    rgba.assign((size_t)w * h * 4, 0);
    // tRNS colour key (types 0 and 2): 16-bit big-endian samples
    int key[3] = {-1, -1, -1};
    if (ctype == 0 && trns.size() >= 2) key[0] = (trns[0] << 8) | trns[1];
    if (ctype == 2 && trns.size() >= 6) for (int c = 0; c < 3; ++c) key[c] = (trns[2 * c] << 8) | trns[2 * c + 1];
    const int gray_scale = depth < 8 ? 255 / ((1 << depth) - 1) : 1;

    size_t off = 0;
    std::vector<uint8_t> rows;
    for (int p = 0; p < npass; ++p) {
        const Pass& ps = interlace ? adam7[p] : whole;
        int pw = (w - ps.x0 + ps.dx - 1) / ps.dx, ph = (h - ps.y0 + ps.dy - 1) / ps.dy;
        if (pw <= 0 || ph <= 0) continue;
        const size_t stride = ((size_t)pw * bits_pp + 7) / 8;
        if (!png_unfilter(raw.data() + off, raw.size() - off, pw, ph, bits_pp, rows, err)) return false;
        off += (stride + 1) * (size_t)ph;
        for (int y = 0; y < ph; ++y) {
            const uint8_t* r = rows.data() + stride * (size_t)y;
            for (int x = 0; x < pw; ++x) {
                int s[4] = {0, 0, 0, 0};  // raw samples at file depth
                for (int c = 0; c < channels; ++c) {
                    if (depth == 8) s[c] = r[(size_t)x * channels + c];
                    else if (depth == 16) s[c] = (r[((size_t)x * channels + c) * 2] << 8) | r[((size_t)x * channels + c) * 2 + 1];
                    else {
                        size_t bit = (size_t)x * depth;
                        s[c] = (r[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                    }
                }
                auto to8 = [&](int v) -> uint8_t { return depth == 16 ? (uint8_t)(v >> 8) : (uint8_t)v; };
                uint8_t px[4] = {0, 0, 0, 255};
                switch (ctype) {
                    case 0: {
                        uint8_t g = depth < 8 ? (uint8_t)(s[0] * gray_scale) : to8(s[0]);
                        px[0] = px[1] = px[2] = g;
                        if (key[0] >= 0 && s[0] == key[0]) px[3] = 0;
                        break;
                    }
                    case 2:
                        px[0] = to8(s[0]); px[1] = to8(s[1]); px[2] = to8(s[2]);
                        if (key[0] >= 0 && s[0] == key[0] && s[1] == key[1] && s[2] == key[2]) px[3] = 0;
                        break;
                    case 3: {
                        size_t i = (size_t)s[0];
                        if (i * 3 + 2 < plte.size()) { px[0] = plte[i * 3]; px[1] = plte[i * 3 + 1]; px[2] = plte[i * 3 + 2]; }
                        if (i < trns.size()) px[3] = trns[i];
                        break;
                    }
                    case 4: px[0] = px[1] = px[2] = to8(s[0]); px[3] = to8(s[1]); break;
                    case 6: px[0] = to8(s[0]); px[1] = to8(s[1]); px[2] = to8(s[2]); px[3] = to8(s[3]); break;
                }
                const int ox = ps.x0 + x * ps.dx, oy = ps.y0 + y * ps.dy;
                memcpy(&rgba[((size_t)oy * w + ox) * 4], px, 4);
            }
        }
    }
    return true;
}

// ---- writers -------------------------------------------------------------------
namespace {
void put_be32(std::vector<uint8_t>& v, uint32_t x) { v.push_back(x >> 24); v.push_back(x >> 16); v.push_back(x >> 8); v.push_back(x); }
void png_chunk(std::vector<uint8_t>& out, const char* type, const std::vector<uint8_t>& body) {
    put_be32(out, (uint32_t)body.size());
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), body.begin(), body.end());
    uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(out.size() - start));
    put_be32(out, crc);
}
}  // namespace

bool save_png_rgba8(const std::string& path, const uint8_t* rgba, int w, int h, bool flip_y, std::string& err) {
    std::vector<uint8_t> raw((size_t)h * ((size_t)w * 4 + 1));
    for (int y = 0; y < h; ++y) {
        int sy = flip_y ? h - 1 - y : y;
        uint8_t* d = &raw[(size_t)y * ((size_t)w * 4 + 1)];
        d[0] = 0;
        memcpy(d + 1, rgba + (size_t)sy * w * 4, (size_t)w * 4);
    }
    uLongf cl = compressBound((uLong)raw.size());
    std::vector<uint8_t> comp(cl);
    if (compress2(comp.data(), &cl, raw.data(), (uLong)raw.size(), 6) != Z_OK) { err = "PNG: deflate failed"; return false; }
    comp.resize(cl);
    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    std::vector<uint8_t> ihdr;
    put_be32(ihdr, (uint32_t)w); put_be32(ihdr, (uint32_t)h);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", comp);
    png_chunk(out, "IEND", std::vector<uint8_t>());
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
    fclose(f);
    if (!ok) err = "short write on " + path;
    return ok;
}

bool save_ppm_rgb8(const std::string& path, const uint8_t* rgba, int w, int h, bool flip_y, std::string& err) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    fprintf(f, "P6\n%d %d\n255\n", w, h);
    std::vector<uint8_t> row((size_t)w * 3);
    for (int y = 0; y < h; ++y) {
        int sy = flip_y ? h - 1 - y : y;
        for (int x = 0; x < w; ++x) memcpy(&row[(size_t)x * 3], rgba + ((size_t)sy * w + x) * 4, 3);
        fwrite(row.data(), 1, row.size(), f);
    }
    fclose(f);
    return true;
}

// ---- OpenEXR (scanline) ----------------------------------------------------------
namespace {

inline float half_to_float(uint16_t hbits) {
    uint32_t sign = (uint32_t)(hbits >> 15) << 31, exp = (hbits >> 10) & 0x1f, man = hbits & 0x3ff, out;
    if (exp == 0) {
        if (man == 0) out = sign;
        else {
            int e = -1;
            do { e++; man <<= 1; } while ((man & 0x400) == 0);
            out = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ff) << 13);
        }
    } else if (exp == 31) out = sign | 0x7f800000u | (man << 13);
    else out = sign | ((exp + 127 - 15) << 23) | (man << 13);
    float f; memcpy(&f, &out, 4); return f;
}

struct ExrChannel { std::string name; int type; int xs, ys; };

// OpenEXR's ZIP/RLE post-process: undo the delta predictor, then interleave halves.
void exr_unpredict_interleave(std::vector<uint8_t>& buf) {
    const size_t n = buf.size();
    if (!n) return;
    for (size_t i = 1; i < n; ++i) buf[i] = (uint8_t)(buf[i - 1] + buf[i] - 128);
    std::vector<uint8_t> out(n);
    const size_t half = (n + 1) / 2;
    size_t a = 0, b = half;
    for (size_t i = 0; i < n;) {
        out[i++] = buf[a++];
        if (i < n) out[i++] = buf[b++];
    }
    buf.swap(out);
}

bool exr_rle_decode(const uint8_t* in, size_t n, std::vector<uint8_t>& out, size_t expected) {
    out.clear(); out.reserve(expected);
    size_t i = 0;
    while (i < n) {
        int8_t c = (int8_t)in[i++];
        if (c < 0) {
            size_t cnt = (size_t)(-c);
            if (i + cnt > n) return false;
            out.insert(out.end(), in + i, in + i + cnt); i += cnt;
        } else {
            if (i >= n) return false;
            out.insert(out.end(), (size_t)c + 1, in[i++]);
        }
    }
    return out.size() == expected;
}

}  // namespace

bool load_exr_float4(const std::string& path, std::vector<float>& rgba, int& w, int& h, std::string& err) {
    std::vector<uint8_t> f;
    if (!read_file(path, f, err)) return false;
    auto need = [&](size_t pos, size_t n) { return pos + n <= f.size(); };
    auto rd32 = [&](size_t pos) { uint32_t v; memcpy(&v, &f[pos], 4); return v; };
    if (f.size() < 8 || rd32(0) != 20000630u) { err = "not an EXR: " + path; return false; }
    const uint32_t version = rd32(4);
    if (version & 0x200u) { err = "EXR: tiled images are not supported"; return false; }
    if (version & 0x1800u) { err = "EXR: deep/multipart images are not supported"; return false; }
    size_t pos = 8;
    std::vector<ExrChannel> chans;
    int compression = -1, line_order = 0;
    int dw[4] = {0, 0, -1, -1};
    for (;;) {
        if (!need(pos, 1)) { err = "EXR: truncated header"; return false; }
        if (f[pos] == 0) { pos++; break; }
        std::string name((const char*)&f[pos]); pos += name.size() + 1;
        if (!need(pos, 1)) { err = "EXR: truncated header"; return false; }
        std::string type((const char*)&f[pos]); pos += type.size() + 1;
        if (!need(pos, 4)) { err = "EXR: truncated header"; return false; }
        uint32_t size = rd32(pos); pos += 4;
        if (!need(pos, size)) { err = "EXR: truncated attribute"; return false; }
        if (name == "channels") {
            size_t p = pos;
            while (p < pos + size && f[p] != 0) {
                ExrChannel c; c.name = (const char*)&f[p]; p += c.name.size() + 1;
                c.type = (int)rd32(p); p += 4; p += 4;  // pLinear + reserved
                c.xs = (int)rd32(p); p += 4; c.ys = (int)rd32(p); p += 4;
                chans.push_back(c);
            }
        } else if (name == "compression") compression = f[pos];
        else if (name == "dataWindow") for (int i = 0; i < 4; ++i) dw[i] = (int)rd32(pos + 4 * (size_t)i);
        else if (name == "lineOrder") line_order = f[pos];
        pos += size;
    }
    (void)line_order;  // chunks carry their own y; any order decodes the same
    w = dw[2] - dw[0] + 1; h = dw[3] - dw[1] + 1;
    if (w <= 0 || h <= 0 || chans.empty()) { err = "EXR: bad header"; return false; }
    int lines_per_block;
    switch (compression) {
        case 0: case 1: case 2: lines_per_block = 1; break;
        case 3: lines_per_block = 16; break;
        default: err = "EXR: compression " + std::to_string(compression) + " is not supported (NONE/RLE/ZIPS/ZIP only)"; return false;
    }
    size_t line_bytes = 0;
    std::vector<size_t> chan_off(chans.size());
    for (size_t c = 0; c < chans.size(); ++c) {
        if (chans[c].xs != 1 || chans[c].ys != 1) { err = "EXR: subsampled channels are not supported"; return false; }
        if (chans[c].type < 0 || chans[c].type > 2) { err = "EXR: bad channel type"; return false; }
        chan_off[c] = line_bytes;
        line_bytes += (size_t)w * (chans[c].type == 1 ? 2 : 4);
    }
    // channel -> output slot
    int slot_of[4] = {-1, -1, -1, -1};
    for (size_t c = 0; c < chans.size(); ++c) {
        const std::string& n = chans[c].name;
        if (n == "R") slot_of[0] = (int)c; else if (n == "G") slot_of[1] = (int)c;
        else if (n == "B") slot_of[2] = (int)c; else if (n == "A") slot_of[3] = (int)c;
    }
    if (slot_of[0] < 0 && slot_of[1] < 0 && slot_of[2] < 0) {
        // single luminance-like channel: replicate
        if (chans.size() == 1 || chans[0].name == "Y") slot_of[0] = slot_of[1] = slot_of[2] = 0;
        else { err = "EXR: no R/G/B channels"; return false; }
    }
    const int nblocks = (h + lines_per_block - 1) / lines_per_block;
    if (!need(pos, (size_t)nblocks * 8)) { err = "EXR: truncated offset table"; return false; }
    std::vector<uint64_t> offsets((size_t)nblocks);
    memcpy(offsets.data(), &f[pos], (size_t)nblocks * 8);
    rgba.assign((size_t)w * h * 4, 0.0f);
    for (size_t i = 3; i < rgba.size(); i += 4) rgba[i] = 1.0f;
    // The chunks (1 or 16 scanlines each) are independent: decode them on all host threads.  The first error in chunk order
    // is the one reported, as in a sequential pass.
    std::vector<std::string> block_err((size_t)nblocks);
    std::atomic<int> next_block(0);
    auto decode = [&]() {
        std::vector<uint8_t> block;
        for (int b = next_block++; b < nblocks; b = next_block++) {
            std::string& berr = block_err[(size_t)b];
            size_t p = (size_t)offsets[(size_t)b];
            if (!need(p, 8)) { berr = "EXR: bad chunk offset"; continue; }
            int y0 = (int)rd32(p) - dw[1];
            uint32_t csize = rd32(p + 4);
            p += 8;
            if (!need(p, csize) || y0 < 0 || y0 >= h) { berr = "EXR: bad chunk"; continue; }
            int nl = std::min(lines_per_block, h - y0);
            size_t expected = line_bytes * (size_t)nl;
            if (compression == 0 || csize == expected) block.assign(&f[p], &f[p] + csize);
            else if (compression == 1) {
                if (!exr_rle_decode(&f[p], csize, block, expected)) { berr = "EXR: RLE decode failed"; continue; }
                exr_unpredict_interleave(block);
            } else {
                if (!zlib_inflate(&f[p], csize, block, expected, berr)) { if (berr.empty()) berr = "EXR: inflate failed"; continue; }
                exr_unpredict_interleave(block);
            }
            if (block.size() != expected) { berr = "EXR: unexpected chunk size"; continue; }
            for (int l = 0; l < nl; ++l) {
                const uint8_t* line = block.data() + line_bytes * (size_t)l;
                float* dst = &rgba[(size_t)(y0 + l) * w * 4];
                for (int s = 0; s < 4; ++s) {
                    int c = slot_of[s];
                    if (c < 0) continue;
                    const uint8_t* src = line + chan_off[(size_t)c];
                    for (int x = 0; x < w; ++x) {
                        float v;
                        if (chans[(size_t)c].type == 1) { uint16_t hb; memcpy(&hb, src + (size_t)x * 2, 2); v = half_to_float(hb); }
                        else if (chans[(size_t)c].type == 2) memcpy(&v, src + (size_t)x * 4, 4);
                        else { uint32_t u; memcpy(&u, src + (size_t)x * 4, 4); v = (float)u; }
                        dst[(size_t)x * 4 + s] = v;
                    }
                }
            }
        }
    };
    {
        unsigned nthreads = std::thread::hardware_concurrency();
        if (nthreads == 0) nthreads = 1;
        if ((int)nthreads > nblocks) nthreads = (unsigned)nblocks;
        if ((size_t)w * h < (1u << 16)) nthreads = 1;
        std::vector<std::thread> pool;
        for (unsigned k = 1; k < nthreads; ++k) pool.emplace_back(decode);
        decode();
        for (std::thread& th : pool) th.join();
    }
    for (int b = 0; b < nblocks; ++b) if (!block_err[(size_t)b].empty()) { err = block_err[(size_t)b]; return false; }
    return true;
}

}  // namespace ptb
