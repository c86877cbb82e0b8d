// host/png_io.cpp — PNG in/out for the host program, on top of zlib.
//
// The reference decodes the environment map with lodepng::decode (RGBA8 output,
// src/Scene.hpp:39-57) and writes the frame with lodepng::encode (src/Renderer.cpp:104-105).
// Only those two uses are needed: decode any non-interlaced PNG to RGBA8, encode RGBA8.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>
#include <zlib.h>

#include "b2pt_host.h"

namespace b2pt_host {
void set_error(const std::string &e);
}
using b2pt_host::set_error;

namespace {
uint32_t be32(const unsigned char *p) { return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3]; }
void put_be32(std::vector<unsigned char> &v, uint32_t x) {
    v.push_back((unsigned char)(x >> 24)); v.push_back((unsigned char)(x >> 16));
    v.push_back((unsigned char)(x >> 8)); v.push_back((unsigned char)x);
}
void put_chunk(std::vector<unsigned char> &out, const char *type, const unsigned char *data, size_t n) {
    put_be32(out, (uint32_t)n);
    size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    if (n) out.insert(out.end(), data, data + n);
    uint32_t crc = (uint32_t)crc32(0L, out.data() + start, (uInt)(n + 4));
    put_be32(out, crc);
}
int paeth(int a, int b, int c) {
    int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
    return (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
}
}  // namespace

extern "C" {

// Renderer.cpp:96-102: raw = clamp(0, 255, 255 * pow(c, 0.45)) assigned to unsigned char.
// clamp(lo,hi,v) = max(lo, min(hi, v)) with std::min/max, so NaN -> 255 (global.hpp:16-18);
// the float -> unsigned char conversion truncates.
void b2pt_host_tonemap_rgba8(const float *rgb, int n_pixels, unsigned char *rgba) {
    const float inv_gamma = 0.45f;  // `float inv_gamma = 0.45;`
    for (int i = 0; i < n_pixels; ++i) {
        for (int c = 0; c < 3; ++c) {
            // unqualified pow() in Renderer.cpp resolves to the C ::pow(double,double) (the reference's
            // object code calls pow@plt); the double product is narrowed by clamp's `const float &v`.
            float v = (float)(255 * ::pow((double)rgb[3 * i + c], (double)inv_gamma));
            float m = (v < 255.f) ? v : 255.f;                    // std::min(hi, v)
            float r = (0.f < m) ? m : 0.f;                        // std::max(lo, m)
            rgba[4 * i + c] = (unsigned char)r;
        }
        rgba[4 * i + 3] = 255;
    }
}

int b2pt_host_write_png_rgba8(const char *path, const unsigned char *rgba, int width, int height) {
    if (!path || !rgba || width <= 0 || height <= 0) { set_error("write_png: bad arguments"); return -1; }
    std::vector<unsigned char> raw((size_t)height * ((size_t)width * 4 + 1));
    for (int y = 0; y < height; ++y) {
        unsigned char *row = raw.data() + (size_t)y * ((size_t)width * 4 + 1);
        row[0] = 0;  // filter: none
        std::memcpy(row + 1, rgba + (size_t)y * width * 4, (size_t)width * 4);
    }
    uLongf zn = compressBound((uLong)raw.size());
    std::vector<unsigned char> z(zn);
    // level 3: 0.07 s instead of 0.16 s for a 1080p frame at +10 % file size (the render itself takes 1.2 s)
    if (compress2(z.data(), &zn, raw.data(), (uLong)raw.size(), 3) != Z_OK) { set_error("write_png: deflate failed"); return -1; }
    std::vector<unsigned char> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    unsigned char ihdr[13];
    ihdr[0] = (unsigned char)(width >> 24); ihdr[1] = (unsigned char)(width >> 16); ihdr[2] = (unsigned char)(width >> 8); ihdr[3] = (unsigned char)width;
    ihdr[4] = (unsigned char)(height >> 24); ihdr[5] = (unsigned char)(height >> 16); ihdr[6] = (unsigned char)(height >> 8); ihdr[7] = (unsigned char)height;
    ihdr[8] = 8; ihdr[9] = 6; ihdr[10] = 0; ihdr[11] = 0; ihdr[12] = 0;
    put_chunk(out, "IHDR", ihdr, 13);
    put_chunk(out, "IDAT", z.data(), zn);
    put_chunk(out, "IEND", nullptr, 0);
    FILE *f = std::fopen(path, "wb");
    if (!f) { set_error(std::string("write_png: cannot open ") + path); return -1; }
    size_t w = std::fwrite(out.data(), 1, out.size(), f);
    std::fclose(f);
    if (w != out.size()) { set_error("write_png: short write"); return -1; }
    return 0;
}

// Decodes any PNG lodepng::decode accepts for the environment map (grey / RGB / palette / grey+alpha / RGBA, 1-16 bits,
// non-interlaced or Adam7) to RGBA8.  Never throws across the C boundary; the header is validated before anything is sized
// from it.
static int read_png_impl(const char *path, unsigned char **rgba, unsigned *width, unsigned *height) {
    FILE *f = std::fopen(path, "rb");
    if (!f) { set_error(std::string("read_png: cannot open ") + path); return -1; }
    std::vector<unsigned char> buf;
    unsigned char tmp[65536];
    size_t n;
    while ((n = std::fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + n);
    std::fclose(f);
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    if (buf.size() < 8 || std::memcmp(buf.data(), sig, 8) != 0) { set_error("read_png: not a PNG"); return -1; }
    uint32_t W = 0, H = 0;
    int depth = 0, ctype = 0, interlace = 0;
    bool have_ihdr = false;
    std::vector<unsigned char> idat, plte, trns;
    size_t pos = 8;
    while (pos + 12 <= buf.size()) {
        uint32_t len = be32(&buf[pos]);
        const unsigned char *type = &buf[pos + 4];
        const unsigned char *data = &buf[pos + 8];
        if ((size_t)len > buf.size() || pos + 12 + len > buf.size()) break;
        if (!std::memcmp(type, "IHDR", 4)) {
            if (len < 13) { set_error("read_png: short IHDR"); return -1; }
            W = be32(data); H = be32(data + 4); depth = data[8]; ctype = data[9]; interlace = data[12];
            have_ihdr = true;
        } else if (!std::memcmp(type, "PLTE", 4)) plte.assign(data, data + len);
        else if (!std::memcmp(type, "tRNS", 4)) trns.assign(data, data + len);
        else if (!std::memcmp(type, "IDAT", 4)) idat.insert(idat.end(), data, data + len);
        else if (!std::memcmp(type, "IEND", 4)) break;
        pos += 12 + len;
    }
    if (!have_ihdr || !W || !H) { set_error("read_png: no image header"); return -1; }
    if (W > 65536u || H > 65536u || (uint64_t)W * H > (1ull << 28)) { set_error("read_png: image too large"); return -1; }
    if (interlace > 1) { set_error("read_png: unknown interlace method"); return -1; }
    const int channels = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
    if (!channels) { set_error("read_png: bad colour type"); return -1; }
    // bit depths the PNG specification allows per colour type
    const bool depth_ok = (ctype == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) ||
                          (ctype == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                          ((ctype == 2 || ctype == 4 || ctype == 6) && (depth == 8 || depth == 16));
    if (!depth_ok) { set_error("read_png: bad bit depth for the colour type"); return -1; }
    const size_t bpp_bits = (size_t)channels * depth, bpp = (bpp_bits + 7) / 8;
    // the (up to seven) sub-images of the stream: one for a non-interlaced file, the Adam7 passes otherwise
    struct Pass { uint32_t xs, ys, dx, dy, w, h; size_t stride; };
    std::vector<Pass> passes;
    if (!interlace) passes.push_back({0, 0, 1, 1, W, H, (W * bpp_bits + 7) / 8});
    else {
        static const uint32_t XS[7] = {0, 4, 0, 2, 0, 1, 0}, YS[7] = {0, 0, 4, 0, 2, 0, 1}, DX[7] = {8, 8, 4, 4, 2, 2, 1}, DY[7] = {8, 8, 8, 4, 4, 2, 2};
        for (int k = 0; k < 7; ++k) {
            const uint32_t pw = W > XS[k] ? (W - XS[k] + DX[k] - 1) / DX[k] : 0, ph = H > YS[k] ? (H - YS[k] + DY[k] - 1) / DY[k] : 0;
            if (pw && ph) passes.push_back({XS[k], YS[k], DX[k], DY[k], pw, ph, (pw * bpp_bits + 7) / 8});
        }
    }
    size_t raw_size = 0;
    for (const Pass &p : passes) raw_size += (p.stride + 1) * p.h;
    std::vector<unsigned char> raw(raw_size);
    uLongf rn = (uLongf)raw.size();
    if (idat.empty() || uncompress(raw.data(), &rn, idat.data(), (uLong)idat.size()) != Z_OK || rn != raw.size()) { set_error("read_png: inflate failed"); return -1; }
    unsigned char *out = (unsigned char *)std::malloc((size_t)W * H * 4);
    if (!out) { set_error("read_png: out of memory"); return -1; }
    auto sample = [&](const unsigned char *row, size_t idx) -> unsigned {  // idx-th sample of the row
        if (depth == 8) return row[idx];
        if (depth == 16) return row[2 * idx];  // high byte
        size_t bit = idx * depth;
        unsigned v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1u << depth) - 1);
        return ctype == 3 ? v : v * 255u / ((1u << depth) - 1);
    };
    size_t off = 0;
    std::vector<unsigned char> img;
    for (const Pass &ps : passes) {
        img.assign(ps.stride * ps.h, 0);
        for (uint32_t y = 0; y < ps.h; ++y) {
            const unsigned char *in = raw.data() + off + y * (ps.stride + 1);
            unsigned char *cur = img.data() + y * ps.stride;
            const unsigned char *up = y ? cur - ps.stride : nullptr;
            int ft = in[0];
            for (size_t x = 0; x < ps.stride; ++x) {
                int a = x >= bpp ? cur[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
                int v = in[1 + x];
                switch (ft) {
                case 0: break;
                case 1: v += a; break;
                case 2: v += b; break;
                case 3: v += (a + b) >> 1; break;
                case 4: v += paeth(a, b, c); break;
                default: std::free(out); set_error("read_png: bad filter"); return -1;
                }
                cur[x] = (unsigned char)v;
            }
        }
        off += (ps.stride + 1) * ps.h;
        for (uint32_t y = 0; y < ps.h; ++y) {
            const unsigned char *row = img.data() + y * ps.stride;
            for (uint32_t x = 0; x < ps.w; ++x) {
                unsigned char *o = out + ((size_t)(ps.ys + y * ps.dy) * W + (ps.xs + x * ps.dx)) * 4;
                switch (ctype) {
                case 0: { unsigned g = sample(row, x); o[0] = o[1] = o[2] = (unsigned char)g; o[3] = 255; break; }
                case 2: o[0] = (unsigned char)sample(row, 3 * x); o[1] = (unsigned char)sample(row, 3 * x + 1); o[2] = (unsigned char)sample(row, 3 * x + 2); o[3] = 255; break;
                case 3: {
                    unsigned i = sample(row, x);
                    if (3 * i + 2 < plte.size()) { o[0] = plte[3 * i]; o[1] = plte[3 * i + 1]; o[2] = plte[3 * i + 2]; }
                    else o[0] = o[1] = o[2] = 0;
                    o[3] = i < trns.size() ? trns[i] : 255;
                    break;
                }
                case 4: { unsigned g = sample(row, 2 * x); o[0] = o[1] = o[2] = (unsigned char)g; o[3] = (unsigned char)sample(row, 2 * x + 1); break; }
                default: for (int c = 0; c < 4; ++c) o[c] = (unsigned char)sample(row, 4 * x + c);
                }
            }
        }
    }
    *rgba = out; *width = W; *height = H;
    return 0;
}
int b2pt_host_read_png_rgba8(const char *path, unsigned char **rgba, unsigned *width, unsigned *height) {
    if (!path || !rgba || !width || !height) { set_error("read_png: bad arguments"); return -1; }
    try {
        return read_png_impl(path, rgba, width, height);
    } catch (const std::exception &e) {  // bad_alloc on a hostile header, nothing may cross the C boundary
        set_error(std::string("read_png: ") + e.what());
        return -1;
    }
}

void b2pt_host_free(void *p) { std::free(p); }

}  // extern "C"
