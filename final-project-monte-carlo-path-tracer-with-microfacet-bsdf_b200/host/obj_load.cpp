// host/obj_load.cpp — mesh input for the scene assembler.
//
// The reference feeds MeshTriangle from objl::Loader (src/OBJ_Loader.hpp) and then consumes
// `LoadedMeshes[0].Vertices` three at a time, ignoring the index buffer
// (src/Triangle.hpp:99-124): triangle k is simply vertices 3k..3k+2 of the stream that the
// `f` lines append in file order (src/OBJ_Loader.hpp:493-505).  This reader produces that
// stream directly: positions via strtof (std::stof in the reference, OBJ_Loader.hpp:463-465),
// texture coordinates for v/vt and v/vt/vn corners, (0,0) otherwise (OBJ_Loader.hpp:664-699),
// negative indices relative to the end (OBJ_Loader.hpp:338-345), `vn` ignored, and only the
// first mesh of a multi-object file (the reference asserts there is exactly one,
// src/Triangle.hpp:90).
//
// `.b2m` is this repository's binary triangle pack (the same stream as raw float32) so the
// named scenes can travel to machines that do not have the reference's models directory.
#include "geom.hpp"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

namespace b2pt_host {

namespace {
const char *skip_ws(const char *p) {
    while (*p == ' ' || *p == '\t') ++p;
    return p;
}
// first whitespace-delimited token of a line
std::string first_token(const char *line, const char **rest) {
    const char *p = skip_ws(line);
    const char *q = p;
    while (*q && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n') ++q;
    *rest = q;
    return std::string(p, q);
}
bool parse_floats(const char *p, float *out, int n) {
    for (int i = 0; i < n; ++i) {
        char *end = nullptr;
        out[i] = std::strtof(p, &end);
        if (end == p) return false;
        p = end;
    }
    return true;
}
}  // namespace

bool load_obj_stream(const std::string &path, MeshData &out, std::string &err) {
    if (path.size() < 4 || path.substr(path.size() - 4) != ".obj") {  // OBJ_Loader.hpp:366-368
        err = "not an .obj path: " + path;
        return false;
    }
    std::ifstream f(path);
    if (!f.is_open()) {
        err = "cannot open " + path;
        return false;
    }
    std::vector<float> P, T;
    out.pos.clear();
    out.uv.clear();
    std::string line;
    while (std::getline(f, line)) {
        const char *rest;
        std::string tok = first_token(line.c_str(), &rest);
        if (tok == "v") {
            float v[3];
            if (!parse_floats(rest, v, 3)) { err = "bad v line in " + path; return false; }
            P.insert(P.end(), v, v + 3);
        } else if (tok == "vt") {
            float v[2];
            if (!parse_floats(rest, v, 2)) { err = "bad vt line in " + path; return false; }
            T.insert(T.end(), v, v + 2);
        } else if (tok == "f") {
            const char *p = rest;
            for (;;) {
                p = skip_ws(p);
                if (!*p || *p == '\r' || *p == '\n') break;
                const char *q = p;
                while (*q && *q != ' ' && *q != '\t' && *q != '\r' && *q != '\n') ++q;
                std::string corner(p, q);
                p = q;
                // v, v/vt, v//vn, v/vt/vn
                int vi = 0, ti = 0;
                bool has_t = false;
                size_t s1 = corner.find('/');
                vi = std::atoi(corner.substr(0, s1).c_str());
                if (s1 != std::string::npos) {
                    size_t s2 = corner.find('/', s1 + 1);
                    std::string ts = corner.substr(s1 + 1, s2 == std::string::npos ? std::string::npos : s2 - s1 - 1);
                    if (!ts.empty()) { has_t = true; ti = std::atoi(ts.c_str()); }
                }
                int np = (int)(P.size() / 3), nt = (int)(T.size() / 2);
                int pi = vi < 0 ? np + vi : vi - 1;
                if (pi < 0 || pi >= np) { err = "face index out of range in " + path; return false; }
                out.pos.insert(out.pos.end(), P.begin() + 3 * pi, P.begin() + 3 * pi + 3);
                if (has_t) {
                    int tj = ti < 0 ? nt + ti : ti - 1;
                    if (tj < 0 || tj >= nt) { err = "texcoord index out of range in " + path; return false; }
                    out.uv.push_back(T[2 * tj]);
                    out.uv.push_back(T[2 * tj + 1]);
                } else {
                    out.uv.push_back(0.f);
                    out.uv.push_back(0.f);
                }
            }
        } else if ((tok == "o" || tok == "g" || tok == "usemtl") && !out.pos.empty()) {
            break;  // a second mesh would start here; the reference only uses LoadedMeshes[0]
        }
    }
    // MeshTriangle walks the stream in steps of three; an incomplete tail is dropped.
    size_t nv = out.pos.size() / 3;
    nv -= nv % 3;
    out.pos.resize(nv * 3);
    out.uv.resize(nv * 2);
    if (nv == 0) { err = "no faces in " + path; return false; }
    return true;
}

// .b2m: "B2M1" | u32 n_vertices | float32 pos[3*n] | float32 uv[2*n]
bool load_b2m(const std::string &path, MeshData &out, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    char magic[4];
    uint32_t n = 0;
    bool ok = std::fread(magic, 1, 4, f) == 4 && std::memcmp(magic, "B2M1", 4) == 0 && std::fread(&n, 4, 1, f) == 1;
    if (ok) {
        out.pos.resize((size_t)n * 3);
        out.uv.resize((size_t)n * 2);
        ok = std::fread(out.pos.data(), 4, out.pos.size(), f) == out.pos.size() &&
             std::fread(out.uv.data(), 4, out.uv.size(), f) == out.uv.size();
    }
    std::fclose(f);
    if (!ok) err = "bad .b2m file " + path;
    return ok;
}
bool save_b2m(const std::string &path, const MeshData &m, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    uint32_t n = (uint32_t)(m.pos.size() / 3);
    std::fwrite("B2M1", 1, 4, f);
    std::fwrite(&n, 4, 1, f);
    std::fwrite(m.pos.data(), 4, m.pos.size(), f);
    std::fwrite(m.uv.data(), 4, m.uv.size(), f);
    std::fclose(f);
    return true;
}
// Triangle-soup OBJ whose %.9g literals read back (strtof / std::stof) to the same float32s.
bool save_obj_soup(const std::string &path, const MeshData &m, std::string &err) {
    FILE *f = std::fopen(path.c_str(), "w");
    if (!f) { err = "cannot write " + path; return false; }
    size_t n = m.pos.size() / 3;
    for (size_t i = 0; i < n; ++i) std::fprintf(f, "v %.9g %.9g %.9g\n", m.pos[3 * i], m.pos[3 * i + 1], m.pos[3 * i + 2]);
    for (size_t i = 0; i < n; ++i) std::fprintf(f, "vt %.9g %.9g\n", m.uv[2 * i], m.uv[2 * i + 1]);
    for (size_t i = 0; i + 2 < n; i += 3)
        std::fprintf(f, "f %zu/%zu %zu/%zu %zu/%zu\n", i + 1, i + 1, i + 2, i + 2, i + 3, i + 3);
    std::fclose(f);
    return true;
}

}  // namespace b2pt_host
