// host/geom.hpp — float3 algebra for the host-side scene assembler.
//
// The reference does all vector algebra through Eigen's Vector3f; the operation order
// below is the one fixed-size-3 Eigen expressions evaluate to (size-3 reductions associate
// as a0 + (a1 + a2); normalized() divides by sqrt(squaredNorm)), because triangle edges,
// normals, areas and boxes computed here must be bit-identical to what Triangle::Triangle
// (src/Triangle.hpp:50-56) and Bounds3 (src/Bounds3.hpp) produce.  Compile with
// -ffp-contract=off.
#pragma once
#include <cmath>
#include <string>
#include <vector>

namespace b2pt_host {

struct V3 {
    float x, y, z;
    float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
inline V3 v3(float x, float y, float z) { return V3{x, y, z}; }
inline V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(float s, V3 a) { return V3{s * a.x, s * a.y, s * a.z}; }
inline V3 operator*(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline V3 operator/(V3 a, float s) { return V3{a.x / s, a.y / s, a.z / s}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + (a.y * b.y + a.z * b.z); }
inline float squared_norm(V3 a) { return dot(a, a); }
inline float norm(V3 a) { return std::sqrt(squared_norm(a)); }
inline V3 normalized(V3 a) {
    float n2 = squared_norm(a);
    return n2 > 0.f ? a / std::sqrt(n2) : a;
}
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

// Bounds3 (src/Bounds3.hpp:14-31,87-94,128-140): fmin/fmax per component.
struct Box {
    V3 mn, mx;
};
inline Box box_of_points(V3 p1, V3 p2) {
    return Box{V3{fminf(p1.x, p2.x), fminf(p1.y, p2.y), fminf(p1.z, p2.z)},
               V3{fmaxf(p1.x, p2.x), fmaxf(p1.y, p2.y), fmaxf(p1.z, p2.z)}};
}
inline Box box_union(const Box &a, const Box &b) {
    return Box{V3{fminf(a.mn.x, b.mn.x), fminf(a.mn.y, b.mn.y), fminf(a.mn.z, b.mn.z)},
               V3{fmaxf(a.mx.x, b.mx.x), fmaxf(a.mx.y, b.mx.y), fmaxf(a.mx.z, b.mx.z)}};
}
inline Box box_union(const Box &a, V3 p) {
    return Box{V3{fminf(a.mn.x, p.x), fminf(a.mn.y, p.y), fminf(a.mn.z, p.z)},
               V3{fmaxf(a.mx.x, p.x), fmaxf(a.mx.y, p.y), fmaxf(a.mx.z, p.z)}};
}
// Bounds3() (src/Bounds3.hpp:16-21): (float)DBL_MAX = +inf, (float)lowest = -inf.
inline Box box_empty() { return Box{V3{INFINITY, INFINITY, INFINITY}, V3{-INFINITY, -INFINITY, -INFINITY}}; }
// Centroid() = 0.5 * pMin + 0.5 * pMax, the double 0.5 converted to float first.
inline V3 box_centroid(const Box &b) { return 0.5f * b.mn + 0.5f * b.mx; }
inline bool box_equal(const Box &a, const Box &b) {
    return a.mn.x == b.mn.x && a.mn.y == b.mn.y && a.mn.z == b.mn.z && a.mx.x == b.mx.x && a.mx.y == b.mx.y &&
           a.mx.z == b.mx.z;
}

// One triangle as Triangle::Triangle leaves it (src/Triangle.hpp:43-56).
struct Tri {
    V3 v0, v1, v2, e1, e2, n;
    float area;
    float uv[6];  // t0 t1 t2
};
inline Tri make_tri(V3 a, V3 b, V3 c) {
    Tri t{};
    t.v0 = a; t.v1 = b; t.v2 = c;
    t.e1 = b - a;
    t.e2 = c - a;
    V3 cr = cross(t.e1, t.e2);
    t.n = normalized(cr);
    t.area = norm(cr) * 0.5f;
    for (float &u : t.uv) u = 0.f;
    return t;
}
inline Box tri_box(const Tri &t) { return box_union(box_of_points(t.v0, t.v1), t.v2); }  // src/Triangle.hpp:220

// Face-vertex stream of an OBJ file in the order the reference's loader emits it.
struct MeshData {
    std::vector<float> pos;  // 3 per vertex, 3 vertices per triangle
    std::vector<float> uv;   // 2 per vertex
};
bool load_obj_stream(const std::string &path, MeshData &out, std::string &err);
bool load_b2m(const std::string &path, MeshData &out, std::string &err);
bool save_b2m(const std::string &path, const MeshData &m, std::string &err);
bool save_obj_soup(const std::string &path, const MeshData &m, std::string &err);

}  // namespace b2pt_host
