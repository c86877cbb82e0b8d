// csrc/pt_pack.hpp — host-side repacking of a b2pt_scene_desc into the arrays a pt::SceneView
// points at (materials with the hasEmission flag resolved, env texels padded to float4).
// Plain C++: used by the CUDA library before its cudaMemcpy calls and by tests/hostcheck.
#pragma once
#include <cmath>
#include <cstring>
#include <algorithm>
#include <string>
#include <utility>
#include <vector>

#include "b2pt.h"
#include "pt_build.hpp"
#include "pt_math.cuh"

namespace pt {

struct PackedScene {
    std::vector<Material> mats;
    std::vector<float4> env;
    std::vector<b2pt_node> nodes_ref;   // the reference's topology, EMPTY boxes rewritten to NaN
    std::vector<b2pt_node> nodes_fast;  // binned-SAH tree over the same leaves (pt_build.hpp)
    int fast_depth = 0;
    float light_sphere[4] = {0.f, 0.f, 0.f, -1.f};  // centre, radius of a sphere around every light primitive (radius < 0: none)
    float light_box[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // and the axis-aligned box around them (min, max)
    QuadTree quads;                     // nodes_fast collapsed four-wide; empty when its stack need exceeds kStackSize4
    // Per light-tree node (leaves only): the primitives that can be hit within the visibility window of a point sampled
    // on that light triangle — the triangle itself first, then every primitive whose box comes within `delta` of its box.
    // Entry = (bmin, prim id) (bmax, kind): the primitive's own reference leaf box, tested before the primitive.
    std::vector<float4> tri;  // (v0, e1, e2) per primitive, interleaved
    std::vector<float4> flat; // small scenes (<= pt::kFlatMax primitives): per primitive (leaf bmin | id, leaf bmax | kind, v0, e1, e2)
    std::vector<float4> lt_entries;
    std::vector<int> lt_off, lt_cnt;  // cnt < 0: too many neighbours, the window is searched by traversal instead
};
constexpr int kMaxLightNeighbours = 12;

constexpr int kMaxLightSamples = B2PT_MAX_LIGHT_SAMPLES;  // n_dir_sample: keeps visibility slots (rays x n_dir) and their phase bit inside 32 bits
// Depth of the sibling-pair tree (root = 0) by an explicit walk; -1 when a node is reachable twice or the links loop (a tree
// visits each of its n nodes once, so more than n visits means a cycle or a shared subtree).
inline int tree_depth(const b2pt_scene_desc *d) {
    std::vector<std::pair<uint32_t, int>> stack;
    stack.push_back({0u, 0});
    size_t visited = 0;
    int depth = 0;
    while (!stack.empty()) {
        const auto [i, dep] = stack.back();
        stack.pop_back();
        if (++visited > d->n_nodes) return -1;
        depth = std::max(depth, dep);
        const b2pt_node &n = d->nodes[i];
        if (n.kind != B2PT_NODE_INTERIOR) continue;
        stack.push_back({2 * n.a, dep + 1});
        stack.push_back({2 * n.a + 1, dep + 1});
    }
    return depth;
}

inline bool validate_material(const b2pt_material &m, std::string &err) {
    // MaterialType has four values (src/Material.hpp:13-18); the kernels file rays under 1 + 2 * type + survives
    if (m.type < B2PT_SMOOTH_CONDUCTOR || m.type > B2PT_ROUGH_DIELECTRIC) { err = "material type outside 0..3"; return false; }
    return true;
}

inline bool validate_scene(const b2pt_scene_desc *d, std::string &err) {
    if (!d) { err = "scene is NULL"; return false; }
    if (d->n_nodes < 2 || (d->n_nodes & 1u) || !d->nodes) { err = "scene needs an even, non-zero number of nodes"; return false; }
    if (d->n_prims == 0 || !d->prim_v0 || !d->prim_e1 || !d->prim_e2 || !d->prim_v1v2 || !d->prim_normal || !d->prim_uv ||
        !d->prim_material || !d->prim_kind) { err = "scene primitive arrays missing"; return false; }
    if (d->n_materials == 0 || d->n_materials > B2PT_MAX_MATERIALS || !d->materials) { err = "bad material table"; return false; }
    for (uint32_t i = 0; i < d->n_materials; ++i)
        if (!validate_material(d->materials[i], err)) return false;
    if (d->n_lights > B2PT_MAX_LIGHTS) { err = "too many lights"; return false; }
    if (d->use_env_map && (!d->env_rgb || d->env_width == 0 || d->env_height == 0)) { err = "env map enabled without texels"; return false; }
    if (d->n_dir_sample < 1 || d->n_dir_sample > kMaxLightSamples) { err = "n_dir_sample must be in 1..1024"; return false; }
    for (uint32_t i = 0; i < d->n_nodes; ++i) {
        const b2pt_node &n = d->nodes[i];
        if (n.kind == B2PT_NODE_INTERIOR) {
            if (2 * (uint64_t)n.a + 1 >= d->n_nodes) { err = "node child index out of range"; return false; }
        } else if (n.kind == B2PT_NODE_TRIANGLE || n.kind == B2PT_NODE_SPHERE) {
            if (n.a >= d->n_prims) { err = "node primitive index out of range"; return false; }
            // the walk picks the primitive test from the leaf's kind, the shading kernels from prim_kind: they must agree
            if (d->prim_kind[n.a] != n.kind) { err = "leaf kind differs from prim_kind of its primitive"; return false; }
        } else if (n.kind != B2PT_NODE_EMPTY) { err = "bad node kind"; return false; }
    }
    // the depth the walk stacks must hold is computed here, never taken from the caller (d->max_depth is informational)
    const int depth = tree_depth(d);
    if (depth < 0) { err = "node links do not form a tree"; return false; }
    if (depth >= B2PT_MAX_TREE_DEPTH || depth + 2 >= kStackSize) { err = "tree too deep for the traversal stack (B2PT_MAX_TREE_DEPTH)"; return false; }
    for (uint32_t i = 0; i < d->n_prims; ++i) {
        if (d->prim_material[i] >= d->n_materials) { err = "primitive material index out of range"; return false; }
        if (d->prim_kind[i] != B2PT_NODE_TRIANGLE && d->prim_kind[i] != B2PT_NODE_SPHERE) { err = "bad prim_kind"; return false; }
    }
    for (uint32_t i = 0; i < d->n_light_nodes; ++i) {
        int l = d->light_node_left[i], r = d->light_node_right[i], p = d->light_node_prim[i];
        if (l >= (int)d->n_light_nodes || r >= (int)d->n_light_nodes) { err = "light tree index out of range"; return false; }
        if ((l < 0 || r < 0) && (p < 0 || p >= (int)d->n_prims)) { err = "light tree leaf without primitive"; return false; }
    }
    for (uint32_t i = 0; i < d->n_lights; ++i)
        if (d->light_root[i] >= d->n_light_nodes || d->light_material[i] >= d->n_materials) { err = "bad light entry"; return false; }
    return true;
}

inline void pack_scene(const b2pt_scene_desc *d, PackedScene &out, bool build_fast_tree = true, const BuildOptions *opt = nullptr) {
    out.tri.resize(3 * (size_t)d->n_prims);
    for (size_t i = 0; i < d->n_prims; ++i) {
        std::memcpy(&out.tri[3 * i], d->prim_v0 + 4 * i, 16);
        std::memcpy(&out.tri[3 * i + 1], d->prim_e1 + 4 * i, 16);
        std::memcpy(&out.tri[3 * i + 2], d->prim_e2 + 4 * i, 16);
    }
    out.flat.clear();
    if (d->n_prims <= (uint32_t)kFlatMax) {
        // every primitive must own exactly one leaf (its box is what decides whether the reference tests it)
        std::vector<int> owner(d->n_prims, -1);
        bool ok = true;
        for (uint32_t i = 0; i < d->n_nodes; ++i) {
            const b2pt_node &n = d->nodes[i];
            if (n.kind != B2PT_NODE_TRIANGLE && n.kind != B2PT_NODE_SPHERE) continue;
            if (owner[n.a] >= 0) ok = false;
            owner[n.a] = (int)i;
        }
        for (uint32_t p = 0; p < d->n_prims; ++p) ok = ok && owner[p] >= 0;
        if (ok) {
            out.flat.resize(5 * (size_t)d->n_prims);
            for (uint32_t p = 0; p < d->n_prims; ++p) {
                const b2pt_node &n = d->nodes[owner[p]];
                float4 lo = make_float4(n.bmin[0], n.bmin[1], n.bmin[2], 0.f), hi = make_float4(n.bmax[0], n.bmax[1], n.bmax[2], 0.f);
                const uint32_t kind = n.kind;
                std::memcpy(&lo.w, &p, 4);
                std::memcpy(&hi.w, &kind, 4);
                out.flat[5 * (size_t)p] = lo; out.flat[5 * (size_t)p + 1] = hi;
                std::memcpy(&out.flat[5 * (size_t)p + 2], d->prim_v0 + 4 * (size_t)p, 16);
                std::memcpy(&out.flat[5 * (size_t)p + 3], d->prim_e1 + 4 * (size_t)p, 16);
                std::memcpy(&out.flat[5 * (size_t)p + 4], d->prim_e2 + 4 * (size_t)p, 16);
            }
        }
    }
    out.nodes_ref.assign(d->nodes, d->nodes + d->n_nodes);
    for (auto &n : out.nodes_ref)
        if (n.kind == B2PT_NODE_EMPTY)
            for (int k = 0; k < 3; ++k) { n.bmin[k] = NAN; n.bmax[k] = NAN; }
    out.nodes_fast.clear();
    out.fast_depth = 0;
    if (build_fast_tree) {
        SahBuilder b;
        if (opt) b.configure(*opt);
        b.run(d);
        if (b.max_depth + 2 < kStackSize) { out.nodes_fast.swap(b.out); out.fast_depth = b.max_depth; }
    }
    if (out.nodes_fast.empty()) { out.nodes_fast = out.nodes_ref; out.fast_depth = (int)d->max_depth; }
    QuadCollapser(out.nodes_fast, out.quads).run();
    // the walks need: the one-lane stack bound, a depth that fits the four-lane walk's per-lane stacks and six-bit depth key, and
    // quad indices below 2^24 (the stack entries carry the depth in the top byte)
    if (out.quads.stack_need + 2 >= kStackSize4 || out.quads.depth + 2 >= kStackSize4 || out.quads.depth >= 60 ||
        out.quads.nodes.size() / 4 >= ((size_t)1 << 24))
        out.quads.nodes.clear();
    {
        // leaf boxes by primitive id
        std::vector<BuildBox> pb(d->n_prims, box_empty_b());
        BuildBox all = box_empty_b();
        for (uint32_t i = 0; i < d->n_nodes; ++i) {
            const b2pt_node &n = d->nodes[i];
            if (n.kind != B2PT_NODE_TRIANGLE && n.kind != B2PT_NODE_SPHERE) continue;
            for (int k = 0; k < 3; ++k) { pb[n.a].mn[k] = n.bmin[k]; pb[n.a].mx[k] = n.bmax[k]; }
            box_grow(all, pb[n.a]);
        }
        float diag = 0.f;
        for (int k = 0; k < 3; ++k) diag += (all.mx[k] - all.mn[k]) * (all.mx[k] - all.mn[k]);
        // A witness hit lies within EPSILON (+ float rounding of dist, ws and the hit point, ~1e-6 of the scene size) of the
        // sampled point; delta is three orders of magnitude above that.
        const float delta = 1e-3f * std::sqrt(diag) + 0.01f;
        out.lt_entries.clear();
        out.lt_off.assign(d->n_light_nodes, 0);
        out.lt_cnt.assign(d->n_light_nodes, 0);
        auto entry = [&](uint32_t prim) {
            float4 a = make_float4(pb[prim].mn[0], pb[prim].mn[1], pb[prim].mn[2], 0.f), b = make_float4(pb[prim].mx[0], pb[prim].mx[1], pb[prim].mx[2], 0.f);
            uint32_t kind = d->prim_kind[prim];
            std::memcpy(&a.w, &prim, 4);
            std::memcpy(&b.w, &kind, 4);
            out.lt_entries.push_back(a);
            out.lt_entries.push_back(b);
        };
        for (uint32_t ln = 0; ln < d->n_light_nodes; ++ln) {
            if (d->light_node_left[ln] >= 0 && d->light_node_right[ln] >= 0) continue;
            const uint32_t L = (uint32_t)d->light_node_prim[ln];
            out.lt_off[ln] = (int)(out.lt_entries.size() / 2);
            entry(L);
            int cnt = 1;
            for (uint32_t q = 0; q < d->n_prims && cnt >= 0; ++q) {
                if (q == L) continue;
                bool overlap = true;
                for (int k = 0; k < 3; ++k)
                    if (pb[q].mn[k] - delta > pb[L].mx[k] + delta || pb[q].mx[k] + delta < pb[L].mn[k] - delta) overlap = false;
                if (!overlap) continue;
                if (cnt > kMaxLightNeighbours) { cnt = -1; break; }
                entry(q);
                cnt++;
            }
            if (cnt < 0) out.lt_entries.resize((size_t)out.lt_off[ln] * 2);
            out.lt_cnt[ln] = cnt;
        }
        if (out.lt_entries.empty()) out.lt_entries.push_back(make_float4(0, 0, 0, 0));
    }
    {
        // bounding sphere of everything a light sample can fall on: the vertices of the light triangles (Triangle::Sample returns
        // a convex combination of them), or centre +- radius of a sphere
        BuildBox lb = box_empty_b();
        std::vector<float> pts;
        for (uint32_t ln = 0; ln < d->n_light_nodes; ++ln) {
            if (d->light_node_prim[ln] < 0) continue;
            const uint32_t L = (uint32_t)d->light_node_prim[ln];
            if (L >= d->n_prims) continue;
            const float *v0 = d->prim_v0 + 4 * (size_t)L, *e1 = d->prim_e1 + 4 * (size_t)L, *e2 = d->prim_e2 + 4 * (size_t)L, *v12 = d->prim_v1v2 + 6 * (size_t)L;
            if (d->prim_kind[L] == B2PT_NODE_SPHERE) {
                for (int sx = -1; sx <= 1; sx += 2) for (int sy = -1; sy <= 1; sy += 2) for (int sz = -1; sz <= 1; sz += 2) {
                    pts.push_back(v0[0] + sx * v0[3]); pts.push_back(v0[1] + sy * v0[3]); pts.push_back(v0[2] + sz * v0[3]);
                }
            } else {
                for (int k = 0; k < 3; ++k) pts.push_back(v0[k]);
                for (int k = 0; k < 3; ++k) pts.push_back(v12[k]);
                for (int k = 0; k < 3; ++k) pts.push_back(v12[3 + k]);
                for (int k = 0; k < 3; ++k) pts.push_back(v0[k] + e1[k]);  // the same vertices through the edges (whichever the sampler uses)
                for (int k = 0; k < 3; ++k) pts.push_back(v0[k] + e2[k]);
            }
        }
        out.light_sphere[3] = -1.f;
        if (!pts.empty()) {
            for (size_t i = 0; i < pts.size(); i += 3)
                for (int k = 0; k < 3; ++k) { lb.mn[k] = std::fmin(lb.mn[k], pts[i + k]); lb.mx[k] = std::fmax(lb.mx[k], pts[i + k]); }
            double c[3], r2 = 0;
            for (int k = 0; k < 3; ++k) c[k] = 0.5 * ((double)lb.mn[k] + (double)lb.mx[k]);
            for (size_t i = 0; i < pts.size(); i += 3) {
                double q = 0;
                for (int k = 0; k < 3; ++k) q += (pts[i + k] - c[k]) * (pts[i + k] - c[k]);
                r2 = std::max(r2, q);
            }
            const double r = std::sqrt(r2) * 1.0001 + 1e-4;
            if (std::isfinite(r) && std::isfinite(c[0]) && std::isfinite(c[1]) && std::isfinite(c[2])) {
                for (int k = 0; k < 3; ++k) out.light_sphere[k] = (float)c[k];
                out.light_sphere[3] = (float)r;
                for (int k = 0; k < 3; ++k) { out.light_box[k] = lb.mn[k]; out.light_box[3 + k] = lb.mx[k]; }
            }
        }
    }
    out.mats.resize(d->n_materials);
    for (uint32_t i = 0; i < d->n_materials; ++i) {
        const b2pt_material &s = d->materials[i];
        Material &m = out.mats[i];
        m.type = s.type;
        for (int c = 0; c < 3; ++c) { m.emission[c] = s.emission[c]; m.refl[c] = s.base_reflectance[c]; }
        m.ior_a = s.ior_a; m.ior_b = s.ior_b; m.roughness = s.roughness;
        m.textured = s.textured;
        // Material::hasEmission: m_emission.norm() > EPSILON (src/Material.hpp:263)
        float n2 = s.emission[0] * s.emission[0] + (s.emission[1] * s.emission[1] + s.emission[2] * s.emission[2]);
        m.emissive = std::sqrt(n2) > kEps ? 1 : 0;
    }
    out.env.clear();
    if (d->use_env_map) {
        size_t n = (size_t)d->env_width * d->env_height;
        out.env.resize(n);
        for (size_t i = 0; i < n; ++i) out.env[i] = make_float4(d->env_rgb[3 * i], d->env_rgb[3 * i + 1], d->env_rgb[3 * i + 2], 0.f);
    }
}

inline Camera make_camera(const b2pt_camera *c) {
    Camera k;
    k.width = c->width; k.height = c->height;
    k.eye = mk3(c->position[0], c->position[1], c->position[2]);
    for (int i = 0; i < 9; ++i) k.O[i] = c->orientation[i];
    k.scale = c->scale; k.aspect = c->aspect;
    k.use_dof = c->use_dof; k.focal_distance = c->focal_distance; k.aperture_radius = c->aperture_radius;
    return k;
}

}  // namespace pt
